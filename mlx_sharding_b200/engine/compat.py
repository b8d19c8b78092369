"""Reference-shaped generation API: ``create_generate_step_with_grpc(stubs) -> generate_step(...)``.

Reference: ``shard/utils.py:111-188``.  The returned ``generate_step(prompt, model, temp, repetition_penalty,
repetition_context_size, top_p, logit_bias)`` is a generator yielding ``(token:int, logprobs:[vocab])``
exactly like the reference (one sequence, hub-and-spoke relay through the stubs, sampling on the
primary).  It exists for drop-in compatibility with code written against the reference; the serving
path uses ``LLMEngine`` instead.
"""
from __future__ import annotations

from typing import Dict, Generator, List, Optional, Sequence, Tuple

import torch

from ..ops import reference as R
from ..ops.meta import BatchMeta
from .kv_cache import PagedKVCache


def create_generate_step_with_grpc(grpc_stubs: Sequence = (), num_pages: int = 512, page_size: int = 64,
                                   wire_dtype: torch.dtype = torch.float16, seed: Optional[int] = None):
    stubs = list(grpc_stubs or [])

    def generate_step(prompt, model, temp: float = 0.0, repetition_penalty: Optional[float] = None,
                      repetition_context_size: Optional[int] = 20, top_p: float = 1.0,
                      logit_bias: Optional[Dict[int, float]] = None) -> Generator[Tuple[int, torch.Tensor], None, None]:
        if repetition_penalty and repetition_penalty < 0:
            raise ValueError(f"repetition_penalty must be a non-negative float, got {repetition_penalty}")
        for s in stubs:  # reference utils.py:122-124
            s.reset_cache()
        dev = model.device
        kv = PagedKVCache.for_model(model, num_pages, page_size)
        pages = list(range(1, num_pages))
        gen = torch.Generator(device="cpu")
        if seed is not None:
            gen.manual_seed(seed)
        ids = [int(t) for t in (prompt.tolist() if hasattr(prompt, "tolist") else prompt)]
        ctx: List[int] = ids[-repetition_context_size:] if repetition_context_size else []
        offset = 0

        def _step(tokens: List[int]):
            nonlocal offset, ctx
            meta = BatchMeta.build([len(tokens)], [offset], [pages], page_size, device=dev)
            x = model.forward(torch.tensor(tokens, dtype=torch.int64, device=dev), meta, kv, all_logits=True)
            offset += len(tokens)
            if stubs:
                x = x.unsqueeze(0)
                for s in stubs:
                    if x.is_floating_point():
                        x = x.to(wire_dtype)
                    x = s.send_tensor(x, dev)
                x = x[0]
            logits = x[-1].float().cpu()
            if repetition_penalty:
                R.apply_repetition_penalty_(logits, torch.tensor(ctx, dtype=torch.int64), repetition_penalty)
            if logit_bias:
                idx = torch.tensor(list(logit_bias.keys()), dtype=torch.int64)
                logits[idx] += torch.tensor(list(logit_bias.values()), dtype=torch.float32)
            toks, _, _, _ = R.sample(logits[None], torch.tensor([float(temp)]), torch.tensor([float(top_p)]), gen)
            logprobs = logits - torch.logsumexp(logits, dim=-1)
            tok = int(toks[0])
            if repetition_penalty:
                ctx.append(tok)
                if repetition_context_size and len(ctx) > repetition_context_size:
                    ctx = ctx[-repetition_context_size:]
            return tok, logprobs

        y, lp = _step(ids)
        while True:
            yield y, lp
            y, lp = _step([y])

    return generate_step
