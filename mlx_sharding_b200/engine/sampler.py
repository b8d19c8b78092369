"""Per-request sampling parameters + the batched sampler front-end.

Reference: the ``sample()`` closure and repetition-penalty bookkeeping inside
``create_generate_step_with_grpc`` (shard/utils.py:126-139,152-177):
logit_bias add -> repetition penalty over the last ``repetition_context_size`` tokens ->
``logprobs = logits - logsumexp`` -> argmax | nucleus | categorical.

In this engine sampling runs on the *last* stage, on device, for the whole micro-batch at once; only
token ids (+ optional top-k logprobs, k <= 10) travel back to stage 0.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch


@dataclass
class SamplingParams:
    temperature: float = 0.0
    top_p: float = 1.0
    repetition_penalty: float = 1.0
    repetition_context_size: int = 20
    logit_bias: Optional[Dict[int, float]] = None
    logprobs: int = 0               # number of top logprobs to return (0 = none, <= 10)
    seed: Optional[int] = None

    def validate(self):
        if self.temperature < 0:
            raise ValueError("temperature must be a non-negative float")
        if not (0 <= self.top_p <= 1):
            raise ValueError("top_p must be a float between 0 and 1")
        if self.repetition_penalty < 0:
            raise ValueError("repetition_penalty must be a non-negative float")
        if self.repetition_context_size < 0:
            raise ValueError("repetition_context_size must be a non-negative integer")
        if not (0 <= self.logprobs <= 10):
            raise ValueError("logprobs must be between 1 and 10")


@dataclass
class SampleOutput:
    tokens: torch.Tensor                      # int64 [B]
    logprobs: torch.Tensor                    # fp32 [B] logprob of the chosen token
    top_ids: Optional[torch.Tensor] = None    # int64 [B, k]
    top_logprobs: Optional[torch.Tensor] = None


class Sampler:
    """Batched sampler; ``ops`` is the backend module (reference or b200)."""

    MAX_CTX = 64    # static width of the repetition-context / bias tables on device
    MAX_BIAS = 64

    def __init__(self, ops, device="cpu", seed: int = 0):
        self.ops = ops
        self.device = torch.device(device)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)

    def __call__(self, logits: torch.Tensor, params: List[SamplingParams],
                 contexts: Optional[List[List[int]]] = None) -> SampleOutput:
        """``logits`` fp32 ``[B, V]``; ``contexts[b]`` = recent token ids for the repetition penalty."""
        B = logits.shape[0]
        assert len(params) == B
        logits = logits.float()
        need_pen = any((p.repetition_penalty not in (0, 1.0) and p.repetition_context_size > 0) or p.logit_bias
                       for p in params)
        if need_pen:
            logits = logits.clone()
            C = max(1, min(self.MAX_CTX, max(p.repetition_context_size for p in params)))
            ctx = torch.full((B, C), -1, dtype=torch.int32)
            pen = torch.ones(B, dtype=torch.float32)
            nb = max(1, max(len(p.logit_bias or {}) for p in params))
            bidx = torch.full((B, nb), -1, dtype=torch.int32)
            bval = torch.zeros(B, nb, dtype=torch.float32)
            for b, p in enumerate(params):
                if p.repetition_penalty not in (0, 1.0) and contexts is not None and p.repetition_context_size > 0:
                    c = contexts[b][-min(p.repetition_context_size, C):]
                    if c:
                        ctx[b, : len(c)] = torch.tensor(c, dtype=torch.int32)
                    pen[b] = p.repetition_penalty
                for j, (k, v) in enumerate((p.logit_bias or {}).items()):
                    bidx[b, j], bval[b, j] = int(k), float(v)
            self.ops.apply_penalties_(logits, ctx.to(self.device), pen.to(self.device),
                                      bidx.to(self.device), bval.to(self.device))
        temps = torch.tensor([p.temperature for p in params], dtype=torch.float32, device=self.device)
        top_p = torch.tensor([p.top_p for p in params], dtype=torch.float32, device=self.device)
        k = max(p.logprobs for p in params)
        toks, lp, ti, tl = self.ops.sample(logits, temps, top_p, generator=self.gen, top_logprobs=k)
        return SampleOutput(toks, lp, ti, tl)
