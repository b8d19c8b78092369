"""CUDA-graph cache for *engine-driven* decode steps (the public serving path).

``decode_loop.py`` is the fully device-resident loop used by the benchmark's device-timed phase.  The serving
engine (``LLMEngine`` -> ``LocalPipeline`` / ``ChainPipeline``) still owns scheduling on the host — sequences join
and leave, so the step metadata comes from the host every step — but the *device work* of a decode step is the
same kernel sequence every time for a given (batch size, block-table width, context bucket).  This module
captures that sequence once per key into a CUDA graph with static input buffers:

    host step  = one pinned H2D copy (packed metadata + token ids) -> graph replay -> one D2H of sampled ids

instead of ~430 eager launches (and their launch latency) per step.  The reference has no equivalent
(``mx.async_eval`` one-token look-ahead, shard/utils.py:180-186).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from ..ops.meta import BatchMeta


def _ctx_bucket(n: int) -> int:
    b = 512
    while b < n:
        b *= 2
    return b


class _Entry:
    def __init__(self, stage, B: int, mb: int, ctx_bucket: int, first: bool, last: bool):
        dev = stage.device
        self.B, self.mb = B, mb
        n = BatchMeta.packed_size(B, B, mb)
        self.flat = torch.zeros(n, dtype=torch.int32, device=dev)
        o = 6
        take = lambda k: (self.flat[o:o + k], o + k)
        pos, o = take(B)
        slots, o = take(B)
        cu, o = take(B + 1)
        ctx, o = take(B)
        last_idx, o = take(B)
        bt, o = take(B * mb)
        self.meta = BatchMeta(pos, slots, cu, ctx, bt.view(B, mb), last_idx, B, B, 1, ctx_bucket, stage.kv.page_size)
        H = stage.model.cfg.hidden_size
        self.x = (torch.zeros(B, dtype=torch.int64, device=dev) if first
                  else torch.zeros(B, H, dtype=stage.model.dtype, device=dev))
        self.temps = torch.zeros(B, dtype=torch.float32, device=dev)
        self.top_p = torch.ones(B, dtype=torch.float32, device=dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.out = None       # hidden [B, H] (non-last) or (tokens, logprobs) (last)
        self.uses = 0


class DecodeGraphCache:
    """Per-stage cache of captured decode graphs keyed by (B, block-table width, context bucket)."""

    WARM_USES = 2   # run eagerly this many times before paying for a capture

    def __init__(self, stage):
        self.stage = stage
        self.entries: Dict[Tuple[int, int, int], _Entry] = {}
        self.enabled = stage.device.type == "cuda" and stage.model.backend_name == "b200"
        self.replays = 0
        self.captures = 0

    def eligible(self, meta: BatchMeta, params=None) -> bool:
        if not self.enabled or meta.max_q_len != 1 or meta.num_tokens != meta.num_seqs:
            return False
        if params is not None and self.stage.model.spec.is_last:
            for p in params:
                if p.temperature != 0 or p.logit_bias or p.repetition_penalty not in (0, 1.0) or p.logprobs:
                    return False   # sampled / penalised requests take the eager path (stateful RNG step counter)
        return True

    def entry(self, B: int, mb: int, max_ctx: int) -> _Entry:
        key = (B, mb, _ctx_bucket(max_ctx))
        e = self.entries.get(key)
        if e is None:
            spec = self.stage.model.spec
            e = _Entry(self.stage, B, mb, key[2], spec.is_first, spec.is_last)
            self.entries[key] = e
            if len(self.entries) > 32:  # bound memory: drop the least used graph
                k = min((k for k in self.entries if k != key), key=lambda k: self.entries[k].uses)
                del self.entries[k]
        return e

    def _body(self, e: _Entry):
        st = self.stage
        out = st.model.forward(e.x, e.meta, st.kv)
        if st.model.spec.is_last:
            toks, lp, _, _ = st.model.ops.sample(out, e.temps, e.top_p, None, 0)
            return toks, lp
        return out

    @torch.inference_mode()
    def run(self, e: _Entry):
        """Inputs are already in ``e.flat`` / ``e.x``.  Returns the static output tensor(s)."""
        e.uses += 1
        if e.graph is None:
            if e.uses <= self.WARM_USES:
                return self._body(e)       # eager (also warms allocator / scratch before capture)
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                e.out = self._body(e)
            e.graph = g
            self.captures += 1
        e.graph.replay()
        self.replays += 1
        return e.out
