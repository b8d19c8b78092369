"""Pipeline runtimes: how one scheduler step (``StepInput``) travels through the stages.

* ``LocalPipeline``  – all stages in this process (1 GPU, or CPU tests).
* ``ChainPipeline``  – **direct chain** ``stage i -> i+1`` across processes; the last stage samples on
  device and only token ids (+ logprobs) return to stage 0.  Replaces the reference's hub-and-spoke
  relay that bounces every hidden state through the primary and ships full ``[1,T,V]`` logits back
  (shard/utils.py:162-166, server/server.py:36-48).
* ``worker_loop``    – what a non-first stage process runs.

Data plane and control plane come from a ``ChainTransport`` (``parallel/transport.py``): gloo on CPU
(BASELINE config 1), NCCL p2p on GPUs (the baseline hand-off), or the fused P2P store path
(``parallel/p2p_fused.py``) where the producing kernel writes straight into the peer's inbox.
"""
from __future__ import annotations

import logging
from collections import deque
from typing import List, Optional

import torch

from ..engine.core import StepInput, StepOutput
from ..engine.kv_cache import PagedKVCache
from ..engine.sampler import Sampler, SamplingParams
from ..ops.meta import BatchMeta

log = logging.getLogger(__name__)


class StageExecutor:
    """A stage model + its paged KV pool (+ the sampler on the last stage)."""

    def __init__(self, model, num_pages: int, page_size: int = 64, seed: int = 0):
        self.model = model
        self.device = model.device
        self.kv = PagedKVCache.for_model(model, num_pages, page_size)
        self.sampler = Sampler(model.ops, self.device, seed) if model.spec.is_last else None
        import os

        from ..utils.tracing import StageTimer

        self.timer = StageTimer(enabled=os.environ.get("MLXB200_STAGE_TIMING", "0") == "1")

    @torch.inference_mode()
    def forward(self, x: torch.Tensor, meta: BatchMeta, all_logits: bool = False) -> torch.Tensor:
        from ..utils.tracing import nvtx_range

        with self.timer.measure(), nvtx_range(f"stage[{self.model.spec.start_layer},{self.model.spec.end_layer}) T={meta.num_tokens}"):
            return self.model.forward(x, meta, self.kv, all_logits)

    @torch.inference_mode()
    def sample(self, logits: torch.Tensor, inp_params, contexts) -> StepOutput:
        so = self.sampler(logits, inp_params, contexts)
        return StepOutput(so.tokens.tolist(), so.logprobs.tolist(),
                          None if so.top_ids is None else so.top_ids.tolist(),
                          None if so.top_logprobs is None else so.top_logprobs.tolist())


def _pinned_to(t: torch.Tensor, device) -> torch.Tensor:
    """Host tensor -> device through pinned memory (async copy on the current stream)."""
    if torch.device(device).type != "cuda":
        return t
    return t.pin_memory().to(device, non_blocking=True)


def stage_inputs(inp: StepInput, device):
    """H2D transfer of one step's inputs (token ids + packed metadata); returns (tokens, meta, bytes)."""
    m = inp.meta
    if torch.device(device).type != "cuda":
        return inp.tokens, m, 0
    parts = [m.positions, m.slot_mapping, m.cu_seqlens, m.context_lens, m.last_idx, m.block_tables.reshape(-1)]
    sizes = [p.numel() for p in parts]
    flat = _pinned_to(torch.cat([p.to(torch.int32) for p in parts]), device)
    v = list(torch.split(flat, sizes))
    meta = BatchMeta(v[0], v[1], v[2], v[3], v[5].view(m.block_tables.shape), v[4], m.num_tokens, m.num_seqs,
                     m.max_q_len, m.max_ctx_len, m.page_size)
    toks = _pinned_to(inp.tokens, device)
    return toks, meta, flat.numel() * 4 + toks.numel() * 8


class LocalPipeline:
    """All stages in-process, executed back to back."""

    def __init__(self, stages: List[StageExecutor]):
        assert stages[0].model.spec.is_first and stages[-1].model.spec.is_last
        self.stages = stages
        self.num_stages = 1  # one executor thread: a single micro-batch group keeps it busy
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        from .graph_decode import DecodeGraphCache

        self.gcache = DecodeGraphCache(stages[0]) if len(stages) == 1 else None

    @classmethod
    def from_models(cls, models, num_pages: int, page_size: int = 64, seed: int = 0):
        return cls([StageExecutor(m, num_pages, page_size, seed) for m in models])

    def _submit_graph(self, inp: StepInput) -> StepOutput:
        """Steady-state decode: one pinned H2D copy of (metadata, token ids), one graph replay, one D2H."""
        m = inp.meta
        e = self.gcache.entry(m.num_seqs, m.block_tables.shape[1], m.max_ctx_len)
        packed = m.pack().pin_memory()
        toks_h = inp.tokens.pin_memory()
        e.flat.copy_(packed, non_blocking=True)
        e.x.copy_(toks_h, non_blocking=True)
        toks, lp = self.gcache.run(e)
        self.h2d_bytes += packed.numel() * 4 + toks_h.numel() * 8
        self.d2h_bytes += m.num_seqs * 12
        return StepOutput(toks.tolist(), lp.tolist())

    def submit(self, inp: StepInput):
        if self.gcache is not None and self.gcache.eligible(inp.meta, inp.params):
            return self._submit_graph(inp)
        x, meta0, nbytes = stage_inputs(inp, self.stages[0].device)
        self.h2d_bytes += nbytes
        for i, st in enumerate(self.stages):
            meta = meta0 if i == 0 or st.device == self.stages[0].device else inp.meta.to(st.device)
            x = st.forward(x.to(st.device), meta)
        out = self.stages[-1].sample(x, inp.params, inp.contexts)
        self.d2h_bytes += len(out.tokens) * 12  # int64 token id + fp32 logprob per sequence
        return out

    def wait(self, handle) -> StepOutput:
        return handle

    def reset(self):
        pass


# -------------------------------------------------------------------------------------------------
# Cross-process chain
# -------------------------------------------------------------------------------------------------
_GREEDY = SamplingParams()   # shared instance for the compact control message


def _plain_greedy(p: SamplingParams) -> bool:
    return (p.temperature == 0 and p.top_p == 1.0 and p.repetition_penalty in (0, 1.0) and not p.logit_bias and p.logprobs == 0
            and p.seed is None)


def _ctrl_of(inp: StepInput) -> dict:
    """Per-step control message that travels down the chain (pickled, gloo).  The common case — every sequence decodes greedily
    with default parameters — is sent as a count instead of B parameter objects + B empty contexts (half the bytes and less
    than half the (un)pickling time per hop, which sits on every stage's critical path)."""
    if all(_plain_greedy(p) for p in inp.params):
        return dict(kind="step", group=inp.group, meta=inp.meta.pack(), params=len(inp.params), contexts=None,
                    is_prefill=inp.is_prefill)
    return dict(kind="step", group=inp.group, meta=inp.meta.pack(), params=inp.params, contexts=inp.contexts,
                is_prefill=inp.is_prefill)


def _params_of(ctrl: dict):
    """(params, contexts) of a control message (expands the compact greedy form)."""
    p = ctrl["params"]
    if isinstance(p, int):
        return [_GREEDY] * p, [[]] * p
    return p, ctrl["contexts"]


class ChainPipeline:
    """Stage-0 side of the cross-process chain."""

    def __init__(self, stage: StageExecutor, transport):
        self.stage = stage
        self.tp = transport
        self.num_stages = transport.world_size
        self.rank = transport.rank
        assert self.rank == 0 and stage.model.spec.is_first
        self._pending = deque()
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        from .graph_decode import DecodeGraphCache

        self.gcache = DecodeGraphCache(stage)

    def submit(self, inp: StepInput):
        tp = self.tp
        if self.num_stages > 1 and self.gcache.eligible(inp.meta, None):
            m = inp.meta
            e = self.gcache.entry(m.num_seqs, m.block_tables.shape[1], m.max_ctx_len)
            packed, toks_h = m.pack().pin_memory(), inp.tokens.pin_memory()
            e.flat.copy_(packed, non_blocking=True)
            e.x.copy_(toks_h, non_blocking=True)
            x = self.gcache.run(e)
            nbytes = packed.numel() * 4 + toks_h.numel() * 8
        else:
            toks, meta, nbytes = stage_inputs(inp, self.stage.device)
            x = self.stage.forward(toks, meta)
        self.h2d_bytes += nbytes
        self.d2h_bytes += len(inp.seq_ids) * 12
        if self.num_stages == 1:
            return ("local", self.stage.sample(x, inp.params, inp.contexts))
        tp.send_ctrl(_ctrl_of(inp), 1)
        tp.send_tensor(x, 1, slot=inp.group)
        h = tp.irecv_ctrl(self.num_stages - 1)  # result comes straight from the last stage
        return ("remote", h)

    def wait(self, handle) -> StepOutput:
        kind, h = handle
        if kind == "local":
            return h
        res = self.tp.wait_ctrl(h)
        if isinstance(res, dict) and res.get("error"):
            raise RuntimeError(f"stage {res.get('rank')} failed: {res['error']}")
        return StepOutput(**res)

    def reset(self):
        pass

    def shutdown(self):
        if self.num_stages > 1:
            self.tp.send_ctrl(dict(kind="shutdown"), 1)
            if hasattr(self.tp, "flush"):
                self.tp.flush()


def worker_loop(stage: StageExecutor, transport):
    """Non-first stage: ``recv (ctrl, hidden) -> forward -> send`` until a shutdown message.
    Errors are reported down the chain to stage 0 instead of wedging the pipeline (the reference
    replies ``success=False`` strings, server/server.py:55-57)."""
    tp = transport
    rank, world = tp.rank, tp.world_size
    last = rank == world - 1
    H = stage.model.cfg.hidden_size
    from .graph_decode import DecodeGraphCache

    gcache = DecodeGraphCache(stage)
    while True:
        ctrl = tp.recv_ctrl(rank - 1)
        if ctrl["kind"] == "shutdown":
            if not last:
                tp.send_ctrl(ctrl, rank + 1)
            if hasattr(tp, "flush"):
                tp.flush()
            return
        if ctrl.get("error"):
            # propagate the failure to stage 0 (drain our payload first to stay in lockstep)
            if last:
                tp.send_ctrl(ctrl, 0)
            else:
                tp.send_ctrl(ctrl, rank + 1)
            continue
        meta = BatchMeta.unpack(ctrl["meta"])
        params, contexts = _params_of(ctrl)
        graphed = gcache.eligible(meta, params)
        if graphed:
            e = gcache.entry(meta.num_seqs, meta.block_tables.shape[1], meta.max_ctx_len)
            x = tp.recv_tensor((meta.num_tokens, H), stage.model.dtype, rank - 1, slot=ctrl["group"], out=e.x)
        else:
            x = tp.recv_tensor((meta.num_tokens, H), stage.model.dtype, rank - 1, slot=ctrl["group"])
        try:
            if graphed:
                e.flat.copy_(ctrl["meta"].pin_memory(), non_blocking=True)
                out = gcache.run(e)
                if last:
                    toks, lp = out
                    res = StepOutput(toks.tolist(), lp.tolist())
                    tp.send_ctrl(dict(tokens=res.tokens, logprobs=res.logprobs, top_ids=None, top_logprobs=None), 0)
                    continue
            else:
                out = stage.forward(x, meta.to(stage.device))
            if last:
                res = stage.sample(out, params, contexts)
                tp.send_ctrl(dict(tokens=res.tokens, logprobs=res.logprobs, top_ids=res.top_ids,
                                  top_logprobs=res.top_logprobs), 0)
            else:
                tp.send_ctrl(ctrl, rank + 1)
                tp.send_tensor(out, rank + 1, slot=ctrl["group"])
        except Exception as e:  # noqa: BLE001
            log.exception("stage %d failed", rank)
            err = dict(kind="step", error=f"{type(e).__name__}: {e}", rank=rank)
            tp.send_ctrl(err, 0 if last else rank + 1)
