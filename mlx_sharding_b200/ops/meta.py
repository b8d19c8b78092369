"""Ragged-batch step metadata shared by every stage of the pipeline.

A step processes ``T`` tokens belonging to ``B`` sequences, flattened along one axis (no padding):
decode micro-batches have one token per sequence, prefill chunks have many.  The reference has no
such structure — it is hard-wired to one sequence (``y[None]``, shard/utils.py:158) with a dense
additive mask (llama.py:48-53); this is the piece that enables micro-batched scheduling and chunked
prefill.

The whole descriptor packs into one int32 vector (``pack``/``unpack``) so it can ride along the
stage-to-stage hand-off in front of the hidden-state payload.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np
import torch


@dataclass
class BatchMeta:
    positions: torch.Tensor      # int32 [T]   absolute position of each token in its sequence
    slot_mapping: torch.Tensor   # int32 [T]   flat KV slot (page * page_size + offset) written by each token
    cu_seqlens: torch.Tensor     # int32 [B+1] offsets of each sequence's query tokens in [0, T]
    context_lens: torch.Tensor   # int32 [B]   KV length of each sequence *after* this step's append
    block_tables: torch.Tensor   # int32 [B, max_blocks] page ids
    last_idx: torch.Tensor       # int32 [B]   row in [0, T) of the last token of each sequence
    num_tokens: int
    num_seqs: int
    max_q_len: int
    max_ctx_len: int
    page_size: int
    fresh: bool = False          # host-side hint: no sequence of this step has cached context (every prompt starts at position 0)

    @property
    def is_decode(self) -> bool:
        return self.max_q_len == 1

    def to(self, device, non_blocking: bool = False) -> "BatchMeta":
        mv = lambda t: t.to(device, non_blocking=non_blocking)
        return BatchMeta(mv(self.positions), mv(self.slot_mapping), mv(self.cu_seqlens), mv(self.context_lens),
                         mv(self.block_tables), mv(self.last_idx), self.num_tokens, self.num_seqs,
                         self.max_q_len, self.max_ctx_len, self.page_size, self.fresh)

    # ------------------------------------------------------------------ wire format
    def pack(self) -> torch.Tensor:
        B, T = self.num_seqs, self.num_tokens
        mb = self.block_tables.shape[1] if self.block_tables.dim() == 2 else 0
        head = torch.tensor([T, B, self.max_q_len, self.max_ctx_len, self.page_size, mb], dtype=torch.int32)
        parts = [head, self.positions.cpu().int(), self.slot_mapping.cpu().int(), self.cu_seqlens.cpu().int(),
                 self.context_lens.cpu().int(), self.last_idx.cpu().int(),
                 self.block_tables.cpu().int().reshape(-1)]
        return torch.cat(parts)

    @staticmethod
    def packed_size(T: int, B: int, max_blocks: int) -> int:
        return 6 + 2 * T + (B + 1) + 2 * B + B * max_blocks

    @staticmethod
    def unpack(buf: torch.Tensor) -> "BatchMeta":
        h = buf[:6].tolist()
        T, B, mq, mc, ps, mb = h
        o = 6
        def take(n):
            nonlocal o
            t = buf[o:o + n]
            o += n
            return t
        positions, slots = take(T), take(T)
        cu, ctx, last = take(B + 1), take(B), take(B)
        bt = take(B * mb).reshape(B, mb)
        return BatchMeta(positions, slots, cu, ctx, bt, last, T, B, mq, mc, ps)

    # ------------------------------------------------------------------ builders
    @staticmethod
    def build(q_lens: Sequence[int], ctx_before: Sequence[int], block_tables: Sequence[Sequence[int]],
              page_size: int, device="cpu", pad_blocks_to: int = 0) -> "BatchMeta":
        """Build from per-sequence query lengths, already-cached lengths and page lists."""
        B = len(q_lens)
        if B and max(q_lens) == 1 and min(q_lens) == 1:
            # decode step: one token per sequence — no per-token inner loop (host fast path, once per generated token)
            pos = [int(c) for c in ctx_before]
            slots = [block_tables[b][p // page_size] * page_size + p % page_size for b, p in enumerate(pos)]
            cu, last, ctx_after = list(range(B + 1)), list(range(B)), [p + 1 for p in pos]
            return BatchMeta._finish_build(pos, slots, cu, ctx_after, last, q_lens, block_tables, page_size, device, pad_blocks_to)
        pos: List[int] = []
        slots: List[int] = []
        cu = [0]
        last = []
        ctx_after = []
        for b in range(B):
            ql, c0 = int(q_lens[b]), int(ctx_before[b])
            pages = block_tables[b]
            for j in range(ql):
                p = c0 + j
                pos.append(p)
                slots.append(pages[p // page_size] * page_size + p % page_size)
            cu.append(cu[-1] + ql)
            last.append(cu[-1] - 1)
            ctx_after.append(c0 + ql)
        m = BatchMeta._finish_build(pos, slots, cu, ctx_after, last, q_lens, block_tables, page_size, device, pad_blocks_to)
        m.fresh = B > 0 and all(int(c) == 0 for c in ctx_before)
        return m

    @staticmethod
    def _finish_build(pos, slots, cu, ctx_after, last, q_lens, block_tables, page_size, device, pad_blocks_to) -> "BatchMeta":
        B = len(q_lens)
        mb = max(max((len(p) for p in block_tables), default=1), pad_blocks_to, 1)
        # one tensor construction for the whole (zero-padded) table: this runs on the host once per step per group
        pad = [0] * mb
        # (numpy converts nested Python lists about twice as fast as torch.tensor)
        i32 = lambda x: torch.from_numpy(np.array(x, dtype=np.int32))
        bt = i32([list(p) + pad[len(p):] for p in block_tables]).reshape(B, mb)
        m = BatchMeta(i32(pos), i32(slots), i32(cu), i32(ctx_after), bt, i32(last), len(pos), B,
                      max(q_lens) if B else 0, max(ctx_after) if B else 0, page_size)
        return m.to(device) if str(device) != "cpu" else m
