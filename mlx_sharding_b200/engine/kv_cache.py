"""Paged KV cache: per-layer K/V pools ``[pages, kv_heads, page_size, d]`` + a page allocator.

Reference: upstream ``KVCache`` is one contiguous ``[1, n_kv, S, d]`` buffer per layer grown
in 256-token steps, a single global list per server (``shard/server/server.py:9-23``) — one sequence
system-wide.  Here every sequence owns a block table into a shared pool, so many sequences /
micro-batches can be in flight and prefill can be chunked.  K and V may have different head dims
(DeepSeek-V2 MLA: 192 / 128; reference deepseek_v2.py:120-125).

Pool geometry is identical on every stage (same ``num_pages`` / ``page_size``) so one block table built
by the scheduler on stage 0 is valid on all stages.
"""
from __future__ import annotations

from typing import Dict, List

import torch


class PageAllocator:
    """Free-list page allocator; page 0 is reserved as the null page (padding slots write there)."""

    def __init__(self, num_pages: int):
        if num_pages < 2:
            raise ValueError("need at least 2 pages")
        self.num_pages = num_pages
        self._free: List[int] = list(range(num_pages - 1, 0, -1))

    @property
    def num_free(self) -> int:
        return len(self._free)

    def alloc(self, n: int) -> List[int]:
        if n > len(self._free):
            raise MemoryError(f"KV pool exhausted: want {n} pages, {len(self._free)} free")
        out = [self._free.pop() for _ in range(n)]
        return out

    def free(self, pages: List[int]):
        self._free.extend(reversed(pages))


class PagedKVCache:
    def __init__(self, num_layers: int, num_pages: int, page_size: int, kv_heads: int, d_k: int, d_v: int,
                 dtype=torch.bfloat16, device="cpu"):
        self.num_layers, self.num_pages, self.page_size = num_layers, num_pages, page_size
        self.kv_heads, self.d_k, self.d_v = kv_heads, d_k, d_v
        self.k = [torch.zeros(num_pages, kv_heads, page_size, d_k, dtype=dtype, device=device)
                  for _ in range(num_layers)]
        self.v = [torch.zeros(num_pages, kv_heads, page_size, d_v, dtype=dtype, device=device)
                  for _ in range(num_layers)]

    @staticmethod
    def bytes_per_page(num_layers, page_size, kv_heads, d_k, d_v, itemsize=2) -> int:
        return num_layers * page_size * kv_heads * (d_k + d_v) * itemsize

    @classmethod
    def for_model(cls, model, num_pages: int, page_size: int = 64):
        L, hk, dk, dv = model.kv_geometry()
        return cls(L, num_pages, page_size, hk, dk, dv, model.dtype, model.device)


class SequenceTable:
    """Host-side block tables: sequence id -> (pages, length)."""

    def __init__(self, allocator: PageAllocator, page_size: int):
        self.alloc = allocator
        self.page_size = page_size
        self.pages: Dict[int, List[int]] = {}
        self.length: Dict[int, int] = {}

    def add(self, seq_id: int):
        self.pages[seq_id] = []
        self.length[seq_id] = 0

    def reserve(self, seq_id: int, new_tokens: int):
        """Make sure pages exist for ``new_tokens`` more positions (does not advance the length)."""
        need = (self.length[seq_id] + new_tokens + self.page_size - 1) // self.page_size
        have = len(self.pages[seq_id])
        if need > have:
            self.pages[seq_id].extend(self.alloc.alloc(need - have))

    def advance(self, seq_id: int, n: int):
        self.length[seq_id] += n

    def release(self, seq_id: int):
        """Free a sequence's pages (the reference's ``ResetCache`` semantics, server.py:59-71)."""
        if seq_id in self.pages:
            self.alloc.free(self.pages.pop(seq_id))
            self.length.pop(seq_id, None)

    def release_all(self):
        for s in list(self.pages):
            self.release(s)
