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

from typing import Dict, List, Optional

import torch


class PageAllocator:
    """Free-list page allocator; page 0 is reserved as the null page (padding slots write there).  An optional
    ``reclaim(n)`` callback (the prefix cache) is asked to give back ``n`` pages before the pool is declared exhausted."""

    def __init__(self, num_pages: int):
        if num_pages < 2:
            raise ValueError("need at least 2 pages")
        self.num_pages = num_pages
        self._free: List[int] = list(range(num_pages - 1, 0, -1))
        self.reclaim = None          # callable(n) -> number of pages returned to the free list
        self.reclaimable = lambda: 0

    @property
    def num_free(self) -> int:
        return len(self._free) + self.reclaimable()

    def alloc(self, n: int) -> List[int]:
        if n > len(self._free) and self.reclaim is not None:
            self.reclaim(n - len(self._free))
        if n > len(self._free):
            raise MemoryError(f"KV pool exhausted: want {n} pages, {len(self._free)} free")
        out = [self._free.pop() for _ in range(n)]
        return out

    def free(self, pages: List[int]):
        self._free.extend(reversed(pages))


class PrefixCache:
    """Content-addressed KV pages of prompt prefixes (automatic prefix caching; no reference counterpart — the reference
    resets its single cache on every request, shard/utils.py:122-124).

    A *full* page of prompt tokens is identified by the chain digest ``H(parent digest, its page_size token ids)``, so equal
    digests mean equal token prefixes and therefore bit-identical K/V on every stage.  A new request re-uses the longest
    chain of cached pages (always leaving at least its last prompt token to compute, whose logits it needs) and only
    prefills the rest.  Pages are reference counted; unreferenced cached pages stay resident in LRU order and are evicted
    only when the allocator runs dry.  Shared pages are read-only by construction: a sequence appends at positions beyond
    the matched (page-aligned) prefix, i.e. into pages it owns alone."""

    def __init__(self, allocator: PageAllocator, page_size: int):
        import collections

        self.alloc, self.page_size = allocator, page_size
        self.by_digest: Dict[bytes, int] = {}
        self.digest_of: Dict[int, bytes] = {}
        self.refs: Dict[int, int] = {}
        self.lru: "collections.OrderedDict[int, None]" = collections.OrderedDict()   # cached pages nobody references
        self.hits = self.misses = self.evictions = 0
        allocator.reclaim = self.evict
        allocator.reclaimable = lambda: len(self.lru)

    def _digests(self, tokens, n_pages: int):
        import hashlib
        import struct

        d = b""
        for i in range(n_pages):
            chunk = tokens[i * self.page_size:(i + 1) * self.page_size]
            d = hashlib.blake2b(d + struct.pack(f"<{len(chunk)}q", *chunk), digest_size=16).digest()
            yield d

    def match(self, tokens) -> List[int]:
        """Longest chain of cached pages covering a proper prefix of ``tokens`` (references are taken for the caller)."""
        limit = (len(tokens) - 1) // self.page_size        # never the page holding the last prompt token
        pages = []
        for d in self._digests(tokens, limit):
            p = self.by_digest.get(d)
            if p is None:
                break
            pages.append(p)
        for p in pages:
            self.refs[p] = self.refs.get(p, 0) + 1
            self.lru.pop(p, None)
        self.hits += len(pages)
        self.misses += limit - len(pages)
        return pages

    def insert(self, tokens, pages: List[int], n_tokens: int, already: int) -> int:
        """Register the full pages of ``tokens[:n_tokens]`` held in ``pages`` (the first ``already`` are known).  Returns the
        new count of registered pages of this sequence.  The owner keeps its reference."""
        full = n_tokens // self.page_size
        if full <= already:
            return already
        for i, d in enumerate(self._digests(tokens, full)):
            if i < already:
                continue
            p = pages[i]
            if d in self.by_digest or p in self.digest_of:
                continue                 # same content cached by someone else meanwhile: ours stays a private page
            self.by_digest[d] = p
            self.digest_of[p] = d
            self.refs[p] = self.refs.get(p, 0) + 1
        return full

    def release(self, pages: List[int]) -> List[int]:
        """Drop one reference of every cached page in ``pages``; returns the pages that are *not* cached (caller frees them)."""
        private = []
        for p in pages:
            if p not in self.digest_of:
                private.append(p)
                continue
            self.refs[p] -= 1
            if self.refs[p] == 0:
                self.lru[p] = None
        return private

    def evict(self, n: int) -> int:
        done = 0
        while done < n and self.lru:
            p, _ = self.lru.popitem(last=False)
            del self.by_digest[self.digest_of.pop(p)]
            del self.refs[p]
            self.alloc._free.append(p)
            done += 1
        self.evictions += done
        return done


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
        self.prefix: Optional[PrefixCache] = None

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
        """Free a sequence's pages (the reference's ``ResetCache`` semantics, server.py:59-71); pages registered in the prefix
        cache only lose this sequence's reference."""
        if seq_id in self.pages:
            pages = self.pages.pop(seq_id)
            self.alloc.free(self.prefix.release(pages) if self.prefix is not None else pages)
            self.length.pop(seq_id, None)

    def release_all(self):
        for s in list(self.pages):
            self.release(s)
