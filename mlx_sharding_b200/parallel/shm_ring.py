"""Single-node control plane of the pipeline: a shared-memory launch ring.

Stage 0 owns scheduling; every other stage has to execute *exactly the same sequence* of steps (same groups, same order — the
device-side flag waits of the fused hand-off are only deadlock-free when all ranks agree on one global launch order).  The
reference ships that decision as one blocking gRPC call per stage per token (shard/utils.py:162-164); round 1 of this repo
shipped it as a pickled control frame over gloo per stage per step.  All stages of a fused pipeline live on one NVSwitch domain,
i.e. on one host, so the launch order can simply be *published in shared memory*:

* rank 0 appends one fixed-header record per step (kind, group, the step's packed metadata + sampling block, ~4 KB for a decode
  micro-batch) to a single-producer ring in ``/dev/shm`` — a ``memcpy`` and two stores, about a microsecond;
* every worker polls the ring head (plain loads), copies the payload out and launches the step — no system call, no pickle,
  no collective on the control path;
* a status page carries per-rank consumed counters (back-pressure), error words / messages and a shutdown flag.

The file is unlinked as soon as every rank has mapped it, so nothing is left behind even if a process dies.
Ordering relies on x86-TSO store ordering (payload first, sequence word last) — the same assumption NCCL's and gloo's shared
memory transports make.
"""
from __future__ import annotations

import mmap
import os
import time
import uuid
from typing import Optional, Tuple

import numpy as np

KIND_STEP = 1        # one scheduler step (payload = step block, see graph_decode.StepLayout)
KIND_SHUTDOWN = 2
KIND_NOP = 3

_HDR_BYTES = 8192
_SLOT_HDR_WORDS = 8   # int32 words: seq_lo, seq_hi, kind, group, nwords, flags, 0, 0
_MAX_RANKS = 64
_ERR_BYTES = 240


class ShmControl:
    """Launch ring + status page shared by all ranks of one node.  ``create`` on rank 0, ``attach`` elsewhere."""

    def __init__(self, path: str, rank: int, world: int, slots: int, slot_bytes: int, create: bool):
        assert world <= _MAX_RANKS
        self.path, self.rank, self.world, self.slots = path, rank, world, slots
        self.slot_words = slot_bytes // 4
        total = _HDR_BYTES + slots * slot_bytes
        if create:
            fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            os.ftruncate(fd, total)
        else:
            fd = os.open(path, os.O_RDWR)
        self._mm = mmap.mmap(fd, total)
        os.close(fd)
        self._u64 = np.frombuffer(self._mm, dtype=np.uint64, count=_HDR_BYTES // 8)
        self._u8 = np.frombuffer(self._mm, dtype=np.uint8, count=_HDR_BYTES)
        self._ring = np.frombuffer(self._mm, dtype=np.int32, offset=_HDR_BYTES).reshape(slots, self.slot_words)
        self._ring_seq = np.frombuffer(self._mm, dtype=np.uint64, offset=_HDR_BYTES).reshape(slots, self.slot_words // 2)
        self._tail = 0          # consumer: records consumed by this rank
        if create:
            self._u64[:] = 0
            self._u64[2] = slots
            self._u64[3] = slot_bytes

    # u64 header words: 0 head | 1 shutdown | 2 slots | 3 slot_bytes | 8+r tail[r] | 80+r error seq[r]; error text at 2048+256*r
    @classmethod
    def create(cls, world: int, slots: int = 128, slot_bytes: int = 256 << 10) -> "ShmControl":
        base = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
        return cls(os.path.join(base, f"mlxb200_ctl_{os.getpid()}_{uuid.uuid4().hex[:8]}"), 0, world, slots, slot_bytes, True)

    @classmethod
    def attach(cls, path: str, rank: int, world: int) -> "ShmControl":
        fd = os.open(path, os.O_RDONLY)
        hdr = np.frombuffer(os.pread(fd, 32, 0), dtype=np.uint64)
        os.close(fd)
        return cls(path, rank, world, int(hdr[2]), int(hdr[3]), False)

    def unlink(self):
        try:
            os.unlink(self.path)
        except OSError:
            pass

    @property
    def max_payload_words(self) -> int:
        return self.slot_words - _SLOT_HDR_WORDS

    # ------------------------------------------------------------------------------------------ producer (rank 0)
    def publish(self, kind: int, group: int, payload: Optional[np.ndarray] = None, timeout_s: float = 60.0) -> int:
        """Append a record; returns its sequence number (1-based).  Blocks while the slowest consumer is a full ring behind."""
        head = int(self._u64[0])
        n = 0 if payload is None else int(payload.size)
        if n > self.max_payload_words:
            raise ValueError(f"step block of {n * 4} bytes exceeds the control-ring slot ({self.max_payload_words * 4} bytes): "
                             "lower max_prefill_tokens or raise slot_bytes")
        if self.world > 1:
            t0 = None
            while head - int(self._u64[8 + 1: 8 + self.world].min()) >= self.slots:
                if t0 is None:
                    t0 = time.monotonic()
                elif time.monotonic() - t0 > timeout_s:
                    raise TimeoutError("control ring full: a stage stopped consuming launch records")
                time.sleep(0.0001)
        slot = self._ring[head % self.slots]
        if n:
            slot[_SLOT_HDR_WORDS:_SLOT_HDR_WORDS + n] = payload
        slot[2], slot[3], slot[4], slot[5] = kind, group, n, 0
        self._ring_seq[head % self.slots, 0] = head + 1     # publication point (after the payload: x86 TSO)
        self._u64[0] = head + 1
        return head + 1

    # ------------------------------------------------------------------------------------------ consumers
    def poll(self) -> Optional[Tuple[int, int, int, np.ndarray]]:
        """Next record ``(seq, kind, group, payload copy)`` or None if nothing new."""
        idx = self._tail
        s = idx % self.slots
        if int(self._ring_seq[s, 0]) != idx + 1:
            return None
        slot = self._ring[s]
        kind, group, n = int(slot[2]), int(slot[3]), int(slot[4])
        payload = slot[_SLOT_HDR_WORDS:_SLOT_HDR_WORDS + n].copy()
        self._tail = idx + 1
        self._u64[8 + self.rank] = idx + 1
        return idx + 1, kind, group, payload

    def next(self, timeout_s: Optional[float] = None):
        """Blocking ``poll``: spins briefly (a decode step is ~1 ms away), then backs off to short sleeps."""
        t0 = time.monotonic()
        spins = 0
        while True:
            r = self.poll()
            if r is not None:
                return r
            spins += 1
            if spins > 2000:
                if self.shutdown_requested():
                    return None
                if timeout_s is not None and time.monotonic() - t0 > timeout_s:
                    raise TimeoutError("no launch record")
                time.sleep(0.0002 if spins < 50000 else 0.002)

    # ------------------------------------------------------------------------------------------ status page
    def set_error(self, seq: int, msg: str):
        raw = msg.encode("utf-8", "replace")[:_ERR_BYTES]
        off = 2048 + 256 * self.rank
        self._u8[off:off + _ERR_BYTES] = 0
        self._u8[off:off + len(raw)] = np.frombuffer(raw, dtype=np.uint8)
        self._u64[80 + self.rank] = max(1, int(seq))

    def first_error(self) -> Optional[Tuple[int, int, str]]:
        """(rank, seq, message) of the lowest-rank stage that reported a failure, else None."""
        errs = self._u64[80:80 + self.world]
        if not errs.any():
            return None
        r = int(np.nonzero(errs)[0][0])
        off = 2048 + 256 * r
        msg = bytes(self._u8[off:off + _ERR_BYTES]).split(b"\0", 1)[0].decode("utf-8", "replace")
        return r, int(errs[r]), msg

    def clear_errors(self):
        self._u64[80:80 + self.world] = 0

    def request_shutdown(self):
        self._u64[1] = 1

    def shutdown_requested(self) -> bool:
        return bool(self._u64[1])

    def consumed(self, rank: int) -> int:
        return int(self._u64[8 + rank])
