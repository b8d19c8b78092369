"""Stage-to-stage transports for the chain pipeline.

Tiers:
  T0  fused P2P store over NVLink (``p2p_fused.py``)      – product path on GPUs, no NCCL call, no host hop
  T1  NCCL ``send/recv`` (``TorchDistTransport`` + nccl)   – baseline + fallback
  T2  gloo ``send/recv`` on CPU tensors                    – plumbing / CI without GPUs (BASELINE config 1)
  T3  the reference's gRPC proto (``grpc_compat.py``)       – wire compatibility with mlx-sharding peers

The reference's only transport is T3: blocking unary gRPC carrying host-staged fp16 bytes
(shard/utils.py:71-109).  Here control (small python objects) always rides a CPU gloo group; tensor
payloads ride the data group of the tier in use and stay bf16 on device (no fp16 down-cast,
SURVEY K15/K16).
"""
from __future__ import annotations

import io
import pickle
from typing import Optional, Tuple

import torch
import torch.distributed as dist


class ChainTransport:
    rank: int
    world_size: int

    def send_ctrl(self, obj, dst: int):
        raise NotImplementedError

    def recv_ctrl(self, src: int):
        raise NotImplementedError

    def irecv_ctrl(self, src: int):
        return src

    def wait_ctrl(self, handle):
        return self.recv_ctrl(handle)

    def send_tensor(self, t: torch.Tensor, dst: int, slot: int = 0):
        raise NotImplementedError

    def recv_tensor(self, shape: Tuple[int, ...], dtype, src: int, slot: int = 0) -> torch.Tensor:
        raise NotImplementedError


class TorchDistTransport(ChainTransport):
    """torch.distributed p2p: control on a gloo group, payload on ``data_backend`` (gloo | nccl)."""

    def __init__(self, device="cpu", data_backend: Optional[str] = None, timeout_s: float = 600.0, ctrl_group=None):
        import datetime

        assert dist.is_initialized(), "init_process_group first (see parallel.launch.init_distributed)"
        self.rank, self.world_size = dist.get_rank(), dist.get_world_size()
        self.device = torch.device(device)
        to = datetime.timedelta(seconds=timeout_s)
        self.ctrl_group = ctrl_group if ctrl_group is not None else dist.new_group(backend="gloo", timeout=to)
        default_backend = dist.get_backend()
        want = data_backend or ("nccl" if self.device.type == "cuda" else "gloo")
        self.data_backend = want
        if want == default_backend:
            self.data_group = dist.group.WORLD
        elif want == "gloo":
            self.data_group = self.ctrl_group
        else:
            self.data_group = dist.new_group(backend=want, timeout=to)
        self.bytes_sent = 0
        self._pending = []  # (work, keep-alive tensor): sends are asynchronous so a chain never deadlocks

    def _isend(self, t: torch.Tensor, dst: int, group):
        self._pending.append((dist.isend(t, dst, group=group), t))
        if len(self._pending) > 64:
            self._pending = [(w, k) for (w, k) in self._pending if not w.is_completed()]

    def flush(self):
        for w, _ in self._pending:
            w.wait()
        self._pending = []

    # control plane ------------------------------------------------------------------------------
    # One gloo message per control object: a fixed-size frame [u32 length | pickle | padding].  Objects that do not fit are
    # announced by the frame (length with the top bit set) and follow in a second, exactly-sized message.
    CTRL_FRAME = 16384

    def send_ctrl(self, obj, dst: int):
        raw = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
        n = len(raw)
        frame = torch.zeros(self.CTRL_FRAME, dtype=torch.uint8)
        if n <= self.CTRL_FRAME - 4:
            frame[:4] = torch.frombuffer(bytearray(n.to_bytes(4, "little")), dtype=torch.uint8)
            frame[4:4 + n] = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
            self._isend(frame, dst, self.ctrl_group)
            return
        frame[:4] = torch.frombuffer(bytearray((n | 0x80000000).to_bytes(4, "little")), dtype=torch.uint8)
        self._isend(frame, dst, self.ctrl_group)
        self._isend(torch.frombuffer(bytearray(raw), dtype=torch.uint8), dst, self.ctrl_group)

    def recv_ctrl(self, src: int):
        frame = torch.empty(self.CTRL_FRAME, dtype=torch.uint8)
        dist.recv(frame, src, group=self.ctrl_group)
        n = int.from_bytes(frame[:4].numpy().tobytes(), "little")
        if n & 0x80000000:
            buf = torch.empty(n & 0x7FFFFFFF, dtype=torch.uint8)
            dist.recv(buf, src, group=self.ctrl_group)
            return pickle.loads(buf.numpy().tobytes())
        return pickle.loads(frame[4:4 + n].numpy().tobytes())

    # data plane ---------------------------------------------------------------------------------
    def send_tensor(self, t: torch.Tensor, dst: int, slot: int = 0):
        t = t.contiguous()
        if self.data_backend == "gloo" and t.is_cuda:
            t = t.cpu()
        if self.data_backend == "gloo" and t.dtype == torch.bfloat16:
            t = t.view(torch.int16)  # gloo has no bf16; reinterpret, do not down-cast
        if self.data_backend == "gloo":
            self._isend(t, dst, self.data_group)
        else:
            dist.send(t, dst, group=self.data_group)  # NCCL: stream-ordered, does not block the host
        self.bytes_sent += t.numel() * t.element_size()

    def recv_tensor(self, shape, dtype, src: int, slot: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is not None and self.data_backend != "gloo":
            dist.recv(out, src, group=self.data_group)  # straight into a static (graph-captured) buffer
            return out
        if self.data_backend == "gloo":
            wire = torch.int16 if dtype == torch.bfloat16 else dtype
            buf = torch.empty(shape, dtype=wire)
            dist.recv(buf, src, group=self.data_group)
            if wire != dtype:
                buf = buf.view(dtype)
            return buf.to(self.device)
        buf = torch.empty(shape, dtype=dtype, device=self.device)
        dist.recv(buf, src, group=self.data_group)
        return buf


def init_distributed(backend: Optional[str] = None, device: Optional[str] = None):
    """Initialise torch.distributed from the torchrun env (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    import os

    if dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    use_cuda = torch.cuda.is_available() and (device is None or str(device).startswith("cuda"))
    if use_cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)) % torch.cuda.device_count())
    backend = backend or ("nccl" if use_cuda else "gloo")
    kw = {}
    if backend == "nccl":
        kw["device_id"] = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world
