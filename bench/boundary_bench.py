#!/usr/bin/env python
"""Fused stage-boundary microbenchmark ("fused stage-boundary send as achieved fraction of its roofline").

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 bench/boundary_bench.py

Rank 0 runs the *last GEMM of a stage* (Llama-3-8B down-proj: [T, 14336] x [4096, 14336]^T + residual) with its
epilogue storing straight into rank 1's resident inbox over NVLink and bumping rank 1's flag; rank 1 runs the
consumer's wait kernel.  Compared with the same GEMM writing locally followed by an NCCL send/recv (the baseline
hand-off).  Roofline = max(GEMM FLOPs / measured bf16 peak, GEMM weight bytes / measured HBM bandwidth,
hand-off bytes / 770 GB/s NVLink).  Also times the DeepSeek-style boundary (MoE weighted-combine with peer store).
Prints one JSON line on rank 0; every time is device-timed (CUDA events), max over the two ranks where it spans both.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mlx_sharding_b200.ops import b200  # noqa: E402
from mlx_sharding_b200.ops.weights import LinearWeight  # noqa: E402
from mlx_sharding_b200.parallel.p2p_fused import FusedP2PBoundary  # noqa: E402
from mlx_sharding_b200.utils.timing import max_over_ranks  # noqa: E402


def ev_time(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    rank, world = dist.get_rank(), dist.get_world_size()
    assert world == 2
    C = b200.load_extension()
    H, I = 4096, 14336
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm, pf = peaks.get("hbm_gbs", 6650.0) * 1e9, peaks.get("bf16_tflops", 1590.0) * 1e12
    results = {"roofline_basis": ("measured" if peaks else "fallback") + f" HBM {hbm / 1e9:.0f} GB/s, bf16 {pf / 1e12:.0f} TFLOP/s, NVLink 770 GB/s"}
    g = torch.Generator(device="cuda").manual_seed(1)
    W = LinearWeight(weight=(torch.randn(H, I, device="cuda", generator=g) * 0.02).to(torch.bfloat16))
    for T in (64, 1024):
        p2p = FusedP2PBoundary(H, 1, T, 64)
        x = (torch.randn(T, I, device="cuda", generator=g)).to(torch.bfloat16)
        res = (torch.randn(T, H, device="cuda", generator=g)).to(torch.bfloat16)
        local_out = torch.empty(T, H, dtype=torch.bfloat16, device="cuda")
        recv_buf = torch.empty(T, H, dtype=torch.bfloat16, device="cuda")

        def fused():
            if rank == 0:
                b200.linear(x, W, residual=res, out=p2p.next_hidden(0, T), signal=(p2p.next_hidden_flag(0), 0))
            else:
                p2p.wait_hidden(0)

        def local_only():
            if rank == 0:
                b200.linear(x, W, residual=res, out=local_out)

        def nccl_baseline():
            if rank == 0:
                b200.linear(x, W, residual=res, out=local_out)
                dist.send(local_out, 1)
            else:
                dist.recv(recv_buf, 0)

        t_fused = max_over_ranks(ev_time(fused))
        t_local = max_over_ranks(ev_time(local_only))
        t_nccl = max_over_ranks(ev_time(nccl_baseline))
        # correctness of the hand-off: rank 1's inbox == rank 0's local result
        if rank == 0:
            dist.send(local_out, 1)
            ok = True
        else:
            dist.recv(recv_buf, 0)
            ok = bool(torch.equal(recv_buf, p2p.hidden_inbox(0, T)))
        okt = torch.tensor([1.0 if ok else 0.0], device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        flops, wbytes, nvl = 2.0 * T * H * I, H * I * 2.0, T * H * 2.0
        roof = max(flops / pf, wbytes / hbm, nvl / 770e9) * 1e6
        results[f"T{T}"] = {"fused_gemm_p2p_us": round(t_fused, 2), "gemm_local_us": round(t_local, 2),
                            "gemm_plus_nccl_sendrecv_us": round(t_nccl, 2), "roofline_us": round(roof, 2),
                            "roofline_fraction_fused": round(roof / t_fused, 3), "handoff_bytes": int(nvl),
                            "handoff_overhead_vs_local_us": round(t_fused - t_local, 2), "inbox_matches": bool(okt.item())}
    if rank == 0:
        print(json.dumps({"bench": "fused stage boundary: last GEMM of a stage + P2P store + flag vs GEMM + NCCL send/recv",
                          "shape": "Llama-3-8B down-proj [T,14336]x[4096,14336]^T + residual", **results}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
