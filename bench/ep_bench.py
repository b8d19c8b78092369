#!/usr/bin/env python
"""BASELINE config 5: DeepSeek-Coder-V2-Lite MoE blocks with the fused expert all-to-all (expert parallelism).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench/ep_bench.py [--tokens 64]

Every rank routes its own ``--tokens`` tokens through one MoE block (router -> 6 of 64 experts -> weighted combine):

* ``local``: all 64 experts resident on the rank (what a pipeline stage does today) — the rank streams the whole
  1.1 GB expert bank per block;
* ``ep``   : experts sharded over the N ranks (64/N each); tokens travel through the fused dispatch / return kernels
  over NVLink peer memory (``ops/csrc/ep.cu``) — each rank streams 1/N of the bank but serves N x tokens.

Prints one JSON line (rank 0): device-timed microseconds per block (CUDA graph of ``--layers`` chained blocks, max over
ranks), aggregate tokens/s per block, NVLink bytes per block and the achieved fraction of the roofline
max(weight-stream time at measured HBM bandwidth, NVLink bytes / 770 GB/s).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mlx_sharding_b200.ops import b200  # noqa: E402
from mlx_sharding_b200.ops.weights import LinearWeight  # noqa: E402
from mlx_sharding_b200.parallel.ep import EPBuffers, ExpertParallelMoE  # noqa: E402
from mlx_sharding_b200.utils.timing import max_over_ranks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=64, help="tokens per rank per step")
    ap.add_argument("--layers", type=int, default=8, help="chained MoE blocks inside the timed graph")
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    rank, world = dist.get_rank(), dist.get_world_size()
    H, I, E, k, T = 2048, 1408, 64, 6, args.tokens
    g = torch.Generator(device="cuda").manual_seed(5)
    mk = lambda *s, sc=0.03: (torch.randn(*s, device="cuda", generator=g) * sc).to(torch.bfloat16)
    L = args.layers
    banks = [(LinearWeight(weight=mk(E, I, H)), LinearWeight(weight=mk(E, I, H)), LinearWeight(weight=mk(E, H, I)), mk(E, H, sc=0.05))
             for _ in range(L)]
    g2 = torch.Generator(device="cuda").manual_seed(100 + rank)
    x0 = torch.randn(T, H, device="cuda", generator=g2).to(torch.bfloat16)
    bufs = EPBuffers(H, max(T, 64), k)
    eps = [ExpertParallelMoE(bufs, Wg, Wu, Wd, E) for (Wg, Wu, Wd, _) in banks]

    def run_local(x):
        for (Wg, Wu, Wd, gate) in banks:
            idx, w = b200.moe_route(x, gate, k)
            x = b200.moe_experts(x, idx, w, Wg, Wu, Wd, "silu", residual=x)
        return x

    def run_ep(x):
        for ep, (_, _, _, gate) in zip(eps, banks):
            idx, w = b200.moe_route(x, gate, k)
            x = ep.forward(x, idx, w, residual=x)
        return x

    def timed(fn):
        for _ in range(2):
            fn(x0)
        torch.cuda.synchronize()
        dist.barrier()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            out = fn(x0)
        gr.replay()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        return max_over_ranks(e0.elapsed_time(e1)) * 1e3 / (args.iters * L), out

    us_local, out_l = timed(run_local)
    us_ep, out_e = timed(run_ep)
    err = (out_l.float() - out_e.float()).abs().max().item()
    bank_bytes = 3 * E * I * H * 2
    nvl_bytes = T * k * (world - 1) / world * (H * 2 + H * 4)          # dispatch bf16 + return fp32, remote share
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm = peaks.get("hbm_gbs", 6650.0) * 1e9
    roof_us = max(bank_bytes / world / hbm, nvl_bytes / 770e9) * 1e6
    if rank == 0:
        print(json.dumps({
            "bench": "DeepSeek-V2-Lite MoE block, expert parallel fused all-to-all", "n_gpus": world, "tokens_per_rank": T,
            "us_per_block_local": round(us_local, 2), "us_per_block_ep": round(us_ep, 2),
            "tokens_per_s_local": round(world * T / us_local * 1e6), "tokens_per_s_ep": round(world * T / us_ep * 1e6),
            "speedup_ep_vs_local": round(us_local / us_ep, 3), "nvlink_bytes_per_block_per_rank": int(nvl_bytes),
            "roofline_us": round(roof_us, 2), "roofline_fraction": round(roof_us / us_ep, 3),
            "roofline_basis": f"max(expert bank / N at {'measured' if peaks else 'fallback'} HBM {hbm / 1e9:.0f} GB/s, NVLink bytes at 770 GB/s)",
            "max_abs_diff_vs_local": err, "p2p_error": bool(bufs.error())}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
