"""torchrun worker: whole-model expert parallelism (data-parallel attention + EP MoE, CUDA-graph decode loop) must emit
the same greedy tokens as the same rank running the un-sharded model.  Prints EP_MODEL_OK on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from helpers import GPU_DSV2  # noqa: E402
from mlx_sharding_b200.ops.meta import BatchMeta  # noqa: E402
from mlx_sharding_b200.parallel.decode_loop import DecodeLoop  # noqa: E402
from mlx_sharding_b200.parallel.ep import enable_expert_parallel  # noqa: E402
from mlx_sharding_b200.parallel.pipeline import StageExecutor  # noqa: E402
from mlx_sharding_b200.utils.loader import random_model  # noqa: E402

if __name__ == "__main__":
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    B, S, PS, steps = 16, 24, 16, 6
    pages_per_seq = (S + steps + 4 + PS - 1) // PS
    bts = [[1 + b * pages_per_seq + i for i in range(pages_per_seq)] for b in range(B)]
    prompts = torch.randint(3, GPU_DSV2["vocab_size"] - 1, (B, S), generator=torch.Generator().manual_seed(50 + rank))

    def decode(model, use_loop):
        stage = StageExecutor(model, B * pages_per_seq + 1, PS)
        meta = BatchMeta.build([S] * B, [0] * B, bts, PS, device=dev)
        logits = stage.forward(prompts.reshape(-1).to(dev), meta)
        toks = logits.argmax(-1)
        hist, lg = [toks.clone()], [logits.float().clone()]
        if use_loop:
            loop = DecodeLoop(stage, 1, B, pages_per_seq, transport="local", standalone=True)
            loop.groups[0].load(torch.full((B,), S, dtype=torch.int32), torch.tensor(bts, dtype=torch.int32), toks, S + steps + 4)
            loop.capture()
            for _ in range(steps):
                loop.step_all()
                torch.cuda.synchronize()
                hist.append(loop.groups[0].tokens.clone())
        else:
            for s in range(steps):
                meta = BatchMeta.build([1] * B, [S + s] * B, bts, PS, device=dev)
                logits = stage.forward(toks, meta)
                toks = logits.argmax(-1)
                hist.append(toks.clone())
                lg.append(logits.float().clone())
        return torch.stack(hist), (torch.stack(lg) if not use_loop else None)

    model = random_model(GPU_DSV2, device=dev, backend="b200", seed=3)
    ref, ref_logits = decode(model, False)
    bufs = enable_expert_parallel(model, max_tokens=B * S)
    assert all(w.get("e_gate") is None for w in model.layer_weights.values() if "router" in w)
    got, _ = decode(model, True)
    # Greedy decoding amplifies a single near-tie into a diverged suffix, and the two paths round differently (the un-sharded model
    # runs the shared experts inside the routed bank with a bf16 combine, the EP path adds them as an fp32 residual).  So: a
    # sequence may diverge only at a step where the reference's own margin between the two candidate tokens is within rounding
    # noise (5 % of the logit row's spread), and at most 2 of the 16 sequences may do so.
    diverged, worst = 0, 0.0
    for b in range(B):
        neq = (ref[:, b] != got[:, b]).nonzero()
        if neq.numel() == 0:
            continue
        s0 = int(neq[0])
        row = ref_logits[s0, b]
        margin = float((row[ref[s0, b]] - row[got[s0, b]]) / row.std())
        diverged, worst = diverged + 1, max(worst, margin)
    bad = int((ref != got).sum().item())
    ok = torch.tensor([1.0 if (diverged <= 2 and worst < 0.05 and not bufs.error()) else 0.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"mismatching tokens on rank 0: {bad} of {ref.numel()} (per step: {(ref != got).sum(1).tolist()}); "
              f"{diverged} sequence(s) diverged, worst reference margin at the divergence {worst:.4f} of the logit spread")
        if ok.item() == 1.0:
            print("EP_MODEL_OK")
    dist.barrier()
    dist.destroy_process_group()
