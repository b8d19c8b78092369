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
        toks = stage.forward(prompts.reshape(-1).to(dev), meta).argmax(-1)
        hist = [toks.clone()]
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
                toks = stage.forward(toks, meta).argmax(-1)
                hist.append(toks.clone())
        return torch.stack(hist)

    model = random_model(GPU_DSV2, device=dev, backend="b200", seed=3)
    ref = decode(model, False)
    bufs = enable_expert_parallel(model, max_tokens=B * S)
    assert all(w.get("e_gate") is None for w in model.layer_weights.values() if "router" in w)
    got = decode(model, True)
    bad = int((ref != got).sum().item())
    ok = torch.tensor([1.0 if (bad <= 1 and not bufs.error()) else 0.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"mismatching tokens on rank 0: {bad} of {ref.numel()}")
        if ok.item() == 1.0:
            print("EP_MODEL_OK")
    dist.barrier()
    dist.destroy_process_group()
