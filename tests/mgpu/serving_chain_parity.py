"""torchrun worker: the *serving* path on GPUs — ``LLMEngine`` -> ``ChainPipeline`` (shared-memory launch ring + fused P2P hand-off,
CUDA-graph replay per stage per decode step) — must emit the same tokens as a single-GPU ``LocalPipeline`` engine, for greedy
requests and for seeded sampled requests (temperature / top-p / repetition penalty / logprobs ride the step block, so they take the
graph path too).  Prints SERVING_OK on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from helpers import GPU_DSV2, GPU_LLAMA  # noqa: E402
from mlx_sharding_b200.config import ModelConfig, ShardSpec  # noqa: E402
from mlx_sharding_b200.engine.core import LLMEngine  # noqa: E402
from mlx_sharding_b200.engine.sampler import SamplingParams  # noqa: E402
from mlx_sharding_b200.parallel.pipeline import ChainPipeline, LocalPipeline, StageExecutor, build_chain, worker_loop  # noqa: E402
from mlx_sharding_b200.utils.loader import random_model  # noqa: E402

N_NEW, PS, PAGES = 12, 16, 512


def requests(cfg):
    gen = torch.Generator().manual_seed(5)
    out = []
    for i in range(12):
        n = 5 + 3 * i
        prompt = torch.randint(3, cfg.vocab_size - 1, (n,), generator=gen).tolist()
        if i % 3 == 0:
            p = SamplingParams(temperature=0.0)
        elif i % 3 == 1:
            p = SamplingParams(temperature=0.8, top_p=0.9, seed=100 + i, logprobs=3)
        else:
            p = SamplingParams(temperature=1.0, repetition_penalty=1.3, repetition_context_size=20, seed=7 * i, logit_bias={5: -2.0})
        out.append((prompt, p))
    return out


def drive(eng, reqs):
    rs = [eng.submit(p, sp, max_tokens=N_NEW) for p, sp in reqs]
    eng.drain()
    assert all(r.error is None for r in rs), [r.error for r in rs]
    return [r.output for r in rs]


if __name__ == "__main__":
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    arch = sys.argv[1] if len(sys.argv) > 1 else "dsv2"
    transport = sys.argv[2] if len(sys.argv) > 2 else "fused"
    cfgd = GPU_DSV2 if arch == "dsv2" else GPU_LLAMA
    cfg = ModelConfig.from_dict(cfgd)
    spec = ShardSpec.even_split(cfg.num_hidden_layers, world)[rank]
    stage = StageExecutor(random_model(cfgd, device=dev, backend="b200", seed=3, spec=spec), PAGES, PS)
    ctl, plane = build_chain(stage, num_groups=world, max_tokens=256, max_seqs=8, transport=transport)
    if rank != 0:
        worker_loop(stage, ctl, plane)
    else:
        pipe = ChainPipeline(stage, ctl, plane)
        eng = LLMEngine(pipe, PAGES, PS, num_groups=world, max_seqs_per_group=8, max_prefill_tokens=256)
        got = drive(eng, requests(cfg))
        again = drive(eng, requests(cfg))          # same seeds -> same streams, now on captured graphs
        replays = pipe.gcache.replays
        pipe.shutdown()
        full = StageExecutor(random_model(cfgd, device=dev, backend="b200", seed=3), PAGES, PS)
        ref = drive(LLMEngine(LocalPipeline([full]), PAGES, PS, num_groups=1, max_seqs_per_group=8, max_prefill_tokens=256), requests(cfg))
        tot = sum(len(r) for r in ref)
        # a bf16 near-tie may flip a token (and then the rest of that sequence): count sequence-level agreement
        same = sum(int(a == b) for a, b in zip(got, ref))
        same2 = sum(int(a == b) for a, b in zip(again, got))
        print(f"plane={plane.name} ctl={type(ctl).__name__} replays={replays} seq-agree {same}/{len(ref)} rerun-agree {same2}/{len(ref)} tokens={tot}")
        if same >= len(ref) - 2 and same2 >= len(ref) - 1 and (replays > 0 or transport != "fused") and plane.name == ("fused" if transport == "fused" else "dist"):
            print("SERVING_OK", arch, transport)
    dist.barrier()
    dist.destroy_process_group()
