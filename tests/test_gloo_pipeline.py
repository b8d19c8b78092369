"""BASELINE config 1: 2-layer tiny-Llama, 2 CPU processes, gloo send/recv hidden-state relay."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from helpers import TINY_LLAMA, run_sequence


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, C, prompts, n_new, q, half=False):
    import torch.distributed as dist

    from mlx_sharding_b200.config import ModelConfig, ShardSpec
    from mlx_sharding_b200.engine.core import LLMEngine
    from mlx_sharding_b200.engine.sampler import SamplingParams
    from mlx_sharding_b200.models import build_stage
    from mlx_sharding_b200.parallel.pipeline import ChainPipeline, StageExecutor, worker_loop
    from mlx_sharding_b200.parallel.transport import TorchDistTransport
    from mlx_sharding_b200.utils.checkpoint import random_state_dict

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = ModelConfig.from_dict(C)
    if half:
        # every interior stage boundary between the attention and the MLP block of a layer (parallel/partition.py)
        L = cfg.num_hidden_layers
        cuts = [0] + [2 * (L * r // world) + 1 for r in range(1, world)] + [2 * L]
        spec = ShardSpec(cuts[rank] // 2, (cuts[rank + 1] + 1) // 2, L, skip_first_attn=bool(cuts[rank] % 2),
                         defer_last_mlp=bool(cuts[rank + 1] % 2))
    else:
        spec = ShardSpec.even_split(cfg.num_hidden_layers, world)[rank]
    sd = dict(random_state_dict(cfg, spec, dtype=torch.float32))
    model = build_stage(cfg, spec, torch.float32).load_state(sd)
    stage = StageExecutor(model, num_pages=32, page_size=16)
    tp = TorchDistTransport("cpu")
    if rank == 0:
        pipe = ChainPipeline(stage, tp)
        eng = LLMEngine(pipe, num_pages=32, page_size=16, num_groups=world)
        reqs = [eng.submit(p, SamplingParams(), max_tokens=n_new) for p in prompts]
        eng.drain()
        pipe.shutdown()
        q.put([r.output for r in reqs])
    else:
        worker_loop(stage, tp)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world,half", [(2, False), (3, False), (3, True)])
def test_gloo_chain_matches_single_process(world, half):
    C = dict(TINY_LLAMA, num_hidden_layers=2 if world == 2 else 3)
    prompts = [[5, 6, 7, 8], [100, 50, 3], [9] * 6]
    n_new = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, C, prompts, n_new, q, half)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # oracle: same weights, single process
    from mlx_sharding_b200.config import ModelConfig
    from mlx_sharding_b200.models import build_stage
    from mlx_sharding_b200.utils.checkpoint import random_state_dict

    cfg = ModelConfig.from_dict(C)
    sd = dict(random_state_dict(cfg, dtype=torch.float32))
    full = build_stage(cfg, cfg.shard(), torch.float32).load_state(sd)
    for p, g in zip(prompts, got):
        ref = [int(o.argmax()) for o in run_sequence([full], p, n_new - 1)]
        assert g == ref
