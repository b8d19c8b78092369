"""Expert parallelism on CPU (gloo): the backend-agnostic ``all_to_all`` path of ``parallel/ep.py`` — whole-model EP with experts
sharded at load must reproduce the un-sharded model token for token, with a *different* batch on every rank."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from helpers import TINY_DSV2, run_sequence


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, quantized, q):
    import torch.distributed as dist

    from mlx_sharding_b200.parallel.ep import ExpertParallelMoERef, enable_expert_parallel
    from mlx_sharding_b200.utils.loader import random_model

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    qcfg = dict(group_size=32, bits=4) if quantized else None
    kw = dict(dtype=torch.float32, quantization=qcfg)
    full = random_model(TINY_DSV2, **kw)
    # rank-dependent prompt lengths: the exchange must cope with different token counts per rank
    toks = [5, 17, 200, 31, 8, 99, 100, 42, 7][: 5 + 2 * rank]
    ref = [int(o.argmax()) for o in run_sequence([full], toks, 3)]
    part = random_model(TINY_DSV2, expert_shard=(rank, world), **kw)            # experts sliced at load
    assert enable_expert_parallel(part) is None                                 # reference backend: no IPC buffers
    assert all(isinstance(l, ExpertParallelMoERef) for l in part.ep_layers.values())
    got = [int(o.argmax()) for o in run_sequence([part], toks, 3)]
    late = random_model(TINY_DSV2, **kw)                                        # full banks, sliced by enable_expert_parallel
    enable_expert_parallel(late)
    got2 = [int(o.argmax()) for o in run_sequence([late], toks, 3)]
    q.put((rank, ref, got, got2))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(240)
@pytest.mark.parametrize("world,quantized", [(2, False), (4, False), (2, True)])
def test_expert_parallel_reference_path_matches_unsharded(world, quantized):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, quantized, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=200) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ref, got, got2 in results:
        assert got == ref and got2 == ref, (rank, ref, got, got2)


def _serve_worker(rank, world, port, q):
    import torch.distributed as dist

    from mlx_sharding_b200.engine.sampler import SamplingParams
    from mlx_sharding_b200.parallel.ep_serving import build_lockstep_group
    from mlx_sharding_b200.utils.loader import random_model

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = random_model(TINY_DSV2, dtype=torch.float32, expert_shard=(rank, world))
    group = build_lockstep_group(model, num_pages=64, page_size=16, max_seqs=4, max_prefill_tokens=8)
    if rank != 0:
        group.serve_forever()
    else:
        group.start()
        prompts = [[5, 6, 7, 8], [100, 50, 3], [9] * 11, [42], [7, 7, 7, 7, 7]]          # 5 requests on 2 ranks: uneven
        reqs = [group.submit(p, SamplingParams(temperature=0.0), max_tokens=5 + i) for i, p in enumerate(prompts)]
        late = None
        outs = []
        for i, r in enumerate(reqs):
            outs.append([ev.token for ev in r])                                        # blocking iterator, like the HTTP handlers
            if i == 1:
                late = group.submit([1, 2, 3], SamplingParams(temperature=0.0), max_tokens=4)   # arrives while others run
        outs.append([ev.token for ev in late])
        cancelled = group.submit([8] * 6, SamplingParams(temperature=0.0), max_tokens=500)
        next(iter(cancelled))
        cancelled.cancel()
        tail = [ev for ev in cancelled]
        stats = group.stats
        group.shutdown()
        q.put((prompts + [[1, 2, 3]], outs, tail[-1].finish_reason, stats))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_lockstep_group_serves_uneven_traffic():
    """Requests assigned unevenly to 2 expert-parallel ranks (one rank idles part of the time and runs dummy steps): every
    request gets exactly the tokens of a single-process greedy run; cancellation and late arrivals work."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_serve_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    prompts, outs, cancel_reason, stats = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from mlx_sharding_b200.utils.loader import random_model

    full = random_model(TINY_DSV2, dtype=torch.float32)
    n_new = [5, 6, 7, 8, 9, 4]
    for p, o, n in zip(prompts, outs, n_new):
        ref = [int(x.argmax()) for x in run_sequence([full], p, n - 1)]
        assert o == ref, (p, o, ref)
    assert cancel_reason == "cancelled"
    assert stats["lockstep_dummy_steps"] >= 0 and stats["lockstep_assigned_rank0"] + stats["lockstep_assigned_rank1"] == 7
