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


def _worker(rank, world, port, C, prompts, n_new, q, half=False, control="auto", fail_at=None):
    import torch.distributed as dist

    from mlx_sharding_b200.config import ModelConfig, ShardSpec
    from mlx_sharding_b200.engine.core import LLMEngine
    from mlx_sharding_b200.engine.sampler import SamplingParams
    from mlx_sharding_b200.models import build_stage
    from mlx_sharding_b200.parallel.pipeline import ChainPipeline, StageExecutor, build_chain, worker_loop
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
    ctl, plane = build_chain(stage, num_groups=world, max_tokens=64, max_seqs=8, control=control)
    if rank == 0:
        pipe = ChainPipeline(stage, ctl, plane)
        eng = LLMEngine(pipe, num_pages=32, page_size=16, num_groups=world)
        if fail_at is None:
            reqs = [eng.submit(p, SamplingParams(), max_tokens=n_new) for p in prompts]
            eng.drain()
            out = [r.output for r in reqs]
        else:
            out = _drive_with_failure(eng, prompts, n_new)
        pipe.shutdown()
        q.put(out)
    else:
        if fail_at is not None and rank == world - 1:
            # inject one failure into this stage's forward after ``fail_at`` steps (two groups are in flight at that point)
            calls, real = [0], stage.model.forward

            def flaky(*a, **k):
                calls[0] += 1
                if calls[0] == fail_at:
                    raise RuntimeError("injected stage failure")
                return real(*a, **k)

            stage.model.forward = flaky
        worker_loop(stage, ctl, plane)
    dist.barrier()
    dist.destroy_process_group()


def _drive_with_failure(eng, prompts, n_new):
    """Two groups in flight, a worker raises once: every active request fails cleanly, pages come back, no stale result is
    attributed to a later step, and the pipeline serves new requests afterwards (step ids + drain before release)."""
    from mlx_sharding_b200.engine.sampler import SamplingParams

    free0 = eng.table.alloc.num_free
    reqs = [eng.submit(p, SamplingParams(), max_tokens=50) for p in prompts]
    for _ in range(200):
        if all(r.finished for r in reqs):
            break
        try:
            eng.step()
        except BaseException as e:  # the engine thread does this in _loop
            eng._fail_all(e)
    assert all(r.finished and r.error is not None for r in reqs), [(r.finished, r.error) for r in reqs]
    assert eng.table.alloc.num_free == free0
    assert not eng.pipe._outstanding
    again = [eng.submit(p, SamplingParams(), max_tokens=n_new) for p in prompts]
    eng.drain()
    assert all(r.error is None for r in again)
    return [r.output for r in again]


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world,half,control,fail_at", [(2, False, "shm", None), (3, False, "shm", None), (3, True, "dist", None),
                                                        (2, False, "shm", 4), (2, False, "dist", 4)])
def test_gloo_chain_matches_single_process(world, half, control, fail_at):
    """Chain over gloo with the shared-memory launch ring (``shm``) and with the broadcast fallback (``dist``); ``fail_at``:
    the last stage raises once mid-generation (see ``_drive_with_failure``)."""
    C = dict(TINY_LLAMA, num_hidden_layers=2 if world == 2 else 3)
    prompts = [[5, 6, 7, 8], [100, 50, 3], [9] * 6]
    n_new = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, C, prompts, n_new, q, half, control, fail_at)) for r in range(world)]
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


def _ctrl_worker(rank, world, port, q):
    import torch.distributed as dist

    from mlx_sharding_b200.parallel.transport import TorchDistTransport

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tp = TorchDistTransport("cpu")
    msgs = [dict(kind="step", n=1), dict(blob=list(range(20000))), "x" * (tp.CTRL_FRAME - 4 - 20), dict(t=torch.arange(7))]
    if rank == 0:
        for m in msgs:
            tp.send_ctrl(m, 1)
        tp.flush()
        q.put(tp.recv_ctrl(1))
    else:
        got = [tp.recv_ctrl(0) for _ in msgs]
        ok = got[0] == msgs[0] and got[1] == msgs[1] and got[2] == msgs[2] and torch.equal(got[3]["t"], msgs[3]["t"])
        tp.send_ctrl(ok, 0)
        tp.flush()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_control_messages_small_and_large_frames():
    """Control objects travel as one fixed-size gloo frame; objects larger than the frame use the announced second message."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ctrl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    assert q.get(timeout=90) is True
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
