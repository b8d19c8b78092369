"""Failure handling: a stage that dies mid-generation produces a clean request error (no hang,
no crashed engine), malformed / oversize inputs are rejected, cancelled requests free their KV pages."""
import threading
import time

import pytest
import torch

from helpers import TINY_LLAMA
from mlx_sharding_b200.config import ModelConfig
from mlx_sharding_b200.engine.core import LLMEngine
from mlx_sharding_b200.engine.sampler import SamplingParams
from mlx_sharding_b200.models import build_stage
from mlx_sharding_b200.parallel import grpc_compat as G
from mlx_sharding_b200.parallel.pipeline import LocalPipeline, StageExecutor
from mlx_sharding_b200.utils.checkpoint import random_state_dict


def _stages(ranges):
    cfg = ModelConfig.from_dict(TINY_LLAMA)
    sd = dict(random_state_dict(cfg, dtype=torch.float32))
    return [build_stage(cfg, cfg.shard(s, e), torch.float32).load_state(sd) for s, e in ranges]


@pytest.mark.timeout(90)
def test_shard_death_mid_generation_is_a_clean_error():
    s0, s1 = _stages([(0, 2), (2, 4)])
    srv, port = G.start_server(G.StageServicer(s1, num_pages=16, page_size=16, wire_dtype=torch.float32), 0, host="127.0.0.1")
    stubs = [G.StageStub(f"127.0.0.1:{port}", timeout_s=3.0)]
    eng = LLMEngine(G.GrpcRelayPipeline(StageExecutor(s0, 16, 16), stubs, wire_dtype=torch.float32), 16, 16,
                    num_groups=1, max_seqs_per_group=1).start()
    try:
        r = eng.submit([5, 6, 7], SamplingParams(), max_tokens=200)
        it = iter(r)
        first = next(it)
        assert first.token >= 0
        srv.stop(0)  # the remote stage dies while the request is in flight
        with pytest.raises(Exception):
            for _ in it:
                pass
        assert r.finished and r.error is not None
        assert eng.table.alloc.num_free == 15  # pages reclaimed
        # engine thread survived: a new request fails fast (shard still down) instead of hanging
        r2 = eng.submit([1, 2], SamplingParams(), max_tokens=2)
        t0 = time.time()
        with pytest.raises(Exception):
            list(r2)
        assert time.time() - t0 < 30
    finally:
        eng.shutdown()


def test_cancel_releases_pages_and_rejects_bad_requests():
    (m,) = _stages([(0, 4)])
    eng = LLMEngine(LocalPipeline.from_models([m], 32, 16), 32, 16, max_model_len=64)
    with pytest.raises(ValueError):
        eng.submit([], SamplingParams())
    with pytest.raises(ValueError):
        eng.submit([1] * 60, SamplingParams(), max_tokens=10)  # exceeds max_model_len
    with pytest.raises(ValueError):
        eng.submit([1], SamplingParams(temperature=-1.0))
    r = eng.submit([1, 2, 3], SamplingParams(), max_tokens=40)
    eng.step(); eng.step()
    r.cancel()
    eng.drain()
    assert r.finished and r.finish_reason == "cancelled" and len(r.output) < 40
    assert eng.table.alloc.num_free == 31
