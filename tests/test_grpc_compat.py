"""gRPC compatibility plane: wire format, servicer semantics (single slot, offset tracking, full logits),
hub-and-spoke relay through stubs == single-process run, error replies (SURVEY §4 plumbing/fault)."""
import pytest
import torch

from helpers import TINY_DSV2, TINY_LLAMA, run_sequence
from mlx_sharding_b200.config import ModelConfig
from mlx_sharding_b200.models import build_stage
from mlx_sharding_b200.parallel import grpc_compat as G
from mlx_sharding_b200.utils.checkpoint import random_state_dict


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16, torch.int32, torch.int64])
def test_tensor_roundtrip(dtype):
    t = (torch.randn(1, 5, 7) * 10).to(dtype)
    m = G.tensor_to_message(t)
    assert m.dtype == "mlx.core." + str(dtype).split(".")[-1] and list(m.shape) == [1, 5, 7]
    blob = m.SerializeToString()
    back = G.message_to_tensor(G.messages().Tensor.FromString(blob))
    assert back.dtype == dtype and torch.equal(back, t)


def test_bad_dtype_and_shape():
    M = G.messages()
    with pytest.raises(ValueError):
        G.message_to_tensor(M.Tensor(tensor_data=b"\0" * 4, shape=[1], dtype="mlx.core.complex64"))
    with pytest.raises(ValueError):
        G.message_to_tensor(M.Tensor(tensor_data=b"\0" * 8, shape=[3], dtype="float32"))


def _stages(C, ranges):
    cfg = ModelConfig.from_dict(C)
    sd = dict(random_state_dict(cfg, dtype=torch.float32))
    return cfg, sd, [build_stage(cfg, cfg.shard(s, e), torch.float32).load_state(sd) for s, e in ranges]


@pytest.mark.timeout(120)
def test_relay_through_two_grpc_shards_matches_local():
    cfg, sd, (s0, s1, s2) = _stages(TINY_DSV2, [(0, 1), (1, 3), (3, 4)])
    full = build_stage(cfg, cfg.shard(), torch.float32).load_state(sd)
    servers = []
    try:
        addrs = []
        for m in (s1, s2):
            srv, port = G.start_server(G.StageServicer(m, num_pages=16, page_size=16, wire_dtype=torch.float32), 0,
                                       host="127.0.0.1")
            servers.append(srv)
            addrs.append(f"127.0.0.1:{port}")
        stubs = G.connect_stubs(",".join(addrs))
        from mlx_sharding_b200.engine.core import LLMEngine
        from mlx_sharding_b200.engine.sampler import SamplingParams
        from mlx_sharding_b200.parallel.pipeline import StageExecutor

        pipe = G.GrpcRelayPipeline(StageExecutor(s0, 16, 16), stubs, wire_dtype=torch.float32)
        eng = LLMEngine(pipe, 16, 16, num_groups=1, max_seqs_per_group=1, max_prefill_tokens=4)
        prompt = [5, 6, 7, 8, 9, 10]
        ref = [int(o.argmax()) for o in run_sequence([full], prompt, 4)]
        assert eng.generate(prompt, SamplingParams(), max_tokens=5) == ref
        # second request: ResetCache must have cleared the remote offsets
        assert eng.generate(prompt, SamplingParams(), max_tokens=5) == ref
        # reference-shaped generator API
        from mlx_sharding_b200.engine.compat import create_generate_step_with_grpc

        gen = create_generate_step_with_grpc(stubs, num_pages=16, page_size=16, wire_dtype=torch.float32)
        out = []
        for (tok, lp), _ in zip(gen(torch.tensor(prompt), s0, temp=0.0), range(5)):
            out.append(tok)
            assert lp.shape == (cfg.vocab_size,) and abs(float(lp.exp().sum()) - 1) < 1e-3
        assert out == ref
        # last stage returns logits for all positions [1, T, V] like the reference
        stubs[0].reset_cache(); stubs[1].reset_cache()
        h = s0.forward(torch.tensor(prompt), __import__("mlx_sharding_b200.ops.meta", fromlist=["BatchMeta"]).BatchMeta.build(
            [6], [0], [list(range(1, 16))], 16), __import__("mlx_sharding_b200.engine.kv_cache", fromlist=["x"]).PagedKVCache.for_model(s0, 16, 16))
        y = stubs[1].send_tensor(stubs[0].send_tensor(h.unsqueeze(0)))
        assert list(y.shape) == [1, 6, cfg.vocab_size]
    finally:
        for s in servers:
            s.stop(0)


@pytest.mark.timeout(60)
def test_servicer_reports_errors_instead_of_crashing():
    cfg, sd, (s1,) = _stages(TINY_LLAMA, [(2, 4)])
    srv, port = G.start_server(G.StageServicer(s1, num_pages=4, page_size=16), 0, host="127.0.0.1")
    try:
        stub = G.StageStub(f"127.0.0.1:{port}")
        with pytest.raises(RuntimeError, match="shard"):
            stub.send_tensor(torch.zeros(1, 3, 7))  # wrong hidden size -> success=False reply
        with pytest.raises(RuntimeError, match="KV pool"):
            stub.send_tensor(torch.zeros(1, 100, cfg.hidden_size))  # exceeds 3 pages * 16
        stub.reset_cache()
        y = stub.send_tensor(torch.zeros(1, 2, cfg.hidden_size))
        assert list(y.shape) == [1, 2, cfg.vocab_size] and y.dtype == torch.float16
    finally:
        srv.stop(0)


@pytest.mark.timeout(30)
def test_dead_shard_gives_clean_error():
    stub = G.StageStub("127.0.0.1:1", timeout_s=2.0)
    with pytest.raises(Exception):
        stub.reset_cache()
