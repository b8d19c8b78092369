"""Model parity (CPU, fp32): our stage models vs HuggingFace transformers with the same random weights,
plus the sharded == unsharded and chunked-prefill invariances."""
import pytest
import torch

from helpers import TINY_DSV2, TINY_GEMMA2, TINY_LLAMA, run_sequence
from mlx_sharding_b200.config import ModelConfig
from mlx_sharding_b200.models import build_stage
from mlx_sharding_b200.utils.checkpoint import random_state_dict

TOKS = [5, 17, 200, 31, 8, 99, 100, 42, 7]


def _ours(C, ranges=None, **kw):
    cfg = ModelConfig.from_dict(C)
    sd = dict(random_state_dict(cfg, dtype=torch.float32, **kw))
    ranges = ranges or [(0, cfg.num_hidden_layers)]
    return cfg, sd, [build_stage(cfg, cfg.shard(s, e), torch.float32).load_state(sd) for s, e in ranges]


def _hf_logits(model, toks, n_decode):
    model.eval()
    outs = []
    ids = torch.tensor([toks])
    with torch.no_grad():
        for _ in range(n_decode + 1):
            lg = model(ids).logits[0, -1].float()
            outs.append(lg)
            ids = torch.cat([ids, lg.argmax().view(1, 1)], dim=1)
    return outs


def test_llama_vs_hf():
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg, sd, stages = _ours(TINY_LLAMA)
    hf = LlamaForCausalLM(LlamaConfig(**{k: v for k, v in TINY_LLAMA.items() if k != "model_type"},
                                      attn_implementation="eager"))
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and all("rotary" in m for m in missing)
    ref = _hf_logits(hf, TOKS, 3)
    got = run_sequence(stages, TOKS, 3)
    for a, b in zip(ref, got):
        assert torch.allclose(a, b, atol=2e-4, rtol=1e-4), (a - b).abs().max()


def test_llama3_rope_scaling_vs_hf():
    from transformers import LlamaConfig, LlamaForCausalLM

    C = dict(TINY_LLAMA, rope_scaling=dict(rope_type="llama3", factor=8.0, low_freq_factor=1.0,
                                           high_freq_factor=4.0, original_max_position_embeddings=64),
             rope_theta=500000.0)
    cfg, sd, stages = _ours(C)
    hf = LlamaForCausalLM(LlamaConfig(**{k: v for k, v in C.items() if k != "model_type"},
                                      attn_implementation="eager"))
    hf.load_state_dict(sd, strict=False)
    toks = list(range(3, 100))
    ref = _hf_logits(hf, toks, 1)
    got = run_sequence(stages, toks, 1)
    for a, b in zip(ref, got):
        assert torch.allclose(a, b, atol=3e-4, rtol=1e-4), (a - b).abs().max()


def test_gemma2_vs_hf():
    from transformers import Gemma2Config, Gemma2ForCausalLM

    cfg, sd, stages = _ours(TINY_GEMMA2)
    hc = Gemma2Config(**{k: v for k, v in TINY_GEMMA2.items() if k != "model_type"},
                      attn_implementation="eager")
    hf = Gemma2ForCausalLM(hc)
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected
    hf.tie_weights()
    ref = _hf_logits(hf, TOKS, 3)
    got = run_sequence(stages, TOKS, 3)
    for a, b in zip(ref, got):
        assert torch.allclose(a, b, atol=3e-4, rtol=1e-4), (a - b).abs().max()


def _dsv2_to_hf(sd, cfg):
    out = {}
    for k, v in sd.items():
        if ".switch_mlp." in k:
            continue
        out[k] = v
    for i in range(cfg.num_hidden_layers):
        p = f"model.layers.{i}.mlp"
        if f"{p}.switch_mlp.gate_proj.weight" in sd:
            g, u = sd[f"{p}.switch_mlp.gate_proj.weight"], sd[f"{p}.switch_mlp.up_proj.weight"]
            out[f"{p}.experts.gate_up_proj"] = torch.cat([g, u], dim=1)
            out[f"{p}.experts.down_proj"] = sd[f"{p}.switch_mlp.down_proj.weight"]
    return out


@pytest.mark.parametrize("topk_method", ["greedy", "group_limited_greedy"])
def test_deepseek_v2_vs_hf(topk_method):
    from transformers import DeepseekV2Config, DeepseekV2ForCausalLM

    C = dict(TINY_DSV2, topk_method=topk_method)
    if topk_method == "group_limited_greedy":
        C.update(n_group=4, topk_group=2)
    cfg, sd, stages = _ours(C)
    hc = DeepseekV2Config(**{k: v for k, v in C.items() if k != "model_type"}, attn_implementation="eager")
    hf = DeepseekV2ForCausalLM(hc)
    missing, unexpected = hf.load_state_dict(_dsv2_to_hf(sd, cfg), strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in m for m in missing), missing
    for layer in hf.model.layers:  # transformers 5.5 forgets this attribute on the group-limited path
        if hasattr(layer.mlp, "gate") and not hasattr(layer.mlp, "num_experts"):
            layer.mlp.num_experts = cfg.n_routed_experts
    ref = _hf_logits(hf, TOKS, 3)
    got = run_sequence(stages, TOKS, 3)
    for a, b in zip(ref, got):
        assert torch.allclose(a, b, atol=3e-4, rtol=1e-4), (a - b).abs().max()


@pytest.mark.parametrize("C", [TINY_LLAMA, TINY_GEMMA2, TINY_DSV2], ids=lambda c: c["model_type"])
def test_sharded_equals_unsharded(C):
    cfg, sd, full = _ours(C)
    _, _, parts = _ours(C, ranges=[(0, 1), (1, 3), (3, 4)])
    a = run_sequence(full, TOKS, 4)
    b = run_sequence(parts, TOKS, 4, chunk=4)  # also exercises chunked prefill
    for x, y in zip(a, b):
        assert torch.allclose(x, y, atol=1e-5)


@pytest.mark.parametrize("C", [TINY_LLAMA, TINY_GEMMA2, TINY_DSV2], ids=lambda c: c["model_type"])
def test_half_layer_boundaries_equal_unsharded(C):
    """A stage boundary may fall between the attention and the MLP block of a layer (ShardSpec.skip_first_attn /
    defer_last_mlp): stages load only their block's tensors and keep KV only for their attention blocks."""
    from mlx_sharding_b200.config import ShardSpec
    from mlx_sharding_b200.parallel.partition import balanced_split

    cfg, sd, full = _ours(C)
    L = cfg.num_hidden_layers
    specs = [ShardSpec(0, 2, L, defer_last_mlp=True), ShardSpec(1, 3, L, skip_first_attn=True, defer_last_mlp=True),
             ShardSpec(2, 3, L, skip_first_attn=True), ShardSpec(3, 4, L)]
    parts = [build_stage(cfg, sp, torch.float32).load_state(sd) for sp in specs]
    assert [m.kv_geometry()[0] for m in parts] == [2, 1, 0, 1]
    assert not parts[0].spec.is_last and parts[-1].spec.is_last and not parts[1].spec.is_first
    assert "o" not in parts[2].layer_weights[2] and ("down" in parts[2].layer_weights[2] or "e_down" in parts[2].layer_weights[2])
    a = run_sequence(full, TOKS, 3)
    b = run_sequence(parts, TOKS, 3, chunk=5)
    for x, y in zip(a, b):
        assert torch.allclose(x, y, atol=1e-5)
    # stage-local random init generates exactly the tensors the stage consumes
    for sp in specs:
        build_stage(cfg, sp, torch.float32).load_state(dict(random_state_dict(cfg, sp, dtype=torch.float32)))
    # the partitioner's half-layer plan covers every block exactly once and is never worse than whole layers
    from mlx_sharding_b200.parallel.partition import stage_cost
    for n in (2, 3, 5):
        plan = balanced_split(cfg, n, half_layers=True)
        units = [(i, blk) for sp in plan for i in sp.layers() for blk in ("a", "m")
                 if (sp.runs_attn(i) if blk == "a" else sp.runs_mlp(i))]
        assert units == [(i, blk) for i in range(L) for blk in ("a", "m")]
        assert plan[0].is_first and plan[-1].is_last and sum(sp.is_last for sp in plan) == 1
        whole = balanced_split(cfg, n)
        assert max(stage_cost(cfg, sp) for sp in plan) <= max(stage_cost(cfg, sp) for sp in whole) + 1e-12


def test_dsv2_unstacked_experts_are_stacked():
    """HF-style per-expert keys are stacked into switch_mlp.* (reference deepseek_v2.py:101-111)."""
    cfg = ModelConfig.from_dict(TINY_DSV2)
    sd_s = dict(random_state_dict(cfg, dtype=torch.float32, stacked_experts=True))
    sd_u = {}
    for k, v in sd_s.items():
        if ".switch_mlp." in k:
            pre, rest = k.split(".switch_mlp.")
            for e in range(cfg.n_routed_experts):
                sd_u[f"{pre}.experts.{e}.{rest}"] = v[e]
        else:
            sd_u[k] = v
    a = build_stage(cfg, cfg.shard(), torch.float32).load_state(sd_s)
    b = build_stage(cfg, cfg.shard(), torch.float32).load_state(sd_u)
    x, y = run_sequence([a], TOKS, 1), run_sequence([b], TOKS, 1)
    assert torch.equal(x[0], y[0])


def test_quantized_model_matches_dequantized():
    """4-bit MLX checkpoints: in-op dequant == loading the dequantised dense weights."""
    from mlx_sharding_b200.utils import quant

    q = dict(group_size=32, bits=4)
    C = dict(TINY_DSV2, quantization=q)
    cfg = ModelConfig.from_dict(C)
    sd = dict(random_state_dict(cfg, dtype=torch.float32, quantization=q))
    assert any(k.endswith(".scales") for k in sd)
    assert "model.layers.1.mlp.gate.scales" not in sd  # router stays unquantised
    mq = build_stage(cfg, cfg.shard(), torch.float32).load_state(sd)
    dense = {}
    for k, v in sd.items():
        if k.endswith(".scales") or k.endswith(".biases"):
            continue
        if k.endswith(".weight") and k[:-7] + ".scales" in sd:
            dense[k] = quant.dequantize(v.view(torch.int32), sd[k[:-7] + ".scales"], sd[k[:-7] + ".biases"], 32, 4)
        else:
            dense[k] = v
    cfg2 = ModelConfig.from_dict(TINY_DSV2)
    md = build_stage(cfg2, cfg2.shard(), torch.float32).load_state(dense)
    x, y = run_sequence([mq], TOKS, 2), run_sequence([md], TOKS, 2)
    for a, b in zip(x, y):
        assert torch.allclose(a, b, atol=1e-5)


def test_yarn_rope_and_mscale():
    """DeepSeek YaRN: frequencies follow the mlx_lm formula; softmax scale gets mscale^2."""
    import math

    from mlx_sharding_b200.config import deepseek_v2_lite_config
    from mlx_sharding_b200.models.base import deepseek_rope_spec

    cfg = ModelConfig.from_dict(deepseek_v2_lite_config())
    spec = deepseek_rope_spec(cfg)
    assert spec.interleaved and spec.rot_dim == 64 and abs(spec.mscale - 1.0) < 1e-6
    inv = spec.inv_freq
    base = 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64))
    # high-frequency dims untouched, low-frequency dims divided by factor 40
    assert torch.allclose(inv[0], base[0]) and torch.allclose(inv[-1], base[-1] / 40, rtol=1e-5)
    m = 0.1 * 0.707 * math.log(40) + 1.0
    assert abs(cfg.attn_scale - (192 ** -0.5) * m * m) < 1e-9


def test_absorbed_latent_mla_matches_decompressed_cache():
    """Opt-in MLA cache layout (latent 512 + roped key 64 per token, kv_b absorbed into the query / output side): same logits
    as the reference's decompressed K/V layout, 4.7x fewer cached values per token; also through a sharded pipeline."""
    from mlx_sharding_b200.models.deepseek_v2 import DeepseekV2Stage

    cfg, sd, full = _ours(TINY_DSV2)
    ref = run_sequence(full, TOKS, 4)

    class Absorbed(DeepseekV2Stage):
        absorbed_mla = True

    parts = [Absorbed(cfg, cfg.shard(s, e), torch.float32).load_state(sd) for s, e in [(0, 2), (2, 4)]]
    L, hk, dk, dv = parts[0].kv_geometry()
    assert (hk, dk, dv) == (1, cfg.kv_lora_rank + cfg.qk_rope_head_dim, cfg.kv_lora_rank)
    got = run_sequence(parts, TOKS, 4, chunk=4)
    for a, b in zip(ref, got):
        assert torch.allclose(a, b, atol=2e-4, rtol=1e-4), (a - b).abs().max()
    std = full[0].kv_geometry()
    assert std[1] * (std[2] + std[3]) > 2 * (dk + dv)
