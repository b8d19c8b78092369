"""Checkpoint layout: MLX affine quant pack/unpack, shard key filter, pre-splitter output, loader."""
import json
import os
import subprocess
import sys

import pytest
import torch

from helpers import TINY_DSV2, TINY_GEMMA2, TINY_LLAMA, run_sequence
from mlx_sharding_b200.config import ModelConfig, ShardSpec
from mlx_sharding_b200.utils import quant
from mlx_sharding_b200.utils.checkpoint import (key_in_shard, save_sharded_weights, copy_other_files,
                                                write_synthetic_checkpoint)
from mlx_sharding_b200.utils.loader import load_model

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("bits", [2, 4, 8])
def test_quant_pack_roundtrip_and_error(bits):
    w = torch.randn(16, 256)
    wq, s, b = quant.quantize(w, 64, bits)
    assert wq.shape == (16, 256 * bits // 32) and s.shape == (16, 4)
    codes = quant.unpack_codes(wq, bits)
    assert codes.max() <= (1 << bits) - 1 and torch.equal(quant.pack_codes(codes, bits), wq)
    d = quant.dequantize(wq, s, b, 64, bits)
    step = (w.view(16, 4, 64).amax(-1) - w.view(16, 4, 64).amin(-1)) / ((1 << bits) - 1)
    if bits >= 4:  # affine grid anchored on the larger-magnitude edge: error stays within ~one step
        assert ((d - w).abs().view(16, 4, 64).amax(-1) <= step * 1.01 + 2e-2).all()


def test_quant_lsb_first_layout():
    codes = torch.arange(8, dtype=torch.uint8)[None]       # 4-bit codes 0..7 -> one uint32
    word = int(quant.pack_codes(codes, 4)[0, 0]) & 0xFFFFFFFF
    assert word == 0x76543210


def test_key_filter():
    first, mid, last = ShardSpec(0, 2, 6), ShardSpec(2, 4, 6), ShardSpec(4, 6, 6)
    assert key_in_shard("model.embed_tokens.weight", first) and not key_in_shard("model.embed_tokens.weight", mid)
    assert key_in_shard("model.embed_tokens.weight", last, tied_embeddings=True)
    assert not key_in_shard("model.embed_tokens.weight", last, tied_embeddings=False)
    assert key_in_shard("model.layers.3.mlp.experts.7.up_proj.weight", mid)
    assert not key_in_shard("model.layers.4.self_attn.q_proj.weight", mid)
    assert key_in_shard("model.norm.weight", last) and key_in_shard("lm_head.weight", last)
    assert not key_in_shard("lm_head.weight", first)
    assert not key_in_shard("model.layers.0.self_attn.rotary_emb.inv_freq", first)


@pytest.mark.parametrize("C,multi", [(TINY_LLAMA, 1), (TINY_DSV2, 3), (TINY_GEMMA2, 1)], ids=["llama", "dsv2", "gemma2"])
def test_presplit_layout_and_equivalence(tmp_path, C, multi):
    src = write_synthetic_checkpoint(str(tmp_path / "full"), C, dtype=torch.float32, shards=multi)
    L = C["num_hidden_layers"]
    outs = []
    for i, (s, e) in enumerate([(0, 2), (2, L)]):
        out = str(tmp_path / f"shard_{i}")
        # through the CLI with the reference's flag spelling
        subprocess.run([sys.executable, os.path.join(REPO, "sharding_weight.py"), "--model", src, "--output_dir", out,
                        "--start_layer", str(s), "--end_layer", str(e), "--total_layers", str(L)], check=True,
                       env=dict(os.environ, PYTHONPATH=REPO))
        assert os.path.exists(os.path.join(out, f"model-{s:05d}-{e:05d}.safetensors"))
        cfg = json.load(open(os.path.join(out, "config.json")))
        assert cfg["start_layer"] == s and cfg["end_layer"] == e
        assert os.path.exists(os.path.join(out, "tokenizer.json"))
        idx = os.path.join(out, f"model-{s:05d}-{e:05d}.safetensors.index.json")
        assert os.path.exists(idx) == (multi > 1)
        if multi > 1:
            wm = json.load(open(idx))["weight_map"]
            assert wm and all(v == f"model-{s:05d}-{e:05d}.safetensors" for v in wm.values())
        from safetensors import safe_open

        with safe_open(os.path.join(out, f"model-{s:05d}-{e:05d}.safetensors"), "pt") as f:
            assert f.metadata() == {"format": "mlx"}
            keys = set(f.keys())
        assert ("model.embed_tokens.weight" in keys) == (s == 0 or (e == L and C["model_type"] == "gemma2"))
        assert ("model.norm.weight" in keys) == (e == L)
        outs.append(out)
    # pre-sharded dirs (range from config.json) == dynamic sharding of the full checkpoint == unsharded
    pre = [load_model(o, dtype=torch.float32, device="cpu") for o in outs]
    dyn = [load_model(src, s, e, dtype=torch.float32, device="cpu") for s, e in [(0, 2), (2, L)]]
    full = load_model(src, dtype=torch.float32, device="cpu")
    assert (pre[0].spec.start_layer, pre[0].spec.end_layer, pre[1].spec.end_layer) == (0, 2, L)
    toks = [3, 14, 15, 92, 65]
    a, b, c = run_sequence(pre, toks, 2), run_sequence(dyn, toks, 2), run_sequence([full], toks, 2)
    for x, y, z in zip(a, b, c):
        assert torch.equal(x, y) and torch.allclose(x, z, atol=1e-5)


def test_quantized_checkpoint_on_disk(tmp_path):
    q = dict(group_size=32, bits=4)
    src = write_synthetic_checkpoint(str(tmp_path / "q4"), TINY_DSV2, dtype=torch.float16, quantization=q)
    from safetensors import safe_open

    with safe_open(os.path.join(src, "model.safetensors"), "pt") as f:
        w = f.get_tensor("model.layers.1.mlp.switch_mlp.gate_proj.weight")
        assert w.dtype == torch.uint32 and list(w.shape) == [8, 48, 64 * 4 // 32]
        assert "model.layers.1.mlp.gate.scales" not in f.keys()
    m = load_model(src, dtype=torch.float32, device="cpu")
    assert m.layer_weights[1]["e_gate"].is_quantized and not m.layer_weights[1]["e_gate"].weight
    assert len(run_sequence([m], [1, 2, 3], 1)) == 2


def test_config_defaults_and_remap():
    cfg = ModelConfig.from_dict(dict(TINY_LLAMA, model_type="mistral"))
    assert cfg.model_type == "llama"
    spec = cfg.shard(start_layer=2)  # either bound alone is accepted
    assert (spec.start_layer, spec.end_layer) == (2, 4)
    assert cfg.shard().end_layer == cfg.num_hidden_layers  # not a hard-wired 32 like the reference
    with pytest.raises(ValueError):
        ModelConfig.from_dict(dict(TINY_LLAMA, model_type="phi-msft"))
    with pytest.raises(ValueError):
        ModelConfig.from_dict(dict(TINY_LLAMA, model_type="gpt2"))
    assert [(s.start_layer, s.end_layer) for s in ShardSpec.even_split(27, 2)] == [(0, 14), (14, 27)]


def test_tied_llama_last_stage_gets_embeddings():
    """The reference crashes here (llama.py:86-87 uses embed_tokens only shard 0 has)."""
    from mlx_sharding_b200.models import build_stage
    from mlx_sharding_b200.utils.checkpoint import random_state_dict

    cfg = ModelConfig.from_dict(dict(TINY_LLAMA, tie_word_embeddings=True))
    sd = dict(random_state_dict(cfg, dtype=torch.float32))
    assert "lm_head.weight" not in sd
    parts = [build_stage(cfg, cfg.shard(0, 2), torch.float32).load_state(sd),
             build_stage(cfg, cfg.shard(2, 4), torch.float32).load_state(sd)]
    full = build_stage(cfg, cfg.shard(), torch.float32).load_state(sd)
    a, b = run_sequence(parts, [1, 2, 3], 1), run_sequence([full], [1, 2, 3], 1)
    assert torch.allclose(a[-1], b[-1], atol=1e-5)


@pytest.mark.parametrize("stacked", [True, False])
def test_expert_sharded_load(tmp_path, stacked):
    """``load_model(expert_shard=(r, world))`` keeps rank r's slice of every routed-expert bank (expert parallelism) — for
    stacked ``switch_mlp.*`` banks and for HF-style per-expert tensors."""
    import torch

    from helpers import TINY_DSV2
    from mlx_sharding_b200.utils.checkpoint import write_synthetic_checkpoint
    from mlx_sharding_b200.utils.loader import load_model

    path = write_synthetic_checkpoint(str(tmp_path / "ckpt"), TINY_DSV2, dtype=torch.float32, stacked_experts=stacked)
    full = load_model(path, dtype=torch.float32, device="cpu")
    E = TINY_DSV2["n_routed_experts"]
    for r in range(2):
        part = load_model(path, dtype=torch.float32, device="cpu", expert_shard=(r, 2))
        assert part.expert_shard == (r, 2)
        lo, hi = r * E // 2, (r + 1) * E // 2
        for i, w in part.layer_weights.items():
            if "router" not in w:
                continue
            for k in ("e_gate", "e_up", "e_down"):
                assert torch.equal(w[k].weight, full.layer_weights[i][k].weight[lo:hi])
            assert torch.equal(w["router"], full.layer_weights[i]["router"])       # router stays whole (routes over all E)
            assert torch.equal(w["s_down"].weight, full.layer_weights[i]["s_down"].weight)
