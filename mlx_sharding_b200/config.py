"""Model configuration + layer-range shard specification.

Parity notes (reference = /root/reference/shard):

* The reference extends each upstream ``ModelArgs`` with ``start_layer`` / ``end_layer`` fields
  (server/model/llama.py:11-14, gemma2.py:9-12, deepseek_v2.py:11-14) whose *defaults are hard-wired*
  to 32 / 46 / 27.  We default ``end_layer`` to ``num_hidden_layers`` and accept either
  bound on its own (the reference only injects them when both are given, utils.py:37-39).
* ``model_type`` remapping ``mistral -> llama`` follows utils.py:14-17.  ``phi-msft -> phixtral`` in the
  reference points at a module that does not exist; we raise a clear error instead.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field
from typing import Any, Dict, Optional

MODEL_REMAPPING = {
    "mistral": "llama",  # reference utils.py:15
}

SUPPORTED_ARCHS = ("llama", "gemma2", "deepseek_v2")


@dataclass(frozen=True)
class ShardSpec:
    """A pipeline stage = the contiguous layer range ``start_layer <= i < end_layer``.

    Placement rules (reference llama.py:26-36,74-77; gemma2.py:23-24; deepseek_v2.py:23-35):
    embedding on the first stage, final norm + LM head on the last stage; when the head is tied to the
    embedding the last stage also holds ``embed_tokens`` (the reference does this only for Gemma-2 and
    would crash for tied Llama — fixed here).
    """

    start_layer: int
    end_layer: int
    num_layers: int
    # Half-layer boundaries (no reference counterpart; used by ``parallel.partition.balanced_split(half_layers=True)``).
    # A decoder layer is two residual blocks (attention, MLP) that only exchange the residual stream ``h``, so a
    # stage boundary may also fall *between* them: ``skip_first_attn`` = the attention block of ``start_layer`` ran
    # on the previous stage; ``defer_last_mlp`` = the MLP block of ``end_layer - 1`` runs on the next stage.
    skip_first_attn: bool = False
    defer_last_mlp: bool = False

    def __post_init__(self):
        if not (0 <= self.start_layer < self.end_layer <= self.num_layers):
            raise ValueError(
                f"invalid layer range [{self.start_layer}, {self.end_layer}) for {self.num_layers} layers"
            )
        if self.skip_first_attn and self.defer_last_mlp and self.num_local_layers == 1:
            raise ValueError("empty stage: both blocks of its only layer live elsewhere")

    @property
    def is_first(self) -> bool:
        return self.start_layer == 0 and not self.skip_first_attn

    @property
    def is_last(self) -> bool:
        return self.end_layer == self.num_layers and not self.defer_last_mlp

    @property
    def num_local_layers(self) -> int:
        return self.end_layer - self.start_layer

    def owns_layer(self, i: int) -> bool:
        return self.start_layer <= i < self.end_layer

    def layers(self):
        return range(self.start_layer, self.end_layer)

    def runs_attn(self, i: int) -> bool:
        return self.owns_layer(i) and not (self.skip_first_attn and i == self.start_layer)

    def runs_mlp(self, i: int) -> bool:
        return self.owns_layer(i) and not (self.defer_last_mlp and i == self.end_layer - 1)

    def attn_layers(self):
        """Layers whose attention block (and therefore KV cache) lives on this stage."""
        return [i for i in self.layers() if self.runs_attn(i)]

    @property
    def num_kv_layers(self) -> int:
        return len(self.attn_layers())

    def describe(self) -> str:
        a = f"{self.start_layer}{'.5' if self.skip_first_attn else ''}"
        b = f"{self.end_layer - 1}.5" if self.defer_last_mlp else f"{self.end_layer}"
        return f"[{a}, {b})"

    @staticmethod
    def even_split(num_layers: int, num_stages: int):
        """Balanced contiguous split, e.g. 27 layers / 2 stages -> [0,14), [14,27)."""
        base, rem = divmod(num_layers, num_stages)
        out, s = [], 0
        for r in range(num_stages):
            n = base + (1 if r < rem else 0)
            out.append(ShardSpec(s, s + n, num_layers))
            s += n
        return out


@dataclass
class ModelConfig:
    """Architecture hyper-parameters parsed from an mlx-community / HF ``config.json``.

    Unknown keys are kept in ``extra`` (the upstream ``ModelArgs.from_dict`` silently drops them).
    """

    model_type: str = "llama"
    hidden_size: int = 4096
    num_hidden_layers: int = 32
    intermediate_size: int = 14336
    num_attention_heads: int = 32
    num_key_value_heads: Optional[int] = None
    head_dim: Optional[int] = None
    rms_norm_eps: float = 1e-5
    vocab_size: int = 32000
    rope_theta: float = 10000.0
    rope_traditional: bool = False
    rope_scaling: Optional[Dict[str, Any]] = None
    max_position_embeddings: int = 8192
    tie_word_embeddings: bool = False
    attention_bias: bool = False
    mlp_bias: bool = False
    # gemma2
    query_pre_attn_scalar: Optional[float] = None
    attn_logit_softcapping: Optional[float] = None
    final_logit_softcapping: Optional[float] = None
    # deepseek_v2
    moe_intermediate_size: Optional[int] = None
    n_shared_experts: Optional[int] = None
    n_routed_experts: Optional[int] = None
    routed_scaling_factor: float = 1.0
    kv_lora_rank: Optional[int] = None
    q_lora_rank: Optional[int] = None
    qk_rope_head_dim: Optional[int] = None
    v_head_dim: Optional[int] = None
    qk_nope_head_dim: Optional[int] = None
    topk_method: str = "greedy"
    n_group: Optional[int] = None
    topk_group: Optional[int] = None
    num_experts_per_tok: Optional[int] = None
    moe_layer_freq: int = 1
    first_k_dense_replace: int = 0
    norm_topk_prob: bool = False
    # quantisation (mlx affine) + sharding
    quantization: Optional[Dict[str, int]] = None
    start_layer: Optional[int] = None
    end_layer: Optional[int] = None
    extra: Dict[str, Any] = field(default_factory=dict)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_dict(cls, d: Dict[str, Any]) -> "ModelConfig":
        d = dict(d)
        mt = d.get("model_type", "llama")
        if mt == "phi-msft":
            raise ValueError(
                "model_type 'phi-msft' is remapped to a non-existent 'phixtral' module by the reference "
                "(shard/utils.py:16); it is not supported"
            )
        d["model_type"] = MODEL_REMAPPING.get(mt, mt)
        if d["model_type"] not in SUPPORTED_ARCHS:
            raise ValueError(f"Model type {mt} not supported (supported: {SUPPORTED_ARCHS})")
        known = {f for f in cls.__dataclass_fields__ if f != "extra"}
        kw = {k: v for k, v in d.items() if k in known}
        extra = {k: v for k, v in d.items() if k not in known}
        # HF >=4.45 nests rope params
        rp = extra.get("rope_parameters")
        if isinstance(rp, dict):
            kw.setdefault("rope_theta", rp.get("rope_theta", 10000.0))
            if rp.get("rope_type", "default") != "default" and "rope_scaling" not in kw:
                kw["rope_scaling"] = rp
        cfg = cls(**kw, extra=extra)
        cfg._finalize()
        return cfg

    @classmethod
    def from_path(cls, path: str) -> "ModelConfig":
        with open(os.path.join(path, "config.json")) as f:
            return cls.from_dict(json.load(f))

    def _finalize(self):
        if self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads
        if self.model_type == "gemma2":
            if self.head_dim is None:
                self.head_dim = 256
            if self.query_pre_attn_scalar is None:
                self.query_pre_attn_scalar = float(self.head_dim)
            if self.attn_logit_softcapping is None:
                self.attn_logit_softcapping = 50.0
            if self.final_logit_softcapping is None:
                self.final_logit_softcapping = 30.0
            # Gemma-2 always ties the LM head to the embedding (reference gemma2.py:80-81)
            self.tie_word_embeddings = True
        elif self.model_type == "deepseek_v2":
            self.rope_traditional = True  # interleaved pairs
        if self.head_dim is None and self.model_type != "deepseek_v2":
            self.head_dim = self.hidden_size // self.num_attention_heads

    def to_dict(self) -> Dict[str, Any]:
        d = {k: getattr(self, k) for k in self.__dataclass_fields__ if k != "extra"}
        d = {k: v for k, v in d.items() if v is not None}
        d.update(self.extra)
        return d

    # ------------------------------------------------------------------ derived
    def shard(self, start_layer: Optional[int] = None, end_layer: Optional[int] = None) -> ShardSpec:
        s = start_layer if start_layer is not None else (self.start_layer if self.start_layer is not None else 0)
        e = end_layer if end_layer is not None else (
            self.end_layer if self.end_layer is not None else self.num_hidden_layers
        )
        return ShardSpec(int(s), int(e), self.num_hidden_layers)

    @property
    def qk_head_dim(self) -> int:
        if self.model_type == "deepseek_v2":
            return self.qk_nope_head_dim + self.qk_rope_head_dim
        return self.head_dim

    @property
    def v_dim(self) -> int:
        if self.model_type == "deepseek_v2":
            return self.v_head_dim
        return self.head_dim

    @property
    def kv_heads(self) -> int:
        # the reference caches *decompressed* per-head K/V for MLA (deepseek_v2.py:120-129)
        if self.model_type == "deepseek_v2":
            return self.num_attention_heads
        return self.num_key_value_heads

    @property
    def attn_scale(self) -> float:
        if self.model_type == "gemma2":
            return float(self.query_pre_attn_scalar) ** -0.5
        if self.model_type == "deepseek_v2":
            scale = self.qk_head_dim ** -0.5
            rs = self.rope_scaling
            if rs is not None:
                mad = rs.get("mscale_all_dim", 0)
                if mad:
                    m = yarn_get_mscale(rs["factor"], mad)
                    scale = scale * m * m
            return scale
        return self.head_dim ** -0.5

    def is_moe_layer(self, i: int) -> bool:
        return (
            self.model_type == "deepseek_v2"
            and self.n_routed_experts is not None
            and i >= self.first_k_dense_replace
            and i % self.moe_layer_freq == 0
        )


def yarn_get_mscale(scale: float = 1.0, mscale: float = 1.0) -> float:
    if scale <= 1:
        return 1.0
    return 0.1 * mscale * math.log(scale) + 1.0


# ---------------------------------------------------------------------- canned configs (offline box)
def deepseek_v2_lite_config(**overrides) -> Dict[str, Any]:
    """config.json of DeepSeek-Coder-V2-Lite-Instruct (the reference's demo model, README.md:26)."""
    d = dict(
        model_type="deepseek_v2", architectures=["DeepseekV2ForCausalLM"], vocab_size=102400,
        hidden_size=2048, intermediate_size=10944, moe_intermediate_size=1408, num_hidden_layers=27,
        num_attention_heads=16, num_key_value_heads=16, n_shared_experts=2, n_routed_experts=64,
        routed_scaling_factor=1.0, kv_lora_rank=512, q_lora_rank=None, qk_rope_head_dim=64,
        v_head_dim=128, qk_nope_head_dim=128, topk_method="greedy", n_group=1, topk_group=1,
        num_experts_per_tok=6, moe_layer_freq=1, first_k_dense_replace=1, norm_topk_prob=False,
        hidden_act="silu", max_position_embeddings=163840, rms_norm_eps=1e-6, rope_theta=10000.0,
        rope_scaling=dict(beta_fast=32, beta_slow=1, factor=40, mscale=0.707, mscale_all_dim=0.707,
                          original_max_position_embeddings=4096, type="yarn"),
        attention_bias=False, tie_word_embeddings=False, bos_token_id=100000, eos_token_id=100001,
    )
    d.update(overrides)
    return d


def llama3_8b_config(**overrides) -> Dict[str, Any]:
    d = dict(
        model_type="llama", architectures=["LlamaForCausalLM"], vocab_size=128256, hidden_size=4096,
        intermediate_size=14336, num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=8,
        rms_norm_eps=1e-5, rope_theta=500000.0, max_position_embeddings=8192, tie_word_embeddings=False,
        attention_bias=False, bos_token_id=128000, eos_token_id=128009,
    )
    d.update(overrides)
    return d


def gemma2_9b_config(**overrides) -> Dict[str, Any]:
    d = dict(
        model_type="gemma2", architectures=["Gemma2ForCausalLM"], vocab_size=256000, hidden_size=3584,
        intermediate_size=14336, num_hidden_layers=42, num_attention_heads=16, num_key_value_heads=8,
        head_dim=256, rms_norm_eps=1e-6, rope_theta=10000.0, query_pre_attn_scalar=256,
        attn_logit_softcapping=50.0, final_logit_softcapping=30.0, max_position_embeddings=8192,
    )
    d.update(overrides)
    return d
