"""Sharded Gemma-2.

Reference: ``shard/server/model/gemma2.py`` + upstream block: four norms per block with the
``(1 + w)`` RMSNorm, ``query_pre_attn_scalar^-0.5`` attention scale, attention-logit soft-capping
``tanh(s/50)*50``, GeGLU MLP, ``sqrt(H)``-scaled embeddings, tied head with final-logit soft-cap
``tanh(x/30)*30`` (gemma2.py:80-84).  Like the upstream of that era, every layer is global attention
(no sliding window).
"""
from __future__ import annotations

import math

from .llama import LlamaStage


class Gemma2Stage(LlamaStage):
    arch = "gemma2"
    act = "gelu_tanh"
    gemma = True

    PRE_MLP_NORM = "pre_feedforward_layernorm"

    def _load_layer(self, sd, i, attn=True, mlp=True) -> dict:
        w = super()._load_layer(sd, i, attn, mlp)
        p = f"model.layers.{i}"
        if attn:
            w["post_ln"] = self._vec(sd, p + ".post_attention_layernorm.weight")
        if mlp:
            w["post_ffn_ln"] = self._vec(sd, p + ".post_feedforward_layernorm.weight")
        return w

    def embed(self, ids):
        # reference gemma2.py:42-43: h = embed(ids) * sqrt(hidden_size)
        return self.ops.embed(ids, self.embed_tokens, math.sqrt(self.cfg.hidden_size), self.dtype)

    def attn_block(self, i, h, meta, kpool, vpool):
        O, c, w = self.ops, self.cfg, self.layer_weights[i]
        eps = c.rms_norm_eps
        normed = O.rmsnorm(h, w["in_ln"], eps, True)
        attn = self._attention(w, normed, meta, kpool, vpool)
        a = O.linear(attn, w["o"])
        # a stage may end here (half-layer boundary): the norm's epilogue is then the fused P2P hand-off (ops/b200.py::rmsnorm)
        return O.rmsnorm(a, w["post_ln"], eps, True, residual=h, **self._final_kwargs(i, h.shape[0], "attn"))

    def mlp_block(self, i, h, meta):
        O, c, w = self.ops, self.cfg, self.layer_weights[i]
        eps = c.rms_norm_eps
        normed = O.rmsnorm(h, w["mlp_ln"], eps, True)
        act = O.gated_up(normed, w["gate"], w["up"], self.act)
        m = O.linear(act, w["down"])
        # Gemma-2's stage-final kernel is this norm, not a GEMM: it stores into the next stage's inbox and raises its flag itself
        return O.rmsnorm(m, w["post_ffn_ln"], eps, True, residual=h, **self._final_kwargs(i, h.shape[0], "mlp"))


Model = Gemma2Stage
