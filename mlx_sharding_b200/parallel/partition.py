"""Cost-balanced contiguous layer partition for the pipeline.

The reference asks the user for ``--start-layer/--end-layer`` per process (README.md:76-90, e.g. 0-14 / 14-27
for DeepSeek-V2-Lite).  Those flags still work; when they are omitted the stages are balanced by the bytes
each stage streams per decode step — the LM head (vocab x hidden) counts as a fraction of a layer and the
dense first layer of DeepSeek is much cheaper than an MoE layer, so an even layer count is not an even split.
"""
from __future__ import annotations

from typing import List

from ..config import ModelConfig, ShardSpec


def layer_cost(cfg: ModelConfig, i: int) -> float:
    H = cfg.hidden_size
    if cfg.model_type == "deepseek_v2":
        nh = cfg.num_attention_heads
        qd = cfg.qk_nope_head_dim + cfg.qk_rope_head_dim
        attn = H * nh * qd + H * (cfg.kv_lora_rank + cfg.qk_rope_head_dim) + \
            cfg.kv_lora_rank * nh * (cfg.qk_nope_head_dim + cfg.v_head_dim) + nh * cfg.v_head_dim * H
        if cfg.is_moe_layer(i):
            mlp = 3 * H * cfg.moe_intermediate_size * (cfg.n_routed_experts + (cfg.n_shared_experts or 0))
        else:
            mlp = 3 * H * cfg.intermediate_size
        return float(attn + mlp)
    hd = cfg.head_dim
    attn = H * hd * (cfg.num_attention_heads * 2 + cfg.num_key_value_heads * 2)
    return float(attn + 3 * H * cfg.intermediate_size)


def balanced_split(cfg: ModelConfig, num_stages: int) -> List[ShardSpec]:
    """Contiguous partition minimising the most expensive stage (exact DP; L <= a few hundred)."""
    L = cfg.num_hidden_layers
    if num_stages >= L:
        return ShardSpec.even_split(L, min(num_stages, L))
    cost = [layer_cost(cfg, i) for i in range(L)]
    head = float(cfg.vocab_size * cfg.hidden_size)
    pre = [0.0]
    for c in cost:
        pre.append(pre[-1] + c)

    def seg(a, b, last):
        return pre[b] - pre[a] + (head if last else 0.0)

    INF = float("inf")
    best = [[INF] * (L + 1) for _ in range(num_stages + 1)]
    cut = [[0] * (L + 1) for _ in range(num_stages + 1)]
    best[0][0] = 0.0
    for s in range(1, num_stages + 1):
        for e in range(s, L + 1):
            if s < num_stages and e == L:
                continue
            for a in range(s - 1, e):
                if best[s - 1][a] == INF:
                    continue
                v = max(best[s - 1][a], seg(a, e, s == num_stages and e == L))
                if v < best[s][e]:
                    best[s][e], cut[s][e] = v, a
    bounds, e = [], L
    for s in range(num_stages, 0, -1):
        a = cut[s][e]
        bounds.append((a, e))
        e = a
    return [ShardSpec(a, b, L) for a, b in reversed(bounds)]
