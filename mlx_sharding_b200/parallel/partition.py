"""Cost-balanced contiguous partition of the decoder stack for the pipeline.

The reference asks the user for ``--start-layer/--end-layer`` per process (README.md:76-90, e.g. 0-14 / 14-27
for DeepSeek-V2-Lite).  Those flags still work; when they are omitted the stages are balanced by a cost model of
one decode step: bytes streamed from HBM (weights dominate at decode batch sizes) plus a fixed per-kernel latency.
The LM head (vocab x hidden) counts as a fraction of a layer and the dense first layer of DeepSeek is much cheaper
than an MoE layer, so an even layer count is not an even split.

``half_layers=True`` additionally allows a stage boundary *between* the attention block and the MLP block of a
layer (both only exchange the residual stream, so the hand-off message is the same ``[T, H]`` tensor).  With 27
layers on 8 GPUs whole-layer granularity leaves the busiest stage with 4 layers against an ideal of 3.4; half-layer
granularity brings the busiest stage to within a few percent of the mean.
"""
from __future__ import annotations

from typing import List, Tuple

from ..config import ModelConfig, ShardSpec

# cost model constants (seconds): measured HBM stream rate of the grouped / dense swap-AB GEMMs and the per-kernel
# latency floor inside a CUDA graph with programmatic dependent launch (profiles/results.md)
_BYTES_PER_S = 6.5e12
_LAUNCH_S = 3.0e-6


def block_costs(cfg: ModelConfig, i: int, wbytes: float = 2.0) -> Tuple[float, float]:
    """(attention block, MLP block) cost of layer ``i`` in seconds per decode step."""
    H = cfg.hidden_size
    if cfg.model_type == "deepseek_v2":
        nh = cfg.num_attention_heads
        qd = cfg.qk_nope_head_dim + cfg.qk_rope_head_dim
        attn = H * nh * qd + H * (cfg.kv_lora_rank + cfg.qk_rope_head_dim) + \
            cfg.kv_lora_rank * nh * (cfg.qk_nope_head_dim + cfg.v_head_dim) + nh * cfg.v_head_dim * H
        attn_launches = 8
        if cfg.is_moe_layer(i):
            mlp = 3 * H * cfg.moe_intermediate_size * (cfg.n_routed_experts + (cfg.n_shared_experts or 0))
            mlp_launches = 10
        else:
            mlp = 3 * H * cfg.intermediate_size
            mlp_launches = 3
    else:
        hd = cfg.head_dim
        attn = H * hd * (cfg.num_attention_heads * 2 + cfg.num_key_value_heads * 2)
        attn_launches = 7 if cfg.model_type != "gemma2" else 8
        mlp = 3 * H * cfg.intermediate_size
        mlp_launches = 3 if cfg.model_type != "gemma2" else 4
    return (attn * wbytes / _BYTES_PER_S + attn_launches * _LAUNCH_S,
            mlp * wbytes / _BYTES_PER_S + mlp_launches * _LAUNCH_S)


def layer_cost(cfg: ModelConfig, i: int) -> float:
    a, m = block_costs(cfg, i)
    return a + m


def head_cost(cfg: ModelConfig, wbytes: float = 2.0) -> float:
    return cfg.vocab_size * cfg.hidden_size * wbytes / _BYTES_PER_S + 3 * _LAUNCH_S   # norm, LM head, sampler


def _min_max_partition(cost: List[float], tail: float, parts: int) -> List[Tuple[int, int]]:
    """Contiguous partition of ``cost`` into ``parts`` non-empty segments minimising the largest segment sum
    (``tail`` is added to the last segment).  Exact DP; len(cost) <= a few hundred."""
    n = len(cost)
    pre = [0.0]
    for c in cost:
        pre.append(pre[-1] + c)
    INF = float("inf")
    best = [[INF] * (n + 1) for _ in range(parts + 1)]
    cut = [[0] * (n + 1) for _ in range(parts + 1)]
    best[0][0] = 0.0
    for s in range(1, parts + 1):
        for e in range(s, n + 1):
            if s < parts and e == n:
                continue
            for a in range(s - 1, e):
                if best[s - 1][a] == INF:
                    continue
                v = max(best[s - 1][a], pre[e] - pre[a] + (tail if (s == parts and e == n) else 0.0))
                if v < best[s][e]:
                    best[s][e], cut[s][e] = v, a
    bounds, e = [], n
    for s in range(parts, 0, -1):
        a = cut[s][e]
        bounds.append((a, e))
        e = a
    return list(reversed(bounds))


def balanced_split(cfg: ModelConfig, num_stages: int, half_layers: bool = False) -> List[ShardSpec]:
    """Contiguous partition minimising the most expensive stage."""
    L = cfg.num_hidden_layers
    wbytes = 2.0
    if cfg.quantization:
        wbytes = cfg.quantization["bits"] / 8.0 + 4.0 / cfg.quantization["group_size"]
    if not half_layers:
        if num_stages >= L:
            return ShardSpec.even_split(L, min(num_stages, L))
        cost = [sum(block_costs(cfg, i, wbytes)) for i in range(L)]
        return [ShardSpec(a, b, L) for a, b in _min_max_partition(cost, head_cost(cfg, wbytes), num_stages)]
    units: List[float] = []           # unit 2i = attention block of layer i, unit 2i+1 = its MLP block
    for i in range(L):
        units.extend(block_costs(cfg, i, wbytes))
    num_stages = min(num_stages, 2 * L)
    out = []
    for a, b in _min_max_partition(units, head_cost(cfg, wbytes), num_stages):
        out.append(ShardSpec(a // 2, (b + 1) // 2, L, skip_first_attn=bool(a % 2), defer_last_mlp=bool(b % 2)))
    return out


def stage_cost(cfg: ModelConfig, spec: ShardSpec) -> float:
    """Modelled seconds per decode step of one stage (for logs / tests)."""
    t = 0.0
    for i in spec.layers():
        a, m = block_costs(cfg, i)
        t += (a if spec.runs_attn(i) else 0.0) + (m if spec.runs_mlp(i) else 0.0)
    return t + (head_cost(cfg) if spec.is_last else 0.0)
