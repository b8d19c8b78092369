"""Shared test helpers: tiny configs, single-sequence driver."""
import torch

from mlx_sharding_b200.config import ModelConfig
from mlx_sharding_b200.engine.kv_cache import PagedKVCache
from mlx_sharding_b200.ops.meta import BatchMeta

TINY_LLAMA = dict(model_type="llama", vocab_size=320, hidden_size=64, intermediate_size=128,
                  num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, rms_norm_eps=1e-5,
                  rope_theta=10000.0, max_position_embeddings=512, tie_word_embeddings=False)
TINY_GEMMA2 = dict(model_type="gemma2", vocab_size=320, hidden_size=64, intermediate_size=128,
                   num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=32,
                   rms_norm_eps=1e-6, rope_theta=10000.0, query_pre_attn_scalar=32,
                   attn_logit_softcapping=50.0, final_logit_softcapping=30.0, max_position_embeddings=512,
                   sliding_window=4096)
TINY_DSV2 = dict(model_type="deepseek_v2", vocab_size=320, hidden_size=64, intermediate_size=160,
                 moe_intermediate_size=48, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=4,
                 n_shared_experts=2, n_routed_experts=8, routed_scaling_factor=1.0, kv_lora_rank=32,
                 q_lora_rank=None, qk_rope_head_dim=16, v_head_dim=24, qk_nope_head_dim=24,
                 topk_method="greedy", n_group=1, topk_group=1, num_experts_per_tok=3, moe_layer_freq=1,
                 first_k_dense_replace=1, norm_topk_prob=False, rms_norm_eps=1e-6, rope_theta=10000.0,
                 max_position_embeddings=512, tie_word_embeddings=False)


def run_sequence(stages, tokens, n_decode=0, page_size=16, chunk=None, greedy_decode=True):
    """Run one sequence through a chain of stage models (list), prefill (optionally chunked) then
    ``n_decode`` greedy steps.  Returns list of last-position logits (one per step)."""
    tokens = list(tokens)
    total = len(tokens) + n_decode
    npages = (total + page_size - 1) // page_size + 2
    kvs = [PagedKVCache.for_model(m, npages, page_size) for m in stages]
    pages = list(range(1, npages))
    outs = []
    done = 0
    chunk = chunk or len(tokens)
    dev = stages[0].device

    def step(ids, c0):
        meta = BatchMeta.build([len(ids)], [c0], [pages], page_size, device=dev)
        x = torch.tensor(ids, dtype=torch.int64, device=dev)
        for m, kv in zip(stages, kvs):
            x = m(x, meta, kv)
        return x

    while done < len(tokens):
        ids = tokens[done:done + chunk]
        logits = step(ids, done)
        done += len(ids)
    outs.append(logits[0].float().cpu())
    cur = tokens
    for _ in range(n_decode):
        nxt = int(outs[-1].argmax())
        logits = step([nxt], done)
        done += 1
        outs.append(logits[0].float().cpu())
    return outs

# GPU-shaped tiny configs: head dims are multiples of 64, hidden/intermediate multiples of 128 (kernel tiles)
GPU_LLAMA = dict(model_type="llama", vocab_size=512, hidden_size=256, intermediate_size=512,
                 num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, rms_norm_eps=1e-5,
                 rope_theta=10000.0, max_position_embeddings=4096, tie_word_embeddings=False)
GPU_GEMMA2 = dict(model_type="gemma2", vocab_size=512, hidden_size=256, intermediate_size=512,
                  num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=64,
                  rms_norm_eps=1e-6, rope_theta=10000.0, query_pre_attn_scalar=64,
                  attn_logit_softcapping=50.0, final_logit_softcapping=30.0, max_position_embeddings=4096)
GPU_DSV2 = dict(model_type="deepseek_v2", vocab_size=512, hidden_size=256, intermediate_size=512,
                moe_intermediate_size=128, num_hidden_layers=4, num_attention_heads=2, num_key_value_heads=2,
                n_shared_experts=2, n_routed_experts=8, routed_scaling_factor=1.0, kv_lora_rank=64,
                q_lora_rank=None, qk_rope_head_dim=64, v_head_dim=128, qk_nope_head_dim=128,
                topk_method="greedy", n_group=1, topk_group=1, num_experts_per_tok=3, moe_layer_freq=1,
                first_k_dense_replace=1, norm_topk_prob=False, rms_norm_eps=1e-6, rope_theta=10000.0,
                max_position_embeddings=4096, tie_word_embeddings=False,
                rope_scaling=dict(beta_fast=32, beta_slow=1, factor=40, mscale=0.707, mscale_all_dim=0.707,
                                  original_max_position_embeddings=4096, type="yarn"))
