// Paged-KV attention: split-KV flash-decoding with online softmax + LSE combine.
//
// The reference runs `mx.fast.scaled_dot_product_attention` over one contiguous, ever-growing KVCache
// with a dense additive mask (shard/server/model/llama.py:48-59 via mlx_lm) for exactly one sequence.
// This kernel serves a ragged batch against a *paged* cache: every query token carries its own causal
// limit (kv_len = position + 1) and its sequence's block table, so the same kernel handles decode
// micro-batches, chunked prefill and mixed batches.  d_qk != d_v (MLA 192/128), GQA and Gemma-2 logit
// soft-capping are supported.
//
// Work decomposition: CTA = (query token, kv head, kv split); the G = Hq/Hk query heads of the group
// share every K/V row read.  A warp owns a strided subset of the KV positions; lanes split the head
// dimension in bf16x2 pairs (coalesced 128 B per warp per load), dot products are finished with an
// xor-shuffle tree, and each lane keeps the slice of the output accumulator matching its V pairs.
// Bandwidth-bound by design: every K/V byte is read once per (token, kv head).
#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"

namespace b200 {

namespace {

constexpr int kAttnWarps = 4;
constexpr int kAttnThreads = kAttnWarps * 32;
constexpr int kTokPerIter = 4;  // KV positions processed per warp iteration (ILP)

template <int DK, int DV, int G>
__global__ void __launch_bounds__(kAttnThreads)
paged_attn_kernel(const __nv_bfloat16* __restrict__ q, long long q_ld_t, long long q_ld_h,
                  const __nv_bfloat16* __restrict__ kpool, const __nv_bfloat16* __restrict__ vpool,
                  const int* __restrict__ block_tables, int max_blocks, const int* __restrict__ positions,
                  const int* __restrict__ token_seq, int kv_heads, int page, float scale, float softcap,
                  int nsplit, __nv_bfloat16* __restrict__ out, long long o_ld_t, float* __restrict__ part_acc,
                  float* __restrict__ part_ml) {
  pdl_sync();  // PDL: predecessor's writes visible; let the successor start its prologue
  constexpr int PK = DK / 64;  // bf16x2 pairs per lane for K
  constexpr int PV = DV / 64;  // pairs per lane for V
  static_assert(DK % 64 == 0 && DV % 64 == 0, "head dims must be multiples of 64");
  const int t = blockIdx.x, hk = blockIdx.y, sp = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = token_seq[t];
  const int kv_len = positions[t] + 1;
  const int per = (kv_len + nsplit - 1) / nsplit;
  const int begin = sp * per;
  int end = begin + per;
  if (end > kv_len) end = kv_len;

  // q fragments of the G heads of this group, pre-scaled
  float qf[G][2 * PK];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const __nv_bfloat16* qp = q + (size_t)t * q_ld_t + (size_t)(hk * G + g) * q_ld_h;
#pragma unroll
    for (int i = 0; i < PK; ++i) {
      const uint32_t u = reinterpret_cast<const uint32_t*>(qp)[lane + 32 * i];
      qf[g][2 * i] = bf16_lo(u) * scale;
      qf[g][2 * i + 1] = bf16_hi(u) * scale;
    }
  }
  float m[G], l[G], acc[G][2 * PV];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    m[g] = -INFINITY; l[g] = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * PV; ++i) acc[g][i] = 0.f;
  }
  const int* bt = block_tables + (size_t)seq * max_blocks;

  for (int base = begin + warp * kTokPerIter; base < end; base += kAttnWarps * kTokPerIter) {
    uint32_t kreg[kTokPerIter][PK], vreg[kTokPerIter][PV];
#pragma unroll
    for (int j = 0; j < kTokPerIter; ++j) {
      int pos = base + j;
      if (pos >= end) pos = end - 1;  // clamp: loads stay in-bounds, contribution masked below
      const size_t pg = bt[pos / page];
      const size_t row = (pg * kv_heads + hk) * page + (pos % page);
      const uint32_t* kp = reinterpret_cast<const uint32_t*>(kpool + row * DK);
      const uint32_t* vp = reinterpret_cast<const uint32_t*>(vpool + row * DV);
#pragma unroll
      for (int i = 0; i < PK; ++i) kreg[j][i] = kp[lane + 32 * i];
#pragma unroll
      for (int i = 0; i < PV; ++i) vreg[j][i] = vp[lane + 32 * i];
    }
#pragma unroll
    for (int j = 0; j < kTokPerIter; ++j) {
      const bool valid = (base + j) < end;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < PK; ++i) {
          s += qf[g][2 * i] * bf16_lo(kreg[j][i]);
          s += qf[g][2 * i + 1] * bf16_hi(kreg[j][i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (softcap > 0.f) s = softcap * tanhf(s / softcap);
        if (valid) {
          const float mn = fmaxf(m[g], s);
          const float corr = __expf(m[g] - mn);
          const float pexp = __expf(s - mn);
          l[g] = l[g] * corr + pexp;
#pragma unroll
          for (int i = 0; i < PV; ++i) {
            acc[g][2 * i] = acc[g][2 * i] * corr + pexp * bf16_lo(vreg[j][i]);
            acc[g][2 * i + 1] = acc[g][2 * i + 1] * corr + pexp * bf16_hi(vreg[j][i]);
          }
          m[g] = mn;
        }
      }
    }
  }

  // ---- merge the 4 warps through shared memory (log-sum-exp combine)
  __shared__ float s_m[kAttnWarps][G], s_l[kAttnWarps][G];
  __shared__ float s_acc[kAttnWarps][G][DV];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    if (lane == 0) { s_m[warp][g] = m[g]; s_l[warp][g] = l[g]; }
#pragma unroll
    for (int i = 0; i < PV; ++i) {
      s_acc[warp][g][2 * (lane + 32 * i)] = acc[g][2 * i];
      s_acc[warp][g][2 * (lane + 32 * i) + 1] = acc[g][2 * i + 1];
    }
  }
  __syncthreads();
  const int Hq = kv_heads * G;
  for (int idx = threadIdx.x; idx < G * DV; idx += kAttnThreads) {
    const int g = idx / DV, d = idx % DV;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) mm = fmaxf(mm, s_m[w][g]);
    float ll = 0.f, a = 0.f;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) {
      const float c = (s_m[w][g] == -INFINITY) ? 0.f : __expf(s_m[w][g] - mm);
      ll += s_l[w][g] * c;
      a += s_acc[w][g][d] * c;
    }
    const int hq = hk * G + g;
    if (nsplit == 1) {
      out[(size_t)t * o_ld_t + (size_t)hq * DV + d] = __float2bfloat16_rn(ll > 0.f ? a / ll : 0.f);
    } else {
      const size_t pidx = ((size_t)t * Hq + hq) * nsplit + sp;
      part_acc[pidx * DV + d] = a;
      if (d == 0) { part_ml[pidx * 2] = mm; part_ml[pidx * 2 + 1] = ll; }
    }
  }
}

template <int DV>
__global__ void attn_combine_kernel(const float* __restrict__ part_acc, const float* __restrict__ part_ml, int nsplit,
                                    __nv_bfloat16* __restrict__ out, long long o_ld_t, int Hq) {
  pdl_sync();  // PDL: predecessor's writes visible; let the successor start its prologue
  const int t = blockIdx.x, hq = blockIdx.y;
  const size_t base = ((size_t)t * Hq + hq) * nsplit;
  float mm = -INFINITY;
  for (int s = 0; s < nsplit; ++s) mm = fmaxf(mm, part_ml[(base + s) * 2]);
  for (int d = threadIdx.x; d < DV; d += blockDim.x) {
    float ll = 0.f, a = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float ms = part_ml[(base + s) * 2];
      const float c = (ms == -INFINITY) ? 0.f : __expf(ms - mm);
      ll += part_ml[(base + s) * 2 + 1] * c;
      a += part_acc[(base + s) * DV + d] * c;
    }
    out[(size_t)t * o_ld_t + (size_t)hq * DV + d] = __float2bfloat16_rn(ll > 0.f ? a / ll : 0.f);
  }
}

template <int DK, int DV, int G>
cudaError_t launch_attn(const PagedAttnArgs& a, cudaStream_t s) {
  dim3 grid(a.T, a.kv_heads, a.nsplit);
  (void)launch_pdl(paged_attn_kernel<DK, DV, G>, dim3(grid), dim3(kAttnThreads), 0, s, 
      static_cast<const __nv_bfloat16*>(a.q), a.q_ld_t, a.q_ld_h, static_cast<const __nv_bfloat16*>(a.kpool),
      static_cast<const __nv_bfloat16*>(a.vpool), a.block_tables, a.max_blocks, a.positions, a.token_seq, a.kv_heads,
      a.page, a.scale, a.softcap, a.nsplit, static_cast<__nv_bfloat16*>(a.out), a.o_ld_t, a.part_acc, a.part_ml);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (a.nsplit > 1) {
    dim3 g2(a.T, a.kv_heads * G);
    (void)launch_pdl(attn_combine_kernel<DV>, dim3(g2), dim3(64), 0, s, a.part_acc, a.part_ml, a.nsplit, static_cast<__nv_bfloat16*>(a.out), a.o_ld_t,
                                              a.kv_heads * G);
    e = cudaGetLastError();
  }
  return e;
}

}  // namespace

cudaError_t paged_attention_launch(const PagedAttnArgs& a, cudaStream_t s) {
  if (a.T == 0) return cudaSuccess;
  if ((a.q_ld_t % 2) || (a.q_ld_h % 2)) return cudaErrorInvalidValue;
  const int G = a.q_heads / a.kv_heads;
#define ATTN_CASE(dk, dv, g) if (a.dk_ == dk && a.dv_ == dv && G == g) return launch_attn<dk, dv, g>(a, s);
  ATTN_CASE(192, 128, 1)   // DeepSeek-V2 MLA, decompressed cache
  ATTN_CASE(128, 128, 1)
  ATTN_CASE(128, 128, 2)
  ATTN_CASE(128, 128, 4)   // Llama-3-8B: 32 q heads / 8 kv heads
  ATTN_CASE(128, 128, 8)
  ATTN_CASE(256, 256, 1)
  ATTN_CASE(256, 256, 2)   // Gemma-2-9B
  ATTN_CASE(64, 64, 1)
  ATTN_CASE(64, 64, 2)
  ATTN_CASE(64, 64, 4)
#undef ATTN_CASE
  return cudaErrorNotSupported;
}

}  // namespace b200
