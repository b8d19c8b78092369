// Absorbed-latent MLA decode attention on tcgen05 / TMEM (SURVEY §2.6 K7, K5 "weight absorption").
//
// DeepSeek-V2's MLA caches one 512-dim latent c_t (+ a 64-dim roped key) per token.  The reference decompresses it to 16 heads of
// K 192 / V 128 and caches *that* (shard/server/model/deepseek_v2.py:120-125 -> 5120 values per token per layer).  With W_UK folded
// into the query and W_UV into the output projection, attention runs directly on the cached latent:
//
//     s[t, h] = q_abs[h] . c_t + q_pe[h] . k_pe_t            (576-dim dot, one shared "KV head")
//     o_lat[h] = sum_t softmax_t(s)[t, h] * c_t              (512-dim)
//
// i.e. 576 + 0 extra values per token instead of 5120: 8.9x fewer KV bytes, and both contractions are GEMM-shaped over the SAME
// shared-memory tile, so they run on the tensor cores:
//
//   * one CTA per (sequence, KV split); KV tile = one 64-token page = [64 x 576] bf16 fetched by TMA (9 boxes of 64 x 64, 128B
//     swizzle) into a double-buffered ring;
//   * S^T[64 tokens, 16 heads] = K_tile[64 x 576] . Q^T: UMMA M=64 (tokens), N=16 (heads), K-major operands, 36 MMAs, fp32 in TMEM;
//   * the 4 softmax warps read S^T with tcgen05.ld (thread = token row), do the online softmax across the tile with warp shuffles,
//     write P^T as bf16 into shared memory in the K-major swizzled layout;
//   * O^T[512 dims, 16 heads] = V^T . P^T where V^T is the *same* tile read as an MN-major operand (dims contiguous) — no
//     transpose, no second copy: 4 UMMA M-tiles of 128 dims x 4 k-steps of 16 tokens;
//   * S and O accumulators are double-buffered in TMEM, P in shared memory: QK of tile i+1 is issued before PV of tile i, so the
//     tensor pipe, the softmax warps and the TMA stream overlap; the running O lives in registers (thread = dim row, 64 fp32);
//   * long contexts are split over CTAs (flash-decoding): partial (o, m, l) in fp32 + a small combine kernel.
//
// Decode only (one query token per sequence); prefill chunks take the decompressing path (models/deepseek_v2.py).
#include <cuda.h>

#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"

namespace b200 {

namespace {

constexpr int kHeads = 16;
constexpr int kLat = 512;          // latent / value dim
constexpr int kRope = 64;
constexpr int kDk = kLat + kRope;  // 576
constexpr int kChunks = kDk / 64;  // 9 chunks of 64 dims (128 B rows)
constexpr int kTileTok = 64;       // tokens per tile = one KV page
constexpr int kChunkBytes = kTileTok * 128;          // 8 KB
constexpr int kTileBytes = kChunks * kChunkBytes;    // 72 KB
constexpr int kQChunkBytes = kHeads * 128;           // 2 KB
constexpr int kQBytes = kChunks * kQChunkBytes;      // 18 KB
constexpr int kPBytes = kHeads * 128;                // [16 heads x 64 tokens] bf16 = 2 KB
constexpr int kThreads = 352;                        // warp 0: TMA, warp 1: QK issuer + TMEM, warp 2: PV issuer, warps 3..10: softmax
constexpr uint32_t kTmemCols = 256;                  // S[2] x 16 + O[2] x 64 = 160 -> 256
constexpr uint32_t kSCol = 0, kOCol = 32;

constexpr uint32_t idesc(uint32_t M, uint32_t N, uint32_t a_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
constexpr uint32_t kIdescQK = idesc(64, kHeads, 0);
constexpr uint32_t kIdescPV = idesc(128, kHeads, 1);

// MN-major, 128B-swizzled operand (cute canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): 64 MN-elements are
// contiguous (one 128 B line), 8 consecutive K rows are 128 B apart, groups of 8 K rows are SBO apart, 64-element MN blocks LBO apart.
B200_DEVICE uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

B200_DEVICE void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

struct MlaParams {
  const int* block_tables; int max_blocks;
  const int* context_lens;                 // tokens in the cache per sequence (including the one written this step)
  int nsplit, tiles_per_split;
  float scale_log2;                        // softmax scale * log2(e)
  __nv_bfloat16* out; long long o_ld_t;    // [B, 16, 512] (row stride o_ld_t) when nsplit == 1
  float* part_o; float* part_ml;           // [B, nsplit, 16, 512], [B, nsplit, 16, 2]
  long long* dbg;                          // optional [6][16] clock64 timestamps of CTA (0,0) (MLXB200_MLA_TRACE, bench/mla_bench.py)
};

#define MLA_TRACE(ev, tile) \
  do { if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && (tile) < 16) p.dbg[(ev) * 16 + (tile)] = clock64(); } while (0)

__global__ void __launch_bounds__(kThreads, 1)
mla_decode_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv, const MlaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* kv_smem = smem;                                   // 2 x 72 KB
  uint8_t* q_smem = smem + 2 * kTileBytes;                   // 18 KB
  uint8_t* p_smem = q_smem + kQBytes;                        // 2 x 2 KB
  float* red = reinterpret_cast<float*>(p_smem + 2 * kPBytes);   // [2][4][16] cross-warp max / sum
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 2 * 4 * kHeads);
  uint64_t* kv_full = bars;          // [2]
  uint64_t* kv_free = bars + 2;      // [2]
  uint64_t* q_full = bars + 4;
  uint64_t* s_full = bars + 5;       // [2]
  uint64_t* p_ready = bars + 7;      // [2]
  uint64_t* o_full = bars + 9;       // [2]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, b = blockIdx.y;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    for (int i = 0; i < 11; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_smem, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_launch_dependents();
  pdl_wait();  // q, the latent pool and the step metadata are written by predecessor kernels

  const int ctx = p.context_lens[b];
  const int ntiles_seq = (ctx + kTileTok - 1) / kTileTok;
  const int t_begin = split * p.tiles_per_split;
  int t_end = t_begin + p.tiles_per_split;
  if (t_end > ntiles_seq) t_end = ntiles_seq;
  const int n = t_end > t_begin ? t_end - t_begin : 0;

  if (warp == 0) {
    // ============================================================== TMA producer (whole warp waits, one elected lane issues)
    if (n > 0) {
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, kQBytes);
        for (int c = 0; c < kChunks; ++c) tma_load_3d(q_smem + c * kQChunkBytes, &tmap_q, q_full, c * 64, 0, b);
      }
      const int* bt = p.block_tables + (size_t)b * p.max_blocks;
      int page_next = bt[t_begin];
      for (int i = 0; i < n; ++i) {
        const int st = i & 1;
        const int page = page_next;
        if (i + 1 < n) page_next = bt[t_begin + i + 1];   // the dependent global load overlaps the wait below
        mbar_wait(&kv_free[st], ((i >> 1) & 1) ^ 1);
        if (elect_one()) {
          MLA_TRACE(0, i);
          uint8_t* dst = kv_smem + st * kTileBytes;
          mbar_arrive_expect_tx(&kv_full[st], kTileBytes);
#pragma unroll
          for (int c = 0; c < kChunks; ++c)
            tma_load_2d(dst + c * kChunkBytes, &tmap_kv, &kv_full[st], c * 64, page * kTileTok, kEvictFirst);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ============================================================== QK issuer (single thread)
    // The MMAs of this kernel are tiny (64 x 16 x 16: 8 tensor-pipe cycles) while issuing one costs ~50 cycles (descriptor set-up
    // + the compiler's uniform-datapath election loop), so ONE issuing thread is the bottleneck (measured: 2.1 us per tile).  QK
    // and PV write different TMEM accumulators and are ordered through mbarriers only, so they get an issuing thread each.
    if (n > 0) {
      mbar_wait(q_full, 0);
      for (int j = 0; j < n; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_full[st], (j >> 1) & 1);
        if (j >= 2) mbar_wait(&p_ready[st], ((j - 2) >> 1) & 1);   // softmax has drained S[st] of tile j-2
        tc_fence_after();
        if (elect_one()) {
          MLA_TRACE(1, j);
          const uint32_t a0 = smem_u32(kv_smem + st * kTileBytes), b0 = smem_u32(q_smem);
          const uint32_t d = tmem_base + kSCol + st * kHeads;
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            const uint64_t ad = umma_desc_sw128(a0 + c * kChunkBytes), bd = umma_desc_sw128(b0 + c * kQChunkBytes);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_f16(d, ad + 2 * kk, bd + 2 * kk, kIdescQK, (c | kk) ? 1u : 0u);
          }
          umma_commit(&s_full[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ============================================================== PV issuer (single thread)
    if (n > 0) {
      for (int i = 0; i < n; ++i) {
        const int st = i & 1;
        mbar_wait(&p_ready[st], (i >> 1) & 1);
        tc_fence_after();
        if (elect_one()) {
          MLA_TRACE(2, i);
          const uint32_t v0 = smem_u32(kv_smem + st * kTileBytes), p0 = smem_u32(p_smem + st * kPBytes);
          const uint64_t pd = umma_desc_sw128(p0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t d = tmem_base + kOCol + st * 64 + j * kHeads;
#pragma unroll
            for (int ks = 0; ks < kTileTok / 16; ++ks) {
              const uint64_t ad = umma_desc_mn_sw128(v0 + (2 * j) * kChunkBytes + ks * 2048, kChunkBytes, 1024);
              umma_f16(d, ad, pd + 2 * ks, kIdescPV, ks ? 1u : 0u);
            }
          }
          umma_commit(&o_full[st]);
          // QK(i) finished reading this stage before S(i) was published, so "PV(i) done" frees it
          umma_commit(&kv_free[st]);
        }
        __syncwarp();
      }
    }
  } else {
    // ============================================================== softmax + running output (8 warps, 256 threads)
    // Two warps share each TMEM lane quarter and split the 16 heads (TMEM columns) between them: every thread carries 8 heads,
    // and each SM sub-partition has two softmax warps to interleave (the per-tile chain is latency-bound with one).
    constexpr int HH = kHeads / 2;              // heads per thread
    const int q4 = warp & 3;                    // TMEM lane quarter of this warp
    const int half = (warp - 3) >> 2;           // 0: heads 0..7, 1: heads 8..15
    const int h0 = half * HH;
    const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
    const int row = q4 * 16 + lane;             // token row of the tile held by this thread (lanes 0..15 only, UMMA M = 64)
    float m_run[HH], l_run[HH], alpha_prev[HH], o[4 * HH];
#pragma unroll
    for (int h = 0; h < HH; ++h) { m_run[h] = -INFINITY; l_run[h] = 0.f; alpha_prev[h] = 1.f; }
#pragma unroll
    for (int j = 0; j < 4 * HH; ++j) o[j] = 0.f;
    // byte offsets of this thread's P^T elements (K-major 128B-swizzled [16 heads][64 tokens] bf16)
    uint32_t p_off[HH];
#pragma unroll
    for (int h = 0; h < HH; ++h) {
      const int hg = h0 + h;
      p_off[h] = hg * 128 + ((((row >> 3) ^ (hg & 7)) & 7) << 4) + (row & 7) * 2;
    }

    auto consume_o = [&](int i) {
      const int st = i & 1;
      mbar_wait(&o_full[st], (i >> 1) & 1);
      tc_fence_after();
      uint32_t v[4][HH];
#pragma unroll
      for (int j = 0; j < 4; ++j) tmem_ld8(tmem_base + lane_addr + kOCol + st * 64 + j * kHeads + h0, v[j]);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int h = 0; h < HH; ++h) o[j * HH + h] = o[j * HH + h] * alpha_prev[h] + __uint_as_float(v[j][h]);
      tc_fence_before();
    };

    for (int i = 0; i < n; ++i) {
      const int st = i & 1;
      mbar_wait(&s_full[st], (i >> 1) & 1);
      if (threadIdx.x == 96) MLA_TRACE(3, i);
      tc_fence_after();
      uint32_t v[HH];
      tmem_ld8(tmem_base + lane_addr + kSCol + st * kHeads + h0, v);
      tmem_ld_wait();
      tc_fence_before();
      const bool valid = lane < 16 && ((t_begin + i) * kTileTok + row) < ctx;
      float s[HH];
#pragma unroll
      for (int h = 0; h < HH; ++h) s[h] = valid ? __uint_as_float(v[h]) * p.scale_log2 : -INFINITY;
      // tile max per head over the 16 token rows of this warp: transpose-reduce butterfly — after the step with offset `off` a
      // lane keeps the half of its values selected by that bit of its lane id; lanes 2k, 2k+1 end with the warp max of head k
      float r4[4], r2[2], r1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool up = lane & 8;
        const float send = up ? s[j] : s[j + 4], keep = up ? s[j + 4] : s[j];
        r4[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 8));
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const bool up = lane & 4;
        const float send = up ? r4[j] : r4[j + 2], keep = up ? r4[j + 2] : r4[j];
        r2[j] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 4));
      }
      {
        const bool up = lane & 2;
        const float send = up ? r2[0] : r2[1], keep = up ? r2[1] : r2[0];
        r1 = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 2));
      }
      r1 = fmaxf(r1, __shfl_xor_sync(0xffffffffu, r1, 1));
      // red: [tile parity][half][quarter][8 heads]
      float* rmax = red + (i & 1) * 64 + half * 32;
      if (lane < 16 && (lane & 1) == 0) rmax[q4 * HH + (lane >> 1)] = r1;
      named_bar_sync(1, 256);
      float alpha[HH];
#pragma unroll
      for (int h4 = 0; h4 < HH; h4 += 4) {
        const float4 a = *reinterpret_cast<const float4*>(rmax + h4), b4 = *reinterpret_cast<const float4*>(rmax + HH + h4);
        const float4 c4 = *reinterpret_cast<const float4*>(rmax + 2 * HH + h4), d4 = *reinterpret_cast<const float4*>(rmax + 3 * HH + h4);
        const float tm[4] = {fmaxf(fmaxf(a.x, b4.x), fmaxf(c4.x, d4.x)), fmaxf(fmaxf(a.y, b4.y), fmaxf(c4.y, d4.y)),
                             fmaxf(fmaxf(a.z, b4.z), fmaxf(c4.z, d4.z)), fmaxf(fmaxf(a.w, b4.w), fmaxf(c4.w, d4.w))};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int h = h4 + e;
          const float mn = fmaxf(m_run[h], tm[e]);
          alpha[h] = (mn == -INFINITY) ? 1.f : exp2f(m_run[h] - mn);
          const float pv = valid ? exp2f(s[h] - mn) : 0.f;   // mn is finite whenever any row of the tile is valid
          s[h] = pv;
          // the row sum stays thread-private (all threads share m_run, so partial sums add linearly): reduced once at the end
          l_run[h] = l_run[h] * alpha[h] + pv;
          m_run[h] = mn;
        }
      }
      if (lane < 16) {
        uint8_t* pb = p_smem + st * kPBytes;
#pragma unroll
        for (int h = 0; h < HH; ++h) *reinterpret_cast<__nv_bfloat16*>(pb + p_off[h]) = __float2bfloat16_rn(s[h]);
        fence_proxy_async_smem();
      }
      named_bar_sync(1, 256);
      if (threadIdx.x == 96) { mbar_arrive(&p_ready[st]); MLA_TRACE(4, i); }
      if (i > 0) consume_o(i - 1);          // overlaps with PV(i) / QK(i+1) on the tensor pipe
      if (threadIdx.x == 96) MLA_TRACE(5, i);
#pragma unroll
      for (int h = 0; h < HH; ++h) alpha_prev[h] = alpha[h];
    }
    if (n > 0) consume_o(n - 1);

    // ---- row sums: thread-private partials -> CTA totals (once per kernel)
    named_bar_sync(1, 256);
#pragma unroll
    for (int h = 0; h < HH; ++h) {
      float t = l_run[h];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
      l_run[h] = t;
    }
    float* rsum = red + half * 32;              // [quarter][8 heads]
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < HH; ++h) rsum[q4 * HH + h] = l_run[h];
    }
    named_bar_sync(1, 256);
#pragma unroll
    for (int h = 0; h < HH; ++h) l_run[h] = (rsum[h] + rsum[HH + h]) + (rsum[2 * HH + h] + rsum[3 * HH + h]);

    // ---- write-out: thread = dim row (32*q4 + lane) of each of the 4 M-tiles, heads h0..h0+7
    const int dim0 = q4 * 32 + lane;
    if (p.nsplit == 1) {
#pragma unroll
      for (int h = 0; h < HH; ++h) {
        const float inv = l_run[h] > 0.f ? 1.0f / l_run[h] : 0.f;
        __nv_bfloat16* orow = p.out + (size_t)b * p.o_ld_t + (size_t)(h0 + h) * kLat;
#pragma unroll
        for (int j = 0; j < 4; ++j) orow[j * 128 + dim0] = __float2bfloat16_rn(o[j * HH + h] * inv);
      }
    } else {
      float* po = p.part_o + ((size_t)b * p.nsplit + split) * kHeads * kLat;
#pragma unroll
      for (int h = 0; h < HH; ++h)
#pragma unroll
        for (int j = 0; j < 4; ++j) po[(size_t)(h0 + h) * kLat + j * 128 + dim0] = o[j * HH + h];
      if (q4 == 0 && lane == 0) {               // one thread per head half
        float* ml = p.part_ml + ((size_t)b * p.nsplit + split) * kHeads * 2;
#pragma unroll
        for (int h = 0; h < HH; ++h) { ml[2 * (h0 + h)] = m_run[h]; ml[2 * (h0 + h) + 1] = l_run[h]; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// LSE combine of the KV splits: out[b, h, :] = sum_s o_s * 2^(m_s - m) / sum_s l_s * 2^(m_s - m)
__global__ void mla_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml, int nsplit,
                                   __nv_bfloat16* __restrict__ out, long long o_ld_t) {
  pdl_sync();
  const int b = blockIdx.x, h = blockIdx.y;
  float m = -INFINITY;
  for (int s = 0; s < nsplit; ++s) m = fmaxf(m, part_ml[(((size_t)b * nsplit + s) * kHeads + h) * 2]);
  float l = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float* ml = part_ml + (((size_t)b * nsplit + s) * kHeads + h) * 2;
    l += (ml[0] == -INFINITY) ? 0.f : ml[1] * exp2f(ml[0] - m);
  }
  const float inv = l > 0.f ? 1.0f / l : 0.f;
  for (int d = threadIdx.x; d < kLat; d += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float ms = part_ml[(((size_t)b * nsplit + s) * kHeads + h) * 2];
      if (ms == -INFINITY) continue;
      acc += part_o[(((size_t)b * nsplit + s) * kHeads + h) * kLat + d] * exp2f(ms - m);
    }
    out[(size_t)b * o_ld_t + (size_t)h * kLat + d] = __float2bfloat16_rn(acc * inv);
  }
}

// Prologue of the absorbed path, one CTA per token: latent = RMSNorm(c_kv) * w -> pool[slot][0:512]; rope(k_pe) -> pool[slot][512:576];
// rope(q_pe) in place for the 16 heads (q rows are [q_abs 512 | q_pe 64]).  Interleaved (traditional) rope pairs, YaRN frequencies.
__global__ void __launch_bounds__(128)
mla_absorbed_prologue_kernel(__nv_bfloat16* __restrict__ q, long long q_ld_t, const __nv_bfloat16* __restrict__ ckv, long long ckv_ld_t,
                             const __nv_bfloat16* __restrict__ kpe, long long pe_ld_t, const __nv_bfloat16* __restrict__ norm_w,
                             float eps, __nv_bfloat16* __restrict__ pool, const int* __restrict__ slots,
                             const int* __restrict__ positions, const float* __restrict__ inv_freq, float mscale) {
  pdl_sync();
  const int t = blockIdx.x, tid = threadIdx.x;
  __shared__ float red[4];
  __nv_bfloat16* dst = pool + (size_t)slots[t] * kDk;
  float f[8];
  float ss = 0.f;
  if (tid < kLat / 8) {
    const uint4 r = reinterpret_cast<const uint4*>(ckv + (size_t)t * ckv_ld_t)[tid];
    const uint32_t w4[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[2 * j] = bf16_lo(w4[j]); f[2 * j + 1] = bf16_hi(w4[j]); }
#pragma unroll
    for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  const float rinv = rsqrtf((red[0] + red[1] + red[2] + red[3]) / (float)kLat + eps);
  if (tid < kLat / 8) {
    const uint4 wr = reinterpret_cast<const uint4*>(norm_w)[tid];
    const uint32_t w4[4] = {wr.x, wr.y, wr.z, wr.w};
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // same rounding points as rmsnorm_kernel: (x * rinv) in fp32, times w, one bf16 rounding
      ow[j] = pack_bf16(f[2 * j] * rinv * bf16_lo(w4[j]), f[2 * j + 1] * rinv * bf16_hi(w4[j]));
    }
    reinterpret_cast<uint4*>(dst)[tid] = o;
  }
  const float pos = (float)positions[t];
  for (int i = tid; i < (kHeads + 1) * (kRope / 2); i += 128) {
    const int pair = i % (kRope / 2), hh = i / (kRope / 2);
    float sn, cs;
    sincosf(pos * inv_freq[pair], &sn, &cs);
    if (hh < kHeads) {
      uint32_t* w = reinterpret_cast<uint32_t*>(q + (size_t)t * q_ld_t + (size_t)hh * kDk + kLat) + pair;
      const uint32_t u = *w;
      const float a = bf16_lo(u) * mscale, bb = bf16_hi(u) * mscale;
      *w = pack_bf16(a * cs - bb * sn, a * sn + bb * cs);
    } else {
      const uint32_t u = reinterpret_cast<const uint32_t*>(kpe + (size_t)t * pe_ld_t)[pair];
      const float a = bf16_lo(u) * mscale, bb = bf16_hi(u) * mscale;
      reinterpret_cast<uint32_t*>(dst + kLat)[pair] = pack_bf16(a * cs - bb * sn, a * sn + bb * cs);
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

}  // namespace

cudaError_t mla_absorbed_prologue_launch(void* q, long long q_ld_t, const void* ckv, long long ckv_ld_t, const void* kpe,
                                         long long pe_ld_t, const void* norm_w, float eps, void* pool, const int* slots,
                                         const int* positions, const float* inv_freq, float mscale, int T, cudaStream_t s) {
  if (T == 0) return cudaSuccess;
  if ((q_ld_t % 8) || (ckv_ld_t % 8) || (pe_ld_t % 2)) return cudaErrorInvalidValue;
  (void)launch_pdl(mla_absorbed_prologue_kernel, dim3(T), dim3(128), 0, s, static_cast<__nv_bfloat16*>(q), q_ld_t,
                   static_cast<const __nv_bfloat16*>(ckv), ckv_ld_t, static_cast<const __nv_bfloat16*>(kpe), pe_ld_t,
                   static_cast<const __nv_bfloat16*>(norm_w), eps, static_cast<__nv_bfloat16*>(pool), slots, positions, inv_freq, mscale);
  return cudaGetLastError();
}

size_t mla_decode_workspace_floats(int B, int nsplit) {
  return nsplit > 1 ? (size_t)B * nsplit * kHeads * (kLat + 2) : 0;
}

cudaError_t mla_decode_launch(const void* q, long long q_ld_t, int B, const void* pool, long long num_pages, int page,
                              const int* block_tables, int max_blocks, const int* context_lens, int max_ctx, float scale,
                              int nsplit, float* workspace, void* out, long long o_ld_t, long long* dbg, cudaStream_t s) {
  if (B == 0) return cudaSuccess;
  if (page != kTileTok || (q_ld_t % 8) != 0) return cudaErrorInvalidValue;
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return cudaErrorUnknown;
  CUtensorMap tq, tkv;
  {
    // q: [B tokens][16 heads][576] bf16, head stride 576, token stride q_ld_t -> boxes of [1][16][64]
    cuuint64_t dims[3] = {(cuuint64_t)kDk, (cuuint64_t)kHeads, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)kDk * 2, (cuuint64_t)q_ld_t * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)kHeads, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (fn(&tq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(q), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)kDk, (cuuint64_t)num_pages * page};
    cuuint64_t strides[1] = {(cuuint64_t)kDk * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)kTileTok};
    cuuint32_t estr[2] = {1, 1};
    if (fn(&tkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(pool), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  MlaParams p;
  p.block_tables = block_tables; p.max_blocks = max_blocks; p.context_lens = context_lens;
  const int max_tiles = (max_ctx + kTileTok - 1) / kTileTok;
  if (nsplit < 1) nsplit = 1;
  if (nsplit > max_tiles) nsplit = max_tiles > 0 ? max_tiles : 1;
  p.nsplit = nsplit;
  p.tiles_per_split = (max_tiles + nsplit - 1) / nsplit;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = static_cast<__nv_bfloat16*>(out); p.o_ld_t = o_ld_t;
  p.dbg = dbg;
  p.part_o = workspace;
  p.part_ml = workspace != nullptr ? workspace + (size_t)B * nsplit * kHeads * kLat : nullptr;
  if (nsplit > 1 && workspace == nullptr) return cudaErrorInvalidValue;
  constexpr int smem = 2 * kTileBytes + kQBytes + 2 * kPBytes + 2 * 4 * kHeads * 4 + 12 * 8 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget exceeded");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mla_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  (void)launch_pdl(mla_decode_kernel, dim3(nsplit, B), dim3(kThreads), smem, s, tq, tkv, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || nsplit == 1) return e;
  (void)launch_pdl(mla_combine_kernel, dim3(B, kHeads), dim3(128), 0, s, (const float*)p.part_o, (const float*)p.part_ml, nsplit,
                   static_cast<__nv_bfloat16*>(out), o_ld_t);
  return cudaGetLastError();
}

}  // namespace b200
