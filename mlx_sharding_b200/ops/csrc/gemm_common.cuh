// Shared pieces of the swap-AB tcgen05 GEMM family: tile constants and the TMEM -> registers -> shared ->
// global epilogue (bias / activation-gating / soft-cap / residual / split-K reduction / fused P2P signal).
// Used by the dense (TMA-fed bf16 weights, gemm_tcgen05.cu) and the quantised (in-kernel MLX-affine
// dequant, gemm_q_tcgen05.cu) kernels.
#pragma once
#include "gemm_tcgen05.h"
#include "ptx.cuh"

namespace b200 {
namespace gemm {

constexpr int kTileM = 128;      // output features per CTA (UMMA M)
constexpr int kBlockK = 64;      // bf16 elements per k-block = one 128B swizzle row
constexpr int kUmmaK = 16;
constexpr int kNumThreads = 192;  // warp0: TMA, warp1: MMA + TMEM alloc, warps 2..5: epilogue
constexpr int kEpiThreads = 128;
constexpr int kATileBytes = kTileM * kBlockK * 2;  // 16 KB

__host__ __device__ constexpr int stage_bytes(int BN, bool dual) {
  return kATileBytes * (dual ? 2 : 1) + BN * kBlockK * 2;
}
__host__ __device__ constexpr int num_stages(int BN, bool dual, int out_bytes) {
  // leave room for barriers; the epilogue staging buffer aliases the (by then idle) stage ring
  int s = (200 * 1024) / stage_bytes(BN, dual);
  s = s > 8 ? 8 : s;
  // the ring must be at least as large as the epilogue staging tile
  while (s * stage_bytes(BN, dual) < BN * kTileM * out_bytes) ++s;
  return s;
}
__host__ __device__ constexpr uint32_t tmem_cols(int BN, bool dual) {
  int c = BN * (dual ? 2 : 1);
  return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512;
}

__device__ __forceinline__ float apply_act(int act, float g) {
  if (act == kActSilu) return g / (1.0f + __expf(-g));
  if (act == kActGeluTanh) {
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    float u = k0 * (g + k1 * g * g * g);
    float t = 1.0f - 2.0f / (1.0f + __expf(2.0f * u));  // tanh(u)
    return 0.5f * g * (1.0f + t);
  }
  return g;
}

template <typename OutT>
__device__ __forceinline__ void stage_store(OutT* p, float v);
template <>
__device__ __forceinline__ void stage_store<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ void stage_store<float>(float* p, float v) { *p = v; }


// Runs on the 4 epilogue warps (threads [epi_base, epi_base + 128)); `warp` is the CTA-wide warp index.
template <int BN, bool DUAL, typename OutT>
__device__ __forceinline__ void run_epilogue(const GemmParams& p, uint8_t* smem, uint32_t tmem_base, uint64_t* tmem_full_bar,
                                             uint32_t* flag_smem, int warp, int lane, int epi_base, int n0, int mt, int split,
                                             int row_base, int rows_valid, int num_kb) {
    // ============================================================== epilogue warps (128 threads)
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int f_local = q * 32 + lane;  // feature within the tile == TMEM lane
    const int f_glob = n0 + f_local;
    const int et = threadIdx.x - epi_base;  // 0..127
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    bool do_epilogue = true;
    if (p.splits > 1) {
      // ---- write the raw fp32 partial, take a ticket, the last arriver reduces
      float* ws = p.workspace + (static_cast<size_t>((blockIdx.y * gridDim.x + blockIdx.x) * (gridDim.z / p.splits) + mt) *
                                 p.splits + split) * (BN * (DUAL ? 2 : 1) * kTileM);
#pragma unroll 1
      for (int c = 0; c < BN * (DUAL ? 2 : 1); c += 16) {
        uint32_t v[16];
        if (num_kb > 0) {
          tmem_ld16(taddr + c, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) ws[(c + j) * kTileM + f_local] = __uint_as_float(v[j]);
      }
      __threadfence();
      named_bar_sync(1, kEpiThreads);
      if (et == 0) {
        unsigned int* ctr = p.tile_counters + ((blockIdx.y * gridDim.x + blockIdx.x) * (gridDim.z / p.splits) + mt);
        const unsigned int old = atomicAdd(ctr, 1u);
        const bool last = (old == static_cast<unsigned int>(p.splits - 1));
        if (last) *ctr = 0u;  // self-reset for the next launch
        *flag_smem = last ? 1u : 0u;
      }
      named_bar_sync(1, kEpiThreads);
      do_epilogue = (*flag_smem != 0u);
      if (do_epilogue) __threadfence();
    }

    if (do_epilogue) {
      OutT* stg = reinterpret_cast<OutT*>(smem);  // aliases the idle stage ring: [BN][128]
      const float bias = (p.bias != nullptr && f_glob < p.n) ? __bfloat162float(p.bias[f_glob]) : 0.0f;
      const float* ws0 = nullptr;
      if (p.splits > 1)
        ws0 = p.workspace + static_cast<size_t>((blockIdx.y * gridDim.x + blockIdx.x) * (gridDim.z / p.splits) + mt) *
                                p.splits * (BN * (DUAL ? 2 : 1) * kTileM);
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        if (c >= rows_valid) break;
        float g[16], u[16];
        if (p.splits > 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) { g[j] = 0.f; u[j] = 0.f; }
          // deterministic split order, but 2 splits x 16 columns of independent L2 loads are in flight
          // per thread before they are consumed (the reduction is latency-, not bandwidth-bound)
          for (int s = 0; s < p.splits; s += 2) {
            const float* w0 = ws0 + static_cast<size_t>(s) * (BN * (DUAL ? 2 : 1) * kTileM);
            const bool two = (s + 1) < p.splits;
            const float* w1 = two ? w0 + (BN * (DUAL ? 2 : 1) * kTileM) : w0;
            float a0[16], a1[16], b0[16], b1[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              a0[j] = __ldcg(&w0[(c + j) * kTileM + f_local]);
              a1[j] = two ? __ldcg(&w1[(c + j) * kTileM + f_local]) : 0.f;
              if (DUAL) {
                b0[j] = __ldcg(&w0[(BN + c + j) * kTileM + f_local]);
                b1[j] = two ? __ldcg(&w1[(BN + c + j) * kTileM + f_local]) : 0.f;
              }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              g[j] = (g[j] + a0[j]) + a1[j];
              if (DUAL) u[j] = (u[j] + b0[j]) + b1[j];
            }
          }
        } else {
          uint32_t v[16];
          tmem_ld16(taddr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) g[j] = __uint_as_float(v[j]);
          if (DUAL) {
            tmem_ld16(taddr + BN + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) u[j] = __uint_as_float(v[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float y = g[j] + bias;
          if (DUAL) y = apply_act(p.act, y) * u[j];
          if (p.softcap > 0.f) y = p.softcap * tanhf(y / p.softcap);
          stage_store<OutT>(&stg[(c + j) * kTileM + f_local], y);
        }
      }
      tc_fence_before();
      named_bar_sync(1, kEpiThreads);
      // ---- coalesced write-out: each token row of the tile is 128 features = 16 chunks of 8 elements
      constexpr int kVec = 8;
      constexpr int kChunks = kTileM / kVec;                  // 16
      constexpr int kRowsPerIter = kEpiThreads / kChunks;     // 8
      const int ch = et % kChunks;
      const int f0 = n0 + ch * kVec;
      if (f0 < p.n) {
        for (int r = et / kChunks; r < rows_valid; r += kRowsPerIter) {
          const size_t row = static_cast<size_t>(row_base + r);
          float vals[kVec];
          if (sizeof(OutT) == 2) {
            const uint4 raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(stg) + r * kTileM + ch * kVec);
            const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { vals[2 * j] = bf16_lo(w4[j]); vals[2 * j + 1] = bf16_hi(w4[j]); }
          } else {
            const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(stg) + r * kTileM + ch * kVec);
            const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(stg) + r * kTileM + ch * kVec + 4);
            vals[0] = a.x; vals[1] = a.y; vals[2] = a.z; vals[3] = a.w;
            vals[4] = b.x; vals[5] = b.y; vals[6] = b.z; vals[7] = b.w;
          }
          if (p.residual != nullptr) {
            const uint4 rr = *reinterpret_cast<const uint4*>(p.residual + row * p.ld_res + f0);
            const uint32_t w4[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { vals[2 * j] += bf16_lo(w4[j]); vals[2 * j + 1] += bf16_hi(w4[j]); }
          }
          if (sizeof(OutT) == 2) {
            uint4 o;
            o.x = pack_bf16(vals[0], vals[1]); o.y = pack_bf16(vals[2], vals[3]);
            o.z = pack_bf16(vals[4], vals[5]); o.w = pack_bf16(vals[6], vals[7]);
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ld_out + f0) = o;
          } else {
            float* o = reinterpret_cast<float*>(p.out) + row * p.ld_out + f0;
            *reinterpret_cast<float4*>(o) = make_float4(vals[0], vals[1], vals[2], vals[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(vals[4], vals[5], vals[6], vals[7]);
          }
        }
      }
    }
    // ---- fused stage boundary: publish "tile stored" and let the last CTA raise the peer's flag
    if (p.signal_flag != nullptr) {
      __threadfence_system();
      named_bar_sync(1, kEpiThreads);
      if (et == 0 && do_epilogue) {
        const unsigned int done = atomicAdd(p.done_counter, 1u) + 1u;
        if (done == p.signal_tiles) {
          *p.done_counter = 0u;
          __threadfence_system();
          // signal_value == 0: counting flag (graph-replay safe: +1 per completed hand-off)
          if (p.signal_value == 0u) atomicAdd_system(p.signal_flag, 1u);
          else st_release_sys(p.signal_flag, p.signal_value);
        }
      }
    }
    tc_fence_before();
}

}  // namespace gemm
}  // namespace b200
