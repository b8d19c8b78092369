// Persistent quantised-weight GEMM (MLX affine int4 / int8, in-kernel dequant) — the production path for 4-bit /
// 8-bit checkpoints (BASELINE config 2).  Structure = gemm_persistent.cu (static round-robin tile list,
// double-buffered TMEM accumulator, epilogue overlapped with the next tile) plus a *decoupled* dequant stage:
//
//   TMA ring (deep, cheap stages)     : packed codes [128 x 32|64 B] + scales/biases rows + bf16 token tile
//   dequant warps (8, two per row)    : ring stage -> bf16 A tile in a small 128B-swizzled ring (DQ buffers)
//   MMA warp (1 thread)               : A tile ring + token tile of the ring stage -> TMEM accumulator
//   epilogue warps (4)                : TMEM -> registers -> shared -> global (shared epilogue semantics)
//
// HBM sees 4.5 / 8.5 bits per weight; the deep ring hides HBM latency with ~13-17 KB stages while the expensive
// 16 KB bf16 tiles only exist DQ (2-3) at a time.
#include <algorithm>

#include "gemm_q_common.cuh"
#include "launch.h"

namespace b200 {

using namespace gemm;

namespace {

constexpr int kQPThreads = 448;  // warp 0 TMA, 1 MMA, 2..5 epilogue, 6..13 dequant

__host__ __device__ constexpr int qp_packed_bytes(int bits) { return kTileM * (kBlockK * bits / 8); }
__host__ __device__ constexpr int qp_stage_bytes(int BN, bool dual, int bits) {
  const int d = dual ? 2 : 1;
  return (BN * kBlockK * 2 + (qp_packed_bytes(bits) + 512) * d + 1023) / 1024 * 1024;
}
__host__ __device__ constexpr int qp_acc_cols(int BN, bool dual) { return BN * (dual ? 2 : 1); }
__host__ __device__ constexpr int qp_num_acc(int BN, bool dual) { return 2 * qp_acc_cols(BN, dual) <= 512 ? 2 : 1; }
__host__ __device__ constexpr uint32_t qp_tmem_cols(int BN, bool dual) {
  int c = qp_acc_cols(BN, dual) * qp_num_acc(BN, dual);
  return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512;
}
__host__ __device__ constexpr int qp_dq(int BN) { return BN <= 64 ? 3 : 2; }  // bf16 A-tile buffers
__host__ __device__ constexpr int qp_stg_rows(int BN) { return BN < 64 ? BN : 64; }
__host__ __device__ constexpr int qp_staging_bytes(int BN, int out_bytes) { return qp_stg_rows(BN) * kTileM * out_bytes; }
__host__ __device__ constexpr int qp_num_stages(int BN, bool dual, int bits, int out_bytes) {
  const int fixed = qp_dq(BN) * kATileBytes * (dual ? 2 : 1) + qp_staging_bytes(BN, out_bytes) + 1024;
  int s = (225 * 1024 - fixed) / qp_stage_bytes(BN, dual, bits);
  return s > 10 ? 10 : s;
}

struct QTile { int n0, w_row, row_base, rows_valid; };

__device__ __forceinline__ QTile q_decode_tile(const GemmParams& p, int t, int tiles_n, int tiles_m, int BN) {
  QTile ti;
  const int nt = t % tiles_n, rest = t / tiles_n;
  const int mt = rest % tiles_m, expert = rest / tiles_m;
  ti.n0 = nt * kTileM;
  ti.row_base = 0;
  ti.rows_valid = p.m;
  if (p.expert_offsets != nullptr) {
    const int lo = __ldg(p.expert_offsets + expert), hi = __ldg(p.expert_offsets + expert + 1);
    ti.row_base = lo;
    ti.rows_valid = hi - lo;
  }
  ti.rows_valid -= mt * BN;
  ti.row_base += mt * BN;
  if (ti.rows_valid > BN) ti.rows_valid = BN;
  ti.w_row = expert * p.n + ti.n0;
  return ti;
}

}  // namespace

template <int BN, bool DUAL, typename OutT, int BITS>
__global__ void __launch_bounds__(kQPThreads, 1)
gemm_q_persistent_kernel(const __grid_constant__ CUtensorMap tmap_wq, const __grid_constant__ CUtensorMap tmap_wq2,
                         const __grid_constant__ CUtensorMap tmap_s, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_s2, const __grid_constant__ CUtensorMap tmap_b2,
                         const __grid_constant__ CUtensorMap tmap_x, const GemmParams p, const int group_kblocks, const int tiles_n,
                         const int tiles_m, const int num_tiles) {
  constexpr int D = DUAL ? 2 : 1;
  constexpr int STAGES = qp_num_stages(BN, DUAL, BITS, sizeof(OutT));
  constexpr int STAGE_BYTES = qp_stage_bytes(BN, DUAL, BITS);
  constexpr int PACKED_BYTES = qp_packed_bytes(BITS);
  constexpr int OFF_PACKED = BN * kBlockK * 2;             // stage layout: [B tile | packed x D | (scales, biases) x D]
  constexpr int OFF_SCALE = OFF_PACKED + PACKED_BYTES * D;
  constexpr uint32_t TX_BYTES = BN * kBlockK * 2 + (PACKED_BYTES + 512) * D;
  constexpr int DQ = qp_dq(BN);
  constexpr int ACC_COLS = qp_acc_cols(BN, DUAL);
  constexpr int NUM_ACC = qp_num_acc(BN, DUAL);
  constexpr uint32_t TMEM_COLS = qp_tmem_cols(BN, DUAL);
  constexpr uint32_t IDESC = umma_idesc_bf16(kTileM, BN);
  constexpr int STG_ROWS = qp_stg_rows(BN);
  static_assert(STAGES >= 2, "ring too small");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024 B alignment by *pointer arithmetic* on the __shared__ array: an integer round-trip loses the address space and
  // turns every LDS/STS below into a generic LD.E/ST.E (checked with cuobjdump -sass)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_ring = smem;                                        // [DQ][D][16 KB]   (1024-aligned)
  uint8_t* ring = a_ring + DQ * D * kATileBytes;                 // [STAGES][STAGE_BYTES]
  OutT* stg = reinterpret_cast<OutT*>(ring + STAGES * STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stg) + qp_staging_bytes(BN, sizeof(OutT)));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* dq_full = empty_bar + STAGES;    // [DQ]
  uint64_t* dq_empty = dq_full + DQ;         // [DQ]
  uint64_t* tfull_bar = dq_empty + DQ;       // [2]
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool pdl_early = p.expert_offsets != nullptr;
  if (pdl_early) pdl_wait();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_wq);
    tma_prefetch_desc(&tmap_s);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int j = 0; j < DQ; ++j) {
      mbar_init(&dq_full[j], 8);   // one arrival per dequant warp
      mbar_init(&dq_empty[j], 1);  // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], kEpiThreads);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_smem, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  if (!pdl_early) pdl_wait();
  pdl_launch_dependents();

  const int kb_total = p.k / kBlockK;

  if (warp == 0) {
    // ============================================================== TMA producer
    if (elect_one()) {  // one elected lane: uniform-datapath issue, no per-instruction ELECT loop (ptx.cuh)
      uint32_t it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const QTile ti = q_decode_tile(p, t, tiles_n, tiles_m, BN);
        if (ti.rows_valid <= 0) continue;
        for (int kb = 0; kb < kb_total; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = ring + s * STAGE_BYTES;
          const int gi = kb / group_kblocks;
          mbar_arrive_expect_tx(&full_bar[s], TX_BYTES);
          tma_load_2d(st + OFF_PACKED, &tmap_wq, &full_bar[s], kb * (kBlockK * BITS / 32), ti.w_row, kEvictFirst);
          tma_load_2d(st + OFF_SCALE, &tmap_s, &full_bar[s], ti.w_row, gi, kEvictFirst);
          tma_load_2d(st + OFF_SCALE + 256, &tmap_b, &full_bar[s], ti.w_row, gi, kEvictFirst);
          if (DUAL) {
            tma_load_2d(st + OFF_PACKED + PACKED_BYTES, &tmap_wq2, &full_bar[s], kb * (kBlockK * BITS / 32), ti.w_row, kEvictFirst);
            tma_load_2d(st + OFF_SCALE + 512, &tmap_s2, &full_bar[s], ti.w_row, gi, kEvictFirst);
            tma_load_2d(st + OFF_SCALE + 768, &tmap_b2, &full_bar[s], ti.w_row, gi, kEvictFirst);
          }
          tma_load_2d(st, &tmap_x, &full_bar[s], kb * kBlockK, ti.row_base, kEvictLast);
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================== MMA issuer
    if (elect_one()) {  // one elected lane: uniform-datapath issue, no per-instruction ELECT loop (ptx.cuh)
      uint32_t it = 0, tc = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const QTile ti = q_decode_tile(p, t, tiles_n, tiles_m, BN);
        if (ti.rows_valid <= 0) continue;
        const uint32_t ab = tc % NUM_ACC, aph = (tc / NUM_ACC) & 1;
        mbar_wait(&tempty_bar[ab], aph ^ 1);
        tc_fence_after();
        const uint32_t acc_addr = tmem_base + ab * ACC_COLS;
        for (int kb = 0; kb < kb_total; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
          const uint32_t j = it % DQ, jph = (it / DQ) & 1;
          mbar_wait(&full_bar[s], ph);   // token tile landed
          mbar_wait(&dq_full[j], jph);   // weight tile dequantised + fenced
          tc_fence_after();
          const uint32_t a_addr = smem_u32(a_ring + j * D * kATileBytes);
          const uint32_t b_addr = smem_u32(ring + s * STAGE_BYTES);
          const uint64_t adesc = umma_desc_sw128(a_addr);
          const uint64_t bdesc = umma_desc_sw128(b_addr);
#pragma unroll
          for (int kk = 0; kk < kBlockK / kUmmaK; ++kk) {
            const uint32_t acc = (kb > 0 || kk > 0) ? 1u : 0u;
            umma_f16(acc_addr, adesc + 2 * kk, bdesc + 2 * kk, IDESC, acc);
            if (DUAL) {
              const uint64_t a2desc = umma_desc_sw128(a_addr + kATileBytes);
              umma_f16(acc_addr + BN, a2desc + 2 * kk, bdesc + 2 * kk, IDESC, acc);
            }
          }
          umma_commit(&empty_bar[s]);  // ring stage (token tile; packed codes were consumed before dq_full)
          umma_commit(&dq_empty[j]);   // bf16 A tile buffer
        }
        umma_commit(&tfull_bar[ab]);
        ++tc;
      }
    }
  } else if (warp < 6) {
    // ============================================================== epilogue warps (128 threads)
    const int q = warp & 3;
    const int f_local = q * 32 + lane;
    const int et = threadIdx.x - 64;
    constexpr int kVec = 8, kChunks = kTileM / kVec, kRowsPerIter = kEpiThreads / kChunks;
    uint32_t tc = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const QTile ti = q_decode_tile(p, t, tiles_n, tiles_m, BN);
      if (ti.rows_valid <= 0) continue;
      const uint32_t ab = tc % NUM_ACC, aph = (tc / NUM_ACC) & 1;
      mbar_wait(&tfull_bar[ab], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ab * ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
      const int f_glob = ti.n0 + f_local;
      const float bias = (p.bias != nullptr && f_glob < p.n) ? __bfloat162float(p.bias[f_glob]) : 0.0f;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += STG_ROWS) {
        if (c0 >= ti.rows_valid) break;
#pragma unroll 1
        for (int c = c0; c < c0 + STG_ROWS; c += 16) {
          if (c >= ti.rows_valid) break;
          uint32_t v[16];
          float g[16], u[16];
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
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float y = g[j] + bias;
            if (DUAL) y = apply_act(p.act, y) * u[j];
            if (p.softcap > 0.f) y = p.softcap * tanhf(y / p.softcap);
            stage_store<OutT>(&stg[(c - c0 + j) * kTileM + f_local], y);
          }
        }
        if (c0 + STG_ROWS >= ti.rows_valid || c0 + STG_ROWS >= BN) {
          tc_fence_before();
          mbar_arrive(&tempty_bar[ab]);
        }
        named_bar_sync(1, kEpiThreads);
        const int ch = et % kChunks;
        const int f0 = ti.n0 + ch * kVec;
        const int rows_here = min(STG_ROWS, ti.rows_valid - c0);
        if (f0 < p.n) {
          for (int r = et / kChunks; r < rows_here; r += kRowsPerIter) {
            const size_t row = static_cast<size_t>(ti.row_base + c0 + r);
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
        named_bar_sync(1, kEpiThreads);
      }
      if (p.signal_flag != nullptr) {
        __threadfence_system();
        named_bar_sync(1, kEpiThreads);
        if (et == 0) {
          const unsigned int done = atomicAdd(p.done_counter, 1u) + 1u;
          if (done == p.signal_tiles) {
            *p.done_counter = 0u;
            __threadfence_system();
            if (p.signal_value == 0u) atomicAdd_system(p.signal_flag, 1u);
            else st_release_sys(p.signal_flag, p.signal_value);
          }
        }
      }
      ++tc;
    }
    tc_fence_before();
  } else {
    // ============================================================== dequant producers (8 warps, two threads per row)
    const int r = (threadIdx.x - 192) & 127;
    const int half = (threadIdx.x - 192) >> 7;
    uint32_t it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const QTile ti = q_decode_tile(p, t, tiles_n, tiles_m, BN);
      if (ti.rows_valid <= 0) continue;
      for (int kb = 0; kb < kb_total; ++kb, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
        const uint32_t j = it % DQ, jph = (it / DQ) & 1;
        mbar_wait(&full_bar[s], ph);        // packed codes + scales landed
        mbar_wait(&dq_empty[j], jph ^ 1);   // the MMAs that read this A buffer have completed
        const uint8_t* st = ring + s * STAGE_BYTES;
        uint8_t* a_tile = a_ring + j * D * kATileBytes;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const __nv_bfloat16* sb = reinterpret_cast<const __nv_bfloat16*>(st + OFF_SCALE + d * 512);
          const float sc = __bfloat162float(sb[r]), bi = __bfloat162float(sb[128 + r]);
          dequant_half_row<BITS>(st + OFF_PACKED + d * PACKED_BYTES + r * (kBlockK * BITS / 8), sc, bi, a_tile + d * kATileBytes, r, half);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&dq_full[j]);
      }
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ================================================================================================ host side
namespace {

int qp_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, bool DUAL, typename OutT, int BITS>
cudaError_t qp_launch_one(const QMaps& t, const GemmParams& p, int gk, int tiles_n, int tiles_m, int num_tiles, cudaStream_t stream) {
  constexpr int STAGES = qp_num_stages(BN, DUAL, BITS, sizeof(OutT));
  constexpr int smem = qp_dq(BN) * kATileBytes * (DUAL ? 2 : 1) + STAGES * qp_stage_bytes(BN, DUAL, BITS) +
                       qp_staging_bytes(BN, sizeof(OutT)) + (2 * STAGES + 2 * qp_dq(BN) + 4) * 8 + 16 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget exceeded");
  auto kern = gemm_q_persistent_kernel<BN, DUAL, OutT, BITS>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int grid = std::min(num_tiles, qp_sm_count());
  (void)launch_pdl(kern, dim3(grid), dim3(kQPThreads), smem, stream, t.wq, t.wq2, t.s, t.b, t.s2, t.b2, t.x, p, gk, tiles_n, tiles_m,
                   num_tiles);
  return cudaGetLastError();
}

template <bool DUAL, typename OutT, int BITS>
cudaError_t qp_dispatch_bn(int bn, const QMaps& t, const GemmParams& p, int gk, int tn, int tm, int nt, cudaStream_t s) {
  switch (bn) {
    case 16: return qp_launch_one<16, DUAL, OutT, BITS>(t, p, gk, tn, tm, nt, s);
    case 32: return qp_launch_one<32, DUAL, OutT, BITS>(t, p, gk, tn, tm, nt, s);
    case 64: return qp_launch_one<64, DUAL, OutT, BITS>(t, p, gk, tn, tm, nt, s);
    case 128: return qp_launch_one<128, DUAL, OutT, BITS>(t, p, gk, tn, tm, nt, s);
    case 256: return qp_launch_one<256, DUAL, OutT, BITS>(t, p, gk, tn, tm, nt, s);
    default: return cudaErrorInvalidValue;
  }
}

template <int BITS>
cudaError_t qp_dispatch(bool dual, bool fp32, int bn, const QMaps& t, const GemmParams& p, int gk, int tn, int tm, int nt, cudaStream_t s) {
  if (fp32) return qp_dispatch_bn<false, float, BITS>(bn, t, p, gk, tn, tm, nt, s);
  if (dual) return qp_dispatch_bn<true, __nv_bfloat16, BITS>(bn, t, p, gk, tn, tm, nt, s);
  return qp_dispatch_bn<false, __nv_bfloat16, BITS>(bn, t, p, gk, tn, tm, nt, s);
}

}  // namespace

cudaError_t gemm_q_persistent_launch(const GemmArgs& a, cudaStream_t stream) {
  const bool dual = a.w2 != nullptr;
  const bool grouped = a.expert_offsets != nullptr;
  const int bn = a.bn > 0 ? a.bn : gemm_pick_bn(a.max_rows);
  if (grouped && (a.n % kTileM) != 0) return cudaErrorInvalidValue;
  if ((a.n % 8) != 0 || (a.k % kBlockK) != 0) return cudaErrorInvalidValue;
  if (dual && a.out_fp32) return cudaErrorInvalidValue;

  const uint64_t w_rows = static_cast<uint64_t>(a.n) * (grouped ? a.num_experts : 1);
  const uint64_t words = static_cast<uint64_t>(a.k) * a.q_bits / 32;
  const uint64_t ngroups = a.k / a.q_group;
  QMaps t;
  if (!gemm_q_make_tmap(&t.wq, 1, a.w, words, w_rows, words, kBlockK * a.q_bits / 32, kTileM)) return cudaErrorUnknown;
  if (!gemm_q_make_tmap(&t.s, 2, a.q_scales_t, w_rows, ngroups, w_rows, kTileM, 1)) return cudaErrorUnknown;
  if (!gemm_q_make_tmap(&t.b, 2, a.q_biases_t, w_rows, ngroups, w_rows, kTileM, 1)) return cudaErrorUnknown;
  if (dual) {
    if (!gemm_q_make_tmap(&t.wq2, 1, a.w2, words, w_rows, words, kBlockK * a.q_bits / 32, kTileM)) return cudaErrorUnknown;
    if (!gemm_q_make_tmap(&t.s2, 2, a.q_scales2_t, w_rows, ngroups, w_rows, kTileM, 1)) return cudaErrorUnknown;
    if (!gemm_q_make_tmap(&t.b2, 2, a.q_biases2_t, w_rows, ngroups, w_rows, kTileM, 1)) return cudaErrorUnknown;
  } else {
    t.wq2 = t.wq; t.s2 = t.s; t.b2 = t.b;
  }
  if (!gemm_q_make_tmap(&t.x, 0, a.x, a.k, a.x_rows, a.ld_x, kBlockK, bn)) return cudaErrorUnknown;

  GemmParams p;
  p.row_dst = nullptr; p.signal_peers = nullptr; p.num_signal_peers = 0; p.w_sf = p.w2_sf = p.x_sf = nullptr; p.sf_ld_w = p.sf_ld_x = 0; p.ep_arrive = nullptr; p.ep_seq = nullptr; p.ep_error = nullptr; p.ep_world = 0; p.ep_zero_other = 0; p.expert_stride = 0;
  p.m = a.m; p.n = a.n; p.k = a.k; p.splits = 1; p.cluster_splitk = 0;
  p.expert_offsets = a.expert_offsets;
  p.out = a.out; p.ld_out = a.ld_out;
  p.residual = static_cast<const __nv_bfloat16*>(a.residual); p.ld_res = a.ld_res;
  p.bias = static_cast<const __nv_bfloat16*>(a.bias); p.act = a.act; p.softcap = a.softcap;
  p.workspace = nullptr; p.tile_counters = nullptr;
  p.signal_flag = a.signal_flag; p.signal_value = a.signal_value; p.done_counter = a.done_counter;
  const int tiles_n = (a.n + kTileM - 1) / kTileM;
  const int tiles_m = (a.max_rows + bn - 1) / bn;
  const int num_tiles = tiles_n * tiles_m * (grouped ? a.num_experts : 1);
  p.signal_tiles = a.signal_tiles > 0 ? a.signal_tiles : static_cast<unsigned int>(tiles_n * tiles_m);
  const int gk = a.q_group / kBlockK;
  if (a.q_bits == 4) return qp_dispatch<4>(dual, a.out_fp32, bn, t, p, gk, tiles_n, tiles_m, num_tiles, stream);
  return qp_dispatch<8>(dual, a.out_fp32, bn, t, p, gk, tiles_n, tiles_m, num_tiles, stream);
}

}  // namespace b200
