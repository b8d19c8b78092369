// On-device sampler: logit_bias + repetition penalty, log-softmax, argmax /
// temperature / nucleus (top-p) sampling and top-k logprobs — one CTA per sequence, everything stays on
// the last pipeline stage so only token ids (+ <=10 logprobs) travel back to stage 0.
// Reference behaviour: shard/utils.py:126-139,166-177 and mlx_lm `top_p_sampling` (sort ascending,
// keep tokens whose cumulative mass exceeds 1 - top_p); here the nucleus is found without a sort by
// bisection on the logit threshold, and sampling inside it uses the Gumbel-max trick with a
// counter-based RNG (stateless: seed, step, row, token id).
#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"

namespace b200 {

namespace {

constexpr int kSampThreads = 1024;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float uniform01(uint64_t seed, uint64_t step, uint32_t row, uint32_t i) {
  const uint64_t h = mix64(mix64(seed ^ (step * 0xD1342543DE82EF95ull)) ^ ((uint64_t)row << 32 | i));
  return ((float)(h >> 40) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
}

struct ArgMax { float v; int i; };
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
  return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ ArgMax block_argmax(ArgMax x, ArgMax* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax y{__shfl_xor_sync(0xffffffffu, x.v, o), __shfl_xor_sync(0xffffffffu, x.i, o)};
    x = better(x, y);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sm[warp] = x;
  __syncthreads();
  ArgMax r = sm[0];
  for (int w = 1; w < kSampThreads / 32; ++w) r = better(r, sm[w]);
  return r;
}
__device__ float block_sum(float x, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sm[warp] = x;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < kSampThreads / 32; ++w) r += sm[w];
  return r;
}

__global__ void __launch_bounds__(kSampThreads)
sample_kernel(const float* __restrict__ logits, int V, const float* __restrict__ temperature, const float* __restrict__ top_p,
              unsigned long long seed, unsigned long long step, const unsigned long long* __restrict__ row_rng,
              long long* __restrict__ tokens, float* __restrict__ logprobs, int top_k, long long* __restrict__ top_ids,
              float* __restrict__ top_lp, const uint4* __restrict__ tag_src, uint4* __restrict__ tag_dst) {
  pdl_sync();  // PDL: predecessor's writes visible; let the successor start its prologue
  // Random stream of a row: (seed, step) launch scalars, or — for CUDA-graph replay, where scalars are frozen at capture — a
  // per-row (seed, step) pair in device memory (`row_rng[b]` = request seed, number of tokens that request has sampled so far:
  // a request's stream is then independent of which batch / row it decodes in).
  uint32_t rng_row = blockIdx.x;
  if (row_rng != nullptr) {
    seed = row_rng[2 * blockIdx.x];
    step = row_rng[2 * blockIdx.x + 1];
    rng_row = 0;
  }
  // 16-byte tag (step id of the scheduler) copied next to the results so the receiver can match result <-> step
  if (tag_dst != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *tag_dst = *tag_src;
  __shared__ ArgMax sm_a[kSampThreads / 32];
  __shared__ float sm_f[kSampThreads / 32];
  const int b = blockIdx.x;
  const float* lg = logits + (size_t)b * V;
  // pass 1: max / argmax
  ArgMax am{-INFINITY, 0x7fffffff};
  for (int i = threadIdx.x; i < V; i += kSampThreads) am = better(am, ArgMax{lg[i], i});
  am = block_argmax(am, sm_a);
  // pass 2: log-sum-exp (temperature is NOT applied to the reported logprobs, reference utils.py:131)
  float se = 0.f;
  for (int i = threadIdx.x; i < V; i += kSampThreads) se += __expf(lg[i] - am.v);
  se = block_sum(se, sm_f);
  const float lse = am.v + logf(se);

  const float temp = temperature[b];
  const float tp = top_p[b];
  int token = am.i;
  if (temp > 0.f) {
    const float inv_t = 1.0f / temp;
    float thresh = -INFINITY;  // keep logits > thresh
    if (tp > 0.f && tp < 1.f) {
      // Z at temperature
      float zs = 0.f;
      for (int i = threadIdx.x; i < V; i += kSampThreads) zs += __expf((lg[i] - am.v) * inv_t);
      zs = block_sum(zs, sm_f);
      // bisection on x in [lo, hi]: F(x) = mass{logit <= x}; find the smallest kept logit, i.e. the
      // infimum of x with F(x) > 1 - top_p
      float lo = am.v - 80.f * temp, hi = am.v;
      const float target = (1.0f - tp) * zs;
      for (int it = 0; it < 26; ++it) {
        const float mid = 0.5f * (lo + hi);
        float ms = 0.f;
        for (int i = threadIdx.x; i < V; i += kSampThreads) {
          const float x = lg[i];
          if (x <= mid) ms += __expf((x - am.v) * inv_t);
        }
        ms = block_sum(ms, sm_f);
        if (ms > target) hi = mid; else lo = mid;
      }
      thresh = lo;  // F(lo) <= target < F(hi): everything strictly above lo is in the nucleus
    }
    // Gumbel-max over the kept set: argmax (logit / T + g), g = -log(-log(u))
    ArgMax gm{-INFINITY, 0x7fffffff};
    for (int i = threadIdx.x; i < V; i += kSampThreads) {
      const float x = lg[i];
      if (x > thresh) {
        const float u = uniform01(seed, step, rng_row, i);
        const float g = -__logf(-__logf(u));
        gm = better(gm, ArgMax{(x - am.v) * inv_t + g, i});
      }
    }
    gm = block_argmax(gm, sm_a);
    token = gm.i;
  }
  if (threadIdx.x == 0) {
    tokens[b] = token;
    logprobs[b] = lg[token] - lse;
  }
  // top-k logprobs (k <= 10): k selection passes in (value desc, index asc) order
  ArgMax prev{INFINITY, -1};
  for (int r = 0; r < top_k; ++r) {
    ArgMax cur{-INFINITY, 0x7fffffff};
    for (int i = threadIdx.x; i < V; i += kSampThreads) {
      const float x = lg[i];
      if (x < prev.v || (x == prev.v && i > prev.i)) cur = better(cur, ArgMax{x, i});
    }
    cur = block_argmax(cur, sm_a);
    if (threadIdx.x == 0) {
      top_ids[(size_t)b * top_k + r] = cur.i;
      top_lp[(size_t)b * top_k + r] = cur.v - lse;
    }
    prev = cur;
  }
}

// repetition penalty over the (deduplicated) context, then logit_bias (reference order: utils.py:167-170 then :127-130)
__global__ void apply_penalties_kernel(float* __restrict__ logits, int V, const int* __restrict__ rep_ctx, int C,
                                       const float* __restrict__ penalty, const int* __restrict__ bias_idx,
                                       const float* __restrict__ bias_val, int NB) {
  pdl_sync();  // PDL: predecessor's writes visible; let the successor start its prologue
  const int b = blockIdx.x;
  float* lg = logits + (size_t)b * V;
  const float pen = penalty[b];
  if (pen != 1.0f) {
    for (int j = threadIdx.x; j < C; j += blockDim.x) {
      const int id = rep_ctx[(size_t)b * C + j];
      if (id < 0 || id >= V) continue;
      bool dup = false;
      for (int i = 0; i < j; ++i) dup |= (rep_ctx[(size_t)b * C + i] == id);
      if (dup) continue;
      const float x = lg[id];
      lg[id] = x < 0.f ? x * pen : x / pen;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < NB; j += blockDim.x) {
    const int id = bias_idx[(size_t)b * NB + j];
    if (id >= 0 && id < V) atomicAdd(&lg[id], bias_val[(size_t)b * NB + j]);
  }
}

}  // namespace

cudaError_t apply_penalties_launch(float* logits, int B, int V, const int* rep_ctx, int C, const float* penalty,
                                   const int* bias_idx, const float* bias_val, int NB, cudaStream_t s) {
  if (B == 0) return cudaSuccess;
  (void)launch_pdl(apply_penalties_kernel, dim3(B), dim3(128), 0, s, logits, V, rep_ctx, C, penalty, bias_idx, bias_val, NB);
  return cudaGetLastError();
}

cudaError_t sample_launch(const float* logits, int B, int V, const float* temperature, const float* top_p,
                          unsigned long long seed, unsigned long long step, const unsigned long long* row_rng, long long* tokens,
                          float* logprobs, int top_k, long long* top_ids, float* top_lp, const void* tag_src, void* tag_dst,
                          cudaStream_t s) {
  if (B == 0) return cudaSuccess;
  if (top_k > 32) return cudaErrorInvalidValue;
  (void)launch_pdl(sample_kernel, dim3(B), dim3(kSampThreads), 0, s, logits, V, temperature, top_p, seed, step, row_rng, tokens,
                   logprobs, top_k, top_ids, top_lp, static_cast<const uint4*>(tag_src), static_cast<uint4*>(tag_dst));
  return cudaGetLastError();
}

}  // namespace b200
