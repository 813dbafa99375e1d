// Hard-negative sampling (reference SPMM_models.py:154-178) and momentum-queue enqueue (:271-286).
#include <cstdlib>

#include "common.cuh"
#include "spmm_b200.h"

namespace spmm {

// ---------------------------------------------------------------- Philox4x32-10 (counter-based; CPU replica in
// oracle/sampler_ref.py).  key = seed, counter = (row, stream, step_lo, step_hi).
__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t (&c)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// exp(x) for x <= 0 from separately rounded fp32 mul/add only, so numpy reproduces it bit for bit
__device__ __forceinline__ float exact_exp_neg(float x) {
  if (x < -87.f) return 0.f;
  const float t = __fmul_rn(x, 1.44269504f);
  const float n = floorf(__fadd_rn(t, 0.5f));
  float r = __fsub_rn(x, __fmul_rn(n, 0.693359375f));
  r = __fsub_rn(r, __fmul_rn(n, -2.12194440e-4f));
  float p = __fadd_rn(__fmul_rn(r, 1.3888889e-3f), 8.3333333e-3f);
  p = __fadd_rn(__fmul_rn(p, r), 4.1666667e-2f);
  p = __fadd_rn(__fmul_rn(p, r), 1.6666667e-1f);
  p = __fadd_rn(__fmul_rn(p, r), 0.5f);
  p = __fadd_rn(__fmul_rn(p, r), 1.f);
  p = __fadd_rn(__fmul_rn(p, r), 1.f);
  const int e = (int)n + 127;           // x >= -87 -> n >= -126 -> e >= 1
  return __fmul_rn(p, __int_as_float(e << 23));
}

// one warp per (stream, row): stream 0 = t2i (negative property for each text), 1 = i2t (negative text per property)
__global__ void sample_neg_kernel(const float* __restrict__ sim_i2t, const float* __restrict__ sim_t2i, int B,
                                  uint32_t seed_lo, uint32_t seed_hi, uint32_t step_lo, uint32_t step_hi,
                                  int* __restrict__ neg_t2i, int* __restrict__ neg_i2t,
                                  const unsigned long long* __restrict__ salt) {
  extern __shared__ float shw[];  // [warps][B]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int job = blockIdx.x * (blockDim.x >> 5) + warp;
  if (job >= 2 * B) return;
  const int stream = job / B, b = job % B;
  const float* row = (stream == 0 ? sim_t2i : sim_i2t) + (size_t)b * B;
  float* w = shw + warp * B;
  float mx = -INFINITY;
  for (int j = lane; j < B; j += 32) mx = fmaxf(mx, row[j]);
  mx = warp_max(mx);
  for (int j = lane; j < B; j += 32) w[j] = (j == b) ? 0.f : exact_exp_neg(__fsub_rn(row[j], mx));  // fill_diagonal_(0)
  __syncwarp();
  if (lane == 0) {
    // the device salt (bumped once per training step) is folded into the Philox counter's step words
    const unsigned long long step = (((unsigned long long)step_hi << 32) | step_lo) + (salt ? __ldg(salt) : 0ull);
    uint32_t c[4] = {(uint32_t)b, (uint32_t)stream, (uint32_t)step, (uint32_t)(step >> 32)};
    philox4x32_10(seed_lo, seed_hi, c);
    const float u = __fmul_rn((float)(c[0] >> 8), 5.9604644775390625e-8f);  // [0,1), 24 bits
    float total = 0.f;
    for (int j = 0; j < B; ++j) total = __fadd_rn(total, w[j]);
    const float target = __fmul_rn(u, total);
    float cum = 0.f;
    int idx = -1, last = (b == 0 && B > 1) ? 1 : 0;
    for (int j = 0; j < B; ++j) {
      if (w[j] > 0.f) last = j;
      cum = __fadd_rn(cum, w[j]);
      if (idx < 0 && cum > target && w[j] > 0.f) idx = j;
    }
    if (idx < 0) idx = last;
    (stream == 0 ? neg_t2i : neg_i2t)[b] = idx;
  }
}

__global__ void enqueue_kernel(float* __restrict__ pq, float* __restrict__ tq, const float* __restrict__ pf,
                               const float* __restrict__ tf, const int64_t* __restrict__ ptr, int n, int E, int Q,
                               const float* __restrict__ skip) {
  if (skip != nullptr && *skip != 0.f) return;
  const int64_t p = *ptr;
  const int64_t total = (int64_t)n * E;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / E, c = i % E;
    const int64_t dst = ((p + r) % Q) * E + c;
    pq[dst] = pf[i];
    tq[dst] = tf[i];
  }
}
__global__ void enqueue_ptr_kernel(int64_t* ptr, int n, int Q, const float* skip) {
  if (skip != nullptr && *skip != 0.f) return;
  *ptr = (*ptr + n) % Q;
}

}  // namespace spmm
namespace spmm {
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("SPMM_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
}  // namespace spmm
using namespace spmm;

extern "C" int spmm_sample_negatives(const float* sim_i2t, const float* sim_t2i, int B, unsigned long long seed,
                                     unsigned long long step, int* neg_t2i, int* neg_i2t, void* stream) {
  SPMM_ARG(sim_i2t && sim_t2i && neg_t2i && neg_i2t && B >= 2 && B <= 2048);
  const int warps = 4;
  sample_neg_kernel<<<(2 * B + warps - 1) / warps, warps * 32, (size_t)warps * B * sizeof(float), (cudaStream_t)stream>>>(
      sim_i2t, sim_t2i, B, (uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)step, (uint32_t)(step >> 32), neg_t2i,
      neg_i2t, spmm_g_rng_salt);
  SPMM_CHECK_LAUNCH();
  return 0;
}

extern "C" int spmm_enqueue(float* prop_queue, float* text_queue, const float* prop_feats, const float* text_feats,
                            int64_t* queue_ptr, int n, int E, int Q, const float* skip_flag, void* stream) {
  SPMM_ARG(prop_queue && text_queue && prop_feats && text_feats && queue_ptr && n > 0 && E > 0 && Q > 0 && Q % n == 0);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = (int)(((int64_t)n * E + 255) / 256);
  if (grid > 4 * kNumSMs) grid = 4 * kNumSMs;
  enqueue_kernel<<<grid, 256, 0, st>>>(prop_queue, text_queue, prop_feats, text_feats, queue_ptr, n, E, Q, skip_flag);
  SPMM_CHECK_LAUNCH();
  enqueue_ptr_kernel<<<1, 1, 0, st>>>(queue_ptr, n, Q, skip_flag);
  SPMM_CHECK_LAUNCH();
  return 0;
}
