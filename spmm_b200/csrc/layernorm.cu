// LayerNorm forward/backward over bf16 rows (reference: nn.LayerNorm sites xbert.py:184,366,444,670;
// eps 1e-12; fp32 statistics as torch's autocast keeps LayerNorm in fp32).
// One warp per row; a lane owns 16-byte chunks {lane, lane+32, ...} of the row, so a 768-wide row is
// three fully coalesced 512-byte warp transactions.  HBM-bound: 2 B read + 2 B write per element forward.
#include "common.cuh"
#include "spmm_b200.h"

namespace spmm {

constexpr int LN_WARPS = 8;
constexpr int LN_SLOTS = 8;

__device__ __forceinline__ bool ln_keep16(unsigned long long seed, unsigned long long e, uint32_t thresh16) {
  return keep16(seed, e, thresh16);
}

template <int NCH>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int H,
              float eps, unsigned long long seed, uint32_t thresh16, float inv_keep) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const __nv_bfloat16* xr = x + (size_t)row * H;
  float v[NCH][8];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = (lane + 32 * c) * 8;
    if (col < H) {
      const uint4 u = *reinterpret_cast<const uint4*>(xr + col);
      unpack_bf16x2(u.x, v[c][0], v[c][1]); unpack_bf16x2(u.y, v[c][2], v[c][3]);
      unpack_bf16x2(u.z, v[c][4], v[c][5]); unpack_bf16x2(u.w, v[c][6], v[c][7]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[c][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[c][j] = 0.f;
    }
  }
  const float mean = warp_sum(sum) / H;
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = (lane + 32 * c) * 8;
    if (col < H) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[c][j] - mean; sq += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / H + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  __nv_bfloat16* yr = y + (size_t)row * H;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = (lane + 32 * c) * 8;
    if (col < H) {
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + col), g1 = *reinterpret_cast<const float4*>(gamma + col + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + col), b1 = *reinterpret_cast<const float4*>(beta + col + 4);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = (v[c][j] - mean) * rstd * g[j] + b[j];
        if (thresh16) o[j] = ln_keep16(seed, (unsigned long long)row * H + col + j, thresh16) ? o[j] * inv_keep : 0.f;
      }
      uint4 u;
      u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]); u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(yr + col) = u;
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; dgamma += dy * xhat; dbeta += dy.
// A CTA walks rows blockIdx.x, +gridDim.x, ... keeping per-lane column partials in registers, then reduces the
// 8 warps through shared memory and issues one fp32 atomicAdd per column per CTA.
template <int NCH>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean,
              const float* __restrict__ rstd, const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dx,
              float* dgamma, float* dbeta, __nv_bfloat16* __restrict__ dx_branch, float* dbias, int rows, int H,
              unsigned long long out_seed, uint32_t out_thresh, float out_inv_keep, unsigned long long br_seed,
              uint32_t br_thresh, float br_inv_keep, float* ws) {
  extern __shared__ float sh[];  // [LN_WARPS][H]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float pg[NCH][8], pb[NCH][8], ps[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) pg[c][j] = pb[c][j] = ps[c][j] = 0.f;

  for (int row = blockIdx.x * LN_WARPS + w; row < rows; row += gridDim.x * LN_WARPS) {
    const float mu = mean[row], rs = rstd[row];
    float xh[NCH][8], g[NCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = (lane + 32 * c) * 8;
      if (col < H) {
        const uint4 ux = *reinterpret_cast<const uint4*>(x + (size_t)row * H + col);
        const uint4 ud = *reinterpret_cast<const uint4*>(dy + (size_t)row * H + col);
        float d[8];
        unpack_bf16x2(ux.x, xh[c][0], xh[c][1]); unpack_bf16x2(ux.y, xh[c][2], xh[c][3]);
        unpack_bf16x2(ux.z, xh[c][4], xh[c][5]); unpack_bf16x2(ux.w, xh[c][6], xh[c][7]);
        unpack_bf16x2(ud.x, d[0], d[1]); unpack_bf16x2(ud.y, d[2], d[3]);
        unpack_bf16x2(ud.z, d[4], d[5]); unpack_bf16x2(ud.w, d[6], d[7]);
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + col), g1 = *reinterpret_cast<const float4*>(gamma + col + 4);
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (out_thresh)  // forward applied dropout after the affine: dy_affine = dy * mask / keep
            d[j] = ln_keep16(out_seed, (unsigned long long)row * H + col + j, out_thresh) ? d[j] * out_inv_keep : 0.f;
          xh[c][j] = (xh[c][j] - mu) * rs;
          g[c][j] = d[j] * gm[j];
          s1 += g[c][j];
          s2 += g[c][j] * xh[c][j];
          pg[c][j] += d[j] * xh[c][j];
          pb[c][j] += d[j];
        }
      }
    }
    s1 = warp_sum(s1) / H;
    s2 = warp_sum(s2) / H;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = (lane + 32 * c) * 8;
      if (col < H) {
        float o[8], ob[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = rs * (g[c][j] - s1 - xh[c][j] * s2);
          ob[j] = o[j];
          if (br_thresh)
            ob[j] = ln_keep16(br_seed, (unsigned long long)row * H + col + j, br_thresh) ? o[j] * br_inv_keep : 0.f;
        }
        uint4 u;
        u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]); u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
        *reinterpret_cast<uint4*>(dx + (size_t)row * H + col) = u;
        if (dx_branch) {
          uint4 ub;
          ub.x = pack_bf16x2(ob[0], ob[1]); ub.y = pack_bf16x2(ob[2], ob[3]); ub.z = pack_bf16x2(ob[4], ob[5]); ub.w = pack_bf16x2(ob[6], ob[7]);
          *reinterpret_cast<uint4*>(dx_branch + (size_t)row * H + col) = ub;
        }
        if (dbias) {
          // bias grad of the preceding dense = column sums of what flows into it (bf16-rounded like the wgrad operand)
#pragma unroll
          for (int j = 0; j < 8; ++j) ps[c][j] += bf2f(f2bf(ob[j]));
        }
      }
    }
  }
  // Column partials: warps -> smem -> fp32 atomics into one of LN_SLOTS workspace slots (spreads same-address
  // contention 296-way -> 296/LN_SLOTS-way); the last CTA to finish folds the slots into the gradient arena and
  // re-zeroes the workspace, so no second launch and no same-address atomic storm on dgamma/dbeta/dbias.
  float* outs[3] = {dgamma, dbeta, dbias};
  float* slot = ws + (size_t)(blockIdx.x % LN_SLOTS) * 3 * H;
#pragma unroll
  for (int which = 0; which < 3; ++which) {
    if (outs[which] == nullptr) continue;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = (lane + 32 * c) * 8;
      if (col < H) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sh[w * H + col + j] = which == 0 ? pg[c][j] : (which == 1 ? pb[c][j] : ps[c][j]);
      }
    }
    __syncthreads();
    for (int col = threadIdx.x; col < H; col += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int ww = 0; ww < LN_WARPS; ++ww) s += sh[ww * H + col];
      atomicAdd(slot + which * H + col, s);
    }
  }
  __threadfence();
  __syncthreads();
  __shared__ unsigned int s_last;
  unsigned int* counter = reinterpret_cast<unsigned int*>(ws + (size_t)LN_SLOTS * 3 * H);
  if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < LN_SLOTS; ++k) {
        s += __ldcg(ws + (size_t)k * 3 * H + i);
        ws[(size_t)k * 3 * H + i] = 0.f;
      }
      float* o = outs[i / H];
      if (o != nullptr) o[i % H] += s;
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

static inline void drop_params(float p, uint32_t& thresh, float& inv_keep) {
  thresh = p > 0.f ? (uint32_t)(p * 65536.f + 0.5f) : 0u;
  inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
}

}  // namespace spmm
using namespace spmm;

extern "C" int spmm_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                                  float* rstd, int rows, int H, float eps, float dropout_p, unsigned long long seed,
                                  void* stream) {
  SPMM_ARG(x && gamma && beta && y && rows > 0 && H > 0 && H % 8 == 0 && H <= 1024);
  SPMM_ARG((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0);
  uint32_t th; float ik;
  drop_params(dropout_p, th, ik);
  const int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  const int nch = (H + 255) / 256;
  cudaStream_t st = (cudaStream_t)stream;
#define SPMM_LN_FWD(N) ln_fwd_kernel<N><<<grid, LN_WARPS * 32, 0, st>>>((const __nv_bfloat16*)x, gamma, beta, (__nv_bfloat16*)y, mean, rstd, rows, H, eps, seed, th, ik)
  if (nch == 1) SPMM_LN_FWD(1); else if (nch == 2) SPMM_LN_FWD(2); else if (nch == 3) SPMM_LN_FWD(3); else SPMM_LN_FWD(4);
#undef SPMM_LN_FWD
  SPMM_CHECK_LAUNCH();
  return 0;
}

extern "C" int spmm_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                                  const float* gamma, void* dx, float* dgamma, float* dbeta, void* dx_branch,
                                  float* dbias, int rows, int H, float out_dropout_p, unsigned long long out_seed,
                                  float branch_dropout_p, unsigned long long branch_seed, float* workspace, void* stream) {
  SPMM_ARG(workspace != nullptr);
  SPMM_ARG(dy && x && mean && rstd && gamma && dx && rows > 0 && H > 0 && H % 8 == 0 && H <= 1024);
  SPMM_ARG((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)dx_branch | (uintptr_t)gamma) & 15) == 0);
  uint32_t oth, bth; float oik, bik;
  drop_params(out_dropout_p, oth, oik);
  drop_params(branch_dropout_p, bth, bik);
  int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  if (grid > 2 * kNumSMs) grid = 2 * kNumSMs;
  const int nch = (H + 255) / 256;
  const size_t smem = (size_t)LN_WARPS * H * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
#define SPMM_LN_BWD(N) ln_bwd_kernel<N><<<grid, LN_WARPS * 32, smem, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, mean, rstd, gamma, (__nv_bfloat16*)dx, dgamma, dbeta, (__nv_bfloat16*)dx_branch, dbias, rows, H, out_seed, oth, oik, branch_seed, bth, bik, workspace)
  if (nch == 1) SPMM_LN_BWD(1); else if (nch == 2) SPMM_LN_BWD(2); else if (nch == 3) SPMM_LN_BWD(3); else SPMM_LN_BWD(4);
#undef SPMM_LN_BWD
  SPMM_CHECK_LAUNCH();
  return 0;
}
