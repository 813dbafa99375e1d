// Flat-arena kernels: EMA of the momentum encoders, fp32->bf16 weight shadows, grad-norm, fused clip+AdamW,
// and a few bandwidth-bound helpers.  All are HBM-roofline kernels: 128-bit accesses, grid = k * 148 CTAs.
#include "common.cuh"
#include "spmm_b200.h"

const unsigned long long* spmm_g_rng_salt = nullptr;

namespace spmm {

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// p_m = p_m*m + p*(1-m): three separately rounded ops == reference SPMM_models.py:269 bit for bit.
__global__ void __launch_bounds__(256) ema_kernel(const float4* __restrict__ p, float4* __restrict__ pm,
                                                  uint2* __restrict__ p_bf, uint2* __restrict__ pm_bf, int64_t n4,
                                                  float m, float om) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = ldg_stream(p + i);
    float4 b = pm[i];
    b.x = __fadd_rn(__fmul_rn(b.x, m), __fmul_rn(a.x, om));
    b.y = __fadd_rn(__fmul_rn(b.y, m), __fmul_rn(a.y, om));
    b.z = __fadd_rn(__fmul_rn(b.z, m), __fmul_rn(a.z, om));
    b.w = __fadd_rn(__fmul_rn(b.w, m), __fmul_rn(a.w, om));
    pm[i] = b;
    if (p_bf) p_bf[i] = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
    if (pm_bf) pm_bf[i] = make_uint2(pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
  }
}

__global__ void __launch_bounds__(256) cast_kernel(const float4* __restrict__ s, uint2* __restrict__ d, int64_t n4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = ldg_stream(s + i);
    d[i] = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
  }
}

// sum g^2 with a FIXED summation order: per-CTA partials, then the last CTA to finish (ticket counter) adds them in
// index order.  Data-parallel replicas hold bit-identical reduced gradients, so they compute the bit-identical clip
// coefficient and their weights never drift apart (a float atomicAdd per CTA made the total depend on arrival order).
__global__ void __launch_bounds__(256) sumsq_kernel(const float4* __restrict__ g, int64_t n4, float* out,
                                                    float* __restrict__ partials, unsigned int* ticket) {
  __shared__ float sh[32];
  __shared__ bool last;
  float acc = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = ldg_stream(g + i);
    acc += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = acc;
    __threadfence();
    last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;   // wraps to 0: ready for the next call
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += __ldcg(partials + i);
  t = block_sum(t, sh);
  if (threadIdx.x == 0) *out = t;
}

// torch.nn.utils.clip_grad_norm_(params, max_norm) followed by torch.optim.AdamW.step (decoupled decay).
__global__ void __launch_bounds__(256)
adamw_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m1, float4* __restrict__ m2,
             int64_t n4, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
             const float* __restrict__ sumsq, float max_norm, float gscale, const float* __restrict__ skip,
             const float* __restrict__ hyper) {
  if (skip != nullptr && *skip != 0.f) return;
  if (hyper != nullptr) { lr = hyper[0]; bc1 = hyper[1]; bc2_sqrt = hyper[2]; }  // per-step values from device memory (CUDA graphs)
  float coef = gscale;
  if (sumsq != nullptr && max_norm > 0.f) {
    const float total = sqrtf(*sumsq) * gscale;
    const float c = max_norm / (total + 1e-6f);
    coef *= fminf(c, 1.f);
  }
  const float step_size = lr / bc1;
  const float decay = 1.f - lr * wd;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 w = p[i];
    const float4 gr = ldg_stream(g + i);
    float4 a = m1[i], v = m2[i];
#define SPMM_ADAM1(c)                                           \
  {                                                             \
    const float gg = gr.c * coef;                               \
    w.c *= decay;                                               \
    a.c = a.c + (gg - a.c) * (1.f - b1);                        \
    v.c = v.c * b2 + (1.f - b2) * gg * gg;                      \
    const float denom = sqrtf(v.c) / bc2_sqrt + eps;            \
    w.c -= step_size * (a.c / denom);                           \
  }
    SPMM_ADAM1(x) SPMM_ADAM1(y) SPMM_ADAM1(z) SPMM_ADAM1(w)
#undef SPMM_ADAM1
    p[i] = w; m1[i] = a; m2[i] = v;
  }
}

__global__ void __launch_bounds__(256) add_bf16_kernel(uint4* __restrict__ d, const uint4* __restrict__ s, int64_t n8) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    uint4 a = d[i];
    const uint4 b = s[i];
    float x0, x1, y0, y1;
#define SPMM_ADD2(c) unpack_bf16x2(a.c, x0, x1); unpack_bf16x2(b.c, y0, y1); a.c = pack_bf16x2(x0 + y0, x1 + y1);
    SPMM_ADD2(x) SPMM_ADD2(y) SPMM_ADD2(z) SPMM_ADD2(w)
#undef SPMM_ADD2
    d[i] = a;
  }
}

__global__ void __launch_bounds__(256) dgelu_kernel(const uint4* __restrict__ da, const uint4* __restrict__ pre,
                                                    uint4* __restrict__ dp, int64_t n8) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 a = da[i], b = pre[i];
    uint4 o;
    float x0, x1, y0, y1;
#define SPMM_DG2(c) unpack_bf16x2(a.c, x0, x1); unpack_bf16x2(b.c, y0, y1); o.c = pack_bf16x2(x0 * dgelu_erf(y0), x1 * dgelu_erf(y1));
    SPMM_DG2(x) SPMM_DG2(y) SPMM_DG2(z) SPMM_DG2(w)
#undef SPMM_DG2
    dp[i] = o;
  }
}

// out[c] += sum_r x[r][c]; block = 32x8 threads handles a 256-column x rows_per_block slab
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, int ld, float* out, int rows,
                                                     int cols, int rows_per_block) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[8][33 * 8];
  const int c0 = blockIdx.x * 256 + threadIdx.x % 32 * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < cols) {
    for (int r = r0 + threadIdx.x / 32; r < r1; r += 8) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + (size_t)r * ld + c0);
      float a, b;
      unpack_bf16x2(u.x, a, b); acc[0] += a; acc[1] += b;
      unpack_bf16x2(u.y, a, b); acc[2] += a; acc[3] += b;
      unpack_bf16x2(u.z, a, b); acc[4] += a; acc[5] += b;
      unpack_bf16x2(u.w, a, b); acc[6] += a; acc[7] += b;
    }
  }
  const int w = threadIdx.x / 32, l = threadIdx.x % 32;
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[w][l * 8 + j + l / 4] = acc[j];
  __syncthreads();
  if (w == 0 && c0 < cols) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
      for (int ww = 0; ww < 8; ++ww) s += sh[ww][l * 8 + j + l / 4];
      if (c0 + j < cols) atomicAdd(out + c0 + j, s);
    }
  }
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const uint4* __restrict__ src, const int* __restrict__ idx,
                                                          uint4* __restrict__ dst, int64_t row_vec) {
  const int r = blockIdx.x;
  const uint4* s = src + (int64_t)idx[r] * row_vec;
  uint4* d = dst + (int64_t)r * row_vec;
  for (int64_t i = threadIdx.x; i < row_vec; i += blockDim.x) d[i] = s[i];
}

// dst[t] = sum over r with idx[r] == t of src[r] (fp32 accumulation, fixed order, written once; zero when no r maps to
// t).  gridDim.x = destination rows, gridDim.y = column slices of a row.
constexpr int kSegMax = 4096;
__global__ void __launch_bounds__(256) segment_sum_rows_kernel(uint4* __restrict__ dst, const int* __restrict__ idx,
                                                               const uint4* __restrict__ src, int n_idx,
                                                               int64_t row_vec) {
  __shared__ int members[kSegMax];
  __shared__ int n_members;
  const int target = blockIdx.x;
  if (threadIdx.x == 0) n_members = 0;
  __syncthreads();
  for (int r = threadIdx.x; r < n_idx; r += blockDim.x)
    if (idx[r] == target) members[atomicAdd(&n_members, 1)] = r;
  __syncthreads();
  const int m = n_members;
  // ascending source order, whatever order the atomics landed in: the sum is reproducible
  if (threadIdx.x == 0)
    for (int i = 1; i < m; ++i) {
      const int v = members[i];
      int j = i - 1;
      for (; j >= 0 && members[j] > v; --j) members[j + 1] = members[j];
      members[j + 1] = v;
    }
  __syncthreads();
  const int64_t per = (row_vec + gridDim.y - 1) / gridDim.y;
  const int64_t i0 = blockIdx.y * per, i1 = i0 + per < row_vec ? i0 + per : row_vec;
  uint4* d = dst + (int64_t)target * row_vec;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int j = 0;
    for (; j + 4 <= m; j += 4) {          // four rows in flight
      uint4 b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) b[u] = __ldg(src + (int64_t)members[j + u] * row_vec + i);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float x, y;
        unpack_bf16x2(b[u].x, x, y); acc[0] += x; acc[1] += y;
        unpack_bf16x2(b[u].y, x, y); acc[2] += x; acc[3] += y;
        unpack_bf16x2(b[u].z, x, y); acc[4] += x; acc[5] += y;
        unpack_bf16x2(b[u].w, x, y); acc[6] += x; acc[7] += y;
      }
    }
    for (; j < m; ++j) {
      const uint4 b = __ldg(src + (int64_t)members[j] * row_vec + i);
      float x, y;
      unpack_bf16x2(b.x, x, y); acc[0] += x; acc[1] += y;
      unpack_bf16x2(b.y, x, y); acc[2] += x; acc[3] += y;
      unpack_bf16x2(b.z, x, y); acc[4] += x; acc[5] += y;
      unpack_bf16x2(b.w, x, y); acc[6] += x; acc[7] += y;
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    d[i] = o;
  }
}

static inline int flat_grid(int64_t n_items, int per_sm) {
  int64_t blocks = (n_items + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * per_sm;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace spmm
using namespace spmm;

extern "C" int spmm_version(void) { return 100; }

extern "C" int spmm_set_rng_salt_ptr(const unsigned long long* dev_ptr) {
  spmm_g_rng_salt = dev_ptr;
  return 0;
}

extern "C" int spmm_ema_multi(const float* p, float* p_m, void* p_bf16, void* p_m_bf16, int64_t n, float momentum,
                              float one_minus_momentum, void* stream) {
  SPMM_ARG(p && p_m && n >= 0 && n % 4 == 0);
  SPMM_ARG(((uintptr_t)p & 15) == 0 && ((uintptr_t)p_m & 15) == 0 && ((uintptr_t)p_bf16 & 7) == 0 &&
           ((uintptr_t)p_m_bf16 & 7) == 0);
  if (n == 0) return 0;
  ema_kernel<<<flat_grid(n / 4, 8), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)p, (float4*)p_m, (uint2*)p_bf16, (uint2*)p_m_bf16, n / 4, momentum, one_minus_momentum);
  SPMM_CHECK_LAUNCH();
  return 0;
}

extern "C" int spmm_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
  SPMM_ARG(src && dst && n >= 0 && n % 4 == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0);
  if (n == 0) return 0;
  cast_kernel<<<flat_grid(n / 4, 8), 256, 0, (cudaStream_t)stream>>>((const float4*)src, (uint2*)dst, n / 4);
  SPMM_CHECK_LAUNCH();
  return 0;
}

extern "C" int spmm_grad_sumsq(const float* g, int64_t n, float* sumsq_out, float* workspace, void* stream) {
  SPMM_ARG(g && sumsq_out && workspace && n >= 0 && n % 4 == 0 && ((uintptr_t)g & 15) == 0);
  if (n == 0) return (int)cudaMemsetAsync(sumsq_out, 0, sizeof(float), (cudaStream_t)stream);
  // workspace: [0] ticket counter (zero before first use, the kernel leaves it zero), [1 ..] per-CTA partials
  sumsq_kernel<<<flat_grid(n / 4, 4), 256, 0, (cudaStream_t)stream>>>((const float4*)g, n / 4, sumsq_out, workspace + 1,
                                                                      reinterpret_cast<unsigned int*>(workspace));
  SPMM_CHECK_LAUNCH();
  return 0;
}

// Advances the optimiser's step counter ON THE DEVICE and publishes (lr, 1-beta1^t, sqrt(1-beta2^t)) for adamw_kernel.
// The counter used to travel host -> pinned -> device per step; a host that enqueues steps ahead of the GPU (or replays
// a CUDA graph back to back) could overwrite the pinned value before the copy of the previous step had executed.
__global__ void adam_tick_kernel(long long* t_dev, const float* lr_dev, float* hyper, double beta1, double beta2,
                                 const float* skip_flag) {
  if (skip_flag != nullptr && *skip_flag != 0.f) return;   // NaN-guarded step: the reference skips optimizer.step()
  const long long t = *t_dev + 1;
  *t_dev = t;
  hyper[0] = *lr_dev;
  hyper[1] = (float)(1.0 - pow(beta1, (double)t));
  hyper[2] = (float)sqrt(1.0 - pow(beta2, (double)t));
}

extern "C" int spmm_adam_tick(long long* t_dev, const float* lr_dev, float* hyper_dev, float beta1, float beta2,
                              const float* skip_flag, void* stream) {
  SPMM_ARG(t_dev && lr_dev && hyper_dev);
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(t_dev, lr_dev, hyper_dev, (double)beta1, (double)beta2, skip_flag);
  SPMM_CHECK_LAUNCH();
  return 0;
}

extern "C" int spmm_adamw_step(float* p, const float* g, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                               float beta1, float beta2, float eps, float weight_decay, int step, const float* sumsq,
                               float max_norm, float grad_scale, const float* skip_flag, const float* hyper_dev, void* stream) {
  SPMM_ARG(p && g && exp_avg && exp_avg_sq && n >= 0 && n % 4 == 0 && step >= 1);
  SPMM_ARG((((uintptr_t)p | (uintptr_t)g | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0);
  if (n == 0) return 0;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adamw_kernel<<<flat_grid(n / 4, 8), 256, 0, (cudaStream_t)stream>>>(
      (float4*)p, (const float4*)g, (float4*)exp_avg, (float4*)exp_avg_sq, n / 4, lr, beta1, beta2, eps, weight_decay,
      (float)bc1, (float)sqrt(bc2), sumsq, max_norm, grad_scale, skip_flag, hyper_dev);
  SPMM_CHECK_LAUNCH();
  return 0;
}

extern "C" int spmm_add_bf16(void* dst, const void* src, int64_t n, void* stream) {
  SPMM_ARG(dst && src && n >= 0 && n % 8 == 0 && (((uintptr_t)dst | (uintptr_t)src) & 15) == 0);
  if (n == 0) return 0;
  add_bf16_kernel<<<flat_grid(n / 8, 8), 256, 0, (cudaStream_t)stream>>>((uint4*)dst, (const uint4*)src, n / 8);
  SPMM_CHECK_LAUNCH();
  return 0;
}

extern "C" int spmm_dgelu_bf16(const void* d_act, const void* pre_act, void* d_pre, int64_t n, void* stream) {
  SPMM_ARG(d_act && pre_act && d_pre && n >= 0 && n % 8 == 0 &&
           (((uintptr_t)d_act | (uintptr_t)pre_act | (uintptr_t)d_pre) & 15) == 0);
  if (n == 0) return 0;
  dgelu_kernel<<<flat_grid(n / 8, 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)d_act, (const uint4*)pre_act,
                                                                      (uint4*)d_pre, n / 8);
  SPMM_CHECK_LAUNCH();
  return 0;
}

extern "C" int spmm_colsum_bf16(const void* x, int ld, float* out, int rows, int cols, void* stream) {
  SPMM_ARG(x && out && rows > 0 && cols > 0 && ld % 8 == 0 && cols % 8 == 0 && ((uintptr_t)x & 15) == 0);
  const int col_blocks = (cols + 255) / 256;
  int row_blocks = (2 * kNumSMs + col_blocks - 1) / col_blocks;
  if (row_blocks > (rows + 31) / 32) row_blocks = (rows + 31) / 32;
  if (row_blocks < 1) row_blocks = 1;
  const int rpb = (rows + row_blocks - 1) / row_blocks;
  cudaError_t le = launch_pdl(colsum_kernel, dim3(col_blocks, row_blocks), dim3(256), 0, (cudaStream_t)stream,
                              (const __nv_bfloat16*)x, ld, out, rows, cols, rpb);
  if (le != cudaSuccess) return (int)le;
  return 0;
}

extern "C" int spmm_gather_rows_bf16(const void* src, const int* idx, void* dst, int n_idx, int64_t row_elems,
                                     void* stream) {
  SPMM_ARG(src && idx && dst && n_idx > 0 && row_elems % 8 == 0 && (((uintptr_t)dst | (uintptr_t)src) & 15) == 0);
  gather_rows_kernel<<<n_idx, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, idx, (uint4*)dst, row_elems / 8);
  SPMM_CHECK_LAUNCH();
  return 0;
}

extern "C" int spmm_segment_sum_rows_bf16(void* dst, int n_dst, const int* idx, const void* src, int n_idx,
                                          int64_t row_elems, void* stream) {
  SPMM_ARG(src && idx && dst && n_dst > 0 && n_idx > 0 && n_idx <= kSegMax && row_elems % 8 == 0 &&
           (((uintptr_t)dst | (uintptr_t)src) & 15) == 0);
  const int64_t row_vec = row_elems / 8;
  int splits = (int)((row_vec + 1023) / 1024);        // <= 4 vectors per thread and slice
  if (splits < 1) splits = 1;
  if (splits > 64) splits = 64;
  segment_sum_rows_kernel<<<dim3(n_dst, splits), 256, 0, (cudaStream_t)stream>>>((uint4*)dst, idx, (const uint4*)src,
                                                                                  n_idx, row_vec);
  SPMM_CHECK_LAUNCH();
  return 0;
}

