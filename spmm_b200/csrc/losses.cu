// Fused loss heads (forward + gradient in one pass, fp32 math over bf16 activations):
//   LM  : next-token CE over ALL positions + momentum-distillation soft CE (reference SPMM_models.py:233-238)
//   ITM : Linear(2H,2) + CE with labels [1]*B + [0]*2B                      (SPMM_models.py:201-206)
//   MPM : Linear(H,1) + masked MSE, x5                                      (SPMM_models.py:251-256)
#include "common.cuh"
#include "spmm_b200.h"

namespace spmm {

// ------------------------------------------------------------------ LM loss
// `valid_len` (device scalar, optional): the batch's own padded width Lv <= L (= its longest sequence).  Rows are
// stored with stride L (a CUDA-graph length bucket); positions t >= Lv - 1 are bucket padding the reference never saw:
// they get zero gradient and do not count in the CE mean over B * (Lv - 1) positions.
__global__ void lm_count_kernel(const int64_t* __restrict__ ids, int B, int L, float* ws) {
  __shared__ float sh[32];
  float c = 0.f;
  for (int i = threadIdx.x; i < B * (L - 1); i += blockDim.x) {
    const int b = i / (L - 1), t = i % (L - 1);
    c += (ids[(size_t)b * L + t + 1] != 0) ? 1.f : 0.f;   // labels != 0, SPMM_models.py:237 (bucket padding is id 0)
  }
  c = block_sum(c, sh);
  if (threadIdx.x == 0) { ws[0] = c; ws[1] = 0.f; ws[2] = 0.f; }
}

constexpr int LM_MAXV = 320;  // per-lane register budget: 10 logits

__global__ void __launch_bounds__(256)
lm_loss_kernel(const __nv_bfloat16* __restrict__ logits, const __nv_bfloat16* __restrict__ teacher, int ld,
               const int64_t* __restrict__ ids, int B, int L, int V, float alpha, const float* __restrict__ alpha_dev,
               const int* __restrict__ valid_len, float* ws, __nv_bfloat16* __restrict__ dlogits) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B * L) return;
  if (alpha_dev != nullptr) alpha = __ldg(alpha_dev);
  const int Lv = valid_len != nullptr ? min(max(__ldg(valid_len), 2), L) : L;
  const int b = row / L, t = row % L;
  __nv_bfloat16* drow = dlogits + (size_t)row * ld;
  if (t >= Lv - 1) {  // sliced away by [:, :-1] (or bucket padding) -> zero gradient
    for (int c = lane; c < ld; c += 32) drow[c] = f2bf(0.f);
    return;
  }
  const int64_t label = ids[(size_t)b * L + t + 1];
  const float n_valid = ws[0];
  float s[LM_MAXV / 32], m[LM_MAXV / 32];
  float mxs = -INFINITY, mxm = -INFINITY;
#pragma unroll
  for (int i = 0; i < LM_MAXV / 32; ++i) {
    const int c = lane + 32 * i;
    s[i] = c < V ? bf2f(logits[(size_t)row * ld + c]) : -INFINITY;
    m[i] = c < V ? bf2f(teacher[(size_t)row * ld + c]) : -INFINITY;
    mxs = fmaxf(mxs, s[i]);
    mxm = fmaxf(mxm, m[i]);
  }
  mxs = warp_max(mxs);
  mxm = warp_max(mxm);
  float ses = 0.f, sem = 0.f, dot = 0.f, slab = 0.f;
#pragma unroll
  for (int i = 0; i < LM_MAXV / 32; ++i) {
    const int c = lane + 32 * i;
    if (c < V) {
      const float es = __expf(s[i] - mxs), em = __expf(m[i] - mxm);
      ses += es; sem += em; dot += em * s[i];
      if (c == label) slab = s[i];
      s[i] = es; m[i] = em;
    }
  }
  ses = warp_sum(ses); sem = warp_sum(sem); dot = warp_sum(dot); slab = warp_sum(slab);
  const float lse = mxs + __logf(ses);
  const float ce = lse - slab;
  const bool valid = label != 0;
  const float distill = lse - dot / sem;
  if (lane == 0) {
    atomicAdd(ws + 1, ce);
    if (valid) atomicAdd(ws + 2, distill);
  }
  const float w_ce = (1.f - alpha) / (float)(B * (Lv - 1));
  const float w_ds = valid ? alpha / n_valid : 0.f;
#pragma unroll
  for (int i = 0; i < LM_MAXV / 32; ++i) {
    const int c = lane + 32 * i;
    if (c < V) {
      const float ps = s[i] / ses, pm = m[i] / sem;
      const float g = w_ce * (ps - (c == label ? 1.f : 0.f)) + w_ds * (ps - pm);
      drow[c] = f2bf(g);
    } else if (c < ld) {
      drow[c] = f2bf(0.f);
    }
  }
}

__global__ void lm_final_kernel(const float* ws, int B, int L, float alpha, const float* alpha_dev, const int* valid_len,
                                float* loss) {
  if (alpha_dev != nullptr) alpha = *alpha_dev;
  const int Lv = valid_len != nullptr ? min(max(*valid_len, 2), L) : L;
  *loss = (1.f - alpha) * ws[1] / (float)(B * (Lv - 1)) + alpha * ws[2] / ws[0];
}

// ------------------------------------------------------------------ ITM head + CE
// Two small multi-CTA kernels (a single CTA took 0.22 ms for ~1 MFLOP: 288 dependent row loads per column thread).
// Phase 1, warp per row: logits, CE, d logits -> workspace, dx row.  Phase 2, thread per column x row split: dw (atomics
// over the splits), and in CTA (0,0) db and the loss (fixed summation order).
__global__ void __launch_bounds__(256)
itm_rows_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int n_rows,
                int n_pos, int D, float* __restrict__ ws, __nv_bfloat16* __restrict__ dx) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  float a0 = 0.f, a1 = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float xv = bf2f(x[(size_t)r * D + k]);
    a0 += xv * w[k];
    a1 += xv * w[D + k];
  }
  a0 = warp_sum(a0) + bias[0];
  a1 = warp_sum(a1) + bias[1];
  const float mx = fmaxf(a0, a1);
  const float e0 = __expf(a0 - mx), e1 = __expf(a1 - mx), se = e0 + e1;
  const int label = r < n_pos ? 1 : 0;
  const float d0 = (e0 / se - (label == 0 ? 1.f : 0.f)) / n_rows, d1 = (e1 / se - (label == 1 ? 1.f : 0.f)) / n_rows;
  if (lane == 0) {
    ws[2 * r] = d0;
    ws[2 * r + 1] = d1;
    ws[2 * n_rows + r] = mx + __logf(se) - (label ? a1 : a0);
  }
  for (int k = lane; k < D; k += 32) dx[(size_t)r * D + k] = f2bf(d0 * w[k] + d1 * w[D + k]);
}

__global__ void __launch_bounds__(256)
itm_cols_kernel(const __nv_bfloat16* __restrict__ x, int n_rows, int D, const float* __restrict__ ws, float* loss, float* dw,
                float* db) {
  __shared__ float sh[32];
  const int k = blockIdx.x * 256 + threadIdx.x;
  const int per = (n_rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * per, r1 = min(n_rows, r0 + per);
  if (k < D) {
    float g0 = 0.f, g1 = 0.f;
    for (int r = r0; r < r1; ++r) {
      const float xv = bf2f(x[(size_t)r * D + k]);
      g0 += ws[2 * r] * xv;
      g1 += ws[2 * r + 1] * xv;
    }
    atomicAdd(dw + k, g0);
    atomicAdd(dw + D + k, g1);
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    float l = 0.f, b0 = 0.f, b1 = 0.f;
    for (int r = threadIdx.x; r < n_rows; r += 256) { l += ws[2 * n_rows + r]; b0 += ws[2 * r]; b1 += ws[2 * r + 1]; }
    l = block_sum(l, sh); b0 = block_sum(b0, sh); b1 = block_sum(b1, sh);
    if (threadIdx.x == 0) { *loss = l / n_rows; db[0] += b0; db[1] += b1; }
  }
}

// ------------------------------------------------------------------ MPM tail
__global__ void mpm_count_kernel(const float* __restrict__ mpm, int n, float* ws) {
  __shared__ float sh[32];
  float c = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) c += (mpm[i] == 0.f) ? 1.f : 0.f;  // (1 - mpm_mask).bool()
  c = block_sum(c, sh);
  if (threadIdx.x == 0) { ws[0] = c; ws[1] = 0.f; }
}

// rows r = b*(n_prop+1) + j; only j < n_prop with mpm_mask == 0 contribute
__global__ void __launch_bounds__(256)
mpm_kernel(const __nv_bfloat16* __restrict__ t, const float* __restrict__ w, const float* __restrict__ bias,
           const float* __restrict__ pv, const float* __restrict__ mpm, int batch, int n_prop, int H, float* ws,
           __nv_bfloat16* __restrict__ dt, float* dw, float* db) {
  extern __shared__ float shw[];  // [8][H]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = batch * (n_prop + 1);
  const float n_valid = ws[0];
  float pw[32];  // per-lane partial dw for columns lane + 32*i  (H <= 1024)
#pragma unroll
  for (int i = 0; i < 32; ++i) pw[i] = 0.f;
  float pb = 0.f, pl = 0.f;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const int b = r / (n_prop + 1), j = r % (n_prop + 1);
    const bool valid = j < n_prop && mpm[b * n_prop + j] == 0.f;
    float dpred = 0.f;
    if (valid) {
      float a = 0.f;
      for (int k = lane; k < H; k += 32) a += bf2f(t[(size_t)r * H + k]) * w[k];
      a = warp_sum(a) + bias[0];
      const float diff = a - pv[b * n_prop + j];
      pl += diff * diff;                       // identical in every lane
      dpred = 5.f * 2.f * diff / n_valid;      // d(5 * mse)/d pred
      pb += dpred;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k = lane + 32 * i;
      if (k < H) {
        if (valid) pw[i] += dpred * bf2f(t[(size_t)r * H + k]);
        dt[(size_t)r * H + k] = f2bf(dpred * w[k]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int k = lane + 32 * i;
    if (k < H) shw[warp * H + k] = pw[i];
  }
  __shared__ float sb[8], sl[8];
  if (lane == 0) { sb[warp] = pb; sl[warp] = pl; }
  __syncthreads();
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) s += shw[ww * H + k];
    atomicAdd(dw + k, s);
  }
  if (threadIdx.x == 0) {
    float s = 0.f, l = 0.f;
    for (int ww = 0; ww < 8; ++ww) { s += sb[ww]; l += sl[ww]; }
    atomicAdd(db, s);
    atomicAdd(ws + 1, l);
  }
}

__global__ void mpm_final_kernel(const float* ws, float* loss) { *loss = 5.f * ws[1] / ws[0]; }

}  // namespace spmm
using namespace spmm;

extern "C" int spmm_lm_loss_fwd_bwd(const void* logits, const void* teacher_logits, int ld, const int64_t* ids, int B,
                                    int L, int V, float alpha, const float* alpha_dev, const int* valid_len,
                                    float* loss, void* dlogits, float* workspace, void* stream) {
  SPMM_ARG(logits && teacher_logits && ids && loss && dlogits && workspace && B > 0 && L > 1 && V > 0 && V <= LM_MAXV &&
           ld >= V);
  cudaStream_t st = (cudaStream_t)stream;
  lm_count_kernel<<<1, 256, 0, st>>>(ids, B, L, workspace);
  SPMM_CHECK_LAUNCH();
  lm_loss_kernel<<<(B * L + 7) / 8, 256, 0, st>>>((const __nv_bfloat16*)logits, (const __nv_bfloat16*)teacher_logits, ld,
                                                  ids, B, L, V, alpha, alpha_dev, valid_len, workspace,
                                                  (__nv_bfloat16*)dlogits);
  SPMM_CHECK_LAUNCH();
  lm_final_kernel<<<1, 1, 0, st>>>(workspace, B, L, alpha, alpha_dev, valid_len, loss);
  SPMM_CHECK_LAUNCH();
  return 0;
}

extern "C" int spmm_itm_loss_fwd_bwd(const void* x, const float* w, const float* b, int n_rows, int n_pos, int D,
                                     float* loss, void* dx, float* dw, float* db, float* workspace, void* stream) {
  SPMM_ARG(x && w && b && loss && dx && dw && db && workspace && n_rows > 0 && D > 0);
  cudaStream_t st = (cudaStream_t)stream;
  itm_rows_kernel<<<(n_rows + 7) / 8, 256, 0, st>>>((const __nv_bfloat16*)x, w, b, n_rows, n_pos, D, workspace,
                                                    (__nv_bfloat16*)dx);
  SPMM_CHECK_LAUNCH();
  const int splits = n_rows >= 64 ? 8 : 1;
  itm_cols_kernel<<<dim3((D + 255) / 256, splits), 256, 0, st>>>((const __nv_bfloat16*)x, n_rows, D, workspace, loss, dw, db);
  SPMM_CHECK_LAUNCH();
  return 0;
}

extern "C" int spmm_mpm_loss_fwd_bwd(const void* t, const float* w, const float* b, const float* pv,
                                     const float* mpm_mask, int batch, int n_prop, int H, float* loss, void* dt,
                                     float* dw, float* db, float* workspace, void* stream) {
  SPMM_ARG(t && w && b && pv && mpm_mask && loss && dt && dw && db && workspace && batch > 0 && n_prop > 0 && H > 0 &&
           H <= 1024);
  cudaStream_t st = (cudaStream_t)stream;
  mpm_count_kernel<<<1, 256, 0, st>>>(mpm_mask, batch * n_prop, workspace);
  SPMM_CHECK_LAUNCH();
  const int rows = batch * (n_prop + 1);
  int grid = (rows + 7) / 8;
  if (grid > kNumSMs) grid = kNumSMs;
  mpm_kernel<<<grid, 256, (size_t)8 * H * sizeof(float), st>>>((const __nv_bfloat16*)t, w, b, pv, mpm_mask, batch, n_prop,
                                                               H, workspace, (__nv_bfloat16*)dt, dw, db);
  SPMM_CHECK_LAUNCH();
  mpm_final_kernel<<<1, 1, 0, st>>>(workspace, loss);
  SPMM_CHECK_LAUNCH();
  return 0;
}
