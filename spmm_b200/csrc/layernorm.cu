// LayerNorm forward/backward over bf16 rows (reference: nn.LayerNorm sites xbert.py:184,366,444,670;
// eps 1e-12; fp32 statistics as torch's autocast keeps LayerNorm in fp32).
// Forward: one warp per row; a lane owns 16-byte chunks {lane, lane+32, ...} of the row, so a 768-wide row is three
// fully coalesced 512-byte warp transactions.  Backward: ONE kernel (ln_bwd_fused_kernel) produces dx, the masked
// branch gradient and the parameter / bias column sums from a single read of dy and x; the older pair (row-parallel dx
// kernel + column-parallel parameter reduction) remains for H > 768 and for calls that want no parameter gradients
// (SPMM_LN_BWD_SPLIT=1 forces it, for A/B timing).  HBM-bound by design: 4 B/element forward, 8-10 B/element backward.
#include <cstdlib>
#include "common.cuh"
#include "spmm_b200.h"

namespace spmm {

constexpr int LN_WARPS = 8;

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  unpack_bf16x2(u.x, v[0], v[1]); unpack_bf16x2(u.y, v[2], v[3]);
  unpack_bf16x2(u.z, v[4], v[5]); unpack_bf16x2(u.w, v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&o)[8]) {
  uint4 u;
  u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]); u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// 6 CTAs per SM (40 registers): the 768 CTAs of a 6144-row call fit in ONE wave of 888 slots; at 5 per SM the last 28
// CTAs formed a second wave and the latency-bound kernel took two row latencies instead of one.
template <int NCH>
__global__ void __launch_bounds__(LN_WARPS * 32, 6)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int H,
              float eps, unsigned long long seed, uint32_t thresh16, float inv_keep, const unsigned long long* salt) {
  pdl_trigger();
  pdl_wait();
  const uint32_t key = thresh16 ? fold_seed(salted(seed, salt)) : 0u;
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
      load8(xr + col, v[c]);
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
      float g[8], b[8], o[8];
      load8f(gamma + col, g);
      load8f(beta + col, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[c][j] - mean) * rstd * g[j] + b[j];
      if (thresh16) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) drop_pair(key, (uint32_t)row * H + col + j, thresh16, inv_keep, o[j], o[j + 1]);
      }
      store8(yr + col, o);
    }
  }
}

// H == 256 * NCH (768 everywhere in the model): persistent form - a warp walks rows (stride = warps in the grid) with
// the next row's three 16-byte loads already in flight, gamma / beta of its columns live in registers, arithmetic on
// packed fp32 pairs.  The one-row-per-warp kernel above needed 3.5 waves for the 24576-row calls of the batched
// passes and spent 421 instructions per row.
constexpr int LNP_WARPS = 8;
template <int NCH>
__global__ void __launch_bounds__(LNP_WARPS * 32, 2)
ln_fwd_rows_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows,
                   float eps, unsigned long long seed, uint32_t thresh16, float inv_keep, const unsigned long long* salt) {
  constexpr int H = NCH * 256;
  pdl_trigger();
  pdl_wait();
  const uint32_t key = thresh16 ? fold_seed(salted(seed, salt)) : 0u;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nw = gridDim.x * LNP_WARPS;
  f32x2 g2[NCH][4], b2[NCH][4];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = (lane + 32 * c) * 8;
    const ulonglong2 g01 = *reinterpret_cast<const ulonglong2*>(gamma + col), g23 = *reinterpret_cast<const ulonglong2*>(gamma + col + 4);
    const ulonglong2 b01 = *reinterpret_cast<const ulonglong2*>(beta + col), b23 = *reinterpret_cast<const ulonglong2*>(beta + col + 4);
    g2[c][0] = g01.x; g2[c][1] = g01.y; g2[c][2] = g23.x; g2[c][3] = g23.y;
    b2[c][0] = b01.x; b2[c][1] = b01.y; b2[c][2] = b23.x; b2[c][3] = b23.y;
  }
  int row = blockIdx.x * LNP_WARPS + w;
  uint4 cur[NCH], nxt[NCH];
  if (row < rows) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) cur[c] = *reinterpret_cast<const uint4*>(x + (size_t)row * H + (lane + 32 * c) * 8);
  }
  for (; row < rows; row += nw) {
    const int nrow = row + nw;
    if (nrow < rows) {
#pragma unroll
      for (int c = 0; c < NCH; ++c) nxt[c] = *reinterpret_cast<const uint4*>(x + (size_t)nrow * H + (lane + 32 * c) * 8);
    }
    f32x2 v[NCH][4];
    f32x2 sum2 = 0ull;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      v[c][0] = bf16x2_to_f32x2(cur[c].x); v[c][1] = bf16x2_to_f32x2(cur[c].y);
      v[c][2] = bf16x2_to_f32x2(cur[c].z); v[c][3] = bf16x2_to_f32x2(cur[c].w);
#pragma unroll
      for (int q = 0; q < 4; ++q) sum2 = add2(sum2, v[c][q]);
    }
    float lo, hi;
    upk2(sum2, lo, hi);
    const float mean = warp_sum(lo + hi) * (1.f / H);
    const f32x2 nmean2 = splat2(-mean);
    f32x2 sq2 = 0ull;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[c][q] = add2(v[c][q], nmean2);
        sq2 = fma2(v[c][q], v[c][q], sq2);
      }
    upk2(sq2, lo, hi);
    const float rstd = rsqrtf(warp_sum(lo + hi) * (1.f / H) + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
    const f32x2 rstd2 = splat2(rstd);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = (lane + 32 * c) * 8;
      f32x2 o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        o[q] = fma2(mul2(v[c][q], rstd2), g2[c][q], b2[c][q]);
        if (thresh16) {
          const uint32_t hb = drop_bits2(key, (uint32_t)row * H + col + 2 * q);
          o[q] = mul2(o[q], pk2((hb & 0xFFFFu) >= thresh16 ? inv_keep : 0.f, (hb >> 16) >= thresh16 ? inv_keep : 0.f));
        }
      }
      uint4 pk;
      pk.x = f32x2_to_bf16x2(o[0]); pk.y = f32x2_to_bf16x2(o[1]); pk.z = f32x2_to_bf16x2(o[2]); pk.w = f32x2_to_bf16x2(o[3]);
      *reinterpret_cast<uint4*>(y + (size_t)row * H + col) = pk;
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) cur[c] = nxt[c];
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; optionally dx_branch = dx * dropout mask.
template <int NCH>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_dx_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                 const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                 __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dx_branch, int rows, int H,
                 unsigned long long out_seed, uint32_t out_thresh, float out_inv_keep, unsigned long long br_seed,
                 uint32_t br_thresh, float br_inv_keep, const unsigned long long* salt) {
  pdl_trigger();
  pdl_wait();
  const uint32_t out_key = out_thresh ? fold_seed(salted(out_seed, salt)) : 0u;
  const uint32_t br_key = br_thresh ? fold_seed(salted(br_seed, salt)) : 0u;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float mu = mean[row], rs = rstd[row];
  float xh[NCH][8], g[NCH][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = (lane + 32 * c) * 8;
    if (col < H) {
      float d[8], gm[8];
      load8(x + (size_t)row * H + col, xh[c]);
      load8(dy + (size_t)row * H + col, d);
      load8f(gamma + col, gm);
#pragma unroll
      if (out_thresh) {  // forward applied dropout after the affine: dy_affine = dy * mask / keep
#pragma unroll
        for (int j = 0; j < 8; j += 2) drop_pair(out_key, (uint32_t)row * H + col + j, out_thresh, out_inv_keep, d[j], d[j + 1]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[c][j] = (xh[c][j] - mu) * rs;
        g[c][j] = d[j] * gm[j];
        s1 += g[c][j];
        s2 += g[c][j] * xh[c][j];
      }
    }
  }
  s1 = warp_sum(s1) / H;
  s2 = warp_sum(s2) / H;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int col = (lane + 32 * c) * 8;
    if (col < H) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rs * (g[c][j] - s1 - xh[c][j] * s2);
      store8(dx + (size_t)row * H + col, o);
      if (dx_branch) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) drop_pair(br_key, (uint32_t)row * H + col + j, br_thresh, br_inv_keep, o[j], o[j + 1]);
        store8(dx_branch + (size_t)row * H + col, o);
      }
    }
  }
}

// Column-parallel parameter gradients: dgamma[c] += sum_r dy*xhat, dbeta[c] += sum_r dy, dbias[c] += sum_r branch[r][c].
// block = 32 lanes (8 columns each = 256 columns) x 8 warps striding rows; grid = (ceil(H/256), row blocks).
__global__ void __launch_bounds__(256)
ln_bwd_param_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                    const float* __restrict__ mean, const float* __restrict__ rstd,
                    const __nv_bfloat16* __restrict__ branch, float* dgamma, float* dbeta, float* dbias, int rows, int H,
                    int rows_per_block, unsigned long long out_seed, uint32_t out_thresh, float out_inv_keep,
                    const unsigned long long* salt) {
  pdl_trigger();
  pdl_wait();
  const uint32_t out_key = out_thresh ? fold_seed(salted(out_seed, salt)) : 0u;
  __shared__ float sh[3][8][33 * 8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float ag[8], ab[8], as[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ag[j] = ab[j] = as[j] = 0.f;
  if (c0 < H) {
    // two rows per iteration, all six 16-byte loads issued before the first use: the loop is bound by load latency
    // (a warp walks ~8 rows), so memory-level parallelism is what shortens it
    auto accumulate = [&](int r, const uint4& pd, const uint4& pxv, const uint4& pb, float mu, float rs) {
      float d[8], xv[8];
      unpack_bf16x2(pd.x, d[0], d[1]); unpack_bf16x2(pd.y, d[2], d[3]); unpack_bf16x2(pd.z, d[4], d[5]); unpack_bf16x2(pd.w, d[6], d[7]);
      unpack_bf16x2(pxv.x, xv[0], xv[1]); unpack_bf16x2(pxv.y, xv[2], xv[3]); unpack_bf16x2(pxv.z, xv[4], xv[5]); unpack_bf16x2(pxv.w, xv[6], xv[7]);
      if (out_thresh) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) drop_pair(out_key, (uint32_t)r * H + c0 + j, out_thresh, out_inv_keep, d[j], d[j + 1]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ag[j] += d[j] * (xv[j] - mu) * rs;
        ab[j] += d[j];
      }
      if (dbias != nullptr) {
        float bv[8];
        unpack_bf16x2(pb.x, bv[0], bv[1]); unpack_bf16x2(pb.y, bv[2], bv[3]); unpack_bf16x2(pb.z, bv[4], bv[5]); unpack_bf16x2(pb.w, bv[6], bv[7]);
#pragma unroll
        for (int j = 0; j < 8; ++j) as[j] += bv[j];
      }
    };
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    for (int r = r0 + w; r < r1; r += 16) {
      const int rb = r + 8;
      const bool two = rb < r1;
      const uint4 d0 = *reinterpret_cast<const uint4*>(dy + (size_t)r * H + c0);
      const uint4 x0 = *reinterpret_cast<const uint4*>(x + (size_t)r * H + c0);
      const uint4 b0 = dbias != nullptr ? *reinterpret_cast<const uint4*>(branch + (size_t)r * H + c0) : z4;
      const uint4 d1 = two ? *reinterpret_cast<const uint4*>(dy + (size_t)rb * H + c0) : z4;
      const uint4 x1 = two ? *reinterpret_cast<const uint4*>(x + (size_t)rb * H + c0) : z4;
      const uint4 b1 = (two && dbias != nullptr) ? *reinterpret_cast<const uint4*>(branch + (size_t)rb * H + c0) : z4;
      const float mu0 = mean[r], rs0 = rstd[r];
      const float mu1 = two ? mean[rb] : 0.f, rs1 = two ? rstd[rb] : 0.f;
      accumulate(r, d0, x0, b0, mu0, rs0);
      if (two) accumulate(rb, d1, x1, b1, mu1, rs1);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sh[0][w][lane * 8 + j + lane / 4] = ag[j];
    sh[1][w][lane * 8 + j + lane / 4] = ab[j];
    sh[2][w][lane * 8 + j + lane / 4] = as[j];
  }
  __syncthreads();
  // 256 threads: thread t reduces column t of this slab for the three outputs
  const int t = threadIdx.x, col = blockIdx.x * 256 + t;
  if (col < H) {
    const int idx = t + (t / 8) / 4;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) { s0 += sh[0][ww][idx]; s1 += sh[1][ww][idx]; s2 += sh[2][ww][idx]; }
    if (dgamma) atomicAdd(dgamma + col, s0);
    if (dbeta) atomicAdd(dbeta + col, s1);
    if (dbias) atomicAdd(dbias + col, s2);
  }
}

// One pass over dy and x for BOTH results of the backward: a warp walks rows (stride = warps in the grid), writes dx
// (and the dropout-masked branch gradient) of each, and keeps the column sums dgamma / dbeta / dbias of all its rows in
// registers; they leave through a shared-memory fold per CTA and one atomicAdd per column and CTA.  The two-kernel form
// read dy and x twice and the branch gradient once more (6 + 6..8 B/element); this one moves 4 B/element in, 2..4 out.
// A warp's rows arrive through its own ring of LNF_STAGES shared-memory slots filled by cp.async (each lane copies and
// later reads the same 16-byte chunks, so no cross-lane synchronisation is needed): with ~14 rows per warp and one row
// in flight the kernel was bound by load latency (48 us for 24576 rows), not by HBM.
constexpr int LNF_WARPS = 12;                            // 384 threads, one CTA per SM: ~170 registers per thread
constexpr int LNF_STAGES = 4;
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// H == 256 * NCH exactly (768 for every LayerNorm of the model): no column predicates.  The per-element arithmetic runs
// on packed fp32 pairs (fma/mul/add.rn.f32x2): the kernel is bound by issue slots at its 12 warps per SM.
template <int NCH>
__global__ void __launch_bounds__(LNF_WARPS * 32, 1)
ln_bwd_fused_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                    __nv_bfloat16* __restrict__ dx, __nv_bfloat16* __restrict__ dx_branch, float* dgamma, float* dbeta,
                    float* dbias, int rows, unsigned long long out_seed, uint32_t out_thresh, float out_inv_keep,
                    unsigned long long br_seed, uint32_t br_thresh, float br_inv_keep, const unsigned long long* salt) {
  constexpr int H = NCH * 256;
  extern __shared__ uint4 ln_ring[];                     // [warp][stage][x | dy][NCH * 32] 16-byte chunks
  pdl_trigger();
  __shared__ float sh[3][NCH * 256];
  for (int i = threadIdx.x; i < 3 * NCH * 256; i += LNF_WARPS * 32) (&sh[0][0])[i] = 0.f;
  __syncthreads();
  pdl_wait();
  const uint32_t out_key = out_thresh ? fold_seed(salted(out_seed, salt)) : 0u;
  const uint32_t br_key = br_thresh ? fold_seed(salted(br_seed, salt)) : 0u;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nw = gridDim.x * LNF_WARPS;
  uint4* ring = ln_ring + (size_t)w * LNF_STAGES * 2 * NCH * 32;
  auto issue = [&](int row, int slot) {
    if (row < rows) {
      uint4* dst = ring + (size_t)slot * 2 * NCH * 32;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col = (lane + 32 * c) * 8;
        cp_async16(dst + lane + 32 * c, x + (size_t)row * H + col);
        cp_async16(dst + NCH * 32 + lane + 32 * c, dy + (size_t)row * H + col);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // column sums of this lane's 8 * NCH columns over all rows of the warp: pairs (col, col + 1)
  f32x2 ag[NCH][4], ab[NCH][4], as[NCH][4];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int q = 0; q < 4; ++q) ag[c][q] = ab[c][q] = as[c][q] = 0ull;
  const int row0 = blockIdx.x * LNF_WARPS + w;
#pragma unroll
  for (int s = 0; s < LNF_STAGES - 1; ++s) issue(row0 + s * nw, s);
  float mu_nx = row0 < rows ? __ldg(mean + row0) : 0.f, rs_nx = row0 < rows ? __ldg(rstd + row0) : 0.f;
  int k = 0;
  for (int row = row0; row < rows; row += nw, ++k) {
    issue(row + (LNF_STAGES - 1) * nw, (k + LNF_STAGES - 1) % LNF_STAGES);
    const float mu = mu_nx, rs = rs_nx;
    if (row + nw < rows) { mu_nx = __ldg(mean + row + nw); rs_nx = __ldg(rstd + row + nw); }
    asm volatile("cp.async.wait_group %0;" ::"n"(LNF_STAGES - 1) : "memory");
    const uint4* src = ring + (size_t)(k % LNF_STAGES) * 2 * NCH * 32;
    uint4 xr[NCH], dr[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      xr[c] = src[lane + 32 * c];
      dr[c] = src[NCH * 32 + lane + 32 * c];
    }
    const f32x2 rs2 = splat2(rs), nmurs2 = splat2(-mu * rs);
    f32x2 s1v = 0ull, s2v = 0ull;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = (lane + 32 * c) * 8;
      const uint32_t xw[4] = {xr[c].x, xr[c].y, xr[c].z, xr[c].w};
      uint32_t dw[4] = {dr[c].x, dr[c].y, dr[c].z, dr[c].w};
      const ulonglong2 g01 = *reinterpret_cast<const ulonglong2*>(gamma + col);
      const ulonglong2 g23 = *reinterpret_cast<const ulonglong2*>(gamma + col + 4);
      const f32x2 gm[4] = {g01.x, g01.y, g23.x, g23.y};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        f32x2 d = bf16x2_to_f32x2(dw[q]);
        if (out_thresh) {   // forward applied dropout after the affine (embedding LayerNorm): dy_affine = dy * mask / keep
          const uint32_t hb = drop_bits2(out_key, (uint32_t)row * H + col + 2 * q);
          d = mul2(d, pk2((hb & 0xFFFFu) >= out_thresh ? out_inv_keep : 0.f, (hb >> 16) >= out_thresh ? out_inv_keep : 0.f));
        }
        const f32x2 xn = fma2(bf16x2_to_f32x2(xw[q]), rs2, nmurs2);
        const f32x2 g = mul2(d, gm[q]);
        s1v = add2(s1v, g);
        s2v = fma2(g, xn, s2v);
        ag[c][q] = fma2(d, xn, ag[c][q]);
        ab[c][q] = add2(ab[c][q], d);
      }
    }
    float s1a, s1b, s2a, s2b;
    upk2(s1v, s1a, s1b);
    upk2(s2v, s2a, s2b);
    const float s1 = warp_sum(s1a + s1b) * (1.f / H), s2 = warp_sum(s2a + s2b) * (1.f / H);
    const f32x2 ns2 = splat2(-s2), nb = splat2(-s1 * rs);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = (lane + 32 * c) * 8;
      const uint32_t xw[4] = {xr[c].x, xr[c].y, xr[c].z, xr[c].w};
      const uint32_t dw[4] = {dr[c].x, dr[c].y, dr[c].z, dr[c].w};
      const ulonglong2 g01 = *reinterpret_cast<const ulonglong2*>(gamma + col);
      const ulonglong2 g23 = *reinterpret_cast<const ulonglong2*>(gamma + col + 4);
      const f32x2 gm[4] = {g01.x, g01.y, g23.x, g23.y};
      f32x2 o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        f32x2 d = bf16x2_to_f32x2(dw[q]);
        if (out_thresh) {
          const uint32_t hb = drop_bits2(out_key, (uint32_t)row * H + col + 2 * q);
          d = mul2(d, pk2((hb & 0xFFFFu) >= out_thresh ? out_inv_keep : 0.f, (hb >> 16) >= out_thresh ? out_inv_keep : 0.f));
        }
        const f32x2 xn = fma2(bf16x2_to_f32x2(xw[q]), rs2, nmurs2);
        // rs * (g - s1 - xn * s2) with g = d * gamma
        o[q] = fma2(fma2(xn, ns2, mul2(d, gm[q])), rs2, nb);
      }
      uint4 pk;
      pk.x = f32x2_to_bf16x2(o[0]); pk.y = f32x2_to_bf16x2(o[1]); pk.z = f32x2_to_bf16x2(o[2]); pk.w = f32x2_to_bf16x2(o[3]);
      *reinterpret_cast<uint4*>(dx + (size_t)row * H + col) = pk;
      if (dx_branch) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t hb = drop_bits2(br_key, (uint32_t)row * H + col + 2 * q);
          o[q] = mul2(o[q], pk2((hb & 0xFFFFu) >= br_thresh ? br_inv_keep : 0.f, (hb >> 16) >= br_thresh ? br_inv_keep : 0.f));
        }
        pk.x = f32x2_to_bf16x2(o[0]); pk.y = f32x2_to_bf16x2(o[1]); pk.z = f32x2_to_bf16x2(o[2]); pk.w = f32x2_to_bf16x2(o[3]);
        *reinterpret_cast<uint4*>(dx_branch + (size_t)row * H + col) = pk;
      }
      if (dbias != nullptr) {
        // the bias gradient of the dense is the column sum of the STORED (bf16) branch gradient, as autograd forms it
        as[c][0] = add2(as[c][0], bf16x2_to_f32x2(pk.x)); as[c][1] = add2(as[c][1], bf16x2_to_f32x2(pk.y));
        as[c][2] = add2(as[c][2], bf16x2_to_f32x2(pk.z)); as[c][3] = add2(as[c][3], bf16x2_to_f32x2(pk.w));
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // fold the CTA's warps: element j of chunk c of a lane lives at [j][lane + 32 c] so a warp's adds hit 32 different banks
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = (2 * q) * (NCH * 32) + lane + 32 * c;
      float lo, hi;
      upk2(ag[c][q], lo, hi); atomicAdd(&sh[0][idx], lo); atomicAdd(&sh[0][idx + NCH * 32], hi);
      upk2(ab[c][q], lo, hi); atomicAdd(&sh[1][idx], lo); atomicAdd(&sh[1][idx + NCH * 32], hi);
      if (dbias != nullptr) { upk2(as[c][q], lo, hi); atomicAdd(&sh[2][idx], lo); atomicAdd(&sh[2][idx + NCH * 32], hi); }
    }
  __syncthreads();
  for (int col = threadIdx.x; col < H; col += LNF_WARPS * 32) {
    const int idx = (col & 7) * (NCH * 32) + (col >> 3);
    if (dgamma) atomicAdd(dgamma + col, sh[0][idx]);
    if (dbeta) atomicAdd(dbeta + col, sh[1][idx]);
    if (dbias) atomicAdd(dbias + col, sh[2][idx]);
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
  cudaError_t le = cudaSuccess;
#define SPMM_LN_FWD(N) le = launch_pdl(ln_fwd_kernel<N>, dim3(grid), dim3(LN_WARPS * 32), 0, st, (const __nv_bfloat16*)x, gamma, beta, (__nv_bfloat16*)y, mean, rstd, rows, H, eps, seed, th, ik, spmm_g_rng_salt)
  static const int one_row = [] { const char* e = getenv("SPMM_LN_FWD_ONE_ROW"); return e && e[0] == '1' ? 1 : 0; }();
  if (!one_row && nch <= 3 && H == nch * 256 && rows >= 2 * kNumSMs * LNP_WARPS) {
    const int ctas = 2 * kNumSMs;
#define SPMM_LN_FWDP(N) le = launch_pdl(ln_fwd_rows_kernel<N>, dim3(ctas), dim3(LNP_WARPS * 32), 0, st, (const __nv_bfloat16*)x, gamma, beta, (__nv_bfloat16*)y, mean, rstd, rows, eps, seed, th, ik, spmm_g_rng_salt)
    if (nch == 1) SPMM_LN_FWDP(1); else if (nch == 2) SPMM_LN_FWDP(2); else SPMM_LN_FWDP(3);
#undef SPMM_LN_FWDP
    return le != cudaSuccess ? (int)le : 0;
  }
  if (nch == 1) SPMM_LN_FWD(1); else if (nch == 2) SPMM_LN_FWD(2); else if (nch == 3) SPMM_LN_FWD(3); else SPMM_LN_FWD(4);
#undef SPMM_LN_FWD
  if (le != cudaSuccess) return (int)le;
  return 0;
}

extern "C" int spmm_layernorm_bwd(const void* dy, const void* x, const float* mean, const float* rstd,
                                  const float* gamma, void* dx, float* dgamma, float* dbeta, void* dx_branch,
                                  float* dbias, int rows, int H, float out_dropout_p, unsigned long long out_seed,
                                  float branch_dropout_p, unsigned long long branch_seed, float* workspace,
                                  void* stream) {
  (void)workspace;  // kept in the ABI for callers that pre-allocate scratch; the two-kernel scheme needs none
  SPMM_ARG(dy && x && mean && rstd && gamma && dx && rows > 0 && H > 0 && H % 8 == 0 && H <= 1024);
  SPMM_ARG((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)dx_branch | (uintptr_t)gamma) & 15) == 0);
  SPMM_ARG(!(dx_branch != nullptr && branch_dropout_p <= 0.f));
  uint32_t oth, bth; float oik, bik;
  drop_params(out_dropout_p, oth, oik);
  drop_params(branch_dropout_p, bth, bik);
  const int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  const int nch = (H + 255) / 256;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t le = cudaSuccess;
  static const int split = [] { const char* e = getenv("SPMM_LN_BWD_SPLIT"); return e && e[0] == '1' ? 1 : 0; }();
  if (!split && nch <= 3 && H == nch * 256 && (dgamma || dbeta || dbias)) {
    int ctas = (rows + LNF_WARPS - 1) / LNF_WARPS;
    if (ctas > kNumSMs) ctas = kNumSMs;
    static bool configured[4] = {false, false, false, false};
#define SPMM_LN_BWDF(N)                                                                                                  \
  do {                                                                                                                   \
    constexpr int ring_bytes = LNF_WARPS * LNF_STAGES * 2 * N * 32 * 16;                                                 \
    if (!configured[N]) {                                                                                                \
      le = cudaFuncSetAttribute(ln_bwd_fused_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, ring_bytes);        \
      configured[N] = le == cudaSuccess;                                                                                 \
    }                                                                                                                    \
    if (le == cudaSuccess)                                                                                               \
      le = launch_pdl(ln_bwd_fused_kernel<N>, dim3(ctas), dim3(LNF_WARPS * 32), ring_bytes, st,                          \
                      (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, mean, rstd, gamma, (__nv_bfloat16*)dx,          \
                      (__nv_bfloat16*)dx_branch, dgamma, dbeta, dbias, rows, out_seed, oth, oik, branch_seed, bth,       \
                      bik, spmm_g_rng_salt);                                                                             \
  } while (0)
    if (nch == 1) SPMM_LN_BWDF(1); else if (nch == 2) SPMM_LN_BWDF(2); else SPMM_LN_BWDF(3);
#undef SPMM_LN_BWDF
    return le != cudaSuccess ? (int)le : 0;
  }
#define SPMM_LN_BWD(N) le = launch_pdl(ln_bwd_dx_kernel<N>, dim3(grid), dim3(LN_WARPS * 32), 0, st, (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, mean, rstd, gamma, (__nv_bfloat16*)dx, (__nv_bfloat16*)dx_branch, rows, H, out_seed, oth, oik, branch_seed, bth, bik, spmm_g_rng_salt)
  if (nch == 1) SPMM_LN_BWD(1); else if (nch == 2) SPMM_LN_BWD(2); else if (nch == 3) SPMM_LN_BWD(3); else SPMM_LN_BWD(4);
#undef SPMM_LN_BWD
  if (le != cudaSuccess) return (int)le;
  if (dgamma || dbeta || dbias) {
    const int col_blocks = (H + 255) / 256;
    int row_blocks = (2 * kNumSMs + col_blocks - 1) / col_blocks;
    if (row_blocks > (rows + 31) / 32) row_blocks = (rows + 31) / 32;
    if (row_blocks < 1) row_blocks = 1;
    const int rpb = (rows + row_blocks - 1) / row_blocks;
    const __nv_bfloat16* branch = (const __nv_bfloat16*)(dx_branch ? dx_branch : dx);
    le = launch_pdl(ln_bwd_param_kernel, dim3(col_blocks, row_blocks), dim3(256), 0, st, (const __nv_bfloat16*)dy,
                    (const __nv_bfloat16*)x, mean, rstd, branch, dgamma, dbeta, dbias, rows, H, rpb, out_seed, oth, oik,
                    spmm_g_rng_salt);
    if (le != cudaSuccess) return (int)le;
  }
  return 0;
}
