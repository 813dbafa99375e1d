// Fused contrastive (SPC / ITC) head -- reference SPMM_models.py:92-131 (+ the in-batch sims of :157-158).
//
// The reference materialises 8 similarity matrices [B, B+Q] (113 MB at B=96, Q=36864) and ~30 elementwise
// kernels.  Here the two key sets ([own momentum feats | queue]) are streamed tile by tile; the 8 blocks are
// formed in registers and reduced immediately, never written to memory:
//   pass 1  per (row-chunk, key-split) CTA: online-softmax statistics of student and teacher rows
//   combine LSEs, loss_ita, d/d temp
//   pass 2  same tiling, G = dL/d sim recomputed from the LSEs, dF += G . keys accumulated in registers
//   finish  chain rule through F.normalize
// All arithmetic is fp32 (d/d temp is ill-conditioned in bf16).  Queues are key-major
// [Q][E] so a key is one contiguous 1 KB row.
#include "common.cuh"
#include "spmm_b200.h"

namespace spmm {

constexpr int E_ = 256;
constexpr int TK = 64;     // keys per tile
constexpr int RP = 32;     // (student, teacher) row pairs per CTA
constexpr int LDS_ = 260;  // padded smem row (floats): conflict-free LDS.128 across rows
constexpr int GLD = 65;
constexpr float NEG_BIG = -1e30f;

struct ItcArgs {
  const float* feats;  // [4][B][E] normalised: f_prop, f_text, m_prop, m_text
  const float* queue0; // text queue (key set 0)
  const float* queue1; // prop queue (key set 1)
  int B, Q, N, splits, tiles_per_split;
  const float* temp;  // device scalar (the clamped nn.Parameter)
  float alpha;
  float* part;     // [2][splits][2B][6]
  float* rowstat;  // [2][2B][4]: lse_s, lse_m, unused, unused
  float* sdiag;    // [2][2B]
  float* sim_i2t;  // [B][B]
  float* sim_t2i;  // [B][B]
  float* dF;       // [2][B][E]: d loss / d f_prop, d f_text
};

__device__ __forceinline__ const float* student_ptr(const ItcArgs& a, int ks, int r) {
  if (ks == 0) return a.feats + (size_t)r * E_;
  return r < a.B ? a.feats + (size_t)(a.B + r) * E_ : a.feats + (size_t)(r - a.B) * E_;
}
__device__ __forceinline__ const float* key_ptr(const ItcArgs& a, int ks, int j) {
  if (j < a.B) return a.feats + (size_t)((ks == 0 ? 3 : 2) * a.B + j) * E_;
  return (ks == 0 ? a.queue0 : a.queue1) + (size_t)(j - a.B) * E_;
}

struct Stat { float mx, sum, w; };
__device__ __forceinline__ void stat_add(Stat& st, float logit, float weight_val) {
  if (logit > st.mx) {
    const float sc = __expf(st.mx - logit);
    st.sum *= sc; st.w *= sc; st.mx = logit;
  }
  const float e = __expf(logit - st.mx);
  st.sum += e; st.w += e * weight_val;
}
__device__ __forceinline__ void stat_merge(Stat& a, const Stat& b) {
  const float m = fmaxf(a.mx, b.mx);
  const float sa = __expf(a.mx - m), sb = __expf(b.mx - m);
  a.sum = a.sum * sa + b.sum * sb; a.w = a.w * sa + b.w * sb; a.mx = m;
}

// Computes the 2x2x4 micro-tile of logits for this thread: row pairs {ty, ty+16}, keys {tx + 16*k}.
__device__ __forceinline__ void micro_dots(const float* sq, const float* sk, int ty, int tx, float (&s)[2][4],
                                           float (&m)[2][4]) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) s[i][k] = m[i][k] = 0.f;
  const float* qs0 = sq + (ty)*LDS_;
  const float* qs1 = sq + (ty + 16) * LDS_;
  const float* qm0 = sq + (RP + ty) * LDS_;
  const float* qm1 = sq + (RP + ty + 16) * LDS_;
#pragma unroll 4
  for (int e = 0; e < E_; e += 4) {
    const float4 a0 = *reinterpret_cast<const float4*>(qs0 + e), a1 = *reinterpret_cast<const float4*>(qs1 + e);
    const float4 b0 = *reinterpret_cast<const float4*>(qm0 + e), b1 = *reinterpret_cast<const float4*>(qm1 + e);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 kv = *reinterpret_cast<const float4*>(sk + (tx + 16 * k) * LDS_ + e);
      s[0][k] += a0.x * kv.x + a0.y * kv.y + a0.z * kv.z + a0.w * kv.w;
      s[1][k] += a1.x * kv.x + a1.y * kv.y + a1.z * kv.z + a1.w * kv.w;
      m[0][k] += b0.x * kv.x + b0.y * kv.y + b0.z * kv.z + b0.w * kv.w;
      m[1][k] += b1.x * kv.x + b1.y * kv.y + b1.z * kv.z + b1.w * kv.w;
    }
  }
}

template <bool GRAD>
__global__ void __launch_bounds__(256, 1) itc_pass_kernel(const ItcArgs a) {
  extern __shared__ float smem[];
  float* sq = smem;                  // [2*RP][LDS_]  students then teachers
  float* sk = sq + 2 * RP * LDS_;    // [TK][LDS_]
  float* sg = sk + TK * LDS_;        // [RP][GLD]   (GRAD only)
  const int n_chunks = (2 * a.B + RP - 1) / RP;
  const int ks = blockIdx.x / n_chunks, chunk = blockIdx.x % n_chunks;
  const int split = blockIdx.y;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int r_base = chunk * RP;

  // resident query rows for this CTA
  for (int i = tid; i < 2 * RP * (E_ / 4); i += 256) {
    const int row = i / (E_ / 4), c4 = i % (E_ / 4);
    const int r = r_base + (row % RP);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < 2 * a.B) {
      const float* src = student_ptr(a, ks, r) + (row >= RP ? (size_t)2 * a.B * E_ : 0);
      v = *reinterpret_cast<const float4*>(src + c4 * 4);
    }
    *reinterpret_cast<float4*>(sq + row * LDS_ + c4 * 4) = v;
  }

  Stat ss[2], sm[2];
  float lse_s[2], lse_m[2];
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    ss[i] = {NEG_BIG, 0.f, 0.f};
    sm[i] = {NEG_BIG, 0.f, 0.f};
    const int r = r_base + ty + 16 * i;
    lse_s[i] = lse_m[i] = 0.f;
    if (GRAD && r < 2 * a.B) {
      lse_s[i] = a.rowstat[((size_t)ks * 2 * a.B + r) * 4 + 0];
      lse_m[i] = a.rowstat[((size_t)ks * 2 * a.B + r) * 4 + 1];
    }
  }
  if (GRAD) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  }
  const float inv_temp = 1.f / __ldg(a.temp);
  const float gscale = inv_temp / (2.f * a.B);
  const int tile0 = split * a.tiles_per_split;
  const int n_tiles = (a.N + TK - 1) / TK;
  const int tile1 = min(n_tiles, tile0 + a.tiles_per_split);

  for (int tile = tile0; tile < tile1; ++tile) {
    const int j0 = tile * TK;
    __syncthreads();  // previous tile fully consumed (and sq visible on first iteration)
    for (int i = tid; i < TK * (E_ / 4); i += 256) {
      const int row = i / (E_ / 4), c4 = i % (E_ / 4);
      const int j = j0 + row;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < a.N) v = __ldg(reinterpret_cast<const float4*>(key_ptr(a, ks, j) + c4 * 4));
      *reinterpret_cast<float4*>(sk + row * LDS_ + c4 * 4) = v;
    }
    __syncthreads();
    float s[2][4], m[2][4];
    micro_dots(sq, sk, ty, tx, s, m);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = r_base + ty + 16 * i;
      const int b = r < a.B ? r : r - a.B;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = j0 + tx + 16 * k;
        const float sv = s[i][k] * inv_temp, mv = m[i][k] * inv_temp;
        const bool ok = (j < a.N) && (r < 2 * a.B);
        if (!GRAD) {
          if (ok) {
            stat_add(ss[i], sv, sv);   // sum exp(s), sum exp(s) * s
            stat_add(sm[i], mv, sv);   // sum exp(m), sum exp(m) * s
            if (j == b) a.sdiag[(size_t)ks * 2 * a.B + r] = sv;
            if (j < a.B && r < a.B) (ks == 0 ? a.sim_i2t : a.sim_t2i)[(size_t)r * a.B + j] = sv;
          }
        } else {
          float g = 0.f;
          if (ok) {
            g = __expf(sv - lse_s[i]) - a.alpha * __expf(mv - lse_m[i]) - ((j == b) ? (1.f - a.alpha) : 0.f);
            g *= gscale;
          }
          sg[(ty + 16 * i) * GLD + tx + 16 * k] = g;
        }
      }
    }
    if (GRAD) {
      __syncthreads();
      // dF[r][e] += sum_j G[r][j] * key[j][e];  thread: rows rg*8..+7, columns c4*4..+3
      const int c4 = tid & 63, rg = tid >> 6;
#pragma unroll 4
      for (int j = 0; j < TK; ++j) {
        const float4 kv = *reinterpret_cast<const float4*>(sk + j * LDS_ + c4 * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float g = sg[(rg * 8 + i) * GLD + j];
          acc[i][0] += g * kv.x; acc[i][1] += g * kv.y; acc[i][2] += g * kv.z; acc[i][3] += g * kv.w;
        }
      }
    }
  }

  if (!GRAD) {
    // merge the 16 threads (tx) that share a row
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        Stat t;
        t.mx = __shfl_xor_sync(0xffffffffu, ss[i].mx, o); t.sum = __shfl_xor_sync(0xffffffffu, ss[i].sum, o);
        t.w = __shfl_xor_sync(0xffffffffu, ss[i].w, o);
        stat_merge(ss[i], t);
        t.mx = __shfl_xor_sync(0xffffffffu, sm[i].mx, o); t.sum = __shfl_xor_sync(0xffffffffu, sm[i].sum, o);
        t.w = __shfl_xor_sync(0xffffffffu, sm[i].w, o);
        stat_merge(sm[i], t);
      }
      const int r = r_base + ty + 16 * i;
      if (tx == 0 && r < 2 * a.B) {
        float* p = a.part + (((size_t)ks * a.splits + split) * 2 * a.B + r) * 6;
        p[0] = ss[i].mx; p[1] = ss[i].sum; p[2] = ss[i].w; p[3] = sm[i].mx; p[4] = sm[i].sum; p[5] = sm[i].w;
      }
    }
  } else {
    const int c4 = tid & 63, rg = tid >> 6;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = r_base + rg * 8 + i;
      if (r >= 2 * a.B) continue;
      // key set 0: rows [0,B) -> f_prop, [B,2B) -> f_text;  key set 1: rows [0,B) -> f_text, [B,2B) -> f_prop
      const int b = r < a.B ? r : r - a.B;
      const int which = (ks == 0) ? (r < a.B ? 0 : 1) : (r < a.B ? 1 : 0);
      float* d = a.dF + ((size_t)which * a.B + b) * E_ + c4 * 4;
      atomicAdd(d + 0, acc[i][0]); atomicAdd(d + 1, acc[i][1]); atomicAdd(d + 2, acc[i][2]); atomicAdd(d + 3, acc[i][3]);
    }
  }
}

// F.normalize(z, dim=-1) for the four feature matrices (eps 1e-12); one warp per row
__global__ void itc_normalize_kernel(const float* z0, const float* z1, const float* z2, const float* z3, float* feats,
                                     float* norms, float* out_m_prop, float* out_m_text, int B) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= 4 * B) return;
  const int which = row / B, b = row % B;
  const float* z = (which == 0 ? z0 : which == 1 ? z1 : which == 2 ? z2 : z3) + (size_t)b * E_;
  float v[8], ssq = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = z[lane + 32 * i]; ssq += v[i] * v[i]; }
  const float n = fmaxf(sqrtf(warp_sum(ssq)), 1e-12f);
  if (lane == 0) norms[row] = n;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float f = v[i] / n;
    feats[(size_t)row * E_ + lane + 32 * i] = f;
    if (which == 2) out_m_prop[(size_t)b * E_ + lane + 32 * i] = f;
    if (which == 3) out_m_text[(size_t)b * E_ + lane + 32 * i] = f;
  }
}

// one thread per (key set, row): merge split partials -> LSEs, row loss, row d/dtemp; then block-reduce
__global__ void itc_combine_kernel(ItcArgs a, float* loss, float* dtemp, float* nan_flag) {
  __shared__ float sh[32];
  const int total = 4 * a.B;
  float l = 0.f, dt = 0.f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int ks = i / (2 * a.B), r = i % (2 * a.B);
    Stat s = {NEG_BIG, 0.f, 0.f}, m = {NEG_BIG, 0.f, 0.f};
    for (int sp = 0; sp < a.splits; ++sp) {
      const float* p = a.part + (((size_t)ks * a.splits + sp) * 2 * a.B + r) * 6;
      Stat t1 = {p[0], p[1], p[2]}, t2 = {p[3], p[4], p[5]};
      stat_merge(s, t1);
      stat_merge(m, t2);
    }
    const float lse_s = s.mx + __logf(s.sum), lse_m = m.mx + __logf(m.sum);
    a.rowstat[(size_t)i * 4 + 0] = lse_s;
    a.rowstat[(size_t)i * 4 + 1] = lse_m;
    const float sd = a.sdiag[i];
    const float teacher_dot = m.w / m.sum;   // sum_j softmax(m)_j * s_j
    const float student_dot = s.w / s.sum;   // sum_j softmax(s)_j * s_j
    l += lse_s - a.alpha * teacher_dot - (1.f - a.alpha) * sd;
    dt += student_dot - a.alpha * teacher_dot - (1.f - a.alpha) * sd;
  }
  l = block_sum(l, sh);
  dt = block_sum(dt, sh);
  if (threadIdx.x == 0) {
    const float L = l / (2.f * a.B);                 // (sum of 4 row-means) / 2
    *loss = L;
    *dtemp = -dt / (__ldg(a.temp) * 2.f * a.B);         // d s / d temp = -s / temp
    if (nan_flag) *nan_flag = (L != L) ? 1.f : 0.f;  // reference NaN guard, SPMM_models.py:132
  }
}

// dz = (dF - f (f . dF)) / ||z||   (backward of F.normalize); one warp per row of the two student matrices
__global__ void itc_finish_kernel(const float* feats, const float* norms, const float* dF, float* dz_prop,
                                  float* dz_text, int B) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= 2 * B) return;
  const float* f = feats + (size_t)row * E_;
  const float* g = dF + (size_t)row * E_;
  float fv[8], gv[8], dot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { fv[i] = f[lane + 32 * i]; gv[i] = g[lane + 32 * i]; dot += fv[i] * gv[i]; }
  dot = warp_sum(dot);
  const float inv = 1.f / norms[row];
  float* out = (row < B ? dz_prop + (size_t)row * E_ : dz_text + (size_t)(row - B) * E_);
#pragma unroll
  for (int i = 0; i < 8; ++i) out[lane + 32 * i] = (gv[i] - fv[i] * dot) * inv;
}

static inline int itc_splits(int B, int N) {
  const int n_chunks = (2 * B + RP - 1) / RP;
  const int n_tiles = (N + TK - 1) / TK;
  int splits = (2 * kNumSMs + 2 * n_chunks - 1) / (2 * n_chunks);
  if (splits > n_tiles) splits = n_tiles;
  if (splits < 1) splits = 1;
  return splits;
}

}  // namespace spmm
using namespace spmm;

extern "C" int64_t spmm_itc_workspace_bytes(int B, int E, int Q) {
  if (E != E_) return -1;
  const int splits = itc_splits(B, B + Q);
  int64_t floats = (int64_t)4 * B * E_      // feats
                   + 4 * B                  // norms
                   + (int64_t)2 * splits * 2 * B * 6   // partials
                   + (int64_t)2 * 2 * B * 4             // rowstat
                   + 2 * 2 * B                          // sdiag
                   + (int64_t)2 * B * E_;               // dF
  return floats * 4 + 256;
}

extern "C" int spmm_itc_fwd_bwd(const float* z_prop, const float* z_text, const float* z_prop_m, const float* z_text_m,
                                const float* prop_queue, const float* text_queue, const float* temp,
                                float alpha, int B, int E, int Q, float* loss, float* dz_prop, float* dz_text,
                                float* dtemp, float* sim_i2t, float* sim_t2i, float* feat_prop_m, float* feat_text_m,
                                float* nan_flag, void* workspace, int64_t workspace_bytes, void* stream) {
  SPMM_ARG(z_prop && z_text && z_prop_m && z_text_m && prop_queue && text_queue && temp);
  SPMM_ARG(loss && dz_prop && dz_text && dtemp && sim_i2t && sim_t2i && feat_prop_m && feat_text_m && workspace);
  SPMM_ARG(E == E_ && B >= 1 && Q >= 0);
  SPMM_ARG(workspace_bytes >= spmm_itc_workspace_bytes(B, E, Q));
  cudaStream_t st = (cudaStream_t)stream;
  ItcArgs a{};
  float* w = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  const int N = B + Q;
  const int splits = itc_splits(B, N);
  const int n_tiles = (N + TK - 1) / TK;
  float* feats = w; w += (size_t)4 * B * E_;
  float* norms = w; w += 4 * B;
  a.part = w; w += (size_t)2 * splits * 2 * B * 6;
  a.rowstat = w; w += (size_t)2 * 2 * B * 4;
  a.sdiag = w; w += 2 * 2 * B;
  a.dF = w;
  a.feats = feats; a.queue0 = text_queue; a.queue1 = prop_queue;
  a.B = B; a.Q = Q; a.N = N; a.splits = splits; a.tiles_per_split = (n_tiles + splits - 1) / splits;
  a.alpha = alpha; a.sim_i2t = sim_i2t; a.sim_t2i = sim_t2i;
  a.temp = temp;

  cudaError_t e = cudaMemsetAsync(a.dF, 0, (size_t)2 * B * E_ * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  itc_normalize_kernel<<<(4 * B + 7) / 8, 256, 0, st>>>(z_prop, z_text, z_prop_m, z_text_m, feats, norms, feat_prop_m,
                                                        feat_text_m, B);
  SPMM_CHECK_LAUNCH();
  const int n_chunks = (2 * B + RP - 1) / RP;
  const size_t smem1 = (size_t)(2 * RP + TK) * LDS_ * sizeof(float);
  const size_t smem2 = smem1 + (size_t)RP * GLD * sizeof(float);
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(itc_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    cudaFuncSetAttribute(itc_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    configured = true;
  }
  dim3 grid(2 * n_chunks, splits);
  itc_pass_kernel<false><<<grid, 256, smem1, st>>>(a);
  SPMM_CHECK_LAUNCH();
  itc_combine_kernel<<<1, 256, 0, st>>>(a, loss, dtemp, nan_flag);
  SPMM_CHECK_LAUNCH();
  itc_pass_kernel<true><<<grid, 256, smem2, st>>>(a);
  SPMM_CHECK_LAUNCH();
  itc_finish_kernel<<<(2 * B + 7) / 8, 256, 0, st>>>(feats, norms, a.dF, dz_prop, dz_text, B);
  SPMM_CHECK_LAUNCH();
  return 0;
}
