// Fused attention core for Tq,Tk <= 128, head_dim 64 (reference xbert.py:305-354: QK^T/8 + additive mask ->
// softmax -> dropout -> PV, plus the head split/merge permutes :265-268,352-354 folded into the addressing).
// One CTA per (batch, head): the whole K/V of a head fits in shared memory, so softmax is single pass (no online
// rescale) and the [B,12,Tq,Tk] probability tensor of the reference is never materialised; backward recomputes P
// from the saved log-sum-exp.  Masks are generated in-kernel from kv_len (pad) and the causal flag.
// These are the round-0 warp-level mma.sync kernels.  The product path is attention_tc.cu (tcgen05 forward and backward);
// this file keeps the C entry points, argument checks and - behind SPMM_ATTN_LEGACY=1, for A/B measurements only - the
// old kernels.
#include <cstdlib>

#include "common.cuh"
#include "spmm_b200.h"

namespace spmm {

constexpr int HD = 64;  // head dim

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// [rows][64] bf16 tile, 16-byte chunks XOR-swizzled by (row & 7)
__device__ __forceinline__ uint32_t tile64_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }
// [rows][KT] bf16 tile (P / dS): row stride KT*2 bytes
template <int KT>
__device__ __forceinline__ uint32_t tileP_off(int row, int chunk) { return (uint32_t)(row * (KT * 2) + ((chunk ^ (row & 7)) << 4)); }

// global [T rows][ld] (head slice of 64 columns) -> swizzled smem tile; rows >= T zero-filled up to rows_pad
__device__ __forceinline__ void load_tile64(uint8_t* smem, const __nv_bfloat16* g, int ld, int T, int rows_pad) {
  for (int i = threadIdx.x; i < rows_pad * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < T) v = *reinterpret_cast<const uint4*>(g + (size_t)r * ld + c * 8);
    *reinterpret_cast<uint4*>(smem + tile64_off(r, c)) = v;
  }
}

__device__ __forceinline__ bool attn_keep(unsigned long long seed, int bh, int i, int j, uint32_t thresh16) {
  const unsigned long long e = ((unsigned long long)bh << 14 | (unsigned long long)i << 7 | (unsigned long long)j);
  return keep16(seed, e, thresh16);
}

// S[16 x KT] = Q[16 rows of this warp] . K^T for the warp's 16 query rows
template <int KT>
__device__ __forceinline__ void qk_scores(uint32_t sQ, uint32_t sK, int q0, int lane, float (&s)[KT / 8][4]) {
#pragma unroll
  for (int n = 0; n < KT / 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) {
    uint32_t a0, a1, a2, a3;
    ldsm_x4(sQ + tile64_off(q0 + (lane & 15), kk * 2 + (lane >> 4)), a0, a1, a2, a3);
#pragma unroll
    for (int np = 0; np < KT / 16; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4(sK + tile64_off(np * 16 + (lane & 7) + ((lane >> 4) << 3), kk * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
      mma16816(s[2 * np], a0, a1, a2, a3, b0, b1);
      mma16816(s[2 * np + 1], a0, a1, a2, a3, b2, b3);
    }
  }
}

template <int QT, int KT>
__global__ void __launch_bounds__(QT * 2)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k, int ldk,
                const __nv_bfloat16* __restrict__ v, int ldv, __nv_bfloat16* __restrict__ o, int ldo,
                float* __restrict__ lse, int heads, int Tq, int Tk, const int* __restrict__ kv_len, int causal,
                int kv_bstride, float scale, unsigned long long seed, uint32_t thresh16, float inv_keep,
                const unsigned long long* salt) {
  if (thresh16) seed = salted(seed, salt);
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* pQ = smem;
  uint8_t* pK = pQ + QT * 128;
  uint8_t* pV = pK + KT * 128;
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Tq_pad = (Tq + 15) & ~15;
  load_tile64(pQ, q + (size_t)b * Tq * ldq + h * HD, ldq, Tq, Tq_pad);
  load_tile64(pK, k + (size_t)b * kv_bstride * ldk + h * HD, ldk, Tk, KT);
  load_tile64(pV, v + (size_t)b * kv_bstride * ldv + h * HD, ldv, Tk, KT);
  __syncthreads();
  const int q0 = warp * 16;
  if (q0 >= Tq) return;
  const uint32_t sQ = smem_u32(pQ), sK = smem_u32(pK), sV = smem_u32(pV);
  const int klen = kv_len ? min(kv_len[b], Tk) : Tk;
  const int g = lane >> 2, t = lane & 3;

  float s[KT / 8][4];
  qk_scores<KT>(sQ, sK, q0, lane, s);
  // mask + softmax: thread owns rows q0+g (c0,c1) and q0+g+8 (c2,c3), keys 8n+2t, 8n+2t+1
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int n = 0; n < KT / 8; ++n) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = n * 8 + 2 * t + (c & 1), i = q0 + g + ((c >> 1) << 3);
      const bool ok = (j < klen) && (!causal || j <= i);
      s[n][c] = ok ? s[n][c] * scale : -INFINITY;
      mx[c >> 1] = fmaxf(mx[c >> 1], s[n][c]);
    }
  }
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    if (mx[r] == -INFINITY) mx[r] = 0.f;  // fully masked row (kv_len == 0): output zeros
  }
#pragma unroll
  for (int n = 0; n < KT / 8; ++n) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      s[n][c] = __expf(s[n][c] - mx[c >> 1]);
      sum[c >> 1] += s[n][c];
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
    sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
  }
  const float inv[2] = {sum[0] > 0.f ? 1.f / sum[0] : 0.f, sum[1] > 0.f ? 1.f / sum[1] : 0.f};
  const int bh = b * heads + h;
  if (t == 0 && lse != nullptr) {
    if (q0 + g < Tq) lse[(size_t)bh * Tq + q0 + g] = mx[0] + __logf(sum[0]);
    if (q0 + g + 8 < Tq) lse[(size_t)bh * Tq + q0 + g + 8] = mx[1] + __logf(sum[1]);
  }
  // O = P V
  float acc[HD / 8][4];
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < KT / 16; ++kk) {
    float p[2][4];
#pragma unroll
    for (int half = 0; half < 2; ++half)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float pv = s[2 * kk + half][c] * inv[c >> 1];
        if (thresh16) {
          const int j = (2 * kk + half) * 8 + 2 * t + (c & 1), i = q0 + g + ((c >> 1) << 3);
          pv = attn_keep(seed, bh, i, j, thresh16) ? pv * inv_keep : 0.f;
        }
        p[half][c] = pv;
      }
    const uint32_t a0 = pack_bf16x2(p[0][0], p[0][1]), a1 = pack_bf16x2(p[0][2], p[0][3]);
    const uint32_t a2 = pack_bf16x2(p[1][0], p[1][1]), a3 = pack_bf16x2(p[1][2], p[1][3]);
#pragma unroll
    for (int np = 0; np < HD / 16; ++np) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(sV + tile64_off(kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), np * 2 + (lane >> 4)), b0, b1, b2, b3);
      mma16816(acc[2 * np], a0, a1, a2, a3, b0, b1);
      mma16816(acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
    }
  }
  __nv_bfloat16* og = o + (size_t)b * Tq * ldo + h * HD;
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) {
    const int col = n * 8 + 2 * t;
    if (q0 + g < Tq) *reinterpret_cast<uint32_t*>(og + (size_t)(q0 + g) * ldo + col) = pack_bf16x2(acc[n][0], acc[n][1]);
    if (q0 + g + 8 < Tq) *reinterpret_cast<uint32_t*>(og + (size_t)(q0 + g + 8) * ldo + col) = pack_bf16x2(acc[n][2], acc[n][3]);
  }
}

// Backward.  Phase 1 (warp = 16 query rows): recompute P, dP = dO V^T, dS = P o (dP - D), dQ = dS K * scale;
// P (after dropout) and dS are parked in smem as bf16.  Phase 2 (warp = 16 keys): dV = P^T dO, dK = dS^T Q * scale.
template <int QT, int KT>
__global__ void __launch_bounds__(QT * 2)
attn_bwd_kernel(const __nv_bfloat16* __restrict__ dO, int lddo, const __nv_bfloat16* __restrict__ q, int ldq,
                const __nv_bfloat16* __restrict__ k, int ldk, const __nv_bfloat16* __restrict__ v, int ldv,
                const __nv_bfloat16* __restrict__ o, int ldo, const float* __restrict__ lse,
                __nv_bfloat16* __restrict__ dq, int lddq, __nv_bfloat16* __restrict__ dk, int lddk,
                __nv_bfloat16* __restrict__ dv, int lddv, int heads, int Tq, int Tk, const int* __restrict__ kv_len,
                int causal, float scale, unsigned long long seed, uint32_t thresh16, float inv_keep,
                const unsigned long long* salt) {
  pdl_trigger();
  pdl_wait();
  if (thresh16) seed = salted(seed, salt);
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* pQ = smem;               // [QT][64]
  uint8_t* pdO = pQ + QT * 128;     // [QT][64]
  uint8_t* pK = pdO + QT * 128;     // [KT][64]
  uint8_t* pV = pK + KT * 128;      // [KT][64]
  uint8_t* pP = pV + KT * 128;      // [QT][KT] bf16 (dropped P)
  uint8_t* pdS = pP + QT * KT * 2;  // [QT][KT] bf16
  float* sD = reinterpret_cast<float*>(pdS + QT * KT * 2);  // [QT]
  const int nwarps = blockDim.x >> 5;
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Tq_pad = (Tq + 15) & ~15;
  const size_t qrow0 = (size_t)b * Tq, krow0 = (size_t)b * Tk;
  load_tile64(pQ, q + qrow0 * ldq + h * HD, ldq, Tq, Tq_pad);
  load_tile64(pdO, dO + qrow0 * lddo + h * HD, lddo, Tq, Tq_pad);
  load_tile64(pK, k + krow0 * ldk + h * HD, ldk, Tk, KT);
  load_tile64(pV, v + krow0 * ldv + h * HD, ldv, Tk, KT);
  __syncthreads();
  const uint32_t sQ = smem_u32(pQ), sdO = smem_u32(pdO), sK = smem_u32(pK), sV = smem_u32(pV), sP = smem_u32(pP),
                 sdS = smem_u32(pdS);
  const int klen = kv_len ? min(kv_len[b], Tk) : Tk;
  const int g = lane >> 2, t = lane & 3;
  const int bh = b * heads + h;
  const int q0 = warp * 16;

  if (q0 < Tq) {
    float s[KT / 8][4];
    qk_scores<KT>(sQ, sK, q0, lane, s);
    float dp[KT / 8][4];
    qk_scores<KT>(sdO, sV, q0, lane, dp);  // dP = dO . V^T has the same operand structure
    const int i0 = q0 + g, i1 = q0 + g + 8;
    const float l0 = i0 < Tq ? lse[(size_t)bh * Tq + i0] : 0.f, l1 = i1 < Tq ? lse[(size_t)bh * Tq + i1] : 0.f;
    // pass A: p (in s[]), dropout keep factor folded into dp[], D_i = sum_j (p keep)_ij dP_ij accumulated on the fly
    // (O = Pd V  =>  sum_d dO_id O_id = sum_j Pd_ij (dO_i . V_j): no second read of O, no serial global-load loop)
    float d0 = 0.f, d1 = 0.f;
    uint32_t kept[KT / 64] = {};   // dropout keep bits of this thread's KT/2 elements (one hash per element, reused below)
#pragma unroll
    for (int n = 0; n < KT / 8; ++n) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = n * 8 + 2 * t + (c & 1), i = (c >> 1) ? i1 : i0;
        const bool ok = (i < Tq) && (j < klen) && (!causal || j <= i);
        const float p = ok ? __expf(s[n][c] * scale - ((c >> 1) ? l1 : l0)) : 0.f;
        float keep = 1.f;
        if (thresh16) {
          const bool kp = attn_keep(seed, bh, i, j, thresh16);
          keep = kp ? inv_keep : 0.f;
          kept[(n * 4 + c) >> 5] |= (kp ? 1u : 0u) << ((n * 4 + c) & 31);
        }
        s[n][c] = p;
        dp[n][c] *= keep;
        if (c >> 1) d1 += p * dp[n][c]; else d0 += p * dp[n][c];
      }
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
#pragma unroll
    for (int n = 0; n < KT / 8; ++n) {
      float pd[4], ds[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float p = s[n][c];
        // dp[] already carries keep (= mask / keep_prob): Pd = p * keep needs the factor once more
        float keep = 1.f;
        if (thresh16) keep = ((kept[(n * 4 + c) >> 5] >> ((n * 4 + c) & 31)) & 1u) ? inv_keep : 0.f;
        pd[c] = p * keep;
        ds[c] = p * (dp[n][c] - ((c >> 1) ? d1 : d0));
      }
      // park bf16 P / dS: element (row, key) -> tile128 chunk key/8, within-chunk offset (key%8)*2 bytes
      const int key = n * 8 + 2 * t;
      const uint32_t off0 = tileP_off<KT>(i0, key >> 3) + ((key & 7) << 1);
      const uint32_t off1 = tileP_off<KT>(i1, key >> 3) + ((key & 7) << 1);
      *reinterpret_cast<uint32_t*>(pP + off0) = pack_bf16x2(pd[0], pd[1]);
      *reinterpret_cast<uint32_t*>(pP + off1) = pack_bf16x2(pd[2], pd[3]);
      *reinterpret_cast<uint32_t*>(pdS + off0) = pack_bf16x2(ds[0], ds[1]);
      *reinterpret_cast<uint32_t*>(pdS + off1) = pack_bf16x2(ds[2], ds[3]);
      s[n][0] = ds[0]; s[n][1] = ds[1]; s[n][2] = ds[2]; s[n][3] = ds[3];  // keep dS for dQ
    }
    // dQ = dS . K * scale
    float acc[HD / 8][4];
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < KT / 16; ++kk) {
      const uint32_t a0 = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]), a1 = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      const uint32_t a2 = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]), a3 = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < HD / 16; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sK + tile64_off(kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), np * 2 + (lane >> 4)), b0, b1, b2, b3);
        mma16816(acc[2 * np], a0, a1, a2, a3, b0, b1);
        mma16816(acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    __nv_bfloat16* dqg = dq + qrow0 * lddq + h * HD;
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) {
      const int col = n * 8 + 2 * t;
      if (i0 < Tq) *reinterpret_cast<uint32_t*>(dqg + (size_t)i0 * lddq + col) = pack_bf16x2(acc[n][0] * scale, acc[n][1] * scale);
      if (i1 < Tq) *reinterpret_cast<uint32_t*>(dqg + (size_t)i1 * lddq + col) = pack_bf16x2(acc[n][2] * scale, acc[n][3] * scale);
    }
  }
  __syncthreads();
  // Phase 2: this warp owns keys [k0, k0+16)
  const int k0 = warp * 16;
  if (k0 < Tk) {
    float av[HD / 8][4], ak[HD / 8][4];
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) {
      av[n][0] = av[n][1] = av[n][2] = av[n][3] = 0.f;
      ak[n][0] = ak[n][1] = ak[n][2] = ak[n][3] = 0.f;
    }
    for (int qq = 0; qq < Tq_pad; qq += 16) {
      // A = P^T / dS^T fragments: stored [q][key]; transposed ldmatrix
      uint32_t p0, p1, p2, p3, e0, e1, e2, e3;
      const int srow = qq + (lane & 7) + ((lane >> 4) << 3), schunk = (k0 >> 3) + ((lane >> 3) & 1);
      ldsm_x4_t(sP + tileP_off<KT>(srow, schunk), p0, p1, p2, p3);
      ldsm_x4_t(sdS + tileP_off<KT>(srow, schunk), e0, e1, e2, e3);
#pragma unroll
      for (int np = 0; np < HD / 16; ++np) {
        uint32_t b0, b1, b2, b3;
        const uint32_t boff = tile64_off(qq + (lane & 7) + (((lane >> 3) & 1) << 3), np * 2 + (lane >> 4));
        ldsm_x4_t(sdO + boff, b0, b1, b2, b3);
        mma16816(av[2 * np], p0, p1, p2, p3, b0, b1);
        mma16816(av[2 * np + 1], p0, p1, p2, p3, b2, b3);
        ldsm_x4_t(sQ + boff, b0, b1, b2, b3);
        mma16816(ak[2 * np], e0, e1, e2, e3, b0, b1);
        mma16816(ak[2 * np + 1], e0, e1, e2, e3, b2, b3);
      }
    }
    __nv_bfloat16* dvg = dv + krow0 * lddv + h * HD;
    __nv_bfloat16* dkg = dk + krow0 * lddk + h * HD;
    const int j0 = k0 + g, j1 = k0 + g + 8;
#pragma unroll
    for (int n = 0; n < HD / 8; ++n) {
      const int col = n * 8 + 2 * t;
      if (j0 < Tk) {
        *reinterpret_cast<uint32_t*>(dvg + (size_t)j0 * lddv + col) = pack_bf16x2(av[n][0], av[n][1]);
        *reinterpret_cast<uint32_t*>(dkg + (size_t)j0 * lddk + col) = pack_bf16x2(ak[n][0] * scale, ak[n][1] * scale);
      }
      if (j1 < Tk) {
        *reinterpret_cast<uint32_t*>(dvg + (size_t)j1 * lddv + col) = pack_bf16x2(av[n][2], av[n][3]);
        *reinterpret_cast<uint32_t*>(dkg + (size_t)j1 * lddk + col) = pack_bf16x2(ak[n][2] * scale, ak[n][3] * scale);
      }
    }
  }
}

int attn_fwd_tc_launch(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, float* lse,
                       int batch, int heads, int Tq, int Tk, const int* kv_len, int causal, int kv_bstride, float scale,
                       uint32_t thresh16, float inv_keep, unsigned long long seed, cudaStream_t st);   // attention_tc.cu
void attn_set_trace(void* p);
int attn_bwd_tc_launch(const void* d_o, int lddo, const void* q, int ldq, const void* k, int ldk, const void* v, int ldv,
                       const float* lse, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv, int batch, int heads,
                       int Tq, int Tk, const int* kv_len, int causal, float scale, uint32_t thresh16, float inv_keep,
                       unsigned long long seed, cudaStream_t st);   // attention_tc.cu

static inline void attn_drop(float p, uint32_t& th, float& ik) {
  th = p > 0.f ? (uint32_t)(p * 65536.f + 0.5f) : 0u;
  ik = p > 0.f ? 1.f / (1.f - p) : 1.f;
}

}  // namespace spmm
using namespace spmm;

extern "C" int spmm_attn_debug_trace(void* buf) {   /* 32 x u64 per CTA; NULL = off */
  attn_set_trace(buf);
  return 0;
}

extern "C" int spmm_attn_fwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo,
                             float* lse, int batch, int heads, int Tq, int Tk, const int* kv_len, int causal,
                             int kv_batch_stride_rows, float scale, float dropout_p, unsigned long long seed,
                             void* stream) {
  SPMM_ARG(q && k && v && o && batch > 0 && heads > 0 && Tq > 0 && Tk > 0 && Tq <= 128 && Tk <= 128);
  SPMM_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0);
  SPMM_ARG((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) & 15) == 0 && ((uintptr_t)o & 3) == 0);
  uint32_t th; float ik;
  attn_drop(dropout_p, th, ik);
  cudaStream_t st = (cudaStream_t)stream;
  SPMM_ARG(ldo % 8 == 0 && ((uintptr_t)o & 15) == 0);
  if (!getenv("SPMM_ATTN_LEGACY"))
    return attn_fwd_tc_launch(q, ldq, k, ldk, v, ldv, o, ldo, lse, batch, heads, Tq, Tk, kv_len, causal, kv_batch_stride_rows,
                              scale, th, ik, seed, st);
  const int KT = Tk <= 64 ? 64 : 128, QT = Tq <= 64 ? 64 : 128;
  dim3 grid(heads, batch);
#define SPMM_ATTN_FWD(QTV, KTV)                                                                                      \
  {                                                                                                                  \
    constexpr int kSmem = QTV * 128 + 2 * KTV * 128;                                                                 \
    static bool cfg = false;                                                                                         \
    if (!cfg) { cudaFuncSetAttribute(attn_fwd_kernel<QTV, KTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem); cfg = true; } \
    attn_fwd_kernel<QTV, KTV><<<grid, QTV * 2, kSmem, st>>>((const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk, \
                                                  (const __nv_bfloat16*)v, ldv, (__nv_bfloat16*)o, ldo, lse, heads,  \
                                                  Tq, Tk, kv_len, causal, kv_batch_stride_rows, scale, seed, th, ik, spmm_g_rng_salt); \
  }
  if (QT == 64 && KT == 64) SPMM_ATTN_FWD(64, 64) else if (QT == 64) SPMM_ATTN_FWD(64, 128)
  else if (KT == 64) SPMM_ATTN_FWD(128, 64) else SPMM_ATTN_FWD(128, 128)
#undef SPMM_ATTN_FWD
  SPMM_CHECK_LAUNCH();
  return 0;
}

extern "C" int spmm_attn_bwd(const void* d_o, int lddo, const void* q, int ldq, const void* k, int ldk, const void* v,
                             int ldv, const void* o, int ldo, const float* lse, void* dq, int lddq, void* dk, int lddk,
                             void* dv, int lddv, int batch, int heads, int Tq, int Tk, const int* kv_len, int causal,
                             float scale, float dropout_p, unsigned long long seed, void* stream) {
  SPMM_ARG(d_o && q && k && v && o && lse && dq && dk && dv);
  SPMM_ARG(batch > 0 && heads > 0 && Tq > 0 && Tk > 0 && Tq <= 128 && Tk <= 128);
  SPMM_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && lddo % 8 == 0 && ldo % 2 == 0 && lddq % 2 == 0 &&
           lddk % 2 == 0 && lddv % 2 == 0);
  SPMM_ARG((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)d_o) & 15) == 0);
  uint32_t th; float ik;
  attn_drop(dropout_p, th, ik);
  cudaStream_t st0 = (cudaStream_t)stream;
  if (!getenv("SPMM_ATTN_LEGACY") && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0 && lddo % 8 == 0 &&
      (((uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) & 15) == 0)
    return attn_bwd_tc_launch(d_o, lddo, q, ldq, k, ldk, v, ldv, lse, dq, lddq, dk, lddk, dv, lddv, batch, heads, Tq, Tk, kv_len,
                              causal, scale, th, ik, seed, st0);
  const int KT = Tk <= 64 ? 64 : 128;
  // phase 2 assigns 16 keys per warp, phase 1 16 queries per warp: the CTA needs max(Tq, Tk)/16 warps
  const int QT = (Tq <= 64 && Tk <= 64) ? 64 : 128;
  dim3 grid(heads, batch);
  cudaStream_t st = (cudaStream_t)stream;
#define SPMM_ATTN_BWD(QTV, KTV)                                                                                       \
  {                                                                                                                   \
    constexpr int kSmem = 2 * QTV * 128 + 2 * KTV * 128 + 2 * QTV * KTV * 2 + QTV * 4;                                \
    static bool cfg = false;                                                                                          \
    if (!cfg) { cudaFuncSetAttribute(attn_bwd_kernel<QTV, KTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem); cfg = true; } \
    cudaError_t le = launch_pdl(attn_bwd_kernel<QTV, KTV>, grid, dim3(QTV * 2), kSmem, st,                          \
        (const __nv_bfloat16*)d_o, lddo, (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk,                  \
        (const __nv_bfloat16*)v, ldv, (const __nv_bfloat16*)o, ldo, lse, (__nv_bfloat16*)dq, lddq, (__nv_bfloat16*)dk, \
        lddk, (__nv_bfloat16*)dv, lddv, heads, Tq, Tk, kv_len, causal, scale, seed, th, ik, spmm_g_rng_salt);         \
    if (le != cudaSuccess) return (int)le;                                                                            \
  }
  if (QT == 64) SPMM_ATTN_BWD(64, 64) else if (KT == 64) SPMM_ATTN_BWD(128, 64) else SPMM_ATTN_BWD(128, 128)
#undef SPMM_ATTN_BWD
  SPMM_CHECK_LAUNCH();
  return 0;
}
