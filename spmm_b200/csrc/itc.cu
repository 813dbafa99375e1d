// Fused contrastive (SPC / ITC) head -- reference SPMM_models.py:92-131 (+ the in-batch sims of :157-158).
//
// The reference materialises 8 similarity matrices [B, B+Q] (113 MB at B=96, Q=36864) and ~30 elementwise kernels.
// Here the head is ONE attention-shaped problem per key set: 4B query rows (2B student + 2B momentum-teacher
// features) against N = B + Q keys ([own momentum features | momentum queue]), with the keys also acting as values:
//
//   pass 1   S = Q K^T / temp on the tensor cores, per-row online (max, sum)               -> LSE of every row
//   pass 2   S again, P = exp(S - LSE), O = P K on the tensor cores                        -> O_r = sum_j softmax_rj k_j
//   finish   everything the loss needs is linear in O:  sum_j softmax(s)_rj s_rj = f_r.O_r / temp,
//            sum_j softmax(m)_rj s_rj = f_r.O_teacher(r) / temp,  dL/df_r = (O_r - alpha O_teacher(r) - (1-alpha) k_r+) / (2B temp)
//
// so neither the similarities nor the probabilities are ever written, and student / teacher rows need no pairing
// inside the scan.  Both GEMMs run as tcgen05.mma kind::tf32 straight from the fp32 queue, accumulators in TMEM.
// The key tile is the K-major B operand of S (TMA SWIZZLE_128B) and the MN-major B operand of O; MN-major 32-bit
// operands only exist in the 32-byte-atom swizzle (UMMA layout SWIZZLE_128B_BASE32B = TMA SWIZZLE_128B_ATOM_32B; with
// the 16-byte-atom layout the MMA was observed to write nothing), so pass 2 fetches every key tile in both forms.
// TF32 keeps 10 mantissa bits of the operands -- the precision of the reference's fp16-autocast `@` -- and
// accumulates in fp32; queries and P are rounded to nearest, statistics and the chain rule are fp32.
// Queues are key-major [Q][E]: a key is one contiguous 1 KB row.
#include <cuda.h>
#include <mutex>
#include <cstdlib>

#include "common.cuh"
#include "spmm_b200.h"

namespace spmm {

constexpr int E_ = 256;
constexpr int IT_TM = 128;                      // query rows per CTA (UMMA M)
constexpr int IT_TK = 32;                       // keys per tile (UMMA N of S, K extent of O)
constexpr int IT_KC = E_ / 32;                  // 32-float (128 B) swizzle chunks along E
constexpr int IT_Q_BYTES = IT_TM * E_ * 4;      // 128 KB resident queries
constexpr int IT_QCH_BYTES = IT_TM * 128;       // one chunk of the query tile
constexpr int IT_K_BYTES = IT_TK * E_ * 4;      // 32 KB per key stage
constexpr int IT_KCH_BYTES = IT_TK * 128;       // one chunk of a key tile
constexpr int IT_STAGES = 2;
constexpr int IT_P_BYTES = IT_TM * IT_TK * 4;   // 16 KB probability tile (A operand of O)
constexpr int IT_SMEM = 1024 + IT_Q_BYTES + IT_STAGES * IT_K_BYTES + 2 * IT_P_BYTES + 256;
constexpr int IT_THREADS = 192;                 // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2..5 softmax / epilogue
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct ItcMaps {
  CUtensorMap q;    // Qm [2 * 4B][E] fp32, box 32 x 128 (query tiles)
  CUtensorMap h;    // same buffer, box 32 x 32 (own-momentum head keys)
  CUtensorMap k0;   // text queue [Q][E], box 32 x 32 (key set 0)
  CUtensorMap k1;   // prop queue (key set 1)
  CUtensorMap h_mn, k0_mn, k1_mn;   // the same three key sources in the 32-byte-atom swizzle (pass 2, MN-major operand)
};

struct ItcArgs {
  int B, Q, rows;               // rows = 4B query rows per key set
  int nh, ntiles, tiles_per_split, splits;
  const float* temp;            // device scalar (the clamped nn.Parameter)
  float* part;                  // [2][splits][mtiles*128][2]  (max, sum) in log2 units
  const float* lse;             // [2][mtiles*128]             log2-scaled LSE per row (pass 2)
  float* oacc;                  // [2][4B][E]                  O accumulated over the key splits
  unsigned long long* trace;    // debug: 64 x u64 %globaltimer stamps written by CTA (0,0,0); null in production
};
__device__ __forceinline__ void it_mark(const ItcArgs& a, int slot) {
  if (a.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[slot] = t;
  }
}

__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// MN-major 32-bit operand: [k rows][128 B of MN], 32-byte units XOR (row & 3)  (layout type 1 = SWIZZLE_128B_BASE32B);
// LBO = stride between 128-byte MN blocks, SBO = stride between 4-row K groups.
__device__ __forceinline__ uint64_t umma_smem_desc_mn32(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// PASS 1: row statistics.  PASS 2: O = softmax(S) K.
template <int PASS>
__global__ void __launch_bounds__(IT_THREADS, 1) itc_scan_kernel(const __grid_constant__ ItcMaps maps, const ItcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + IT_Q_BYTES;
  uint8_t* sP = sK + IT_STAGES * IT_K_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * IT_P_BYTES);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* k_empty = bars + 3;       // [2]
  uint64_t* s_full = bars + 5;        // [2]
  uint64_t* s_empty = bars + 7;       // [2]
  uint64_t* p_full = bars + 9;        // [2]
  uint64_t* p_empty = bars + 11;      // [2]
  uint64_t* o_full = bars + 13;       // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, mt = blockIdx.y, ks = blockIdx.z;
  const int t0 = split * a.tiles_per_split, t1 = min(a.ntiles, t0 + a.tiles_per_split);
  const int nt = t1 - t0;
  constexpr uint32_t TMEM_COLS = (PASS == 1) ? 64 : 512;
  constexpr uint32_t O_COL = 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.h);
    tma_prefetch_desc(ks ? &maps.k1 : &maps.k0);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 4);
      mbar_init(&p_full[s], 4);
      mbar_init(&p_empty[s], 1);
    }
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nt > 0) {
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer =====================
      mbar_expect_tx(q_full, IT_Q_BYTES);
#pragma unroll
      for (int c = 0; c < IT_KC; ++c) tma_load_2d(sQ + c * IT_QCH_BYTES, &maps.q, q_full, c * 32, ks * a.rows + mt * IT_TM);
      for (int i = 0; i < nt; ++i) {
        const int t = t0 + i;
        const bool head = t < a.nh;
        const int row = head ? ks * a.rows + 3 * a.B + t * IT_TK : (t - a.nh) * IT_TK;
        if (PASS == 1) {            // two K-major stages
          const int st = i & 1;
          mbar_wait(&k_empty[st], ((i >> 1) & 1) ^ 1);
          mbar_expect_tx(&k_full[st], IT_K_BYTES);
          const CUtensorMap* m = head ? &maps.h : (ks ? &maps.k1 : &maps.k0);
#pragma unroll
          for (int c = 0; c < IT_KC; ++c) tma_load_2d(sK + st * IT_K_BYTES + c * IT_KCH_BYTES, m, &k_full[st], c * 32, row);
        } else {                    // buffer 0: K-major tile for S, buffer 1: MN-major tile for O
          const CUtensorMap* ma = head ? &maps.h : (ks ? &maps.k1 : &maps.k0);
          const CUtensorMap* mb = head ? &maps.h_mn : (ks ? &maps.k1_mn : &maps.k0_mn);
          mbar_wait(&k_empty[0], (i & 1) ^ 1);
          mbar_expect_tx(&k_full[0], IT_K_BYTES);
#pragma unroll
          for (int c = 0; c < IT_KC; ++c) tma_load_2d(sK + c * IT_KCH_BYTES, ma, &k_full[0], c * 32, row);
          mbar_wait(&k_empty[1], (i & 1) ^ 1);
          mbar_expect_tx(&k_full[1], IT_K_BYTES);
#pragma unroll
          for (int c = 0; c < IT_KC; ++c) tma_load_2d(sK + IT_K_BYTES + c * IT_KCH_BYTES, mb, &k_full[1], c * 32, row);
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_s = umma_idesc_tf32(IT_TM, IT_TK, 0);
      constexpr uint32_t idesc_o = umma_idesc_tf32(IT_TM, E_, 1);
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aP = smem_u32(sP);
      mbar_wait(q_full, 0);
      for (int i = 0; i <= nt; ++i) {
        if (i < nt) {
          const int st = i & 1;                       // S accumulator buffer
          const int kst = (PASS == 1) ? st : 0;       // key buffer holding the K-major tile
          const uint32_t ph = (i >> 1) & 1;
          mbar_wait(&k_full[kst], (PASS == 1) ? ph : (uint32_t)(i & 1));
          mbar_wait(&s_empty[st], ph ^ 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < IT_KC; ++c)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_tf32(tmem_base + st * IT_TK, umma_smem_desc(aQ + c * IT_QCH_BYTES + k * 32, 16, 1024),
                          umma_smem_desc(aK + kst * IT_K_BYTES + c * IT_KCH_BYTES + k * 32, 16, 1024), idesc_s, (c | k) != 0);
          tc_commit(&s_full[st]);
          tc_commit(&k_empty[kst]);                   // the K-major tile is free once S is formed
        }
        if (PASS == 2 && i > 0) {
          const int j = i - 1, sj = j & 1;
          mbar_wait(&p_full[sj], (j >> 1) & 1);
          mbar_wait(&k_full[1], j & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < IT_TK / 8; ++k)   // 8 keys per MMA = two 4-row swizzle atoms of every E chunk
            tc_mma_tf32(tmem_base + O_COL, umma_smem_desc(aP + sj * IT_P_BYTES + k * 32, 16, 1024),
                        umma_smem_desc_mn32(aK + IT_K_BYTES + k * 1024, IT_KCH_BYTES, 512), idesc_o, (j | k) != 0);
          tc_commit(&k_empty[1]);
          tc_commit(&p_empty[sj]);
        }
      }
      if (PASS == 2) tc_commit(o_full);
    } else if (warp >= 2) {
      // ===================== softmax / epilogue warps: thread = one query row =====================
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const int gr = mt * IT_TM + r;                        // row within this key set's 4B rows
      const bool row_ok = gr < a.rows;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      const float c2 = LOG2E / __ldg(a.temp);               // logits in log2 units
      float mx = -INFINITY, sum = 0.f;
      float lse2 = INFINITY;
      if (PASS == 2 && row_ok) lse2 = a.lse[(size_t)ks * gridDim.y * IT_TM + gr];
      for (int i = 0; i < nt; ++i) {
        const int t = t0 + i, st = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        const int nvalid = (t < a.nh) ? min(IT_TK, a.B - t * IT_TK) : min(IT_TK, a.Q - (t - a.nh) * IT_TK);
        mbar_wait(&s_full[st], ph);
        tc_fence_after();
        uint32_t sr[32];
        tmem_ld32(lane_addr + st * IT_TK, sr);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[st]);
        if (PASS == 1) {
          float tmax = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = (j < nvalid) ? __uint_as_float(sr[j]) * c2 : -INFINITY;
            sr[j] = __float_as_uint(x);
            tmax = fmaxf(tmax, x);
          }
          if (tmax > mx) { sum *= fast_exp2(mx - tmax); mx = tmax; }
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; ++j) s4[j & 3] += fast_exp2(__uint_as_float(sr[j]) - mx);
          sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
        } else {
          mbar_wait(&p_empty[st], ph ^ 1);
          uint8_t* prow = sP + st * IT_P_BYTES + r * 128;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float4 pv;
            float* pp = reinterpret_cast<float*>(&pv);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 4 * u + e;
              pp[e] = (j < nvalid) ? round_tf32(fast_exp2(__uint_as_float(sr[j]) * c2 - lse2)) : 0.f;
            }
            *reinterpret_cast<float4*>(prow + ((u ^ (r & 7)) << 4)) = pv;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[st]);
        }
      }
      if (PASS == 1) {
        if (row_ok) {
          float* p = a.part + ((((size_t)ks * a.splits + split) * gridDim.y * IT_TM) + gr) * 2;
          p[0] = mx;
          p[1] = sum;
        }
      } else {
        mbar_wait(o_full, 0);
        tc_fence_after();
        float* dst = a.oacc + ((size_t)ks * a.rows + gr) * E_;
#pragma unroll 1
        for (int c = 0; c < E_ / 32; ++c) {
          uint32_t orr[32];
          tmem_ld32(lane_addr + O_COL + c * 32, orr);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c * 32 + 4 * u),
                           "f"(__uint_as_float(orr[4 * u])), "f"(__uint_as_float(orr[4 * u + 1])),
                           "f"(__uint_as_float(orr[4 * u + 2])), "f"(__uint_as_float(orr[4 * u + 3])) : "memory");
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------- pass 1 (v2)
// Row statistics with the 128 x 256 query tile held in TENSOR MEMORY as the A operand (128 lanes x 256 fp32 columns):
// all 224 KB of shared memory become a 7-deep ring of 32-key tiles.  With the queries in shared memory only two key
// tiles fit, and each SM had < 64 KB in flight against ~1.5 us of HBM latency (72 us per scan, 16 % of HBM peak).
// A tcgen05.mma reads its whole A slice (128 rows x 32 B = 4 KB) whatever N is: at N = 32 the issue rate (~63 cycles
// per MMA measured, TMEM A-read bound) not the math (16 cycles) set the pace.  Pass 1 therefore multiplies against
// 64-key tiles (two 32-key TMA tiles side by side), 3 stages of 64 KB.
constexpr int IT1_STAGES = 3;
constexpr int IT1_TK = 2 * IT_TK;                 // 64 keys per MMA tile
constexpr int IT1_K_BYTES = 2 * IT_K_BYTES;       // 64 KB per stage
constexpr int IT1_KCH_BYTES = 2 * IT_KCH_BYTES;   // one 32-float chunk of a 64-key tile
constexpr int IT1_SMEM = 1024 + IT1_STAGES * IT1_K_BYTES + 256;
constexpr uint32_t IT1_QCOL = 128;  // TMEM columns: S[0], S[1] at 0 / 64, the query tile at 128..383

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc], tf32
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(IT_THREADS, 1) itc_stats_kernel(const __grid_constant__ ItcMaps maps, const ItcArgs a,
                                                                  const float* __restrict__ qm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sK = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sK + IT1_STAGES * IT1_K_BYTES);
  uint64_t* k_full = bars;                       // [7]
  uint64_t* k_empty = bars + IT1_STAGES;         // [7]
  uint64_t* s_full = bars + 2 * IT1_STAGES;      // [2]
  uint64_t* s_empty = s_full + 2;                // [2]
  uint64_t* q_ready = s_empty + 2;               // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) it_mark(a, 0);
  const int split = blockIdx.x, mt = blockIdx.y, ks = blockIdx.z;
  const int t0 = split * a.tiles_per_split, t1 = min(a.ntiles, t0 + a.tiles_per_split);
  const int nt32 = t1 - t0;            // 32-key tiles of this split
  const int nt = (nt32 + 1) / 2;       // 64-key MMA tiles

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.h);
    tma_prefetch_desc(ks ? &maps.k1 : &maps.k0);
    for (int s = 0; s < IT1_STAGES; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 4); }
    mbar_init(q_ready, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nt > 0) {
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer: 7-deep ring of key tiles =====================
      for (int i = 0; i < nt; ++i) {
        const int st = i % IT1_STAGES;
        const int halves = min(2, nt32 - 2 * i);
        mbar_wait(&k_empty[st], ((i / IT1_STAGES) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], halves * IT_K_BYTES);
        for (int hh = 0; hh < halves; ++hh) {          // each half is an independent 32-key tile (head keys or queue keys)
          const int t = t0 + 2 * i + hh;
          const bool head = t < a.nh;
          const int row = head ? ks * a.rows + 3 * a.B + t * IT_TK : (t - a.nh) * IT_TK;
          const CUtensorMap* m = head ? &maps.h : (ks ? &maps.k1 : &maps.k0);
#pragma unroll
          for (int c = 0; c < IT_KC; ++c)
            tma_load_2d(sK + st * IT1_K_BYTES + c * IT1_KCH_BYTES + hh * IT_KCH_BYTES, m, &k_full[st], c * 32, row);
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ===================== MMA issuer: S = Q(tmem) . K^T =====================
      constexpr uint32_t idesc_s = umma_idesc_tf32(IT_TM, IT1_TK, 0);
      const uint32_t aK = smem_u32(sK);
      it_mark(a, 1);
      mbar_wait(q_ready, 0);
      tc_fence_after();
      it_mark(a, 2);
      for (int i = 0; i < nt; ++i) {
        const int st = i % IT1_STAGES, ab = i & 1;
        if (i < 12) it_mark(a, 8 + 3 * i);
        mbar_wait(&k_full[st], (i / IT1_STAGES) & 1);
        if (i < 12) it_mark(a, 9 + 3 * i);
        mbar_wait(&s_empty[ab], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        if (i < 12) it_mark(a, 10 + 3 * i);
#pragma unroll
        for (int c = 0; c < IT_KC; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 8 tf32 of K per MMA = 8 TMEM columns of the query tile
            tc_mma_tf32_ts(tmem_base + ab * IT1_TK, tmem_base + IT1_QCOL + c * 32 + k * 8,
                           umma_smem_desc(aK + st * IT1_K_BYTES + c * IT1_KCH_BYTES + k * 32, 16, 1024), idesc_s, (c | k) != 0);
        tc_commit(&s_full[ab]);
        tc_commit(&k_empty[st]);
      }
      it_mark(a, 3);
    } else if (warp >= 2) {
      // ===================== softmax-statistics warps: thread = one query row =====================
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const int gr = mt * IT_TM + r;
      const bool row_ok = gr < a.rows;
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      // the thread's query row (TF32-rounded by itc_normalize_kernel) goes into its TMEM lane
      {
        const float4* src = reinterpret_cast<const float4*>(qm + ((size_t)ks * a.rows + (row_ok ? gr : 0)) * E_);
#pragma unroll 1
        for (int c = 0; c < E_ / 32; ++c) {
          uint32_t v[32];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 f = row_ok ? __ldg(src + c * 8 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[4 * u] = __float_as_uint(f.x); v[4 * u + 1] = __float_as_uint(f.y);
            v[4 * u + 2] = __float_as_uint(f.z); v[4 * u + 3] = __float_as_uint(f.w);
          }
          tmem_st32(lane_addr + IT1_QCOL + c * 32, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(q_ready);
        if (warp == 2 && lane == 0) it_mark(a, 4);
      }
      const float c2 = LOG2E / __ldg(a.temp);               // logits in log2 units
      float mx = -INFINITY, sum = 0.f;
      for (int i = 0; i < nt; ++i) {
        const int ab = i & 1;
        mbar_wait(&s_full[ab], (i >> 1) & 1);
        tc_fence_after();
        uint32_t sr[2][32];
        tmem_ld32(lane_addr + ab * IT1_TK, sr[0]);
        tmem_ld32(lane_addr + ab * IT1_TK + 32, sr[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[ab]);
        float tmax = -INFINITY;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int t = t0 + 2 * i + hh;
          const int nvalid = (t >= t1) ? 0 : (t < a.nh) ? min(IT_TK, a.B - t * IT_TK) : min(IT_TK, a.Q - (t - a.nh) * IT_TK);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = (j < nvalid) ? __uint_as_float(sr[hh][j]) * c2 : -INFINITY;
            sr[hh][j] = __float_as_uint(x);
            tmax = fmaxf(tmax, x);
          }
        }
        if (tmax > mx) { sum *= fast_exp2(mx - tmax); mx = tmax; }
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int j = 0; j < 32; ++j) s4[j & 3] += fast_exp2(__uint_as_float(sr[hh][j]) - mx);
        sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
      }
      if (warp == 2 && lane == 0) it_mark(a, 5);
      if (row_ok) {
        float* p = a.part + ((((size_t)ks * a.splits + split) * gridDim.y * IT_TM) + gr) * 2;
        p[0] = mx;
        p[1] = sum;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (threadIdx.x == 0) it_mark(a, 6);
}

// F.normalize(z, dim=-1) for the four feature matrices (eps 1e-12); one warp per row.  Writes the exact features
// (outputs + chain rule) and the two TF32-rounded query matrices Qm[ks] = [student 2B | teacher 2B]:
//   ks 0 (keys = [m_text | text queue]): f_prop, f_text, m_prop, m_text
//   ks 1 (keys = [m_prop | prop queue]): f_text, f_prop, m_text, m_prop        (rows 3B..4B of Qm[ks] are the head keys)
__global__ void itc_normalize_kernel(const float* z0, const float* z1, const float* z2, const float* z3, float* feats,
                                     float* norms, float* qm, float* out_m_prop, float* out_m_text, int B) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= 4 * B) return;
  const int which = row / B, b = row % B;
  const float* z = (which == 0 ? z0 : which == 1 ? z1 : which == 2 ? z2 : z3) + (size_t)b * E_;
  float v[8], ssq = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { v[i] = z[lane + 32 * i]; ssq += v[i] * v[i]; }
  const float n = fmaxf(sqrtf(warp_sum(ssq)), 1e-12f);
  if (lane == 0) norms[row] = n;
  const int slot0 = which, slot1 = which ^ 1;   // position of this matrix inside Qm[0] / Qm[1]
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float f = v[i] / n;
    const int e = lane + 32 * i;
    feats[(size_t)row * E_ + e] = f;
    const float ft = round_tf32(f);
    qm[((size_t)slot0 * B + b) * E_ + e] = ft;
    qm[((size_t)(4 + slot1) * B + b) * E_ + e] = ft;
    if (which == 2) out_m_prop[(size_t)b * E_ + e] = f;
    if (which == 3) out_m_text[(size_t)b * E_ + e] = f;
  }
}

// in-batch similarities for the hard-negative sampler (SPMM_models.py:157-158): sim_i2t = f_prop m_text^T / temp,
// sim_t2i = f_text m_prop^T / temp, exact fp32.  grid (B, 2), one thread per column.
__global__ void itc_inbatch_kernel(const float* feats, const float* temp, float* sim_i2t, float* sim_t2i, int B) {
  __shared__ float fq[E_];
  const int r = blockIdx.x, which = blockIdx.y;
  const float* f = feats + ((size_t)which * B + r) * E_;
  for (int e = threadIdx.x; e < E_; e += blockDim.x) fq[e] = f[e];
  __syncthreads();
  const float inv_temp = 1.f / __ldg(temp);
  const float* keys = feats + (size_t)(which == 0 ? 3 : 2) * B * E_;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    const float4* k4 = reinterpret_cast<const float4*>(keys + (size_t)j * E_);
    float acc = 0.f;
#pragma unroll 8
    for (int e = 0; e < E_ / 4; ++e) {
      const float4 kv = __ldg(k4 + e);
      acc += fq[4 * e] * kv.x + fq[4 * e + 1] * kv.y + fq[4 * e + 2] * kv.z + fq[4 * e + 3] * kv.w;
    }
    (which == 0 ? sim_i2t : sim_t2i)[(size_t)r * B + j] = acc * inv_temp;
  }
}

// merge the key-split partials: lse2[ks][row] = max + log2(sum)   (log2 units)
__global__ void itc_combine_kernel(const float* part, float* lse, int splits, int rows_pad, int rows, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ks = i / rows_pad, r = i % rows_pad;
  if (r >= rows) { lse[i] = 0.f; return; }   // padding rows of the last query tile: never written by pass 1
  float mx = -INFINITY, sum = 0.f;
  for (int sp = 0; sp < splits; ++sp) {
    const float* p = part + (((size_t)ks * splits + sp) * rows_pad + r) * 2;
    const float m2 = p[0], s2 = p[1];
    if (s2 > 0.f) {
      const float m = fmaxf(mx, m2);
      sum = sum * exp2f(mx - m) + s2 * exp2f(m2 - m);
      mx = m;
    }
  }
  lse[i] = mx + log2f(sum);
}

// Loss, d/d temp and the chain rule through F.normalize.  One block; warp per student feature row (2B rows:
// f_prop[b], f_text[b]); each row occurs once per key set.
__global__ void __launch_bounds__(1024) itc_finish_kernel(const float* feats, const float* norms, const float* oacc,
                                                           const float* lse, const float* temp, float alpha,
                                                           const float* alpha_dev, int B,
                                                           int rows_pad, float* dz_prop, float* dz_text, float* loss,
                                                           float* dtemp, float* nan_flag) {
  __shared__ float sh[32];
  if (alpha_dev != nullptr) alpha = __ldg(alpha_dev);   // device scalar: one captured graph serves the epoch-0 alpha ramp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float inv_temp = 1.f / __ldg(temp);
  const float gscale = inv_temp / (2.f * B);
  const int rows = 4 * B;
  float l_acc = 0.f, dt_acc = 0.f;
  for (int fr = warp; fr < 2 * B; fr += nw) {
    const int which = fr / B, b = fr % B;       // 0: f_prop[b], 1: f_text[b]
    float f[8], g[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { f[i] = feats[(size_t)fr * E_ + lane + 32 * i]; g[i] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      // key set 0 rows: [f_prop | f_text | m_prop | m_text]; key set 1 rows: [f_text | f_prop | m_text | m_prop]
      const int r = (which == ks) ? b : B + b;
      const float* S = oacc + ((size_t)ks * rows + r) * E_;
      const float* T = oacc + ((size_t)ks * rows + 2 * B + r) * E_;
      const float* kp = feats + ((size_t)(ks == 0 ? 3 : 2) * B + b) * E_;   // positive key: own momentum twin
      float ds = 0.f, dt = 0.f, dk = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = lane + 32 * i;
        const float s = S[e], t = T[e], k = kp[e];
        ds += f[i] * s; dt += f[i] * t; dk += f[i] * k;
        g[i] += gscale * (s - alpha * t - (1.f - alpha) * k);
      }
      ds = warp_sum(ds) * inv_temp; dt = warp_sum(dt) * inv_temp; dk = warp_sum(dk) * inv_temp;
      const float lse_s = lse[(size_t)ks * rows_pad + r] * LN2;
      l_acc += lse_s - alpha * dt - (1.f - alpha) * dk;
      dt_acc += ds - alpha * dt - (1.f - alpha) * dk;
    }
    // dz = (g - f (f . g)) / ||z||
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) dot += f[i] * g[i];
    dot = warp_sum(dot);
    const float inv = 1.f / norms[fr];
    float* out = (which == 0 ? dz_prop : dz_text) + (size_t)b * E_;
#pragma unroll
    for (int i = 0; i < 8; ++i) out[lane + 32 * i] = (g[i] - f[i] * dot) * inv;
  }
  // l_acc / dt_acc are warp-uniform: count each warp once
  float l = block_sum(lane == 0 ? l_acc : 0.f, sh);
  float dt = block_sum(lane == 0 ? dt_acc : 0.f, sh);
  if (threadIdx.x == 0) {
    const float L = l / (2.f * B);                  // (sum of 4 row-means) / 2
    *loss = L;
    *dtemp = -dt / (__ldg(temp) * 2.f * B);         // d s / d temp = -s / temp
    if (nan_flag) *nan_flag = (L != L) ? 1.f : 0.f; // reference NaN guard, SPMM_models.py:132
  }
}

struct ItcPlan { int mtiles, rows_pad, nh, nq, ntiles, splits, tps; };
static inline ItcPlan itc_plan(int B, int Q) {
  ItcPlan p;
  p.mtiles = (4 * B + IT_TM - 1) / IT_TM;
  p.rows_pad = p.mtiles * IT_TM;
  p.nh = (B + IT_TK - 1) / IT_TK;
  p.nq = (Q + IT_TK - 1) / IT_TK;
  p.ntiles = p.nh + p.nq;
  int splits = kNumSMs / (2 * p.mtiles);
  if (splits < 1) splits = 1;
  if (splits > p.ntiles) splits = p.ntiles;
  p.tps = (p.ntiles + splits - 1) / splits;
  p.splits = (p.ntiles + p.tps - 1) / p.tps;
  return p;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn itc_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}
// fp32 [rows][E] row-major, box = 32 floats (128 B, SWIZZLE_128B) x box_rows
static int itc_map(CUtensorMap* m, const float* ptr, uint64_t rows, uint32_t box_rows, bool atom32 = false) {
  EncodeTiledFn fn = itc_encode_fn();
  if (!fn) return -2;
  cuuint64_t dims[2] = {(cuuint64_t)E_, rows};
  cuuint64_t strides[1] = {E_ * 4};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -3;
}

}  // namespace spmm
using namespace spmm;

static unsigned long long* g_itc_trace = nullptr;
extern "C" int spmm_itc_debug_trace(void* buf) {   /* 64 x u64 stamps of CTA (0,0,0) of the pass-1 kernel; NULL = off */
  g_itc_trace = reinterpret_cast<unsigned long long*>(buf);
  return 0;
}

extern "C" int64_t spmm_itc_workspace_bytes(int B, int E, int Q) {
  if (E != E_ || B < 1 || Q < 0) return -1;
  const ItcPlan p = itc_plan(B, Q);
  int64_t floats = (int64_t)4 * B * E_                       // feats (exact)
                   + 4 * B                                   // norms
                   + (int64_t)2 * 4 * B * E_                 // Qm (TF32-rounded query / head-key matrices)
                   + (int64_t)2 * p.splits * p.rows_pad * 2  // pass-1 partials
                   + (int64_t)2 * p.rows_pad                 // lse
                   + (int64_t)2 * 4 * B * E_;                // O accumulators
  return floats * 4 + 1024;
}

extern "C" int spmm_itc_fwd_bwd(const float* z_prop, const float* z_text, const float* z_prop_m, const float* z_text_m,
                                const float* prop_queue, const float* text_queue, const float* temp,
                                float alpha, const float* alpha_dev, int B, int E, int Q, float* loss, float* dz_prop,
                                float* dz_text,
                                float* dtemp, float* sim_i2t, float* sim_t2i, float* feat_prop_m, float* feat_text_m,
                                float* nan_flag, void* workspace, int64_t workspace_bytes, void* stream) {
  SPMM_ARG(z_prop && z_text && z_prop_m && z_text_m && temp);
  SPMM_ARG(loss && dz_prop && dz_text && dtemp && sim_i2t && sim_t2i && feat_prop_m && feat_text_m && workspace);
  SPMM_ARG(E == E_ && B >= 1 && Q >= 0 && (Q == 0 || (prop_queue && text_queue)));
  SPMM_ARG(workspace_bytes >= spmm_itc_workspace_bytes(B, E, Q));
  SPMM_ARG(Q == 0 || (((uintptr_t)prop_queue | (uintptr_t)text_queue) & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  const ItcPlan pl = itc_plan(B, Q);
  float* w = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  float* feats = w; w += (size_t)4 * B * E_;
  float* norms = w; w += (4 * B + 63) / 64 * 64;
  float* qm = w; w += (size_t)2 * 4 * B * E_;
  float* part = w; w += (size_t)2 * pl.splits * pl.rows_pad * 2;
  float* lse = w; w += (size_t)2 * pl.rows_pad;
  float* oacc = w;

  ItcMaps maps;
  int rc = itc_map(&maps.q, qm, (uint64_t)2 * 4 * B, IT_TM);
  if (rc) return rc;
  rc = itc_map(&maps.h, qm, (uint64_t)2 * 4 * B, IT_TK);
  if (rc) return rc;
  rc = itc_map(&maps.h_mn, qm, (uint64_t)2 * 4 * B, IT_TK, true);
  if (rc) return rc;
  if (Q > 0) {
    rc = itc_map(&maps.k0, text_queue, Q, IT_TK);
    if (rc) return rc;
    rc = itc_map(&maps.k1, prop_queue, Q, IT_TK);
    if (rc) return rc;
    rc = itc_map(&maps.k0_mn, text_queue, Q, IT_TK, true);
    if (rc) return rc;
    rc = itc_map(&maps.k1_mn, prop_queue, Q, IT_TK, true);
    if (rc) return rc;
  } else {
    maps.k0 = maps.h;
    maps.k1 = maps.h;
    maps.k0_mn = maps.h_mn;
    maps.k1_mn = maps.h_mn;
  }
  ItcArgs a{};
  a.B = B; a.Q = Q; a.rows = 4 * B;
  a.nh = pl.nh; a.ntiles = pl.ntiles; a.tiles_per_split = pl.tps; a.splits = pl.splits;
  a.temp = temp; a.part = part; a.lse = lse; a.oacc = oacc;
  a.trace = g_itc_trace;

  static bool configured = false;
  if (!configured) {
    cudaError_t e1 = cudaFuncSetAttribute(itc_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, IT_SMEM);
    cudaError_t e2 = cudaFuncSetAttribute(itc_scan_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, IT_SMEM);
    cudaError_t e3 = cudaFuncSetAttribute(itc_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, IT1_SMEM);
    if (e3 != cudaSuccess) return (int)e3;
    if (e1 != cudaSuccess) return (int)e1;
    if (e2 != cudaSuccess) return (int)e2;
    configured = true;
  }
  cudaError_t e = cudaMemsetAsync(oacc, 0, (size_t)2 * 4 * B * E_ * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  itc_normalize_kernel<<<(4 * B + 7) / 8, 256, 0, st>>>(z_prop, z_text, z_prop_m, z_text_m, feats, norms, qm, feat_prop_m,
                                                        feat_text_m, B);
  SPMM_CHECK_LAUNCH();
  itc_inbatch_kernel<<<dim3(B, 2), 128, 0, st>>>(feats, temp, sim_i2t, sim_t2i, B);
  SPMM_CHECK_LAUNCH();
  dim3 grid(pl.splits, pl.mtiles, 2);
  static const bool pass1_smem = getenv("SPMM_ITC_PASS1_SMEM") != nullptr;   // A/B: queries in shared memory (2 key stages)
  if (pass1_smem) itc_scan_kernel<1><<<grid, IT_THREADS, IT_SMEM, st>>>(maps, a);
  else itc_stats_kernel<<<grid, IT_THREADS, IT1_SMEM, st>>>(maps, a, qm);
  SPMM_CHECK_LAUNCH();
  const int total = 2 * pl.rows_pad;
  itc_combine_kernel<<<(total + 255) / 256, 256, 0, st>>>(part, lse, pl.splits, pl.rows_pad, 4 * B, total);
  SPMM_CHECK_LAUNCH();
  itc_scan_kernel<2><<<grid, IT_THREADS, IT_SMEM, st>>>(maps, a);
  SPMM_CHECK_LAUNCH();
  itc_finish_kernel<<<1, 1024, 0, st>>>(feats, norms, oacc, lse, temp, alpha, alpha_dev, B, pl.rows_pad, dz_prop, dz_text, loss, dtemp,
                                        nan_flag);
  SPMM_CHECK_LAUNCH();
  return 0;
}
