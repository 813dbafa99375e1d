// Fused attention forward on 5th-gen tensor cores (reference xbert.py:305-354: QK^T/8 + additive mask -> softmax ->
// dropout -> PV; head split/merge permutes :265-268,352-354 folded into the TMA coordinates).
//
// Sequences here are short (Tq, Tk <= 128, head_dim 64), so one 128-row UMMA tile holds either one (batch, head)
// problem or -- when Tq, Tk <= 64 -- a PAIR of heads of the same sample: slot s owns query rows [64s, 64s+64) and key
// rows [64s, 64s+64) of the tile, S = Q K^T is computed for the whole 128 x 128 tile and only the diagonal blocks are
// used; P is written block-diagonal so one O = P V chain serves both heads.
//
// Persistent CTAs, 576 threads:  warp 0 = TMA producer (3-stage Q/K/V ring), warp 1 = tcgen05.mma issuer (S and O
// accumulators double-buffered in TMEM), warps 2..17 = softmax: two sets of 8 warps alternate tiles, two threads per
// query row read its S row from TMEM, mask (kv_len / causal), exponentiate, write bf16 P into the swizzled A-operand
// tile and later scale the O row (the single-warp-per-scheduler version spent 2.7 us per tile in these warps).
// The probability tensor of the reference ([B,12,Tq,Tk]) never exists; LSE is saved for backward.
#include <cuda.h>
#include <mutex>

#include "common.cuh"
#include "spmm_b200.h"

namespace spmm {

constexpr int AT_TILE_BYTES = 128 * 128;                 // [128 rows][64 bf16]: Q, K, V tile / one 64-key chunk of P
constexpr int AT_STAGE_BYTES = 3 * AT_TILE_BYTES;
constexpr int AT_STAGES = 3;
constexpr int AT_P_BYTES = 2 * AT_TILE_BYTES;            // P: [128 rows][128 keys] bf16
constexpr int AT_XCH_BYTES = 2 * 128 * 4 * 4;             // row max / row sum exchange between the two halves of a row
constexpr int AT_SMEM = 1024 + AT_STAGES * AT_STAGE_BYTES + 2 * AT_P_BYTES + AT_XCH_BYTES + 512;
constexpr int AT_THREADS = 64 + 512;                      // TMA warp, MMA warp, 16 softmax warps
constexpr float AT_LOG2E = 1.4426950408889634f;
constexpr float AT_LN2 = 0.6931471805599453f;

struct AttnMaps {
  CUtensorMap q, k, v;   // bf16 [rows][heads*64], box 64 x 64, SWIZZLE_128B
  CUtensorMap o;         // bf16 [batch][Tq][heads*64], box 64 x 64 x 1: rows >= Tq of a sample are clipped by the store
};

struct AttnFwdArgs {
  __nv_bfloat16* o;
  int ldo;
  float* lse;
  int batch, heads, Tq, Tk;
  const int* kv_len;
  const int* kv_index;   // optional: batch element b reads the K/V of batch element kv_index[b]
  int causal, kv_bstride;
  float scale;
  unsigned long long seed;
  uint32_t thresh16;
  float inv_keep;
  const unsigned long long* salt;
  int pair, hp, num_tiles;
  unsigned long long* trace;   // debug: 32 x u64 %globaltimer stamps per CTA, null in production
};

__device__ __forceinline__ void at_mark(const AttnFwdArgs& a, int slot) {
  if (a.trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[(size_t)blockIdx.x * 32 + slot] = t;
  }
}

struct AttnTile { int b, h0, h1, qrow0, qrow1, krow0, krow1; };
__device__ __forceinline__ AttnTile attn_decode(const AttnFwdArgs& a, int tile) {
  AttnTile t;
  if (a.pair) {
    t.b = tile / a.hp;
    t.h0 = 2 * (tile % a.hp);
    t.h1 = min(t.h0 + 1, a.heads - 1);
    t.qrow0 = t.qrow1 = t.b * a.Tq;
    t.krow0 = t.krow1 = (a.kv_index ? __ldg(a.kv_index + t.b) : t.b) * a.kv_bstride;
  } else {
    t.b = tile / a.heads;
    t.h0 = t.h1 = tile % a.heads;
    t.qrow0 = t.b * a.Tq; t.qrow1 = t.qrow0 + 64;
    t.krow0 = (a.kv_index ? __ldg(a.kv_index + t.b) : t.b) * a.kv_bstride; t.krow1 = t.krow0 + 64;
  }
  return t;
}

__device__ __forceinline__ void bar_sync_at(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void at_tma_store_3d(const void* desc, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ float at_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// dropout keep bits for keys (j, j+1) of query row i of head bh: same function in forward and backward
__device__ __forceinline__ uint32_t at_drop_bits(uint32_t key, int bh, int i, int j_even) {
  return drop_bits2(key, ((uint32_t)bh << 14) | ((uint32_t)i << 7) | (uint32_t)j_even);
}

__global__ void __launch_bounds__(AT_THREADS, 1) attn_fwd_tc_kernel(const __grid_constant__ AttnMaps maps, const AttnFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sP = smem + AT_STAGES * AT_STAGE_BYTES;
  float* xch = reinterpret_cast<float*>(sP + 2 * AT_P_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * AT_P_BYTES + AT_XCH_BYTES);
  uint64_t* full = bars;            // [3] TMA landed
  uint64_t* empty = bars + 3;       // [3] stage consumed (O MMA done)
  uint64_t* s_full = bars + 6;      // [2]
  uint64_t* s_free = bars + 8;      // [2]
  uint64_t* p_full = bars + 10;     // [2]
  uint64_t* p_free = bars + 12;     // [2]
  uint64_t* o_full = bars + 14;     // [2]
  uint64_t* o_free = bars + 16;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) at_mark(a, 0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
    tma_prefetch_desc(&maps.o);
    for (int s = 0; s < AT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1); mbar_init(&s_free[s], 8);
      mbar_init(&p_full[s], 8); mbar_init(&p_free[s], 1);
      mbar_init(&o_full[s], 1); mbar_init(&o_free[s], 8);
    }
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
  pdl_wait();
  if (threadIdx.x == 0) at_mark(a, 1);
  // TMEM columns: S[ab] at 128*ab (128 fp32 columns each), O[ob] at 256 + 64*ob

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int n = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++n) {
      const int st = n % AT_STAGES;
      mbar_wait(&empty[st], ((n / AT_STAGES) & 1) ^ 1);
      const AttnTile t = attn_decode(a, tile);
      uint8_t* sq = smem + st * AT_STAGE_BYTES;
      uint8_t* sk = sq + AT_TILE_BYTES;
      uint8_t* sv = sk + AT_TILE_BYTES;
      mbar_expect_tx(&full[st], AT_STAGE_BYTES);
      tma_load_2d(sq, &maps.q, &full[st], t.h0 * 64, t.qrow0);
      tma_load_2d(sq + AT_TILE_BYTES / 2, &maps.q, &full[st], t.h1 * 64, t.qrow1);
      tma_load_2d(sk, &maps.k, &full[st], t.h0 * 64, t.krow0);
      tma_load_2d(sk + AT_TILE_BYTES / 2, &maps.k, &full[st], t.h1 * 64, t.krow1);
      tma_load_2d(sv, &maps.v, &full[st], t.h0 * 64, t.krow0);
      tma_load_2d(sv + AT_TILE_BYTES / 2, &maps.v, &full[st], t.h1 * 64, t.krow1);
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // S = Q K^T: both K-major (head_dim contiguous)
    const uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);    // O = P V: V is [key][d], d contiguous = MN-major B
    const uint32_t aP = smem_u32(sP);
    auto issue_o = [&](int m) {
      const int pb = m & 1, stm = m % AT_STAGES;
      const uint32_t pph = (m >> 1) & 1;
      mbar_wait(&p_full[pb], pph);
      mbar_wait(&o_free[pb], pph ^ 1);
      tc_fence_after();
      const uint32_t sv = smem_u32(smem + stm * AT_STAGE_BYTES + 2 * AT_TILE_BYTES);
#pragma unroll
      for (int k = 0; k < 8; ++k)   // 16 keys per MMA
        tc_mma_bf16(tmem_base + 256 + 64 * pb, umma_smem_desc(aP + pb * AT_P_BYTES + (k >> 2) * AT_TILE_BYTES + (k & 3) * 32, 16, 1024),
                    umma_smem_desc(sv + k * 2048, 8192, 1024), idesc_o, k != 0);
      tc_commit(&o_full[pb]);
      tc_commit(&empty[stm]);
      tc_commit(&p_free[pb]);
    };
    int n = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++n) {
      const int st = n % AT_STAGES, ab = n & 1;
      mbar_wait(&full[st], (n / AT_STAGES) & 1);
      mbar_wait(&s_free[ab], ((n >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t sq = smem_u32(smem + st * AT_STAGE_BYTES), sk = sq + AT_TILE_BYTES;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc_mma_bf16(tmem_base + 128 * ab, umma_smem_desc(sq + k * 32, 16, 1024), umma_smem_desc(sk + k * 32, 16, 1024),
                    idesc_s, k != 0);
      tc_commit(&s_full[ab]);
      if (n > 0) issue_o(n - 1);
    }
    if (n > 0) issue_o(n - 1);
  } else if (warp >= 2) {
    // ===================== softmax / epilogue =====================
    // 16 warps = 2 sets x 2 halves x 4 TMEM lane quadrants.  Set g takes the tiles with (n & 1) == g (and the TMEM /
    // P buffers of that parity); the two threads (half 0 / 1) of a query row split its key columns and its output
    // columns and exchange row max / row sum through shared memory.
    const int sw = warp - 2;
    const int g = sw >> 3, hf = (sw >> 2) & 1, q = warp & 3;
    const int r = q * 32 + lane;
    const int bar_id = 1 + g;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t drop_key = a.thresh16 ? fold_seed(salted(a.seed, a.salt)) : 0u;
    const float c2 = a.scale * AT_LOG2E;
    const int slot = a.pair ? (r >> 6) : 0;
    const int i = a.pair ? (r & 63) : r;              // query index within the head
    const int ncol = a.pair ? 64 : 128;               // columns of S that belong to this row's head
    const int col0 = a.pair ? 64 * slot : 0;
    float* xrow = xch + (g * 128 + r) * 4;            // [max half0, max half1, sum half0, sum half1]
    // carried to the deferred O epilogue of this set's previous tile
    float prev_inv = 0.f;
    const bool trw = (warp == 2 && lane == 0);

    AttnTile prev_t{};
    const bool elected = (hf == 0 && q == 0 && lane == 0);    // issues this set's bulk stores
    auto epilogue_o = [&](int m) {       // this thread: output columns [32 hf, 32 hf + 32) of its row
      mbar_wait(&o_full[g], m & 1);        // m = this set's tile counter; P V done => the P buffer is free as well
      tc_fence_after();
      uint32_t o0[32];
      tmem_ld32(lane_addr + 256 + 64 * g + 32 * hf, o0);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[g]);
      // stage the bf16 O tile ([128 rows][128 B], swizzled) in the first chunk of this set's P buffer, then ONE bulk
      // tensor store per head: per-thread 16-byte global stores (32 lines per warp instruction) were transaction-bound
      uint8_t* orow = sP + g * AT_P_BYTES + r * 128;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(o0[8 * u]) * prev_inv, __uint_as_float(o0[8 * u + 1]) * prev_inv);
        w.y = pack_bf16x2(__uint_as_float(o0[8 * u + 2]) * prev_inv, __uint_as_float(o0[8 * u + 3]) * prev_inv);
        w.z = pack_bf16x2(__uint_as_float(o0[8 * u + 4]) * prev_inv, __uint_as_float(o0[8 * u + 5]) * prev_inv);
        w.w = pack_bf16x2(__uint_as_float(o0[8 * u + 6]) * prev_inv, __uint_as_float(o0[8 * u + 7]) * prev_inv);
        *reinterpret_cast<uint4*>(orow + (((4 * hf + u) ^ (r & 7)) << 4)) = w;
      }
      fence_proxy_async();
      bar_sync_at(bar_id, 256);
      if (elected) {
        uint8_t* stg = sP + g * AT_P_BYTES;
        if (a.pair) {
          at_tma_store_3d(&maps.o, stg, prev_t.h0 * 64, 0, prev_t.b);
          if (prev_t.h0 + 1 < a.heads) at_tma_store_3d(&maps.o, stg + AT_TILE_BYTES / 2, (prev_t.h0 + 1) * 64, 0, prev_t.b);
        } else {
          at_tma_store_3d(&maps.o, stg, prev_t.h0 * 64, 0, prev_t.b);
          if (a.Tq > 64) at_tma_store_3d(&maps.o, stg + AT_TILE_BYTES / 2, prev_t.h0 * 64, 64, prev_t.b);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };

    int n = 0, nl = 0;   // n: CTA-local tile counter, nl: this set's tile counter
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++n) {
      if ((n & 1) != g) continue;
      const bool tr = trw && nl < 4;
      if (nl > 0) epilogue_o(nl - 1);              // O of this set's previous tile (its P V ran during the other set's tile)
      const AttnTile t = attn_decode(a, tile);
      const int h = a.pair ? t.h0 + slot : t.h0;
      const bool valid = (h < a.heads) && (i < a.Tq);
      const int klen = a.kv_len ? min(__ldg(a.kv_len + t.b), a.Tk) : a.Tk;
      const int jmax = (a.causal > 0 && t.b >= a.causal - 1) ? min(klen, i + 1) : klen;   // keys [0, jmax) are visible to this row
      const int bh = t.b * a.heads + h;
      const uint32_t ph = nl & 1;
      if (tr) at_mark(a, 4 + 6 * nl);
      mbar_wait(&s_full[g], ph);
      tc_fence_after();
      if (tr) at_mark(a, 5 + 6 * nl);
      const uint32_t s_addr = lane_addr + 128 * g + col0;
      // this thread's key chunks (32 columns each) inside the row's head block: cc = 32 hf (+ 64 in single-head mode)
      // pass 1: row maximum over the visible keys (log2 units)
      float mx = -INFINITY;
      for (int cc = 32 * hf; cc < ncol && cc < klen; cc += 64) {     // warp-uniform (tcgen05.ld is warp-collective)
        uint32_t sr[32];
        tmem_ld32(s_addr + cc, sr);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (cc + j < jmax) mx = fmaxf(mx, __uint_as_float(sr[j]));
      }
      xrow[hf] = mx;
      if (elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // O staging (= P buffer) drained
      bar_sync_at(bar_id, 256);
      mx = fmaxf(xrow[0], xrow[1]);
      mx = (mx == -INFINITY) ? 0.f : mx * c2;     // fully masked row: p = 0 everywhere, output zeros
      if (tr) at_mark(a, 6 + 6 * nl);
      // pass 2: p = exp2(s*c2 - mx), partial row sum, bf16 P (with dropout) into the A-operand tile.  The thread writes
      // the P columns [32 hf, +32) and [64 + 32 hf, +32) of its row; columns of the other head's block are zeros.
      mbar_wait(&p_free[g], ph ^ 1);
      uint8_t* prow = sP + g * AT_P_BYTES + r * 128;
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 32 * hf; c < 128; c += 64) {
        const int cc = c - col0;                   // column within this row's head
        const bool blk = cc >= 0 && cc < ncol && cc < klen;   // warp-uniform: chunk of this warp's head with visible keys
        uint32_t w[16];
        if (blk) {
          uint32_t sr[32];
          tmem_ld32(s_addr + cc, sr);
          tmem_ld_wait();
          float p[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            p[j] = (valid && cc + j < jmax) ? at_exp2(__uint_as_float(sr[j]) * c2 - mx) : 0.f;
            s4[j & 3] += p[j];
          }
          if (a.thresh16) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const uint32_t hb = at_drop_bits(drop_key, bh, i, cc + j);
              p[j] = ((hb & 0xFFFFu) >= a.thresh16) ? p[j] * a.inv_keep : 0.f;
              p[j + 1] = ((hb >> 16) >= a.thresh16) ? p[j + 1] * a.inv_keep : 0.f;
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = pack_bf16x2(p[2 * j], p[2 * j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) w[j] = 0u;
        }
        uint8_t* chunk = prow + (c >> 6) * AT_TILE_BYTES;   // 64-key chunk of the P tile
        const int u0 = (c & 63) >> 3;                      // first 16-byte unit (8 keys) inside the chunk row
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<uint4*>(chunk + (((u0 + u) ^ (r & 7)) << 4)) = make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
      }
      if (tr) at_mark(a, 7 + 6 * nl);
      xrow[2 + hf] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&s_free[g]); mbar_arrive(&p_full[g]); }
      bar_sync_at(bar_id, 256);
      const float sum = xrow[2] + xrow[3];
      if (tr) at_mark(a, 8 + 6 * nl);
      if (valid && hf == 0 && a.lse != nullptr) a.lse[(size_t)bh * a.Tq + i] = (sum > 0.f) ? mx * AT_LN2 + __logf(sum) : 0.f;
      prev_inv = (valid && sum > 0.f) ? 1.f / sum : 0.f;
      prev_t = t;
      if (tr) at_mark(a, 9 + 6 * nl);
      ++nl;
    }
    if (nl > 0) {
      epilogue_o(nl - 1);
      if (elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
  if (threadIdx.x == 64) at_mark(a, 2);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (threadIdx.x == 0) at_mark(a, 3);
}

typedef CUresult (*AtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static AtEncodeFn at_encode_fn() {
  static AtEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<AtEncodeFn>(f);
  });
  return fn;
}
// bf16 [rows][cols] with leading dimension ld (elements); box = 64 columns (one head) x 64 rows
int attn_make_map(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld) {
  AtEncodeFn fn = at_encode_fn();
  if (!fn) return -2;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -3;
}

static unsigned long long* g_attn_trace = nullptr;
void attn_set_trace(void* p) { g_attn_trace = reinterpret_cast<unsigned long long*>(p); }

// bf16 [batch][T][cols] view of a [batch*T][ld] buffer; box = 64 columns x 64 rows of one sample (rows >= T are clipped)
int attn_make_map3d(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t T, uint64_t batch, uint64_t ld) {
  AtEncodeFn fn = at_encode_fn();
  if (!fn) return -2;
  cuuint64_t dims[3] = {cols, T, batch};
  cuuint64_t strides[2] = {ld * 2, T * ld * 2};
  cuuint32_t box[3] = {64, 64, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -3;
}

int attn_fwd_tc_launch(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, float* lse,
                       int batch, int heads, int Tq, int Tk, const int* kv_len, int causal, int kv_bstride, float scale,
                       uint32_t thresh16, float inv_keep, unsigned long long seed, const int* kv_index, int kv_batches,
                       cudaStream_t st) {
  AttnMaps maps;
  const int nkv = kv_index ? kv_batches : batch;
  const uint64_t krows = kv_bstride ? (uint64_t)(nkv - 1) * kv_bstride + Tk : (uint64_t)Tk;
  int rc = attn_make_map(&maps.q, q, (uint64_t)heads * 64, (uint64_t)batch * Tq, ldq);
  if (rc) return rc;
  rc = attn_make_map(&maps.k, k, (uint64_t)heads * 64, krows, ldk);
  if (rc) return rc;
  rc = attn_make_map(&maps.v, v, (uint64_t)heads * 64, krows, ldv);
  if (rc) return rc;
  rc = attn_make_map3d(&maps.o, o, (uint64_t)heads * 64, Tq, batch, ldo);
  if (rc) return rc;
  AttnFwdArgs a{};
  a.o = reinterpret_cast<__nv_bfloat16*>(o); a.ldo = ldo; a.lse = lse;
  a.batch = batch; a.heads = heads; a.Tq = Tq; a.Tk = Tk; a.kv_len = kv_len; a.causal = causal; a.kv_bstride = kv_bstride;
  a.kv_index = kv_index;
  a.scale = scale; a.seed = seed; a.thresh16 = thresh16; a.inv_keep = inv_keep; a.salt = spmm_g_rng_salt;
  a.pair = (Tq <= 64 && Tk <= 64) ? 1 : 0;
  a.trace = g_attn_trace;
  a.hp = (heads + 1) / 2;
  a.num_tiles = a.pair ? batch * a.hp : batch * heads;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int ctas = a.num_tiles < kNumSMs ? a.num_tiles : kNumSMs;
  cudaError_t le = launch_pdl(attn_fwd_tc_kernel, dim3(ctas), dim3(AT_THREADS), AT_SMEM, st, maps, a);
  if (le != cudaSuccess) return (int)le;
  return 0;
}

}  // namespace spmm

// =====================================================================================================================
// Backward on the same tile geometry (autograd of xbert.py:305-354).  Per 128 x 128 tile (one head, or a pair of heads
// on the diagonal blocks) five UMMA chains, all operands in the layouts TMA / the softmax threads already produce:
//   S  = Q K^T          A = Q  (K-major)    B = K  (K-major)          recompute; P = exp(S scale - LSE)
//   dP = dO V^T         A = dO (K-major)    B = V  (K-major)
//   dV = Pd^T dO        A = Pd (MN-major: [query][key] tile read with M = keys)    B = dO (MN-major)
//   dK = dS^T Q         A = dS (MN-major)   B = Q  (MN-major)         dS already carries `scale`
//   dQ = dS K           A = dS (K-major)    B = K  (MN-major)
// with Pd = P o dropout-mask / keep, D_i = sum_j Pd_ij dP_ij, dS = P o (dP o mask / keep - D) * scale.
// 576 threads: TMA producer (2-stage Q/K/V/dO ring), MMA issuer, 16 softmax warps = 4 threads per tile row (each
// owns 16 key columns of the row's head when two heads share the tile, 32 otherwise; with 576 threads the register
// file gives 96 per thread, and the 32-column form of the pair case spilled to local memory: 7.4 -> 4.6 us per tile).
// dQ / dK / dV leave through 3-D bulk tensor stores (rows beyond the sequence are clipped), one issued by each of six
// threads; their column sums (projection bias gradients) are kept in registers per head pair and flushed once.
namespace spmm {

constexpr int AB_STAGE_BYTES = 4 * AT_TILE_BYTES;            // Q, K, V, dO
constexpr int AB_STAGES = 2;
constexpr int AB_PDS_BYTES = 4 * AT_TILE_BYTES;              // Pd (2 chunks) + dS (2 chunks); reused as dQ/dK/dV staging
constexpr int AB_XCH_BYTES = 128 * 4 * 4;
constexpr int AB_SMEM = 1024 + AB_STAGES * AB_STAGE_BYTES + AB_PDS_BYTES + AB_XCH_BYTES + 512;
constexpr uint32_t AB_S = 0, AB_DP = 128, AB_DK = 256, AB_DV = 320, AB_DQ = 384;   // TMEM columns

struct AttnBwdMaps { CUtensorMap q, k, v, dout, dq, dk, dv; };

template <int N> struct TmemLd;
template <> struct TmemLd<32> {
  static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
};
template <> struct TmemLd<16> {
  static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
  }
};

struct AttnBwdArgs {
  const float* lse;
  int batch, heads, Tq, Tk;
  const int* kv_len;
  const int* kv_index;   // optional: K/V of batch element b are rows of batch element kv_index[b]; dK/dV stay per b
  int causal;
  float scale;
  unsigned long long seed;
  uint32_t thresh16;
  float inv_keep;
  const unsigned long long* salt;
  int pair, hp, num_tiles;
  float* dbq;   // optional [heads*64] fp32 each: += column sums of dQ / dK / dV (bias gradients of the q/k/v projections)
  float* dbk;
  float* dbv;
  unsigned long long* trace;   // debug: 32 x u64 %globaltimer stamps per CTA, null in production
};

__device__ __forceinline__ void ab_mark(const AttnBwdArgs& a, int slot) {
  if (a.trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    a.trace[(size_t)blockIdx.x * 32 + slot] = t;
  }
}

// PAIR: two heads of one batch element share a tile (Tq, Tk <= 64; every benchmarked shape): a row has 64 live key
// columns, split 16 per softmax thread so all 16 warps work and nothing spills.  !PAIR: 128 columns, 32 per thread.
template <bool PAIR>
__global__ void __launch_bounds__(AT_THREADS, 1) attn_bwd_tc_kernel(const __grid_constant__ AttnBwdMaps maps, const AttnBwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  pdl_trigger();
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  if (threadIdx.x == 0) ab_mark(a, 0);
  uint8_t* sPd = smem + AB_STAGES * AB_STAGE_BYTES;          // [2 chunks][128 rows][128 B]
  uint8_t* sdS = sPd + 2 * AT_TILE_BYTES;
  float* xch = reinterpret_cast<float*>(sPd + AB_PDS_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPd + AB_PDS_BYTES + AB_XCH_BYTES);
  uint64_t* full = bars;            // [2]
  uint64_t* empty = bars + 2;       // [2]
  uint64_t* sdp_full = bars + 4;    // S and dP accumulators ready
  uint64_t* sdp_free = bars + 5;    // softmax threads have read S and dP (16 warps)
  uint64_t* pds_full = bars + 6;    // Pd and dS operand tiles written (16 warps)
  uint64_t* out_full = bars + 7;    // dQ, dK, dV accumulators ready
  uint64_t* acc_free = bars + 8;    // dQ, dK, dV accumulators read (16 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.q); tma_prefetch_desc(&maps.k); tma_prefetch_desc(&maps.v); tma_prefetch_desc(&maps.dout);
    tma_prefetch_desc(&maps.dq); tma_prefetch_desc(&maps.dk); tma_prefetch_desc(&maps.dv);
    for (int s = 0; s < AB_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(sdp_full, 1); mbar_init(sdp_free, 16); mbar_init(pds_full, 16); mbar_init(out_full, 1); mbar_init(acc_free, 16);
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
  pdl_wait();

  // forward-compatible tile decoding (same pairing as attn_fwd_tc_kernel); K/V rows are batch-major (no broadcast)
  auto decode = [&](int tile, int& b, int& h0, int& h1, int& qrow0, int& qrow1, int& krow0, int& krow1) {
    if (PAIR) {
      b = tile / a.hp; h0 = 2 * (tile % a.hp); h1 = min(h0 + 1, a.heads - 1);
      qrow0 = qrow1 = b * a.Tq; krow0 = krow1 = (a.kv_index ? __ldg(a.kv_index + b) : b) * a.Tk;
    } else {
      b = tile / a.heads; h0 = h1 = tile % a.heads;
      qrow0 = b * a.Tq; qrow1 = qrow0 + 64; krow0 = (a.kv_index ? __ldg(a.kv_index + b) : b) * a.Tk; krow1 = krow0 + 64;
    }
  };

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int n = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++n) {
      const int st = n % AB_STAGES;
      mbar_wait(&empty[st], ((n / AB_STAGES) & 1) ^ 1);
      int b, h0, h1, q0, q1, k0, k1;
      decode(tile, b, h0, h1, q0, q1, k0, k1);
      uint8_t* sq = smem + st * AB_STAGE_BYTES;
      uint8_t* sk = sq + AT_TILE_BYTES;
      uint8_t* sv = sk + AT_TILE_BYTES;
      uint8_t* sdo = sv + AT_TILE_BYTES;
      mbar_expect_tx(&full[st], AB_STAGE_BYTES);
      tma_load_2d(sq, &maps.q, &full[st], h0 * 64, q0);
      tma_load_2d(sq + AT_TILE_BYTES / 2, &maps.q, &full[st], h1 * 64, q1);
      tma_load_2d(sk, &maps.k, &full[st], h0 * 64, k0);
      tma_load_2d(sk + AT_TILE_BYTES / 2, &maps.k, &full[st], h1 * 64, k1);
      tma_load_2d(sv, &maps.v, &full[st], h0 * 64, k0);
      tma_load_2d(sv + AT_TILE_BYTES / 2, &maps.v, &full[st], h1 * 64, k1);
      tma_load_2d(sdo, &maps.dout, &full[st], h0 * 64, q0);
      tma_load_2d(sdo + AT_TILE_BYTES / 2, &maps.dout, &full[st], h1 * 64, q1);
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    const uint32_t id_kk = umma_idesc_bf16(128, 128, 0, 0);   // S, dP: both operands K-major (head_dim contiguous)
    const uint32_t id_mm = umma_idesc_bf16(128, 64, 1, 1);    // dV, dK: A = [query][key] tile read MN-major, B MN-major
    const uint32_t id_km = umma_idesc_bf16(128, 64, 0, 1);    // dQ: A = dS K-major, B = K MN-major
    const uint32_t aPd = smem_u32(sPd), adS = smem_u32(sdS);
    // S and dP of tile m: issued as soon as the softmax threads have read tile m-1's, i.e. while they are still turning
    // tile m-1's P / dP into the dS operand - the accumulators are ready when they come back for tile m
    auto issue_sdp = [&](int m) {
      const int st = m % AB_STAGES;
      const uint32_t sq = smem_u32(smem + st * AB_STAGE_BYTES), sk = sq + AT_TILE_BYTES, sv = sk + AT_TILE_BYTES,
                     sdo = sv + AT_TILE_BYTES;
      mbar_wait(&full[st], (m / AB_STAGES) & 1);
      mbar_wait(sdp_free, (m & 1) ^ 1);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc_mma_bf16(tmem_base + AB_S, umma_smem_desc(sq + k * 32, 16, 1024), umma_smem_desc(sk + k * 32, 16, 1024), id_kk, k != 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc_mma_bf16(tmem_base + AB_DP, umma_smem_desc(sdo + k * 32, 16, 1024), umma_smem_desc(sv + k * 32, 16, 1024), id_kk, k != 0);
      tc_commit(sdp_full);
    };
    int n = 0;
    if ((int)blockIdx.x < a.num_tiles) issue_sdp(0);
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++n) {
      const int st = n % AB_STAGES;
      const uint32_t ph = n & 1;
      const uint32_t sq = smem_u32(smem + st * AB_STAGE_BYTES), sk = sq + AT_TILE_BYTES, sdo = sk + 2 * AT_TILE_BYTES;
      if (tile + (int)gridDim.x < a.num_tiles) issue_sdp(n + 1);
      mbar_wait(pds_full, ph);
      mbar_wait(acc_free, ph ^ 1);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 8; ++k)   // 16 queries per MMA
        tc_mma_bf16(tmem_base + AB_DV, umma_smem_desc(aPd + k * 2048, 2 * 8192, 1024), umma_smem_desc(sdo + k * 2048, 8192, 1024), id_mm, k != 0);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        tc_mma_bf16(tmem_base + AB_DK, umma_smem_desc(adS + k * 2048, 2 * 8192, 1024), umma_smem_desc(sq + k * 2048, 8192, 1024), id_mm, k != 0);
#pragma unroll
      for (int k = 0; k < 8; ++k)   // 16 keys per MMA
        tc_mma_bf16(tmem_base + AB_DQ, umma_smem_desc(adS + (k >> 2) * AT_TILE_BYTES + (k & 3) * 32, 16, 1024),
                    umma_smem_desc(sk + k * 2048, 8192, 1024), id_km, k != 0);
      tc_commit(out_full);
      tc_commit(&empty[st]);
    }
  } else if (warp >= 2) {
    // ===================== softmax-backward threads: 4 per tile row =====================
    const int sw = warp - 2;
    const int part = sw >> 2, q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t drop_key = a.thresh16 ? fold_seed(salted(a.seed, a.salt)) : 0u;
    const float c2 = a.scale * AT_LOG2E;
    constexpr int NC = PAIR ? 16 : 32;                     // key columns per softmax thread
    constexpr int NU = NC / 8;                             // = 16-byte units of a bf16 operand row
    const int slot = PAIR ? (r >> 6) : 0;
    const int i = PAIR ? (r & 63) : r;
    const int cc = NC * part;                              // first key (within the head) of this thread's chunk
    const uint32_t tcol = (PAIR ? 64 * slot : 0) + cc;     // its column in the S / dP accumulators
    const bool elected = (sw == 0 && lane == 0);
    const int st_role = (lane == 0 && sw < 6) ? sw : -1;   // issues (and waits for) one of the six output stores
    float bsum[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};       // PAIR: bias-gradient column sums per head pair
    // operand rows: PAIR -> the head's own 64-key chunk (and zeros into the other head's chunk: the tile is block
    // diagonal); !PAIR -> chunk = 64-key half of the row
    const int chunk = PAIR ? slot : (part >> 1);
    const int u0 = PAIR ? 2 * part : (part & 1) * 4;
    uint8_t* pd_row = sPd + chunk * AT_TILE_BYTES + r * 128;
    uint8_t* ds_row = sdS + chunk * AT_TILE_BYTES + r * 128;
    // per-tile scalars from global memory (key count of the batch element, this row's log-sum-exp) are fetched one
    // tile ahead: their latency would otherwise sit between two tiles
    auto fetch = [&](int tile, int& klen_out, float& lse_out) {
      int b, h0, h1, q0r, q1r, k0r, k1r;
      decode(tile, b, h0, h1, q0r, q1r, k0r, k1r);
      const int h = PAIR ? h0 + slot : h0;
      klen_out = a.kv_len ? min(__ldg(a.kv_len + b), a.Tk) : a.Tk;
      lse_out = (h < a.heads && i < a.Tq) ? __ldg(a.lse + (size_t)(b * a.heads + h) * a.Tq + i) : 0.f;
    };
    int klen_nx = 0;
    float lse_nx = 0.f;
    if ((int)blockIdx.x < a.num_tiles) fetch(blockIdx.x, klen_nx, lse_nx);
    int n = 0;
    for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++n) {
      int b, h0, h1, q0r, q1r, k0r, k1r;
      decode(tile, b, h0, h1, q0r, q1r, k0r, k1r);
      const int h = PAIR ? h0 + slot : h0;
      const bool head_ok = h < a.heads;
      const bool valid = head_ok && (i < a.Tq);
      const int klen = klen_nx;
      const int jmax = (a.causal > 0 && b >= a.causal - 1) ? min(klen, i + 1) : klen;
      const int bh = b * a.heads + h;
      const float lse2 = lse_nx * AT_LOG2E;
      if (tile + (int)gridDim.x < a.num_tiles) fetch(tile + gridDim.x, klen_nx, lse_nx);
      const uint32_t ph = n & 1;
      const bool tr = elected && a.trace != nullptr && n >= 1 && n < 4;   // tiles 1..3: steady state
      const int tb = 4 + 8 * (n - 1);
      mbar_wait(sdp_full, ph);
      tc_fence_after();
      if (tr) ab_mark(a, tb);
      float p[NC], dpk[NC];
      float dpart = 0.f;
      uint32_t kmask = 0xFFFFFFFFu;                        // dropout keep bits of this thread's keys
      const bool blk = cc < klen;                          // warp-uniform
      if (blk) {
        uint32_t sr[NC], dr[NC];
        TmemLd<NC>::ld(lane_addr + AB_S + tcol, sr);
        TmemLd<NC>::ld(lane_addr + AB_DP + tcol, dr);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          p[j] = (valid && cc + j < jmax) ? at_exp2(__uint_as_float(sr[j]) * c2 - lse2) : 0.f;
          dpk[j] = __uint_as_float(dr[j]);
        }
        if (a.thresh16) {
          kmask = 0u;
#pragma unroll
          for (int j = 0; j < NC; j += 2) {
            const uint32_t hb = at_drop_bits(drop_key, bh, i, cc + j);
            const bool k0 = (hb & 0xFFFFu) >= a.thresh16, k1 = (hb >> 16) >= a.thresh16;
            kmask |= (k0 ? 1u : 0u) << j;
            kmask |= (k1 ? 1u : 0u) << (j + 1);
            dpk[j] = k0 ? dpk[j] * a.inv_keep : 0.f;
            dpk[j + 1] = k1 ? dpk[j + 1] * a.inv_keep : 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < NC; ++j) dpart += p[j] * dpk[j];     // = Pd_ij dP_ij (keep factor already in dpk)
      } else {
#pragma unroll
        for (int j = 0; j < NC; ++j) { p[j] = 0.f; dpk[j] = 0.f; }
      }
      if (PAIR) {                                          // S and dP are consumed: the next tile's may be issued
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sdp_free);
      }
      xch[r * 4 + part] = dpart;
      // the Pd / dS tiles double as the output staging of the previous tile: wait until its bulk stores have read them
      if (st_role >= 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      bar_sync_at(1, 512);
      if (tr) ab_mark(a, tb + 1);
      const float D = (xch[r * 4] + xch[r * 4 + 1]) + (xch[r * 4 + 2] + xch[r * 4 + 3]);
      if (!PAIR) {
        // 32 columns per thread: p[] and dP together do not fit the register budget of 576 threads, so dP is read from
        // TMEM a second time here instead of being carried across the barrier
        if (blk) {
          uint32_t dr[NC];
          TmemLd<NC>::ld(lane_addr + AB_DP + tcol, dr);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < NC; ++j)
            dpk[j] = a.thresh16 ? (((kmask >> j) & 1u) ? __uint_as_float(dr[j]) * a.inv_keep : 0.f) : __uint_as_float(dr[j]);
        } else {
#pragma unroll
          for (int j = 0; j < NC; ++j) dpk[j] = 0.f;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sdp_free);
      }
      // Pd = p * keep (keep = mask / keep_prob, bits kept in kmask)
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        uint32_t wp[4], wd[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const int j = 8 * u + 2 * e2;
          const float kf = a.thresh16 ? a.inv_keep : 1.f;
          const float k0f = ((kmask >> j) & 1u) ? kf : 0.f, k1f = ((kmask >> (j + 1)) & 1u) ? kf : 0.f;
          wp[e2] = pack_bf16x2(p[j] * k0f, p[j + 1] * k1f);
          wd[e2] = pack_bf16x2(p[j] * (dpk[j] - D) * a.scale, p[j + 1] * (dpk[j + 1] - D) * a.scale);
        }
        const int off = ((u0 + u) ^ (r & 7)) << 4;
        *reinterpret_cast<uint4*>(pd_row + off) = make_uint4(wp[0], wp[1], wp[2], wp[3]);
        *reinterpret_cast<uint4*>(ds_row + off) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
        if (PAIR) {   // the other head's keys: zero (rewritten every tile - the region doubles as output staging)
          const int zoff = (1 - 2 * slot) * AT_TILE_BYTES + off;
          *reinterpret_cast<uint4*>(pd_row + zoff) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(ds_row + zoff) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      if (tr) ab_mark(a, tb + 2);
      // ---- outputs (lanes = query rows of dQ / key rows of dK, dV)
      mbar_wait(out_full, ph);
      tc_fence_after();
      if (tr) ab_mark(a, tb + 3);
      // dQ | dK | dV are 3 x 64 fp32 columns = twelve 16-column blocks, three per thread of a row (part p takes blocks
      // 3p .. 3p+2): 48 live registers and the same work for every warp
      uint32_t o[3][16];
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int g = 3 * part + t, which = g >> 2, cb = g & 3;
        TmemLd<16>::ld(lane_addr + (which == 0 ? AB_DQ : (which == 1 ? AB_DK : AB_DV)) + 16 * cb, o[t]);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_free);
      if (tr) ab_mark(a, tb + 4);
      // staging tiles ([128 rows][128 B], swizzled) in the Pd / dS region (the MMAs that read it are complete): dQ | dK | dV
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int g = 3 * part + t, which = g >> 2, cb = g & 3;
        uint8_t* trow = sPd + which * AT_TILE_BYTES + r * 128;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(o[t][8 * u]), __uint_as_float(o[t][8 * u + 1]));
          w.y = pack_bf16x2(__uint_as_float(o[t][8 * u + 2]), __uint_as_float(o[t][8 * u + 3]));
          w.z = pack_bf16x2(__uint_as_float(o[t][8 * u + 4]), __uint_as_float(o[t][8 * u + 5]));
          w.w = pack_bf16x2(__uint_as_float(o[t][8 * u + 6]), __uint_as_float(o[t][8 * u + 7]));
          *reinterpret_cast<uint4*>(trow + (((2 * cb + u) ^ (r & 7)) << 4)) = w;
        }
      }
      fence_proxy_async();
      bar_sync_at(1, 512);
      if (tr) ab_mark(a, tb + 5);
      // ---- bulk tensor stores of the staged tiles: six threads (lane 0 of warps sw 0..5) issue one each
      if (st_role >= 0) {
        const int which = st_role % 3, second = st_role / 3;            // dQ | dK | dV; first / second 64-row half
        const CUtensorMap* mp = which == 0 ? &maps.dq : (which == 1 ? &maps.dk : &maps.dv);
        const uint8_t* src = sPd + which * AT_TILE_BYTES + second * (AT_TILE_BYTES / 2);
        if (PAIR) {                                                     // halves = the two heads of the pair
          if (!second || h0 + 1 < a.heads) at_tma_store_3d(mp, src, (h0 + second) * 64, 0, b);
        } else {                                                        // halves = rows 0..63 / 64..127 of one head
          if (!second || (which == 0 ? a.Tq : a.Tk) > 64) at_tma_store_3d(mp, src, h0 * 64, 64 * second, b);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (tr) ab_mark(a, tb + 6);
      if (a.dbq != nullptr && sw < 12) {
        // Bias gradients of the q/k/v projections = column sums of dQ / dK / dV (reference: autograd of nn.Linear,
        // xbert.py:280-298), taken from the staged bf16 tiles instead of a separate pass over the stored tensors.
        // Warp = (tensor, 64-row slot, 32-column half); lane = (16-byte unit, row group): 8 conflict-free 16-byte
        // loads per lane, then the 8 row groups are folded by three exchanges that leave lane rg with column rg of
        // its unit.  Rows past the sequence hold zeros (their P / dS rows are zero).
        const int which = sw >> 2, s_ = (sw >> 1) & 1;
        const int un = (sw & 1) * 4 + (lane >> 3), rg = lane & 7;
        const uint8_t* base = sPd + which * AT_TILE_BYTES + (s_ * 64) * 128;
        float c8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rr = rg + 8 * k;
          const uint4 w = *reinterpret_cast<const uint4*>(base + rr * 128 + ((un ^ (rr & 7)) << 4));
          float x, y;
          unpack_bf16x2(w.x, x, y); c8[0] += x; c8[1] += y;
          unpack_bf16x2(w.y, x, y); c8[2] += x; c8[3] += y;
          unpack_bf16x2(w.z, x, y); c8[4] += x; c8[5] += y;
          unpack_bf16x2(w.w, x, y); c8[6] += x; c8[7] += y;
        }
        float d4[4], d2[2];
        const bool b4 = lane & 4, b2 = lane & 2, b1 = lane & 1;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          d4[j] = (b4 ? c8[j + 4] : c8[j]) + __shfl_xor_sync(0xFFFFFFFFu, b4 ? c8[j] : c8[j + 4], 4);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          d2[j] = (b2 ? d4[j + 2] : d4[j]) + __shfl_xor_sync(0xFFFFFFFFu, b2 ? d4[j] : d4[j + 2], 2);
        const float csum = (b1 ? d2[1] : d2[0]) + __shfl_xor_sync(0xFFFFFFFFu, b1 ? d2[0] : d2[1], 1);
        if (PAIR && a.hp <= 6) {
          // per-head-pair register accumulators, flushed once when the CTA is done: no per-tile atomics
          const int hpi = h0 >> 1;
#pragma unroll
          for (int k = 0; k < 6; ++k) bsum[k] += (hpi == k) ? csum : 0.f;
        } else {
          const int hh = PAIR ? h0 + s_ : h0;
          float* dst = which == 0 ? a.dbq : (which == 1 ? a.dbk : a.dbv);
          if (hh < a.heads) atomicAdd(dst + hh * 64 + un * 8 + rg, csum);
        }
      }
      if (tr) ab_mark(a, tb + 7);
    }
    if (st_role >= 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (PAIR && a.hp <= 6 && a.dbq != nullptr && sw < 12) {
      const int which = sw >> 2, s_ = (sw >> 1) & 1, c = ((sw & 1) * 4 + (lane >> 3)) * 8 + (lane & 7);
      float* dst = which == 0 ? a.dbq : (which == 1 ? a.dbk : a.dbv);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int hh = 2 * k + s_;
        if (k < a.hp && hh < a.heads) atomicAdd(dst + hh * 64 + c, bsum[k]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) ab_mark(a, 3);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int attn_bwd_tc_launch(const void* d_o, int lddo, const void* q, int ldq, const void* k, int ldk, const void* v, int ldv,
                       const float* lse, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv, int batch, int heads,
                       int Tq, int Tk, const int* kv_len, int causal, float scale, uint32_t thresh16, float inv_keep,
                       unsigned long long seed, float* dbq, float* dbk, float* dbv, const int* kv_index, int kv_batches,
                       cudaStream_t st) {
  AttnBwdMaps maps;
  const int nkv = kv_index ? kv_batches : batch;
  int rc = attn_make_map(&maps.q, q, (uint64_t)heads * 64, (uint64_t)batch * Tq, ldq);
  if (rc) return rc;
  rc = attn_make_map(&maps.dout, d_o, (uint64_t)heads * 64, (uint64_t)batch * Tq, lddo);
  if (rc) return rc;
  rc = attn_make_map(&maps.k, k, (uint64_t)heads * 64, (uint64_t)nkv * Tk, ldk);
  if (rc) return rc;
  rc = attn_make_map(&maps.v, v, (uint64_t)heads * 64, (uint64_t)nkv * Tk, ldv);
  if (rc) return rc;
  rc = attn_make_map3d(&maps.dq, dq, (uint64_t)heads * 64, Tq, batch, lddq);
  if (rc) return rc;
  rc = attn_make_map3d(&maps.dk, dk, (uint64_t)heads * 64, Tk, batch, lddk);
  if (rc) return rc;
  rc = attn_make_map3d(&maps.dv, dv, (uint64_t)heads * 64, Tk, batch, lddv);
  if (rc) return rc;
  AttnBwdArgs a{};
  a.lse = lse; a.batch = batch; a.heads = heads; a.Tq = Tq; a.Tk = Tk; a.kv_len = kv_len; a.causal = causal; a.scale = scale;
  a.seed = seed; a.thresh16 = thresh16; a.inv_keep = inv_keep; a.salt = spmm_g_rng_salt;
  a.pair = (Tq <= 64 && Tk <= 64) ? 1 : 0;
  a.hp = (heads + 1) / 2;
  a.num_tiles = a.pair ? batch * a.hp : batch * heads;
  a.dbq = dbq; a.dbk = dbk; a.dbv = dbv; a.kv_index = kv_index;
  a.trace = g_attn_trace;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int ctas = a.num_tiles < kNumSMs ? a.num_tiles : kNumSMs;
  cudaError_t le = a.pair ? launch_pdl(attn_bwd_tc_kernel<true>, dim3(ctas), dim3(AT_THREADS), AB_SMEM, st, maps, a)
                          : launch_pdl(attn_bwd_tc_kernel<false>, dim3(ctas), dim3(AT_THREADS), AB_SMEM, st, maps, a);
  if (le != cudaSuccess) return (int)le;
  return 0;
}

}  // namespace spmm
