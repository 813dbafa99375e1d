// Single-token decode kernels for PV -> SMILES beam search (reference d_pv2smiles_batched.py:17-59,
// d_pv2smiles_single.py:26-44).  The reference re-runs the whole causal 12-layer stack on the growing prefix for every
// new token; the stack is causal, so the keys / values of earlier positions never change and are cached here:
//   decode_embed      : x = word[token] + type[0] + pos[t]               (BertEmbeddings for ONE new position, xbert.py:193-217)
//   decode_attn_self  : append this step's K/V to the cache, attend the new query over positions 0..t of ITS beam
//                       (ancestor table instead of physically re-ordering the caches when beams are re-ranked);
//                       keys whose token id is 0 are masked, as `text_atts = where(text == 0, 0, 1)` does
//   decode_attn_cross : the new query against the 54 property tokens of its molecule (K/V projected once per molecule,
//                       shared by the molecule's beams)
//   beam_step         : softmax + top-k per live beam, the [SEP] / finished-list / -1e5 rule, top-k over the k*k
//                       candidates, new token rows, parents, ancestor rows - all on the device (no host sync per token)
// All per-step scalars (position t) are read from device memory, so ONE captured CUDA graph replays for every step.
// HBM-bound SIMT kernels: per step they touch the caches once (<= 100 keys x 128 B per head), the GEMMs stream the weights.
#include "common.cuh"
#include "spmm_b200.h"

namespace spmm {

constexpr int DEC_HD = 64;         // head dim
constexpr int DEC_MAXT = 128;      // max cached positions (the reference stops after 101 tokens)
constexpr float DEC_LOG2E = 1.4426950408889634f;

__global__ void decode_embed_kernel(const int64_t* __restrict__ ids, const int* __restrict__ t_dev,
                                    const float* __restrict__ word, const float* __restrict__ pos,
                                    const float* __restrict__ type0, __nv_bfloat16* __restrict__ x, int H) {
  const int row = blockIdx.x;
  const int t = __ldg(t_dev);
  const int64_t id = ids[row];
  for (int c = threadIdx.x * 2; c < H; c += blockDim.x * 2) {
    const float2 w = *reinterpret_cast<const float2*>(word + id * H + c);
    const float2 p = *reinterpret_cast<const float2*>(pos + (size_t)t * H + c);
    const float2 ty = *reinterpret_cast<const float2*>(type0 + c);
    *reinterpret_cast<uint32_t*>(x + (size_t)row * H + c) = pack_bf16x2((w.x + ty.x) + p.x, (w.y + ty.y) + p.y);
  }
}

// One warp per (row, head).  Scores: lane l owns keys l, l+32, l+64, l+96 (full 64-wide dot product each, the query in
// registers); output: lane l owns dims 2l, 2l+1 and walks the keys with the probabilities broadcast by shuffle.
//   SELF : keys live in cache[j][phys][H], phys = anc[row][j] for j < t and `row` for j == t (appended here first);
//          key j is masked when tokens[row][j] == 0
//   CROSS: keys are rows (row / group) * Tk + j of a [groups * Tk][ldkv] matrix; kv_len optional per group
template <bool SELF>
__global__ void __launch_bounds__(128)
decode_attn_kernel(const __nv_bfloat16* __restrict__ q, int ldq, const __nv_bfloat16* __restrict__ k_new,
                   const __nv_bfloat16* __restrict__ v_new, int ldkv, __nv_bfloat16* __restrict__ cache_k,
                   __nv_bfloat16* __restrict__ cache_v, const int* __restrict__ anc, const int64_t* __restrict__ tokens,
                   int tmax, const int* __restrict__ t_dev, const __nv_bfloat16* __restrict__ ck,
                   const __nv_bfloat16* __restrict__ cv, int Tk, int group, const int* __restrict__ kv_len,
                   __nv_bfloat16* __restrict__ out, int ldo, int rows, int heads, float scale) {
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= rows * heads) return;
  const int row = w / heads, h = w % heads;
  const int H = heads * DEC_HD;
  int nkeys;
  if (SELF) {
    const int t = __ldg(t_dev);
    nkeys = t + 1;
    // append this step's key / value (4 bytes per lane)
    const size_t dst = ((size_t)t * rows + row) * H + h * DEC_HD + 2 * lane;
    *reinterpret_cast<uint32_t*>(cache_k + dst) = *reinterpret_cast<const uint32_t*>(k_new + (size_t)row * ldkv + h * DEC_HD + 2 * lane);
    *reinterpret_cast<uint32_t*>(cache_v + dst) = *reinterpret_cast<const uint32_t*>(v_new + (size_t)row * ldkv + h * DEC_HD + 2 * lane);
    __syncwarp();
  } else {
    const int g = row / group;
    nkeys = kv_len != nullptr ? min(__ldg(kv_len + g), Tk) : Tk;
  }
  // query of this head, replicated in every lane (fp32)
  float qf[DEC_HD];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(q + (size_t)row * ldq + h * DEC_HD);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 u = __ldg(qp + i);
      unpack_bf16x2(u.x, qf[8 * i], qf[8 * i + 1]); unpack_bf16x2(u.y, qf[8 * i + 2], qf[8 * i + 3]);
      unpack_bf16x2(u.z, qf[8 * i + 4], qf[8 * i + 5]); unpack_bf16x2(u.w, qf[8 * i + 6], qf[8 * i + 7]);
    }
  }
  auto key_ptr = [&](const __nv_bfloat16* base_self, const __nv_bfloat16* base_cross, int j) -> const __nv_bfloat16* {
    if (SELF) {
      const int phys = (j == nkeys - 1) ? row : __ldg(anc + (size_t)row * tmax + j);
      return base_self + ((size_t)j * rows + phys) * H + h * DEC_HD;
    }
    return base_cross + ((size_t)(row / group) * Tk + j) * ldkv + h * DEC_HD;
  };
  const float c2 = scale * DEC_LOG2E;
  float s[DEC_MAXT / 32];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < DEC_MAXT / 32; ++i) {
    const int j = lane + 32 * i;
    s[i] = -INFINITY;
    if (j < nkeys) {
      bool ok = true;
      if (SELF) ok = tokens[(size_t)row * tmax + j] != 0;        // text_atts = where(text == 0, 0, 1)
      if (ok) {
        const uint4* kp = reinterpret_cast<const uint4*>(key_ptr(cache_k, ck, j));
        float acc = 0.f;
#pragma unroll
        for (int u8 = 0; u8 < 8; ++u8) {
          const uint4 u = kp[u8];
          float a, b;
          unpack_bf16x2(u.x, a, b); acc += qf[8 * u8] * a + qf[8 * u8 + 1] * b;
          unpack_bf16x2(u.y, a, b); acc += qf[8 * u8 + 2] * a + qf[8 * u8 + 3] * b;
          unpack_bf16x2(u.z, a, b); acc += qf[8 * u8 + 4] * a + qf[8 * u8 + 5] * b;
          unpack_bf16x2(u.w, a, b); acc += qf[8 * u8 + 6] * a + qf[8 * u8 + 7] * b;
        }
        s[i] = acc * c2;
      }
    }
    mx = fmaxf(mx, s[i]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < DEC_MAXT / 32; ++i) {
    s[i] = (s[i] == -INFINITY) ? 0.f : exp2f(s[i] - mx);
    sum += s[i];
  }
  sum = warp_sum(sum);
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  float o0 = 0.f, o1 = 0.f;
#pragma unroll
  for (int i = 0; i < DEC_MAXT / 32; ++i) {
    if (32 * i >= nkeys) break;                                  // warp-uniform
    const int lim = min(32, nkeys - 32 * i);
    for (int jj = 0; jj < lim; ++jj) {
      const float p = __shfl_sync(0xffffffffu, s[i], jj);
      if (p != 0.f) {                                            // warp-uniform (same value in every lane)
        const int j = 32 * i + jj;
        const uint32_t u = *reinterpret_cast<const uint32_t*>(key_ptr(cache_v, cv, j) + 2 * lane);
        float a, b;
        unpack_bf16x2(u, a, b);
        o0 += p * a; o1 += p * b;
      }
    }
  }
  *reinterpret_cast<uint32_t*>(out + (size_t)row * ldo + h * DEC_HD + 2 * lane) = pack_bf16x2(o0 * inv, o1 * inv);
}

// ---------------------------------------------------------------------------------------------------- beam bookkeeping
// One CTA (128 threads) per molecule.  Rows m*k .. m*k + k - 1 are the molecule's live beams.
struct BeamArgs {
  const __nv_bfloat16* logits; int ld, V;
  int k, tmax, n_mol, fin_cap;
  int cls_id, sep_id;
  int* t_dev;               // step counter (position of the token whose logits these are); advanced by the last CTA
  float* scores;            // [n_mol][k]      accumulated log-probability of each live beam
  int64_t* tokens;          // [n_mol*k][tmax] live prefixes
  int* anc;                 // [n_mol*k][tmax] physical cache row of every cached position
  int64_t* next_ids;        // [n_mol*k]       token fed to the next step
  float* fin_scores;        // [n_mol][fin_cap]
  int64_t* fin_tokens;      // [n_mol][fin_cap][tmax]
  int* fin_len;             // [n_mol][fin_cap]
  int* fin_count;           // [n_mol]
  int* done;                // [n_mol]
  float* trace_logp;        // optional [steps][n_mol*k][k] chosen log-probs
  int* trace_tok;           // optional [steps][n_mol*k][k] chosen tokens
  unsigned int* ticket;     // CTA completion counter (self-resetting)
};

constexpr int BEAM_MAXK = 8;

__global__ void __launch_bounds__(128) beam_step_kernel(const BeamArgs a) {
  __shared__ float cand_lp[BEAM_MAXK][BEAM_MAXK];
  __shared__ int cand_tok[BEAM_MAXK][BEAM_MAXK];
  __shared__ float sel_score[BEAM_MAXK];
  __shared__ int sel_flat[BEAM_MAXK];
  __shared__ int64_t par_tok[BEAM_MAXK][DEC_MAXT];
  __shared__ int par_anc[BEAM_MAXK][DEC_MAXT];
  __shared__ int s_done;
  const int m = blockIdx.x, k = a.k;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = *a.t_dev;                      // the prefix fed this step has t + 1 tokens; we pick token t + 1
  const bool first = t == 0;                   // d_pv2smiles_batched.py:29-32: a single [CLS] beam, no [SEP] rule yet
  const bool active = a.done[m] == 0;
  // ---- softmax + top-k per live beam (warp per beam, d_pv2smiles_single.py:37-44)
  for (int b = warp; b < k; b += 4) {
    const __nv_bfloat16* lr = a.logits + (size_t)(m * k + b) * a.ld;
    float v[10];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < a.V ? bf2f(lr[c]) : -INFINITY;
      mx = fmaxf(mx, v[i]);
    }
    mx = warp_max(mx);
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) se += (v[i] == -INFINITY) ? 0.f : __expf(v[i] - mx);
    se = warp_sum(se);
    const float lse = mx + __logf(se);
    for (int r = 0; r < k; ++r) {              // k rounds of warp arg-max (ties -> lowest token id)
      float best = -INFINITY;
      int bi = 0x7fffffff;
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const int c = lane + 32 * i;
        if (v[i] > best) { best = v[i]; bi = c; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) { cand_lp[b][r] = best - lse; cand_tok[b][r] = bi; }
#pragma unroll
      for (int i = 0; i < 10; ++i)
        if (lane + 32 * i == bi) v[i] = -INFINITY;
    }
  }
  __syncthreads();
  if (a.trace_logp != nullptr && threadIdx.x < k * k) {
    const int b = threadIdx.x / k, r = threadIdx.x % k;
    const size_t o = ((size_t)t * a.n_mol * k + (size_t)(m * k + b)) * k + r;
    a.trace_logp[o] = cand_lp[b][r];
    a.trace_tok[o] = cand_tok[b][r];
  }
  // ---- candidates, finished list, next beams (thread 0; k*k <= 64 entries)
  if (threadIdx.x == 0) {
    int done = active ? 0 : 1;
    if (active) {
      float sc[BEAM_MAXK][BEAM_MAXK];
      for (int b = 0; b < k; ++b)
        for (int r = 0; r < k; ++r)
          sc[b][r] = (first && b > 0) ? -INFINITY : a.scores[m * k + b] + cand_lp[b][r];
      if (!first) {
        int cnt = a.fin_count[m];
        for (int b = 0; b < k; ++b)
          for (int r = 0; r < k; ++r)
            if (cand_tok[b][r] == a.sep_id) {  // row-major order of (indices == sep).nonzero(), d_pv2smiles_batched.py:39-44
              if (cnt < a.fin_cap) {
                a.fin_scores[(size_t)m * a.fin_cap + cnt] = sc[b][r];
                a.fin_len[(size_t)m * a.fin_cap + cnt] = -(b + 1);   // parent beam, resolved below (tokens copied by all threads)
                ++cnt;
              }
              sc[b][r] = -1e5f;
            }
        a.fin_count[m] = cnt;
        if (cnt >= k) done = 1;                // `if len(final_output) >= k: break`
      }
      if (!done) {
        for (int j = 0; j < k; ++j) {          // topk(k2_p.flatten(), k): ties -> lowest flat index
          float best = -INFINITY;
          int bf = 0;
          for (int f = 0; f < k * k; ++f) {
            const float x = sc[f / k][f % k];
            if (x > best) { best = x; bf = f; }
          }
          sel_score[j] = best;
          sel_flat[j] = bf;
          sc[bf / k][bf % k] = -INFINITY;
        }
      }
      if (t + 2 >= a.tmax) done = 1;           // out of cache positions
    }
    s_done = done;
  }
  // parents' rows -> shared (every thread), so children can be written in place
  for (int i = threadIdx.x; i < k * a.tmax; i += blockDim.x) {
    const int b = i / a.tmax, j = i % a.tmax;
    par_tok[b][j] = a.tokens[(size_t)(m * k + b) * a.tmax + j];
    par_anc[b][j] = a.anc[(size_t)(m * k + b) * a.tmax + j];
  }
  __syncthreads();
  if (active) {
    // finished sequences recorded this step: prefix of the parent beam + [SEP]
    const int cnt = a.fin_count[m];
    for (int e = 0; e < cnt; ++e) {
      const int tag = a.fin_len[(size_t)m * a.fin_cap + e];
      if (tag < 0) {
        const int b = -tag - 1;
        int64_t* dst = a.fin_tokens + ((size_t)m * a.fin_cap + e) * a.tmax;
        for (int j = threadIdx.x; j < a.tmax; j += blockDim.x) dst[j] = j <= t ? par_tok[b][j] : (j == t + 1 ? (int64_t)a.sep_id : 0);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int e = 0; e < cnt; ++e)
        if (a.fin_len[(size_t)m * a.fin_cap + e] < 0) a.fin_len[(size_t)m * a.fin_cap + e] = t + 2;
    if (!s_done) {
      for (int i = threadIdx.x; i < k * a.tmax; i += blockDim.x) {
        const int jb = i / a.tmax, j = i % a.tmax;
        const int f = sel_flat[jb], pb = f / k;
        int64_t tok = j <= t ? par_tok[pb][j] : 0;
        int an = j < t ? par_anc[pb][j] : 0;
        if (j == t) an = m * k + pb;                      // position t was computed (and cached) by the parent's row
        if (j == t + 1) tok = cand_tok[pb][f % k];
        a.tokens[(size_t)(m * k + jb) * a.tmax + j] = tok;
        a.anc[(size_t)(m * k + jb) * a.tmax + j] = an;
      }
      if (threadIdx.x < k) {
        const int f = sel_flat[threadIdx.x];
        a.scores[m * k + threadIdx.x] = sel_score[threadIdx.x];
        a.next_ids[m * k + threadIdx.x] = cand_tok[f / k][f % k];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (active && s_done) a.done[m] = 1;
    __threadfence();
    if (atomicInc(a.ticket, gridDim.x - 1) == gridDim.x - 1) *a.t_dev = t + 1;   // last CTA: every CTA has read t
  }
}

}  // namespace spmm
using namespace spmm;

extern "C" int spmm_decode_embed(const int64_t* ids, const int* t_dev, const float* word, const float* pos,
                                 const float* type0, void* x, int rows, int H, void* stream) {
  SPMM_ARG(ids && t_dev && word && pos && type0 && x && rows > 0 && H > 0 && H % 2 == 0);
  decode_embed_kernel<<<rows, 128, 0, (cudaStream_t)stream>>>(ids, t_dev, word, pos, type0, (__nv_bfloat16*)x, H);
  SPMM_CHECK_LAUNCH();
  return 0;
}

extern "C" int spmm_decode_attn_self(const void* q, int ldq, const void* k_new, const void* v_new, int ldkv, void* cache_k,
                                     void* cache_v, const int* anc, const int64_t* tokens, int tmax, const int* t_dev,
                                     void* out, int ldo, int rows, int heads, float scale, void* stream) {
  SPMM_ARG(q && k_new && v_new && cache_k && cache_v && anc && tokens && t_dev && out);
  SPMM_ARG(rows > 0 && heads > 0 && tmax > 0 && tmax <= DEC_MAXT && ldq % 8 == 0 && ldkv % 2 == 0 && ldo % 2 == 0);
  SPMM_ARG((((uintptr_t)q | (uintptr_t)cache_k | (uintptr_t)cache_v) & 15) == 0);
  const int warps = rows * heads;
  decode_attn_kernel<true><<<(warps + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k_new, (const __nv_bfloat16*)v_new, ldkv, (__nv_bfloat16*)cache_k,
      (__nv_bfloat16*)cache_v, anc, tokens, tmax, t_dev, nullptr, nullptr, 0, 1, nullptr, (__nv_bfloat16*)out, ldo, rows,
      heads, scale);
  SPMM_CHECK_LAUNCH();
  return 0;
}

extern "C" int spmm_decode_attn_cross(const void* q, int ldq, const void* k, const void* v, int ldkv, int Tk, int group,
                                      const int* kv_len, void* out, int ldo, int rows, int heads, float scale,
                                      void* stream) {
  SPMM_ARG(q && k && v && out && rows > 0 && heads > 0 && Tk > 0 && Tk <= DEC_MAXT && group > 0);
  SPMM_ARG(ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 2 == 0);
  SPMM_ARG((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) & 15) == 0);
  const int warps = rows * heads;
  decode_attn_kernel<false><<<(warps + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)q, ldq, nullptr, nullptr, ldkv, nullptr, nullptr, nullptr, nullptr, 0, nullptr,
      (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, Tk, group, kv_len, (__nv_bfloat16*)out, ldo, rows, heads, scale);
  SPMM_CHECK_LAUNCH();
  return 0;
}

extern "C" int spmm_beam_step(const void* logits, int ld, int V, int k, int tmax, int n_mol, int fin_cap, int cls_id,
                              int sep_id, int* t_dev, float* scores, int64_t* tokens, int* anc, int64_t* next_ids,
                              float* fin_scores, int64_t* fin_tokens, int* fin_len, int* fin_count, int* done,
                              float* trace_logp, int* trace_tok, unsigned int* ticket, void* stream) {
  SPMM_ARG(logits && t_dev && scores && tokens && anc && next_ids && fin_scores && fin_tokens && fin_len && fin_count &&
           done && ticket);
  SPMM_ARG(k >= 1 && k <= BEAM_MAXK && V > 0 && V <= 320 && ld >= V && tmax > 2 && tmax <= DEC_MAXT && n_mol > 0 &&
           fin_cap >= k);
  SPMM_ARG((trace_logp == nullptr) == (trace_tok == nullptr));
  BeamArgs a{};
  a.logits = (const __nv_bfloat16*)logits; a.ld = ld; a.V = V; a.k = k; a.tmax = tmax; a.n_mol = n_mol; a.fin_cap = fin_cap;
  a.cls_id = cls_id; a.sep_id = sep_id; a.t_dev = t_dev; a.scores = scores; a.tokens = tokens; a.anc = anc;
  a.next_ids = next_ids; a.fin_scores = fin_scores; a.fin_tokens = fin_tokens; a.fin_len = fin_len;
  a.fin_count = fin_count; a.done = done; a.trace_logp = trace_logp; a.trace_tok = trace_tok; a.ticket = ticket;
  beam_step_kernel<<<n_mol, 128, 0, (cudaStream_t)stream>>>(a);
  SPMM_CHECK_LAUNCH();
  return 0;
}
