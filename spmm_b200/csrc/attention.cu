// C entry points of the fused attention core (reference xbert.py:305-354: QK^T/8 + additive mask -> softmax -> dropout
// -> PV, with the head split / merge permutes :265-268,352-354 folded into the addressing): argument checks and
// dispatch to the tcgen05 kernels in attention_tc.cu.  Tq, Tk <= 128, head_dim 64; masks are generated in-kernel from
// kv_len (pad) and the causal flag; the [B,12,Tq,Tk] probability tensor of the reference is never materialised and the
// backward recomputes P from the saved log-sum-exp.  There is no second implementation: arguments the tensor-core
// kernels cannot take (unaligned rows / pointers) are rejected, not rerouted.
#include "common.cuh"
#include "spmm_b200.h"

namespace spmm {

void attn_set_trace(void* p);   // attention_tc.cu
int attn_fwd_tc_launch(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* o, int ldo, float* lse,
                       int batch, int heads, int Tq, int Tk, const int* kv_len, int causal, int kv_bstride, float scale,
                       uint32_t thresh16, float inv_keep, unsigned long long seed, const int* kv_index, int kv_batches,
                       cudaStream_t st);
int attn_bwd_tc_launch(const void* d_o, int lddo, const void* q, int ldq, const void* k, int ldk, const void* v, int ldv,
                       const float* lse, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv, int batch, int heads,
                       int Tq, int Tk, const int* kv_len, int causal, float scale, uint32_t thresh16, float inv_keep,
                       unsigned long long seed, float* dbq, float* dbk, float* dbv, const int* kv_index, int kv_batches,
                       cudaStream_t st);

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
                             const int* kv_index, int kv_batches, void* stream) {
  SPMM_ARG(kv_index == nullptr || kv_batches > 0);
  SPMM_ARG(q && k && v && o && batch > 0 && heads > 0 && Tq > 0 && Tk > 0 && Tq <= 128 && Tk <= 128);
  SPMM_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0);           // 16-byte rows: TMA tensor maps
  SPMM_ARG((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o) & 15) == 0);
  uint32_t th; float ik;
  attn_drop(dropout_p, th, ik);
  return attn_fwd_tc_launch(q, ldq, k, ldk, v, ldv, o, ldo, lse, batch, heads, Tq, Tk, kv_len, causal, kv_batch_stride_rows,
                            scale, th, ik, seed, kv_index, kv_batches, (cudaStream_t)stream);
}

extern "C" int spmm_attn_bwd(const void* d_o, int lddo, const void* q, int ldq, const void* k, int ldk, const void* v,
                             int ldv, const void* o, int ldo, const float* lse, void* dq, int lddq, void* dk, int lddk,
                             void* dv, int lddv, int batch, int heads, int Tq, int Tk, const int* kv_len, int causal,
                             float scale, float dropout_p, unsigned long long seed, float* dbias_q, float* dbias_k,
                             float* dbias_v, const int* kv_index, int kv_batches, void* stream) {
  SPMM_ARG(kv_index == nullptr || kv_batches > 0);
  SPMM_ARG(d_o && q && k && v && o && lse && dq && dk && dv);
  SPMM_ARG(batch > 0 && heads > 0 && Tq > 0 && Tk > 0 && Tq <= 128 && Tk <= 128);
  SPMM_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && lddo % 8 == 0 && ldo % 8 == 0 && lddq % 8 == 0 &&
           lddk % 8 == 0 && lddv % 8 == 0);
  SPMM_ARG((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)d_o | (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) & 15) == 0);
  SPMM_ARG((dbias_q == nullptr) == (dbias_k == nullptr) && (dbias_q == nullptr) == (dbias_v == nullptr));
  uint32_t th; float ik;
  attn_drop(dropout_p, th, ik);
  return attn_bwd_tc_launch(d_o, lddo, q, ldq, k, ldk, v, ldv, lse, dq, lddq, dk, lddk, dv, lddv, batch, heads, Tq, Tk, kv_len,
                            causal, scale, th, ik, seed, dbias_q, dbias_k, dbias_v, kv_index, kv_batches, (cudaStream_t)stream);
}
