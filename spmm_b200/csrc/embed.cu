// Embedding-side kernels: BertEmbeddings pre-LN sums (reference xbert.py:193-217) and the property-vector
// tokeniser (SPMM_models.py:82-88).  Tiny, launch/latency-bound; fp32 tables in, bf16 activations out.
#include "common.cuh"
#include "spmm_b200.h"

namespace spmm {

__global__ void embed_text_fwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ word,
                                      const float* __restrict__ pos, const float* __restrict__ type0,
                                      __nv_bfloat16* __restrict__ x, int rows, int T, int H) {
  const int row = blockIdx.x;
  const int t = row % T;
  const int64_t id = ids[row];
  for (int c = threadIdx.x * 2; c < H; c += blockDim.x * 2) {
    const float2 w = *reinterpret_cast<const float2*>(word + id * H + c);
    const float2 p = *reinterpret_cast<const float2*>(pos + (size_t)t * H + c);
    const float2 ty = *reinterpret_cast<const float2*>(type0 + c);
    // same association as the reference: (inputs_embeds + token_type) + position (xbert.py:214-217)
    *reinterpret_cast<uint32_t*>(x + (size_t)row * H + c) = pack_bf16x2((w.x + ty.x) + p.x, (w.y + ty.y) + p.y);
  }
}

// CTA (t, s): position t, batch slice s (gridDim.y slices: a single CTA per position walked the whole batch serially,
// 95 us for 96 x 64 tokens); partial sums meet in dpos / dtype0 / dword through atomics
__global__ void embed_bwd_kernel(const __nv_bfloat16* __restrict__ dx, const int64_t* __restrict__ ids, float* dword,
                                 float* dpos, float* dtype0, int batch, int T, int H, int pad_id) {
  const int t = blockIdx.x;
  const int per = (batch + gridDim.y - 1) / gridDim.y;
  const int b0 = blockIdx.y * per, b1 = min(batch, b0 + per);
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float acc = 0.f;
    for (int b = b0; b < b1; ++b) {
      const size_t row = (size_t)b * T + t;
      const float v = bf2f(dx[row * H + c]);
      acc += v;
      if (dword != nullptr) {
        const int64_t id = ids[row];
        if (id != pad_id) atomicAdd(dword + id * H + c, v);  // padding_idx row receives no embedding-path grad
      }
    }
    atomicAdd(dpos + (size_t)t * H + c, acc);
    atomicAdd(dtype0 + c, acc);
  }
}

__global__ void pv_tokens_fwd_kernel(const float* __restrict__ pv, const float* __restrict__ mpm,
                                     const float* __restrict__ w, const float* __restrict__ bias,
                                     const float* __restrict__ cls, const float* __restrict__ mtok,
                                     __nv_bfloat16* __restrict__ out, int n_prop, int H) {
  const int b = blockIdx.x / (n_prop + 1), j = blockIdx.x % (n_prop + 1);
  __nv_bfloat16* o = out + (size_t)blockIdx.x * H;
  if (j == 0) {
    for (int c = threadIdx.x; c < H; c += blockDim.x) o[c] = f2bf(cls[c]);
    return;
  }
  const float val = pv[b * n_prop + j - 1], m = mpm[b * n_prop + j - 1];
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    const float feat = val * w[c] + bias[c];
    o[c] = f2bf(feat * (1.f - m) + mtok[c] * m);  // SPMM_models.py:87
  }
}

// CTA (j, s): token position j (0 = cls), batch slice s; partial sums meet through atomics
__global__ void pv_tokens_bwd_kernel(const __nv_bfloat16* __restrict__ dprop, const float* __restrict__ pv,
                                     const float* __restrict__ mpm, float* dw, float* db, float* dcls, float* dmtok,
                                     int batch, int n_prop, int H) {
  const int j = blockIdx.x;
  const int per = (batch + gridDim.y - 1) / gridDim.y;
  const int b0 = blockIdx.y * per, b1 = min(batch, b0 + per);
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float aw = 0.f, ab = 0.f, am = 0.f, ac = 0.f;
    for (int b = b0; b < b1; ++b) {
      const float d = bf2f(dprop[((size_t)b * (n_prop + 1) + j) * H + c]);
      if (j == 0) { ac += d; continue; }
      const float m = mpm[b * n_prop + j - 1], val = pv[b * n_prop + j - 1];
      aw += d * (1.f - m) * val;
      ab += d * (1.f - m);
      am += d * m;
    }
    if (j == 0) atomicAdd(dcls + c, ac);
    else { atomicAdd(dw + c, aw); atomicAdd(db + c, ab); atomicAdd(dmtok + c, am); }
  }
}

__global__ void embed_inputs_fwd_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ pos,
                                        const float* __restrict__ type0, __nv_bfloat16* __restrict__ x, int T, int H) {
  const int row = blockIdx.x, t = row % T;
  for (int c = threadIdx.x * 2; c < H; c += blockDim.x * 2) {
    float a, b;
    unpack_bf16x2(*reinterpret_cast<const uint32_t*>(in + (size_t)row * H + c), a, b);
    const float2 p = *reinterpret_cast<const float2*>(pos + (size_t)t * H + c);
    const float2 ty = *reinterpret_cast<const float2*>(type0 + c);
    *reinterpret_cast<uint32_t*>(x + (size_t)row * H + c) = pack_bf16x2((a + ty.x) + p.x, (b + ty.y) + p.y);
  }
}

}  // namespace spmm
using namespace spmm;

extern "C" int spmm_embed_text_fwd(const int64_t* ids, const float* word, const float* pos, const float* type0, void* x,
                                   int rows, int T, int H, void* stream) {
  SPMM_ARG(ids && word && pos && type0 && x && rows > 0 && T > 0 && H % 2 == 0);
  embed_text_fwd_kernel<<<rows, 128, 0, (cudaStream_t)stream>>>(ids, word, pos, type0, (__nv_bfloat16*)x, rows, T, H);
  SPMM_CHECK_LAUNCH();
  return 0;
}
extern "C" int spmm_embed_text_bwd(const void* dx, const int64_t* ids, float* dword, float* dpos, float* dtype0,
                                   int rows, int T, int H, int pad_id, void* stream) {
  SPMM_ARG(dx && ids && dword && dpos && dtype0 && rows > 0 && T > 0 && rows % T == 0);
  embed_bwd_kernel<<<dim3(T, rows / T >= 16 ? 8 : 1), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dx, ids, dword,
                                                                                      dpos, dtype0, rows / T, T, H, pad_id);
  SPMM_CHECK_LAUNCH();
  return 0;
}
extern "C" int spmm_pv_tokens_fwd(const float* pv, const float* mpm_mask, const float* w_embed, const float* b_embed,
                                  const float* cls_tok, const float* mask_tok, void* properties, int batch, int n_prop,
                                  int H, void* stream) {
  SPMM_ARG(pv && mpm_mask && w_embed && b_embed && cls_tok && mask_tok && properties && batch > 0 && n_prop > 0);
  pv_tokens_fwd_kernel<<<batch * (n_prop + 1), 128, 0, (cudaStream_t)stream>>>(
      pv, mpm_mask, w_embed, b_embed, cls_tok, mask_tok, (__nv_bfloat16*)properties, n_prop, H);
  SPMM_CHECK_LAUNCH();
  return 0;
}
extern "C" int spmm_pv_tokens_bwd(const void* dproperties, const float* pv, const float* mpm_mask, float* dw_embed,
                                  float* db_embed, float* dcls, float* dmask_tok, int batch, int n_prop, int H,
                                  void* stream) {
  SPMM_ARG(dproperties && pv && mpm_mask && dw_embed && db_embed && dcls && dmask_tok && batch > 0 && n_prop > 0);
  pv_tokens_bwd_kernel<<<dim3(n_prop + 1, batch >= 16 ? 8 : 1), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dproperties, pv, mpm_mask, dw_embed, db_embed, dcls, dmask_tok, batch, n_prop, H);
  SPMM_CHECK_LAUNCH();
  return 0;
}
extern "C" int spmm_embed_inputs_fwd(const void* inputs, const float* pos, const float* type0, void* x, int rows, int T,
                                     int H, void* stream) {
  SPMM_ARG(inputs && pos && type0 && x && rows > 0 && T > 0 && H % 2 == 0);
  embed_inputs_fwd_kernel<<<rows, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)inputs, pos, type0,
                                                                  (__nv_bfloat16*)x, T, H);
  SPMM_CHECK_LAUNCH();
  return 0;
}
extern "C" int spmm_embed_inputs_bwd(const void* dx, float* dpos, float* dtype0, int rows, int T, int H, void* stream) {
  SPMM_ARG(dx && dpos && dtype0 && rows > 0 && T > 0 && rows % T == 0);
  embed_bwd_kernel<<<dim3(T, rows / T >= 16 ? 8 : 1), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dx, nullptr, nullptr,
                                                                                      dpos, dtype0, rows / T, T, H, -1);
  SPMM_CHECK_LAUNCH();
  return 0;
}
