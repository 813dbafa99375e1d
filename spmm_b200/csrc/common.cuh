// Shared device helpers for the sm_100a kernels (inline PTX wrappers, reductions, RNG).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define SPMM_CHECK_LAUNCH()                                   \
  do {                                                        \
    cudaError_t e__ = cudaPeekAtLastError();                  \
    if (e__ != cudaSuccess) return (int)e__;                  \
  } while (0)

#define SPMM_ARG(cond) \
  do {                 \
    if (!(cond)) return -1; \
  } while (0)

// Device scalar added to every dropout / sampler seed (set with spmm_set_rng_salt_ptr).  A training step bumps it
// once, so a CUDA graph of the step replays with fresh randomness although the per-op seeds are baked in.
extern const unsigned long long* spmm_g_rng_salt;

namespace spmm {

constexpr int kNumSMs = 148;

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// The hot kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel lets its successor
// start launching right away (pdl_trigger at its top) and waits for its own predecessor to finish and flush
// (pdl_wait) before its first global-memory access.  Launch latency, barrier/TMEM set-up and descriptor prefetch of
// kernel N+1 then overlap the tail of kernel N -- a step is ~1600 kernels of 5-60 us.  Both are no-ops when the
// kernel was launched without the attribute.  SPMM_PDL=0 in the environment turns the attribute off.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();   // misc.cu
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ unsigned long long salted(unsigned long long seed, const unsigned long long* salt) {
  return salt ? seed + __ldg(salt) * 0x9E3779B97F4A7C15ull : seed;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug becomes a trap (reported as a launch failure) instead of a hang.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) { __trap(); }
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* desc, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 columns of fp32: thread t of the warp gets lane (base_lane + t), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (sm_100): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// UMMA instruction descriptor, kind::f16, bf16 inputs, fp32 accumulate.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------- math / reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum; `sh` must hold >= 32 floats; result valid in every thread
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float r = (lane < nw) ? sh[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below bf16 output resolution): 2 MUFU + ~10 FMA,
// branch-free -- the GEMM epilogue evaluates it 32768 times per tile.
__device__ __forceinline__ float fast_erf(float x, float* exp_neg_x2) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
  const float e = __expf(-ax * ax);
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = 1.f - p * t * e;
  if (exp_neg_x2) *exp_neg_x2 = e;
  return copysignf(r, x);
}
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.f + fast_erf(x * 0.70710678118654752f, nullptr));
}
// gelu(x) and gelu'(x) from ONE erf evaluation (the A-S form yields exp(-x^2/2) as a by-product)
__device__ __forceinline__ float gelu_and_grad(float x, float* grad) {
  float e;
  const float cdf = 0.5f * (1.f + fast_erf(x * 0.70710678118654752f, &e));
  *grad = fmaf(x * 0.3989422804014327f, e, cdf);
  return x * cdf;
}
__device__ __forceinline__ float dgelu_erf(float x) {
  float e;  // exp(-(x/sqrt2)^2) = exp(-x^2/2)
  const float cdf = 0.5f * (1.f + fast_erf(x * 0.70710678118654752f, &e));
  return fmaf(x * 0.3989422804014327f, e, cdf);
}

// ---------------------------------------------------------------- packed fp32 pairs (sm_100: FFMA2 / FMUL2 / FADD2)
// Blackwell issues two independent fp32 FMAs per instruction on a 64-bit register pair.  The GEMM epilogues are bound by
// the issue slots of their few warps (TMEM reads leave room for ~8 warps), so their per-element math runs on pairs.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 splat2(float c) { return pk2(c, c); }
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// bf16x2 word -> fp32 pair (lo = element 0): a shift and a mask, no conversion instruction
__device__ __forceinline__ f32x2 bf16x2_to_f32x2(uint32_t u) {
  return pk2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u));
}
__device__ __forceinline__ uint32_t f32x2_to_bf16x2(f32x2 v) {
  float lo, hi;
  upk2(v, lo, hi);
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// gelu(x) = x Phi(x) and gelu'(x) = Phi(x) + x phi(x) for a PAIR of pre-activations, exact-erf form (ACT2FN["gelu"],
// reference xbert.py:430) through Abramowitz-Stegun 7.1.26 (|err(erf)| <= 1.5e-7): with z = |x|/sqrt2, t = 1/(1 + p z),
// h = 0.5 erfc(z) = 0.5 t (a1 + t (a2 + ... )) exp(-z^2);  Phi(x) = 0.5 + copysign(0.5 - h, x).
// 2 MUFU (rcp, ex2) per element + 19 pair-wide issue slots per PAIR (the scalar version cost ~35 per element).
__device__ __forceinline__ void gelu_and_grad2(f32x2 x, f32x2& act, f32x2& grad) {
  const f32x2 ax = x & 0x7FFFFFFF7FFFFFFFull;
  const f32x2 d = fma2(ax, splat2(0.3275911f * 0.70710678118654752f), splat2(1.f));
  float d0, d1;
  upk2(d, d0, d1);
  const f32x2 t = pk2(rcp_approx(d0), rcp_approx(d1));
  const f32x2 xc = mul2(x, splat2(-0.72134752044448170f));      // -0.5 log2(e) x
  const f32x2 y = mul2(xc, x);                                    // -x^2/2 in log2 units
  float y0, y1;
  upk2(y, y0, y1);
  const f32x2 e = pk2(ex2_approx(y0), ex2_approx(y1));            // exp(-x^2/2)
  f32x2 pl = fma2(t, splat2(0.5f * 1.061405429f), splat2(0.5f * -1.453152027f));
  pl = fma2(pl, t, splat2(0.5f * 1.421413741f));
  pl = fma2(pl, t, splat2(0.5f * -0.284496736f));
  pl = fma2(pl, t, splat2(0.5f * 0.254829592f));
  const f32x2 h = mul2(mul2(pl, t), e);                           // 0.5 erfc(|x|/sqrt2) in [0, 0.5]
  f32x2 q = fma2(h, splat2(-1.f), splat2(0.5f));                  // 0.5 - h >= 0
  q |= x & 0x8000000080000000ull;                                 // copysign(0.5 - h, x)
  const f32x2 cdf = add2(q, splat2(0.5f));
  act = mul2(x, cdf);
  // x phi(x) = x exp(-x^2/2) / sqrt(2 pi) = (xc e) * (1/sqrt(2 pi)) / (-0.5 log2 e)
  grad = fma2(mul2(xc, e), splat2(0.3989422804014327f / -0.72134752044448170f), cdf);
}

// Counter-based RNG for dropout: one 32-bit draw per (seed, stream, index).  splitmix-style finaliser.
__device__ __forceinline__ uint32_t hash_u32(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}
// Dropout keep test shared by every kernel that applies / re-applies a mask: 16 random bits per element, two elements
// per 32-bit mix (murmur3 finaliser over (seed, pair index)); identical in forward and backward by construction.
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
// 64-bit (seed + step salt) -> 32-bit key, once per kernel
__device__ __forceinline__ uint32_t fold_seed(unsigned long long s) {
  return mix32((uint32_t)s ^ mix32((uint32_t)(s >> 32) + 0x9E3779B9u));
}
// 2 x 16 random bits for elements (e_even, e_even + 1): ONE mix per pair of elements
__device__ __forceinline__ uint32_t drop_bits2(uint32_t key, uint32_t e_even) {
  return mix32((e_even >> 1) * 0x9E3779B1u + key);
}
__device__ __forceinline__ bool keep16(unsigned long long seed, unsigned long long e, uint32_t thresh16) {
  const uint32_t h = drop_bits2(fold_seed(seed), (uint32_t)e & ~1u);
  return ((h >> (16 * (e & 1))) & 0xFFFFu) >= thresh16;
}
// apply the mask to an (even, odd) element pair
__device__ __forceinline__ void drop_pair(uint32_t key, uint32_t e_even, uint32_t thresh16, float inv_keep, float& a,
                                          float& b) {
  const uint32_t h = drop_bits2(key, e_even);
  a = ((h & 0xFFFFu) >= thresh16) ? a * inv_keep : 0.f;
  b = ((h >> 16) >= thresh16) ? b * inv_keep : 0.f;
}
// keep-probability test: returns scale (1/(1-p)) or 0
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint64_t idx, uint32_t thresh, float inv_keep) {
  return (hash_u32(seed, idx) >= thresh) ? inv_keep : 0.f;
}

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ __nv_bfloat16 f2bf(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t u, float& lo, float& hi) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  lo = __low2float(v);
  hi = __high2float(v);
}

}  // namespace spmm
