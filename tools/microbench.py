"""Per-kernel micro-benchmarks on the shapes of one SPMM step (B=96, L=64): CUDA-event timing, L2 flushed between
iterations by cycling through > 126 MB of distinct buffers.  Usage: python tools/microbench.py [gemm|attn|ln|all]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spmm_b200 import kernels as K

DEV = "cuda"
BF = torch.bfloat16


def timeit(fn, iters=20, warm=3):
    """Device time per call: the calls are captured into one CUDA graph and replayed, so Python / ctypes / tensor-map
    encoding on the host does not show up (an eager loop of ~15 us kernels measures the host)."""
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3   # us


def gemm_cases():
    T = 6144
    H, I = 768, 3072
    cases = []
    for name, M, N, Kd, kw in [
        ("fwd qkv      ", T, 3 * H, H, dict(bias=True)),
        ("fwd out+res  ", T, H, H, dict(bias=True, residual=True, dropout_p=0.1)),
        ("fwd ffn-up   ", T, I, H, dict(bias=True, gelu=True, pre=True)),
        ("fwd ffn-down ", T, H, I, dict(bias=True, residual=True, dropout_p=0.1)),
        ("dgrad out    ", T, H, H, dict(b_mn=True)),
        ("dgrad ffn2+dg", T, I, H, dict(b_mn=True, dgelu=True)),
        ("dgrad ffn1+rs", T, H, I, dict(b_mn=True, residual=True)),
        ("dgrad qkv+res", T, H, 3 * H, dict(b_mn=True, residual=True)),
        ("wgrad 768x768", H, H, T, dict(a_mn=True, b_mn=True, wgrad=True)),
        ("wgrad qkv    ", 3 * H, H, T, dict(a_mn=True, b_mn=True, wgrad=True)),
        ("wgrad ffn1   ", I, H, T, dict(a_mn=True, b_mn=True, wgrad=True)),
        ("wgrad ffn2   ", H, I, T, dict(a_mn=True, b_mn=True, wgrad=True)),
        ("fwd 2B out   ", 2 * T, H, H, dict(bias=True, residual=True)),
        ("fwd P54 qkv  ", 5184, 3 * H, H, dict(bias=True)),
    ]:
        cases.append((name, M, N, Kd, kw))
    return cases


def bench_gemm(only=None):
    NB = 6   # rotate buffers so operands do not stay L2-resident artificially
    for ci, (name, M, N, Kd, kw) in enumerate(gemm_cases()):
        if only is not None and ci not in only:
            continue
        a_mn, b_mn = kw.get("a_mn", False), kw.get("b_mn", False)
        As = [torch.randn((Kd, M) if a_mn else (M, Kd), device=DEV).to(BF) for _ in range(NB)]
        Bs = [(torch.randn((Kd, N) if b_mn else (N, Kd), device=DEV) * 0.05).to(BF) for _ in range(NB)]
        wgrad = kw.get("wgrad", False)
        outs = [torch.zeros(M, N, device=DEV, dtype=torch.float32 if wgrad else BF) for _ in range(NB)]
        bias = torch.randn(N, device=DEV) if kw.get("bias") else None
        res = [torch.randn(M, N, device=DEV).to(BF) for _ in range(NB)] if kw.get("residual") else None
        pre = [torch.empty(M, N, device=DEV, dtype=BF) for _ in range(NB)] if kw.get("pre") else None
        dg = [torch.randn(M, N, device=DEV).to(BF) for _ in range(NB)] if kw.get("dgelu") else None

        def run(i):
            j = i % NB
            K.gemm(As[j], Bs[j], M, N, Kd, a_mn=a_mn, b_mn=b_mn, out=outs[j], out_f32=wgrad, accumulate=wgrad, bias=bias,
                   residual=res[j] if res else None, pre_act_out=pre[j] if pre else None, dgelu_pre=dg[j] if dg else None,
                   gelu=kw.get("gelu", False), dropout_p=kw.get("dropout_p", 0.0), seed=123)
        us = timeit(run)
        print("gemm %s M=%5d N=%4d K=%5d  %7.1f us  %6.1f TFLOP/s" % (name, M, N, Kd, us, 2.0 * M * N * Kd / us / 1e6))


def bench_attn():
    for (B, Tq, Tk, causal) in [(96, 64, 64, False), (96, 54, 54, False), (96, 54, 64, False), (96, 64, 54, False), (192, 64, 54, False), (96, 64, 64, True)]:
        H, h = 768, 12
        qkv = torch.randn(B * Tq, 3 * H, device=DEV).to(BF)
        kv = torch.randn(B * Tk, 2 * H, device=DEV).to(BF)
        q = qkv[:, :H]
        k, v = (qkv[:, H:2 * H], qkv[:, 2 * H:]) if Tq == Tk else (kv[:, :H], kv[:, H:])
        o = torch.empty(B * Tq, H, device=DEV, dtype=BF)
        lse = torch.empty(B * h * Tq, device=DEV)
        us_f = timeit(lambda i: K.attn_fwd(q, k, v, o, lse, B, h, Tq, Tk, None, causal, 0.125, 0.1, 5))
        do = torch.randn(B * Tq, H, device=DEV).to(BF)
        dq = torch.empty(B * Tq, H, device=DEV, dtype=BF)
        dkv = torch.empty(B * Tk, 2 * H, device=DEV, dtype=BF)
        us_b = timeit(lambda i: K.attn_bwd(do, q, k, v, o, lse, dq, dkv[:, :H], dkv[:, H:], B, h, Tq, Tk, None, causal, 0.125, 0.1, 5))
        print("attn B=%3d Tq=%3d Tk=%3d causal=%d  fwd %6.1f us  bwd %6.1f us" % (B, Tq, Tk, causal, us_f, us_b))


def bench_ln():
    for rows in (6144, 10368, 12288, 20736, 24576):
        H = 768
        x = torch.randn(rows, H, device=DEV).to(BF)
        g, b = torch.ones(H, device=DEV), torch.zeros(H, device=DEV)
        y, mean, rstd = K.layernorm_fwd(x, g, b, 1e-12)
        us_f = timeit(lambda i: K.layernorm_fwd(x, g, b, 1e-12))
        dy = torch.randn(rows, H, device=DEV).to(BF)
        dg, db, dbias = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
        us_b = timeit(lambda i: K.layernorm_bwd(dy, x, mean, rstd, g, dg, db, dbias=dbias, want_branch=True, branch_dropout_p=0.1, branch_seed=3))
        cs = torch.zeros(3 * H, device=DEV)
        xx = torch.randn(rows, 3 * H, device=DEV).to(BF)
        us_c = timeit(lambda i: K.colsum(xx, cs))
        print("ln rows=%5d  fwd %6.1f us  bwd(+branch,+dbias) %6.1f us  colsum[%d x 2304] %6.1f us" % (rows, us_f, us_b, rows, us_c))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    only = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else None
    if len(sys.argv) > 3:
        from spmm_b200 import _lib
        _lib.lib().spmm_gemm_debug_config(0, 0, int(sys.argv[4], 0) if len(sys.argv) > 4 else 0, int(sys.argv[3]))
        print("max_ctas", sys.argv[3])
    if what in ("gemm", "all"):
        bench_gemm(only)
    if what in ("attn", "all"):
        bench_attn()
    if what in ("ln", "all"):
        bench_ln()
