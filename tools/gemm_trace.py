"""Phase timeline of the 2-CTA GEMM kernel (per-CTA %globaltimer stamps, spmm_gemm_debug_trace).
Usage: python tools/gemm_trace.py   -> per shape: median/max over CTAs of each phase, ns relative to the first CTA entry."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spmm_b200 import kernels as K, _lib

DEV, BF = "cuda", torch.bfloat16
NAMES = ["entry", "setup done", "1st TMA issued", "1st full", "tile0 MMAs committed", "tile0 tfull seen", "tile0 epi done",
         "pre-exit sync", "exit", "producer done", "last epi done", "-", "tile0 sub0 chunks done", "tile0 sub0 barrier passed",
         "tile0 sub1 start (prev store read)", "tile0 sub1 chunks done"]
T, H, I = 6144, 768, 3072
CASES = [("dgrad out (plain)", T, H, H, dict(b_mn=True)),
         ("fwd out+res+drop", T, H, H, dict(bias=True, residual=True, dropout_p=0.1)),
         ("fwd qkv", T, 3 * H, H, dict(bias=True)),
         ("fwd ffn-up gelu+pre", T, I, H, dict(bias=True, gelu=True, pre=True)),
         ("fwd ffn-down", T, H, I, dict(bias=True, residual=True, dropout_p=0.1)),
         ("dgrad ffn2+dgelu", T, I, H, dict(b_mn=True, dgelu=True)),
         ("dgrad ffn2+dgelu+colsum", T, I, H, dict(b_mn=True, dgelu=True, colsum=True)),
         ("wgrad ffn1", I, H, T, dict(a_mn=True, b_mn=True, wgrad=True))]
trace = torch.zeros(148 * 16, dtype=torch.int64, device=DEV)
for name, M, N, Kd, kw in CASES:
    a_mn, b_mn, wgrad = kw.get("a_mn", False), kw.get("b_mn", False), kw.get("wgrad", False)
    A = torch.randn((Kd, M) if a_mn else (M, Kd), device=DEV).to(BF)
    B = (torch.randn((Kd, N) if b_mn else (N, Kd), device=DEV) * 0.05).to(BF)
    out = torch.zeros(M, N, device=DEV, dtype=torch.float32 if wgrad else BF)
    bias = torch.randn(N, device=DEV) if kw.get("bias") else None
    res = torch.randn(M, N, device=DEV).to(BF) if kw.get("residual") else None
    pre = torch.empty(M, N, device=DEV, dtype=BF) if kw.get("pre") else None
    dg = torch.randn(M, N, device=DEV).to(BF) if kw.get("dgelu") else None
    cs = torch.zeros(N, device=DEV) if kw.get("colsum") else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)

    def run():
        K.gemm(A, B, M, N, Kd, a_mn=a_mn, b_mn=b_mn, out=out, out_f32=wgrad, accumulate=wgrad, bias=bias, residual=res,
               pre_act_out=pre, dgelu_pre=dg, gelu=kw.get("gelu", False), dropout_p=kw.get("dropout_p", 0.0), seed=123,
               dgelu_stored=bool(kw.get("dgelu") or kw.get("pre")), colsum_out=cs)
    for _ in range(3):
        run()
    for cold in (0, 1):
        if cold:
            flush.zero_()
        torch.cuda.synchronize()
        trace.zero_()
        _lib.lib().spmm_gemm_debug_trace(trace.data_ptr())
        run()
        torch.cuda.synchronize()
        _lib.lib().spmm_gemm_debug_trace(None)
        t = trace.view(148, 16).cpu()
        used = t[:, 0] > 0
        t0 = int(t[used, 0].min())
        print("%s M=%d N=%d K=%d  CTAs=%d  %s  total %.2f us" % (name, M, N, Kd, int(used.sum()), "COLD-L2" if cold else "warm", (int(t[used, 8].max()) - t0) / 1e3))
        for s, nm in enumerate(NAMES):
            col = t[used, s]
            col = col[col > 0]
            if len(col):
                rel = (col - t0).float() / 1e3
                print("   %-22s n=%3d  min %7.2f  med %7.2f  max %7.2f us" % (nm, len(col), rel.min(), rel.median(), rel.max()))
