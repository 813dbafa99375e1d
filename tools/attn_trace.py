"""Phase timeline of the tcgen05 attention forward kernel (per-CTA %globaltimer stamps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spmm_b200 import kernels as K, _lib
DEV, BF = "cuda", torch.bfloat16
names = {0: "entry", 1: "setup done", 2: "softmax warps done", 3: "exit"}
for n in range(4):
    for k, nm in enumerate(["wait S", "S ready", "pass1 done", "pass2 done", "arrived", "prev O epilogue done"]):
        names[4 + 6 * n + k] = "tile%d %s" % (n, nm)
for (B, Tq, Tk, p) in [(96, 64, 64, 0.1), (96, 64, 64, 0.0), (288, 64, 54, 0.1)]:
    H, h = 768, 12
    qkv = torch.randn(B * Tq, 3 * H, device=DEV).to(BF)
    kv = torch.randn(B * Tk, 2 * H, device=DEV).to(BF)
    q = qkv[:, :H]
    k, v = (qkv[:, H:2 * H], qkv[:, 2 * H:]) if Tq == Tk else (kv[:, :H], kv[:, H:])
    o = torch.empty(B * Tq, H, device=DEV, dtype=BF)
    lse = torch.empty(B * h * Tq, device=DEV)
    for _ in range(3):
        K.attn_fwd(q, k, v, o, lse, B, h, Tq, Tk, None, False, 0.125, p, 5)
    trace = torch.zeros(148 * 32, dtype=torch.int64, device=DEV)
    torch.cuda.synchronize()
    _lib.lib().spmm_attn_debug_trace(trace.data_ptr())
    K.attn_fwd(q, k, v, o, lse, B, h, Tq, Tk, None, False, 0.125, p, 5)
    torch.cuda.synchronize()
    _lib.lib().spmm_attn_debug_trace(None)
    t = trace.view(148, 32).cpu()
    t0 = int(t[:, 0][t[:, 0] > 0].min())
    print("B=%d Tq=%d Tk=%d dropout=%.1f  total %.2f us" % (B, Tq, Tk, p, (int(t[:, 3].max()) - t0) / 1e3))
    for s in sorted(names):
        col = t[:, s]; col = col[col > 0]
        if len(col):
            rel = (col - t0).float() / 1e3
            print("   %-28s n=%3d  min %6.2f  med %6.2f  max %6.2f us" % (names[s], len(col), rel.min(), rel.median(), rel.max()))


# ---------------------------------------------------------------- backward kernel (tiles 1..3 of every CTA)
bnames = {0: "entry", 3: "exit"}
for n in range(3):
    for k, nm in enumerate(["S,dP ready", "P/D pass done (barrier)", "Pd,dS written", "dQ,dK,dV ready", "accumulators read",
                            "staged (barrier)", "stores issued", "bias sums done"]):
        bnames[4 + 8 * n + k] = "tile%d %s" % (n + 1, nm)
for (B, Tq, Tk, p, bias) in [(384, 64, 64, 0.1, True), (384, 64, 64, 0.1, False), (384, 54, 64, 0.1, True), (192, 54, 54, 0.1, True)]:
    H, h = 768, 12
    qkv = torch.randn(B * Tq, 3 * H, device=DEV).to(BF)
    kv = torch.randn(B * Tk, 2 * H, device=DEV).to(BF)
    q = qkv[:, :H]
    k, v = (qkv[:, H:2 * H], qkv[:, 2 * H:]) if Tq == Tk else (kv[:, :H], kv[:, H:])
    o = torch.empty(B * Tq, H, device=DEV, dtype=BF)
    lse = torch.empty(B * h * Tq, device=DEV)
    K.attn_fwd(q, k, v, o, lse, B, h, Tq, Tk, None, False, 0.125, p, 5)
    do = torch.randn(B * Tq, H, device=DEV).to(BF)
    dq = torch.empty(B * Tq, H, device=DEV, dtype=BF)
    dkv = torch.empty(B * Tk, 2 * H, device=DEV, dtype=BF)
    db = torch.zeros(3, H, device=DEV) if bias else None
    def run():
        K.attn_bwd(do, q, k, v, o, lse, dq, dkv[:, :H], dkv[:, H:], B, h, Tq, Tk, None, False, 0.125, p, 5,
                   dbias=None if db is None else (db[0], db[1], db[2]))
    for _ in range(3):
        run()
    trace = torch.zeros(148 * 32, dtype=torch.int64, device=DEV)
    torch.cuda.synchronize()
    _lib.lib().spmm_attn_debug_trace(trace.data_ptr())
    run()
    torch.cuda.synchronize()
    _lib.lib().spmm_attn_debug_trace(None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record(); torch.cuda.synchronize()
    t = trace.view(148, 32).cpu()
    t0 = int(t[:, 0][t[:, 0] > 0].min())
    print("BWD B=%d Tq=%d Tk=%d dropout=%.1f bias_sums=%s  total %.2f us (trace), %.2f us (events, avg of 20)"
          % (B, Tq, Tk, p, bias, (int(t[:, 3].max()) - t0) / 1e3, e0.elapsed_time(e1) * 1e3 / 20))
    prev = None
    for s in sorted(bnames):
        col = t[:, s]; col = col[col > 0]
        if len(col):
            rel = (col - t0).float() / 1e3
            med = float(rel.median())
            print("   %-36s n=%3d  min %6.2f  med %6.2f  max %6.2f us   (+%.2f)" % (bnames[s], len(col), rel.min(), med, rel.max(),
                                                                                  med - prev if prev is not None and s > 3 else 0.0))
            prev = med
