"""Runs one forward+backward of the tiny model after filling the CUDA caching allocator's free blocks with NaN
("dirty") or not, and reports losses plus which gradient tensors contain NaN / differ: finds reads of torch.empty memory."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_spmm_gpu import build_model
from tests.util import load_golden
DEV = "cuda"
mode = sys.argv[1]
g = load_golden("tiny_b6")
model = build_model("tiny_b6")
if mode == "dirty":
    junk = []
    for sz in [1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20, 1 << 22, 1 << 24]:
        for _ in range(24):
            junk.append(torch.full((sz,), float("nan"), device=DEV, dtype=torch.float32))
    del junk
pv, ids, mask = g["pv"].to(DEV), g["ids"].to(DEV), g["mask"].to(DEV)
losses = model(pv, ids, mask, alpha=g["alpha"], mpm_mask=g["mpm_mask"].to(DEV), neg_idx=(g["neg_t2i"], g["neg_i2t"]))
(losses[0] + losses[1] + losses[2] + losses[3]).backward()
torch.cuda.synchronize()
print(mode, "losses", [float(x) for x in losses])
grads = {n: p.grad.detach().float().cpu() for n, p in model.named_parameters() if p.grad is not None}
torch.save(grads, "gpurun_out/grads_%s.pt" % mode)
bad = [(n, int(torch.isnan(v).sum()), v.numel()) for n, v in grads.items() if torch.isnan(v).any()]
print(mode, "tensors with NaN grads:", len(bad))
for b in bad[:40]:
    print("   ", b)
if mode == "dirty" and os.path.exists("gpurun_out/grads_clean.pt"):
    ref = torch.load("gpurun_out/grads_clean.pt")
    worst = sorted(((float((grads[n] - ref[n]).norm() / (ref[n].norm() + 1e-30)), n) for n in ref if not torch.isnan(grads[n]).any()), reverse=True)
    print("largest rel diffs vs clean (non-NaN tensors):")
    for w in worst[:15]:
        print("    %.3e  %s" % w)
