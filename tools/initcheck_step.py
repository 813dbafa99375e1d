"""One eager tiny training step, meant to run under `compute-sanitizer --tool initcheck` with
PYTORCH_NO_CUDA_MEMORY_CACHING=1 (every tensor = a fresh cudaMalloc, so reads of never-written memory are reported)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_spmm_gpu import build_model
from tests.util import load_golden
from spmm_b200 import trainer
from spmm_b200.optim import FusedClipAdamW
DEV = "cuda"
g = load_golden("tiny_b6")
pv, ids, mask, mpm = g["pv"].to(DEV), g["ids"].to(DEV), g["mask"].to(DEV), g["mpm_mask"].to(DEV)
model = build_model("tiny_b6")
opt = FusedClipAdamW(model, lr=2e-4, weight_decay=0.02)
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    l = trainer.train_step(model, opt, pv, ids, mask, 0.4, mpm_mask=mpm)
    torch.cuda.synchronize()
    print("step", it, [float(x) for x in l])
