"""Debug: first-step gradients of the tiny model, eager vs CUDA-graph, repeated, with per-tensor scale factors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_spmm_gpu import build_model
from tests.util import load_golden
from spmm_b200 import ops, trainer
from spmm_b200.optim import FusedClipAdamW
DEV = "cuda"
g = load_golden("tiny_b6")
pv, ids, mask, mpm = g["pv"].to(DEV), g["ids"].to(DEV), g["mask"].to(DEV), g["mpm_mask"].to(DEV)
def run(mode, steps=1):
    model = build_model("tiny_b6")
    opt = FusedClipAdamW(model, lr=2e-4, weight_decay=0.02)
    ops.step_rng(DEV).reset(0)
    st = trainer.GraphedTrainStep(model, opt) if mode == "graph" else None
    for _ in range(steps):
        if st is None:
            l = torch.stack([x.detach() for x in trainer.train_step(model, opt, pv, ids, mask, 0.4, mpm_mask=mpm)])
        else:
            l = st(pv, ids, mask, 0.4, mpm_mask=mpm).clone()
    torch.cuda.synchronize()
    return {k: v.grad.detach().clone() for k, v in model.named_parameters() if v.grad is not None}, l, model.last_aux["neg_t2i"].tolist()
runs = [("eager", run("eager")), ("graph", run("graph")), ("graph", run("graph")), ("eager", run("eager"))]
keys = ["text_encoder.bert.encoder.layer.0.attention.self.query.weight", "text_encoder.bert.encoder.layer.0.attention.self.key.weight",
        "text_encoder.bert.encoder.layer.0.attention.self.value.weight", "text_encoder.bert.encoder.layer.2.crossattention.self.query.weight",
        "property_encoder.encoder.layer.0.attention.self.query.weight", "text_encoder.bert.encoder.layer.0.attention.output.dense.weight"]
ref = runs[0][1][0]
for name, (gr, l, neg) in runs:
    print(name, "losses", [round(float(x), 6) for x in l], "neg", neg)
    for k in keys:
        d = float((gr[k] - ref[k]).norm() / ref[k].norm())
        ratio = float((gr[k] * ref[k]).sum() / (ref[k] * ref[k]).sum())
        print("    rel diff vs eager#1 %.3e   ls-scale %.5f   %s" % (d, ratio, k))
