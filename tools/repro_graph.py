import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spmm_b200 import synth, trainer
from spmm_b200.SPMM_models import SPMM
from spmm_b200.optim import FusedClipAdamW
CFG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spmm_b200", "configs")
full = len(sys.argv) > 1 and sys.argv[1] == "full"
B = 96 if full else 6
cfg = synth.pretrain_config(os.path.join(CFG, "config_bert.json" if full else "config_tiny_text.json"),
                            os.path.join(CFG, "config_bert_property.json" if full else "config_tiny_property.json"),
                            queue_size=36864 if full else 96, batch_size=B)
model = SPMM(config=cfg); synth.fill_by_name(model); model.to("cuda"); model.build_arenas("cuda"); model.train()
opt = FusedClipAdamW(model)
pv, ids, mask, _ = synth.synthetic_batch(B, seed=1234, fixed_len=64)
pv, ids, mask = pv.cuda(), ids.cuda(), mask.cuda()
def sync(tag):
    torch.cuda.synchronize(); print("ok", tag, flush=True)
trainer.train_step(model, opt, pv, ids, mask, 0.4); sync("eager1")
g = trainer.GraphedTrainStep(model, opt)
for i in range(3):
    g(pv, ids, mask, 0.4); sync("graph%d" % i)
trainer.train_step(model, opt, pv, ids, mask, 0.4); sync("eager2")
g(pv, ids, mask, 0.4); sync("graph again")
