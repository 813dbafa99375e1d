"""End-to-end sanity of the public training path on the full-size model: `trainer.fit` over a small synthetic "dataset" of
SMILES strings (ragged lengths, native tokenizer, CUDA-graph replay per length bucket, dropout on, device-side sampler,
epoch-0 alpha ramp, cosine schedule) - the four losses must fall while the model memorises the handful of batches.
Usage: python tools/train_sanity.py [epochs]   ->  one line per epoch with the mean losses."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from spmm_b200 import synth, trainer
from spmm_b200.SPMM_models import SPMM
from spmm_b200.tokenizer import WordPieceTokenizer

CFG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spmm_b200", "configs")
dev = torch.device("cuda", 0)
B, n_batches = 96, 6
epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = synth.pretrain_config(os.path.join(CFG, "config_bert.json"), os.path.join(CFG, "config_bert_property.json"), queue_size=36864, batch_size=B)
cfg["optimizer"]["lr"] = cfg["schedular"]["lr"] = cfg["schedular"]["warmup_lr"] = 1e-4
tok = WordPieceTokenizer(os.path.join(CFG, "vocab_bpe_300.txt"), do_lower_case=False, do_basic_tokenize=False)
torch.manual_seed(0)
model = SPMM(config=cfg, tokenizer=tok, loader_len=n_batches)
model.to(dev)
model.build_arenas(dev)
loader = []
for i in range(n_batches):
    pv, _, _, lens = synth.synthetic_batch(B, seed=100 + i)                 # ragged lengths U{12..99}
    loader.append((pv.pin_memory(), bench.synthetic_smiles(tok, lens, 200 + i)))
t0 = time.perf_counter()
hist = trainer.fit(model, loader, max_epochs=epochs)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
for e, h in enumerate(hist):
    print("epoch %d  mlm %.4f  mpm %.4f  ita %.4f  itm %.4f" % (e, *h))
print("graphs captured: %d (length buckets %s); %.1f s for %d steps" % (
    len(model._stepper.graphs), sorted(k[1] for k in model._stepper.graphs), dt, epochs * n_batches))
ok = all(hist[-1][i] < hist[0][i] for i in (0, 1, 3)) and all(x == x for h in hist for x in h)
print("losses fell:", ok)
sys.exit(0 if ok else 1)
