"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt by running the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

For each case the reference `SPMM` (SPMM_models.py:16) is built, filled with name-seeded
weights (spmm_b200/synth.py -- so the weights never have to be shipped), switched to eval()
(dropout off), and one forward+backward is run (SPMM_models.py:79-256).  Recorded: inputs, the
Bernoulli PV mask and the 2B multinomial negatives the reference drew, the four losses,
per-parameter gradient norms and 256-entry samples, post-forward queue/EMA samples.
"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import ref_shim  # noqa: E402
from spmm_b200 import synth  # noqa: E402

CFG = os.path.join(REPO, "spmm_b200", "configs")
CASES = {
    # name: (text json, prop json, queue, batch, batch seed, embed_dim)
    "tiny_b6": ("config_tiny_text.json", "config_tiny_property.json", 96, 6, 4321, 256),
    "full_b8": ("config_bert.json", "config_bert_property.json", 36864, 8, 1234, 256),
}


def sample_idx(numel, n=256):
    step = max(1, numel // n)
    return torch.arange(0, numel, step)[:n]


def run_case(name):
    tj, pj, q, b, seed, e = CASES[name]
    cfg = synth.pretrain_config(os.path.join(CFG, tj), os.path.join(CFG, pj), queue_size=q, batch_size=b)
    cfg["embed_dim"] = e
    torch.manual_seed(0)
    model = ref_shim.build_reference(cfg)
    synth.fill_by_name(model)
    model.eval()
    pv, ids, mask, lens = synth.synthetic_batch(b, seed=seed)
    alpha = 0.4

    pm_before = {n: p.detach().clone() for n, p in model.named_parameters() if "_m." in n}
    drawn = []
    real_multinomial = torch.multinomial

    def logging_multinomial(w, n, *a, **k):
        r = real_multinomial(w, n, *a, **k)
        drawn.append(int(r))
        return r

    torch.manual_seed(999)
    mpm_mask = torch.bernoulli(torch.ones(b, 53) * 0.5)     # replay of SPMM_models.py:85
    torch.manual_seed(999)
    torch.multinomial = logging_multinomial
    try:
        losses = model(pv, ids, mask, alpha=alpha)
    finally:
        torch.multinomial = real_multinomial
    total = sum(losses)
    total.backward()

    grads = {}
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach().flatten()
        grads[n] = {"norm": float(g.double().norm()), "sample": g[sample_idx(g.numel())].clone()}
    ema = {}
    for n, p in model.named_parameters():
        if "_m." in n:
            f = p.detach().flatten()
            ema[n] = f[sample_idx(f.numel(), 64)].clone()
    out = {
        "case": name, "alpha": alpha, "batch_seed": seed, "lens": lens,
        "pv": pv, "ids": ids, "mask": mask, "mpm_mask": mpm_mask,
        "neg_t2i": drawn[:b], "neg_i2t": drawn[b:2 * b],
        "losses": torch.stack([l.detach() for l in losses]).double(),
        "grads": grads, "ema_sample": ema,
        "queue_ptr": int(model.queue_ptr), "prop_queue_head": model.prop_queue[:, :b].clone(),
        "text_queue_head": model.text_queue[:, :b].clone(),
        "temp_grad": float(model.temp.grad),
        "global_grad_norm": float(torch.sqrt(sum(p.grad.double().pow(2).sum() for p in model.parameters()
                                                 if p.grad is not None))),
        "n_trainable": sum(p.numel() for p in model.parameters() if p.requires_grad),
        "n_all": sum(p.numel() for p in model.parameters()),
        "state_dict_keys": [(k, tuple(v.shape), str(v.dtype)) for k, v in model.state_dict().items()],
        "torch": torch.__version__,
    }
    assert len(drawn) == 2 * b
    path = os.path.join(REPO, "tests", "golden", name + ".pt")
    torch.save(out, path)
    print(name, "losses", out["losses"].tolist(), "gnorm", out["global_grad_norm"], "negs", drawn,
          "size", os.path.getsize(path))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    for c in (sys.argv[1:] or list(CASES)):
        run_case(c)
