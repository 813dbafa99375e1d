"""Pins oracle/spmm_ref.py (the CPU restatement) against golden vectors produced by the
unmodified reference (oracle/make_golden.py, SPMM_models.py:79-256)."""
import pytest
import torch

from oracle import spmm_ref
import os

from tests.util import REPO, load_cfgs, load_golden, oracle_state, sample_idx


def _run(case):
    g = load_golden(case)
    ct, cp, q = load_cfgs(case)
    P = oracle_state(g)
    losses, aux = spmm_ref.forward(P, ct, cp, g["pv"], g["ids"], g["mask"], g["alpha"], g["mpm_mask"],
                                   neg_t2i=g["neg_t2i"], neg_i2t=g["neg_i2t"])
    sum(losses).backward()
    return g, P, losses, aux


def _check(case, tol):
    g, P, losses, aux = _run(case)
    got = torch.stack([l.detach() for l in losses]).double()
    assert torch.allclose(got, g["losses"], rtol=tol, atol=tol), (got, g["losses"])
    assert aux["queue_ptr"] == g["queue_ptr"]
    assert torch.allclose(P["prop_queue"][:, :g["pv"].shape[0]], g["prop_queue_head"], atol=1e-6)
    assert torch.allclose(P["text_queue"][:, :g["pv"].shape[0]], g["text_queue_head"], atol=1e-6)
    # EMA samples: bit-exact restatement of SPMM_models.py:269
    for n, s in g["ema_sample"].items():
        f = P[n].detach().flatten()
        assert torch.equal(f[sample_idx(f.numel(), 64)], s), n
    # gradients
    worst = 0.0
    for n, rec in g["grads"].items():
        gr = P[n].grad
        assert gr is not None, n
        f = gr.flatten()
        ref_norm = rec["norm"]
        assert abs(float(f.double().norm()) - ref_norm) <= 2e-3 * ref_norm + 1e-7 * g["global_grad_norm"], (n, float(f.norm()), ref_norm)
        d = (f[sample_idx(f.numel())] - rec["sample"]).abs().max().item()
        # key biases have mathematically zero gradient (softmax shift invariance): floor the scale
        worst = max(worst, d / max(rec["sample"].abs().max().item(), 1e-7 * g["global_grad_norm"]))
    assert worst < 5e-3, worst
    assert abs(float(P["temp"].grad) - g["temp_grad"]) <= 1e-3 * abs(g["temp_grad"])
    # the reference never produces a grad for the PV word embedding
    assert "property_encoder.embeddings.word_embeddings.weight" not in g["grads"]
    # hard-negative weights reproduce the reference's draws under torch.multinomial's CPU stream
    torch.manual_seed(999)
    torch.bernoulli(torch.ones_like(g["pv"]) * 0.5)
    t2i = [int(torch.multinomial(aux["w_t2i"][b], 1)) for b in range(len(g["neg_t2i"]))]
    i2t = [int(torch.multinomial(aux["w_i2t"][b], 1)) for b in range(len(g["neg_i2t"]))]
    assert t2i == g["neg_t2i"] and i2t == g["neg_i2t"]


def test_oracle_matches_reference_tiny():
    _check("tiny_b6", 2e-5)


@pytest.mark.slow
def test_oracle_matches_reference_full():
    torch.set_num_threads(8)
    _check("full_b8", 5e-5)


def test_generation_oracle_matches_reference_golden():
    """oracle/generate_ref.py (SMILES->PV loop, PV->SMILES decoder step and beam search) against
    tests/golden/generate_tiny.pt, recorded by running the unmodified reference functions
    (oracle/make_golden_generate.py).  fp32 on CPU: values to 1e-5, token ids exact."""
    from oracle import generate_ref
    gg = torch.load(os.path.join(REPO, "tests", "golden", "generate_tiny.pt"), weights_only=False)
    g = load_golden("tiny_b6")
    ct, cp, _ = load_cfgs("tiny_b6")
    P = oracle_state(g, device="cpu")
    got = generate_ref.smiles2pv(P, ct, cp, gg["ids"], gg["mask"])
    assert torch.allclose(got, gg["smiles2pv"], atol=2e-5), float((got - gg["smiles2pv"]).abs().max())
    with torch.no_grad():
        P["text_encoder.cls.predictions.bias"][3] += gg["sep_bias"]      # the [SEP] nudge the golden run applied
    for b in range(3):
        pv = gg["pv"][b:b + 1]
        lp = torch.log_softmax(generate_ref.next_token_logits(P, ct, cp, pv, torch.tensor([[2]])), dim=-1)
        vals, idx = gg["first_step"][b]
        assert torch.topk(lp[0], 5).indices.tolist() == idx.tolist()
        assert torch.allclose(lp[0, idx], vals, atol=2e-5)
        best = generate_ref.pv2smiles_beam(P, ct, cp, pv, k=2)[0][1].tolist()
        assert best[:-1] == gg["beam_best"][b] and best[-1] == 3
