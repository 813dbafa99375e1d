"""Configs #4 / #5 (BASELINE.json) on the reference's config_bert*.json shapes: the KV-cached batched beam decoder
(`generate.PvDecoder`, d_pv2smiles_batched.py:24-59) and the SMILES -> PV decoder with cached text-side cross K/V
(`generate.Smiles2PvDecoder`, d_smiles2pv.py:14-52) against the fp32 oracle (oracle/generate_ref.py, pinned to the
unmodified reference by tests/test_oracle.py) and against the reference-shaped full-prefix loops.

Beam search is a discrete procedure: a bf16-level difference in two nearly tied log-probabilities flips a token and the
sequences diverge, so "identical token sequences" is checked where it is well defined -
  (a) every step's logits of the cached path against the oracle's full-prefix logits on the SAME prefix (<= 6e-2, and the
      same arg-max wherever the oracle's top-2 gap exceeds that), i.e. teacher-forced parity of the decoder itself, and
  (b) the device-side beam bookkeeping against a host restatement of the reference's rules fed with the same logits:
      identical beams at every step, identical finished sequences and scores,
  (c) graph replay == eager launches, bit for bit."""
import json
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.util import CFG  # noqa: E402

DEV = "cuda"


def _build(sep_bias=0.0):
    from spmm_b200 import synth
    from spmm_b200.SPMM_models import SPMM
    tj, pj = os.path.join(CFG, "config_bert.json"), os.path.join(CFG, "config_bert_property.json")
    model = SPMM(config=synth.pretrain_config(tj, pj, queue_size=96, batch_size=8))
    synth.fill_by_name(model)
    if sep_bias:
        with torch.no_grad():
            model.text_encoder.cls.predictions.bias[3] += sep_bias       # makes [SEP] a frequent top-k candidate
    model.to(DEV)
    model.build_arenas(DEV)
    model.eval()
    return model, json.load(open(tj)), json.load(open(pj))


def test_pv2smiles_kv_cached_decoder_matches_oracle_and_reference_beam_rules():
    from oracle import beam_ref, generate_ref, spmm_ref
    from spmm_b200 import generate
    N, k, steps = 6, 2, 24
    pv = torch.randn(N, 53, generator=torch.Generator().manual_seed(5)).to(DEV)

    def run(model, use_graph=False):
        dec = generate.PvDecoder(model, N, k=k, use_graph=use_graph)
        rec = {"tok": [], "logits": [], "done": []}

        def hook(t, logits):
            if logits is None:
                rec["tok"].append(dec.state.tokens.clone()); rec["done"].append(dec.state.done.clone())
            else:
                rec["logits"].append(logits[:, :300].float().clone())
        return dec.generate(pv, max_steps=steps - 1, on_step=hook), rec
    # (a) teacher-forced parity with the oracle's full-prefix logits: name-seeded weights as they are ([SEP] is rare, so
    #     the beams live for all 24 steps and long prefixes are compared)
    model, ct, cp = _build()
    _, rec = run(model)
    T = len(rec["logits"])
    assert T == steps
    P = spmm_ref.state_from_model(model, device=DEV, requires_grad=False)
    worst, checked, flips = 0.0, 0, 0
    for t in (0, 1, 2, 5, 9, 14, 19, T - 1):
        for r in range(N * k):
            m = r // k
            if int(rec["done"][t][m]) or (t == 0 and r % k):
                continue
            prefix = rec["tok"][t][r, :t + 1][None]
            want = generate_ref.next_token_logits(P, ct, cp, pv[m:m + 1], prefix)[0]
            got = rec["logits"][t][r]
            worst = max(worst, float((got - want).abs().max()))
            top2 = torch.topk(want, 2).values
            if float(top2[0] - top2[1]) > 6e-2:
                flips += int(int(got.argmax()) != int(want.argmax()))
            checked += 1
    print("cached decoder vs oracle: %d (step, beam) pairs, max |d logit| %.3e, arg-max flips outside ties %d" % (checked, worst, flips))
    assert checked >= 60 and worst <= 6e-2 and flips == 0, (checked, worst, flips)
    del model, P
    # (b) device beam bookkeeping == the reference's rules on the same logits; a [SEP] bias makes candidates finish
    model, ct, cp = _build(sep_bias=1.2)
    eager, rec = run(model)
    T = len(rec["logits"])
    n_fin = 0
    for m in range(N):
        fin, hist = beam_ref.replay_beams([l[m * k:(m + 1) * k].cpu() for l in rec["logits"]], k, 2, 3)
        for t in range(min(len(hist), T - 1)):
            if int(rec["done"][t + 1][m]):
                break
            live = rec["tok"][t + 1][m * k:(m + 1) * k, :t + 2].tolist()
            assert live == hist[t], (m, t, live, hist[t])
        got = eager[m]
        assert [s.tolist() for _, s in got] == [s for _, s in fin], (m, got, fin)
        assert all(abs(a[0] - b[0]) <= 1e-3 for a, b in zip(got, fin))
        for lp, toks in got:
            assert int(toks[0]) == 2 and int(toks[-1]) == 3 and math.isfinite(lp)
        n_fin += len(fin)
    print("finished candidates over %d molecules: %d" % (N, n_fin))
    assert n_fin >= N                      # the [SEP] bias makes the finished-list path run
    # (c) one captured graph per decoder, replayed per token: same results as the eager launches
    dec_g = generate.PvDecoder(model, N, k=k, use_graph=True)
    graphed = dec_g.generate(pv, max_steps=steps - 1)
    for a, b in zip(eager, graphed):
        assert [x[1].tolist() for x in a] == [x[1].tolist() for x in b]
        assert all(abs(x[0] - y[0]) <= 1e-5 for x, y in zip(a, b))
    again = dec_g.generate(pv, max_steps=steps - 1)          # replay on re-used caches / state
    assert [[x[1].tolist() for x in a] for a in again] == [[x[1].tolist() for x in a] for a in graphed]
    # the reference-shaped full-prefix loop (one molecule at a time) reaches the same candidates up to bf16 ties
    slow = generate.pv2smiles(model, pv[:1], k=k, max_steps=steps - 1)
    print("full-prefix loop:", [(round(p, 3), t.tolist()) for p, t in slow], " cached:", [(round(p, 3), t.tolist()) for p, t in eager[0]])


def test_smiles2pv_cached_decoder_matches_oracle_at_reference_shapes():
    from oracle import generate_ref, spmm_ref
    from spmm_b200 import generate, synth
    model, ct, cp = _build()
    B = 64
    _, ids, mask, _ = synth.synthetic_batch(B, seed=77, min_len=20, max_len=72)
    ids, mask = ids.to(DEV), mask.to(DEV)
    got = generate.smiles2pv_fast(model, ids, mask)
    P = spmm_ref.state_from_model(model, device=DEV, requires_grad=False)
    want = generate_ref.smiles2pv(P, ct, cp, ids, mask)
    assert got.shape == want.shape == (B, 53) and got.dtype == torch.float32
    d = float((got - want).abs().max())
    print("smiles2pv cached decoder vs oracle: max |d| %.3e  rel-L2 %.3e  (|want| max %.3f)" % (
        d, float((got - want).norm() / want.norm()), float(want.abs().max())))
    assert d <= 6e-2                        # bf16 activations fed back 53 times
    slow = generate.smiles2pv(model, ids[:8], mask[:8])          # the reference-shaped loop over the sub-module API
    fast8 = generate.smiles2pv_fast(model, ids[:8], mask[:8])
    print("cached vs full-recompute loop (B=8): max |d| %.3e" % float((slow - fast8).abs().max()))
    assert float((slow - fast8).abs().max()) <= 4e-2
    assert float((fast8 - got[:8]).abs().max()) <= 4e-2          # a different batch size / length bucket, same molecules
