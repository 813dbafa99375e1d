"""TEST INFRASTRUCTURE: fp32 restatement of the reference's SMILES->PV generation loop (d_smiles2pv.py:14-52) and of one
PV->SMILES decoder step / beam search (d_pv2smiles_single.py:26-44, d_pv2smiles_batched.py:24-59) on the oracle's functional
BERT (oracle/spmm_ref.py).  PINNED: tests/test_oracle.py::test_generation_oracle_matches_reference_golden checks it against
tests/golden/generate_tiny.pt, recorded by executing the unmodified reference functions (oracle/make_golden_generate.py).
Never imported by spmm_b200/."""
import torch
import torch.nn.functional as F

from . import spmm_ref as R


def mtr_head(P, x):
    h = F.gelu(R._lin(P, "property_mtr_head.0", x))
    h = F.layer_norm(h, (h.shape[-1],), P["property_mtr_head.2.weight"], P["property_mtr_head.2.bias"], 1e-12)
    return R._lin(P, "property_mtr_head.3", h)


@torch.no_grad()
def smiles2pv(P, cfg_text, cfg_prop, ids, att, n_prop=53):
    text_embeds = R.bert(P, "text_encoder.bert", cfg_text, ids=ids, att=att, mode="text")
    prop_input = P["property_cls"].expand(ids.shape[0], -1, -1)
    preds = []
    for _ in range(n_prop):
        prop_embeds = R.bert(P, "property_encoder", cfg_prop, inputs_embeds=prop_input)
        tok = R.bert(P, "text_encoder.bert", cfg_text, encoder_embeds=prop_embeds, enc=text_embeds, enc_att=att,
                     is_decoder=True, mode="fusion")
        out = mtr_head(P, tok).squeeze(-1)[:, -1].unsqueeze(1)
        preds.append(out)
        prop_input = torch.cat([prop_input, F.linear(out.unsqueeze(2), P["property_embed.weight"], P["property_embed.bias"])], dim=1)
    return torch.stack(preds, dim=-1).squeeze(1)


@torch.no_grad()
def next_token_logits(P, cfg_text, cfg_prop, pv, text):
    feat = F.linear(pv.unsqueeze(2), P["property_embed.weight"], P["property_embed.bias"])
    properties = torch.cat([P["property_cls"].expand(pv.shape[0], -1, -1), feat], dim=1)
    prop_embeds = R.bert(P, "property_encoder", cfg_prop, inputs_embeds=properties)
    att = torch.where(text == 0, 0, 1)
    h = R.bert(P, "text_encoder.bert", cfg_text, ids=text, att=att, enc=prop_embeds.expand(text.shape[0], -1, -1), is_decoder=True)
    return R.lm_head(P, "text_encoder.cls.predictions", h)[:, -1, :]


@torch.no_grad()
def pv2smiles_beam(P, cfg_text, cfg_prop, pv, k=2, cls_id=2, sep_id=3, max_steps=100):
    """d_pv2smiles_batched.py:24-59, deterministic branch: returns the finished (log-prob, ids) list, best first."""
    def step(text):
        lp = torch.log_softmax(next_token_logits(P, cfg_text, cfg_prop, pv, text), dim=-1)
        top = torch.topk(lp, k=k, dim=-1)
        return top.values, top.indices
    beams = torch.full((1, 1), cls_id, dtype=torch.long, device=pv.device)
    scores = torch.zeros(1, device=pv.device)
    finished = []
    for it in range(max_steps + 1):
        logp, tok = step(beams)
        cand_scores = scores[:, None] + logp
        cand = torch.cat([beams[:, None, :].expand(-1, k, -1), tok[:, :, None]], dim=-1)
        if it > 0:
            ended = tok == sep_id
            for b, j in ended.nonzero(as_tuple=False).tolist():
                finished.append((float(cand_scores[b, j]), cand[b, j].clone()))
            cand_scores = cand_scores.masked_fill(ended, -1e5)
            if len(finished) >= k:
                break
        scores, flat = cand_scores.flatten().topk(k)
        beams = cand.flatten(0, 1)[flat]
    finished.sort(key=lambda x: x[0], reverse=True)
    return finished[:k]
