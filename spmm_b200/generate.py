"""Inference drivers over the sub-module API, mirroring the reference's d_smiles2pv.py:14-52 (SMILES -> 53 property
values, autoregressive over the property tokens) and d_pv2smiles_batched.py:17-59 + d_pv2smiles_single.py:26-51
(property vector -> SMILES, beam search with k beams over the fusion decoder).  Host logic only: every encoder pass goes
through the same sm_100a kernels as pre-training (no KV cache yet - SURVEY.md section 8f ranks that next)."""
import torch


def _pv_step(model, prop_input, text_embeds, text_atts):
    """d_smiles2pv.py:14-27: PV encoder (bidirectional over the prefix) -> causal fusion pass -> regression head;
    the prediction for the next property is read at the last position."""
    prop_embeds = model.property_encoder(inputs_embeds=prop_input, return_dict=True).last_hidden_state
    token_output = model.text_encoder.bert(encoder_embeds=prop_embeds, attention_mask=None, encoder_hidden_states=text_embeds,
                                           encoder_attention_mask=text_atts, return_dict=True, is_decoder=True,
                                           mode='fusion').last_hidden_state
    return model.property_mtr_head(token_output).squeeze(-1)[:, -1].unsqueeze(1)


@torch.no_grad()
def smiles2pv(model, input_ids, attention_mask, n_prop=53):
    """d_smiles2pv.py:41-52 for a tokenised batch (`input_ids[:, 1:]` / `attention_mask[:, 1:]` of the tokenizer output):
    returns normalised property predictions [B, n_prop] (fp32)."""
    model.eval()
    text_embeds = model.text_encoder.bert(input_ids, attention_mask=attention_mask, return_dict=True, mode='text').last_hidden_state
    prop_input = model.property_cls.expand(input_ids.shape[0], -1, -1)
    prediction = []
    for _ in range(n_prop):
        output = _pv_step(model, prop_input, text_embeds, attention_mask)
        prediction.append(output)
        prop_input = torch.cat([prop_input, model.property_embed(output.unsqueeze(2))], dim=1)
    return torch.stack(prediction, dim=-1).squeeze(1)


def _next_token_logp(model, prop_embeds, text, k, stochastic):
    """d_pv2smiles_single.py:26-44: causal fusion-decoder pass over the prefix, top-k (or sampled) next tokens."""
    text_atts = torch.where(text == 0, 0, 1)
    enc = prop_embeds.expand(text.shape[0], -1, -1)
    logits = model.text_encoder(text, attention_mask=text_atts, encoder_hidden_states=enc, encoder_attention_mask=None,
                                return_dict=True, is_decoder=True, return_logits=True)[:, -1, :]
    p = torch.softmax(logits, dim=-1)
    if stochastic:
        out = torch.multinomial(p, num_samples=k, replacement=False)
        return torch.log(torch.gather(p, 1, out)), out
    top = torch.topk(p, k=k, dim=-1)
    return torch.log(top.values), top.indices


def encode_properties(model, prop):
    """[B, 53] property vectors -> PV-encoder states [B, 54, H] ([CLS] token + one embedded token per property)."""
    tokens = torch.cat([model.property_cls.expand(prop.shape[0], -1, -1), model.property_embed(prop.unsqueeze(2))], dim=1)
    return model.property_encoder(inputs_embeds=tokens, return_dict=True).last_hidden_state


@torch.no_grad()
def pv2smiles(model, prop, cls_id=2, sep_id=3, k=2, stochastic=False, max_steps=100):
    """Beam search of d_pv2smiles_batched.py:24-59 for ONE property vector `prop` [1, 53].  Semantics kept from the
    reference: every live beam is expanded by its k best (or k sampled) next tokens; from the second expansion on, a
    candidate that ends in [SEP] moves to the finished list with its accumulated log-probability and is masked with
    -1e5 in the live pool; the search stops as soon as k candidates are finished (or after `max_steps` expansions);
    no length normalisation.  Returns up to k (log-prob, token ids incl. [CLS] ... [SEP]) pairs, best first."""
    model.eval()
    enc = encode_properties(model, prop)
    beams = torch.full((1, 1), cls_id, dtype=torch.long, device=prop.device)     # live prefixes [n_beams, T]
    scores = torch.zeros(1, device=prop.device)
    finished = []
    for step in range(max_steps + 1):
        logp, tok = _next_token_logp(model, enc, beams, k, stochastic)            # [n_beams, k] each
        cand_scores = scores[:, None] + logp
        cand = torch.cat([beams[:, None, :].expand(-1, k, -1), tok[:, :, None]], dim=-1)   # [n_beams, k, T + 1]
        if step > 0:
            ended = tok == sep_id
            if bool(ended.any()):
                for b, j in ended.nonzero(as_tuple=False).tolist():
                    finished.append((float(cand_scores[b, j]), cand[b, j].clone()))
                cand_scores = cand_scores.masked_fill(ended, -1e5)
                if len(finished) >= k:
                    break
        scores, flat = cand_scores.flatten().topk(k)
        beams = cand.flatten(0, 1)[flat]
    finished.sort(key=lambda item: item[0], reverse=True)
    return finished[:k]
