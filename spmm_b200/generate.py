"""Inference drivers over the sub-module API, mirroring the reference's d_smiles2pv.py:14-52 (SMILES -> 53 property
values, autoregressive over the property tokens) and d_pv2smiles_batched.py:17-59 + d_pv2smiles_single.py:26-51
(property vector -> SMILES, beam search with k beams over the fusion decoder).  Host logic only: every encoder pass goes
through the same sm_100a kernels as pre-training (no KV cache yet - SURVEY.md section 8f ranks that next)."""
import numpy as np
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


@torch.no_grad()
def pv2smiles(model, prop, cls_id=2, sep_id=3, k=2, stochastic=False, max_steps=100):
    """d_pv2smiles_batched.py:24-59 for ONE property vector `prop` [1, 53]: beam search with k beams; a beam that emits
    [SEP] is moved to the finished list (its slot gets score -1e5), the search stops once k candidates are finished; no
    length normalisation.  Returns the finished (log-prob, token ids incl. [CLS]/[SEP]) list, best first."""
    model.eval()
    dev = prop.device
    property1 = model.property_embed(prop.unsqueeze(2))
    properties = torch.cat([model.property_cls.expand(property1.size(0), -1, -1), property1], dim=1)
    prop_embeds = model.property_encoder(inputs_embeds=properties, return_dict=True).last_hidden_state
    product_input = torch.tensor([cls_id], device=dev).expand(1, 1)
    values, indices = _next_token_logp(model, prop_embeds, product_input, k, stochastic)
    product_input = torch.cat([torch.tensor([cls_id], device=dev).expand(k, 1), indices.squeeze(0).unsqueeze(-1)], dim=-1)
    current_p = values.squeeze(0)
    final_output = []
    for _ in range(max_steps):
        values, indices = _next_token_logp(model, prop_embeds, product_input, k, stochastic)
        k2_p = current_p[:, None] + values
        product_input_k2 = torch.cat([product_input.unsqueeze(1).repeat(1, k, 1), indices.unsqueeze(-1)], dim=-1)
        if bool((indices == sep_id).any()):
            for e in (indices == sep_id).nonzero(as_tuple=False):
                final_output.append((float(k2_p[e[0], e[1]]), product_input_k2[e[0], e[1]].clone()))
                k2_p[e[0], e[1]] = -1e5
            if len(final_output) >= k:
                break
        current_p, flat = torch.topk(k2_p.flatten(), k)
        rows, cols = np.unravel_index(flat.cpu().numpy(), tuple(k2_p.shape))
        product_input = torch.stack([product_input_k2[r, c] for r, c in zip(rows, cols)], dim=0)
    return sorted(final_output, key=lambda x: x[0], reverse=True)[:k]
