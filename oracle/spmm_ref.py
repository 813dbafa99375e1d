"""TEST INFRASTRUCTURE ONLY -- CPU/fp32 restatement ("port") of the reference hot path.

A functional, plain-PyTorch restatement of `SPMM.forward` (reference SPMM_models.py:79-256)
and of the xbert.py blocks it calls, operating on a {reference state-dict name: tensor} dict.
It is the checker for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg import it.  The product path (spmm_b200/) never does.

Parity status: PINNED -- tests/test_oracle.py checks this file against tests/golden/*.pt, which
oracle/make_golden.py produced by running the unmodified reference in the build container.

RNG is injected (mpm_mask, negative indices) so results are comparable across implementations.
"""
import math

import torch
import torch.nn.functional as F


def _lin(P, pre, x):
    return F.linear(x, P[pre + ".weight"], P[pre + ".bias"])


def _ln(P, pre, x, eps):
    return F.layer_norm(x, (x.shape[-1],), P[pre + ".weight"], P[pre + ".bias"], eps)


def attention(P, pre, x, kv, add_mask, n_heads):
    """BertSelfAttention.forward + BertSelfOutput.forward (xbert.py:270-359, 369-373), eval mode."""
    B, Tq, H = x.shape
    d = H // n_heads
    q = _lin(P, pre + ".self.query", x).view(B, Tq, n_heads, d).permute(0, 2, 1, 3)
    k = _lin(P, pre + ".self.key", kv).view(kv.shape[0], kv.shape[1], n_heads, d).permute(0, 2, 1, 3)
    v = _lin(P, pre + ".self.value", kv).view(kv.shape[0], kv.shape[1], n_heads, d).permute(0, 2, 1, 3)
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d)          # xbert.py:305,323
    if add_mask is not None:
        s = s + add_mask                                             # xbert.py:327
    p = torch.softmax(s, dim=-1)                                     # xbert.py:335
    ctx = torch.matmul(p, v).permute(0, 2, 1, 3).reshape(B, Tq, H)   # xbert.py:350-354
    return _ln(P, pre + ".output.LayerNorm", _lin(P, pre + ".output.dense", ctx) + x, 1e-12)


def layer(P, pre, x, self_mask, enc, enc_mask, n_heads, has_cross):
    """BertLayer.forward (xbert.py:469-534)."""
    x = attention(P, pre + ".attention", x, x, self_mask, n_heads)
    if has_cross:
        x = attention(P, pre + ".crossattention", x, enc, enc_mask, n_heads)
    h = F.gelu(_lin(P, pre + ".intermediate.dense", x))              # exact erf GELU, xbert.py:435-436
    return _ln(P, pre + ".output.LayerNorm", _lin(P, pre + ".output.dense", h) + x, 1e-12)


def self_mask(att, is_decoder, dtype):
    """get_extended_attention_mask (xbert.py:889-948): (1-m) * -10000."""
    B, T = att.shape
    if is_decoder:
        ar = torch.arange(T, device=att.device)
        causal = (ar[None, None, :] <= ar[None, :, None]).to(att.dtype)
        m = causal[:, None, :, :] * att[:, None, None, :]
    else:
        m = att[:, None, None, :]
    return (1.0 - m.to(dtype)) * -10000.0


def cross_mask(att, dtype):
    """transformers' invert_attention_mask (call sites xbert.py:1038-1043): (1-m) * finfo.min."""
    return (1.0 - att[:, None, None, :].to(dtype)) * torch.finfo(dtype).min


def embeddings(P, pre, ids=None, inputs_embeds=None):
    """BertEmbeddings.forward (xbert.py:193-220), eval mode."""
    if inputs_embeds is None:
        inputs_embeds = F.embedding(ids, P[pre + ".word_embeddings.weight"], padding_idx=0)   # xbert.py:178
    T = inputs_embeds.shape[1]
    x = inputs_embeds + P[pre + ".token_type_embeddings.weight"][0] + P[pre + ".position_embeddings.weight"][:T]
    return _ln(P, pre + ".LayerNorm", x, 1e-12)


def bert(P, pre, cfg, ids=None, att=None, inputs_embeds=None, encoder_embeds=None, enc=None, enc_att=None,
         is_decoder=False, mode="multi_modal"):
    """BertModel.forward + BertEncoder.forward (xbert.py:950-1091, 543-644)."""
    if encoder_embeds is not None:
        x = encoder_embeds
    else:
        x = embeddings(P, pre + ".embeddings", ids, inputs_embeds)
    if att is None:
        att = torch.ones(x.shape[:2], device=x.device)
    sm = self_mask(att, is_decoder, x.dtype)
    cm = None
    if enc is not None:
        if enc_att is None:
            enc_att = torch.ones(enc.shape[:2], device=x.device)
        cm = cross_mask(enc_att, x.dtype)
    fl, nl = cfg["fusion_layer"], cfg["num_hidden_layers"]
    lo, hi = {"text": (0, fl), "fusion": (fl, nl), "multi_modal": (0, nl)}[mode]
    for i in range(lo, hi):
        x = layer(P, "%s.encoder.layer.%d" % (pre, i), x, sm, enc, cm, cfg["num_attention_heads"], i >= fl)
    return x


def lm_head(P, pre, x):
    """BertLMPredictionHead.forward (xbert.py:693-696) with the tied decoder."""
    h = _ln(P, pre + ".transform.LayerNorm", F.gelu(_lin(P, pre + ".transform.dense", x)), 1e-12)
    return F.linear(h, P[pre + ".decoder.weight"], P[pre + ".bias"])


def ema_update(P, momentum):
    """_momentum_update (SPMM_models.py:265-269): three separately rounded fp32 ops."""
    with torch.no_grad():
        for k in list(P.keys()):
            for a, b in (("property_encoder.", "property_encoder_m."), ("property_proj.", "property_proj_m."),
                         ("text_encoder.", "text_encoder_m."), ("text_proj.", "text_proj_m.")):
                if k.startswith(a) and (b + k[len(a):]) in P and P[k].is_floating_point():
                    km = b + k[len(a):]
                    P[km] = P[km] * momentum + P[k].detach() * (1. - momentum)


def itc_loss(prop_feat, text_feat, prop_feat_m, text_feat_m, prop_queue, text_queue, temp, alpha):
    """SPMM_models.py:102-131."""
    with torch.no_grad():
        prop_all = torch.cat([prop_feat_m.t(), prop_queue], dim=1)
        text_all = torch.cat([text_feat_m.t(), text_queue], dim=1)
        tg = torch.zeros(prop_feat.shape[0], prop_all.shape[1], device=prop_feat.device)
        tg.fill_diagonal_(1)
        t_i2t = alpha * F.softmax(prop_feat_m @ text_all / temp, 1) + (1 - alpha) * tg
        t_t2i = alpha * F.softmax(text_feat_m @ prop_all / temp, 1) + (1 - alpha) * tg
        t_i2i = alpha * F.softmax(prop_feat_m @ prop_all / temp, 1) + (1 - alpha) * tg
        t_t2t = alpha * F.softmax(text_feat_m @ text_all / temp, 1) + (1 - alpha) * tg
    s_i2t = prop_feat @ text_all / temp
    s_t2i = text_feat @ prop_all / temp
    s_i2i = prop_feat @ prop_all / temp
    s_t2t = text_feat @ text_all / temp
    loss = 0
    for s, t in ((s_i2t, t_i2t), (s_t2i, t_t2i), (s_i2i, t_i2i), (s_t2t, t_t2t)):
        loss = loss - torch.sum(F.log_softmax(s, 1) * t, 1).mean()
    return loss / 2, s_i2t, s_t2i


def negative_weights(s_i2t, s_t2i):
    """SPMM_models.py:154-161."""
    bs = s_i2t.shape[0]
    w_i2t = F.softmax(s_i2t[:, :bs].detach(), 1).clone()
    w_t2i = F.softmax(s_t2i[:, :bs].detach(), 1).clone()
    w_i2t.fill_diagonal_(0)
    w_t2i.fill_diagonal_(0)
    return w_i2t, w_t2i


def forward(P, cfg_text, cfg_prop, pv, ids, att, alpha, mpm_mask, neg_t2i=None, neg_i2t=None, momentum=0.995,
            queue_ptr=0, world_feats=None, sampler=None):
    """SPMM.forward.  P is modified in place (EMA, queues) like the reference's buffers.

    Returns (loss_mlm, 5*loss_mpm, loss_ita, loss_itm), aux dict.
    `sampler(w_t2i, w_i2t) -> (neg_t2i, neg_i2t)` is used when indices are not injected."""
    with torch.no_grad():
        P["temp"].clamp_(0.01, 0.5)
    temp = P["temp"]
    B = pv.shape[0]
    feat = F.linear(pv.unsqueeze(2), P["property_embed.weight"], P["property_embed.bias"])
    m = mpm_mask.unsqueeze(2)
    masked = feat * (1 - m) + P["property_mask"].expand(B, pv.shape[1], -1) * m
    properties = torch.cat([P["property_cls"].expand(B, -1, -1), masked], dim=1)

    prop_embeds = bert(P, "property_encoder", cfg_prop, inputs_embeds=properties)
    prop_atts = torch.ones(prop_embeds.shape[:2], dtype=torch.long, device=pv.device)
    prop_feat = F.normalize(_lin(P, "property_proj", prop_embeds[:, 0]), dim=-1)
    text_embeds = bert(P, "text_encoder.bert", cfg_text, ids=ids, att=att, mode="text")
    text_feat = F.normalize(_lin(P, "text_proj", text_embeds[:, 0]), dim=-1)

    with torch.no_grad():
        ema_update(P, momentum)
        prop_embeds_m = bert(P, "property_encoder_m", cfg_prop, inputs_embeds=properties)
        prop_feat_m = F.normalize(_lin(P, "property_proj_m", prop_embeds_m[:, 0]), dim=-1)
        text_embeds_m = bert(P, "text_encoder_m.bert", cfg_text, ids=ids, att=att, mode="text")
        text_feat_m = F.normalize(_lin(P, "text_proj_m", text_embeds_m[:, 0]), dim=-1)
    loss_ita, s_i2t, s_t2i = itc_loss(prop_feat, text_feat, prop_feat_m, text_feat_m,
                                      P["prop_queue"].clone(), P["text_queue"].clone(), temp, alpha)

    def fusion(q, q_att, kv, kv_att, dec=False):
        return bert(P, "text_encoder.bert", cfg_text, att=q_att, encoder_embeds=q, enc=kv, enc_att=kv_att,
                    is_decoder=dec, mode="fusion")

    pos_prop = fusion(prop_embeds, prop_atts, text_embeds, att)[:, 0]
    pos_text = fusion(text_embeds, att, prop_embeds, prop_atts)[:, 0]
    w_i2t, w_t2i = negative_weights(s_i2t, s_t2i)
    if neg_t2i is None:
        neg_t2i, neg_i2t = sampler(w_t2i, w_i2t)
    it = torch.as_tensor(neg_t2i, device=pv.device)
    ii = torch.as_tensor(neg_i2t, device=pv.device)
    prop_neg = prop_embeds[it]
    text_neg, text_att_neg = text_embeds[ii], att[ii]
    text_all = torch.cat([text_embeds, text_neg]); text_att_all = torch.cat([att, text_att_neg])
    prop_all = torch.cat([prop_neg, prop_embeds]); prop_att_all = torch.cat([prop_atts, prop_atts])
    neg_prop = fusion(prop_all, prop_att_all, text_all, text_att_all)[:, 0]
    neg_text = fusion(text_all, text_att_all, prop_all, prop_att_all)[:, 0]
    vl = torch.cat([torch.cat([pos_prop, pos_text], -1), torch.cat([neg_prop, neg_text], -1)], 0)
    labels = torch.cat([torch.ones(B, dtype=torch.long), torch.zeros(2 * B, dtype=torch.long)]).to(pv.device)
    loss_itm = F.cross_entropy(_lin(P, "itm_head", vl), labels)

    # _dequeue_and_enqueue (SPMM_models.py:271-286); world_feats emulates concat_all_gather
    with torch.no_grad():
        pf, tf = (prop_feat_m, text_feat_m) if world_feats is None else world_feats
        n = pf.shape[0]
        assert P["prop_queue"].shape[1] % n == 0
        P["prop_queue"][:, queue_ptr:queue_ptr + n] = pf.t()
        P["text_queue"][:, queue_ptr:queue_ptr + n] = tf.t()
        new_ptr = (queue_ptr + n) % P["prop_queue"].shape[1]

    lab = ids[:, 1:]
    with torch.no_grad():
        hm = bert(P, "text_encoder_m.bert", cfg_text, ids=ids, att=att, enc=prop_embeds_m, enc_att=prop_atts,
                  is_decoder=True)
        logits_m = lm_head(P, "text_encoder_m.cls.predictions", hm)[:, :-1]
    h = bert(P, "text_encoder.bert", cfg_text, ids=ids, att=att, enc=prop_embeds, enc_att=prop_atts, is_decoder=True)
    logits = lm_head(P, "text_encoder.cls.predictions", h)[:, :-1]
    loss_ce = F.cross_entropy(logits.permute(0, 2, 1), lab)              # includes PAD labels (ignore_index -100)
    distill = -torch.sum(F.log_softmax(logits, -1) * F.softmax(logits_m, -1), -1)
    loss_mlm = (1 - alpha) * loss_ce + alpha * distill[lab != 0].mean()

    pc = bert(P, "property_encoder", cfg_prop, inputs_embeds=properties, is_decoder=True)
    po = fusion(pc, prop_atts, text_embeds, att, dec=True)[:, :-1]
    t = F.gelu(_lin(P, "property_mtr_head.0", po))
    t = _ln(P, "property_mtr_head.2", t, 1e-12)
    pred = _lin(P, "property_mtr_head.3", t).squeeze(-1)
    keep = (1 - mpm_mask).bool()
    loss_mpm = F.mse_loss(pred[keep], pv[keep])
    aux = {"neg_t2i": list(map(int, it.tolist())), "neg_i2t": list(map(int, ii.tolist())), "queue_ptr": new_ptr,
           "prop_feat_m": prop_feat_m, "text_feat_m": text_feat_m, "w_t2i": w_t2i, "w_i2t": w_i2t,
           "prop_feat": prop_feat, "text_feat": text_feat}
    return (loss_mlm, loss_mpm * 5, loss_ita, loss_itm), aux


def state_from_model(model, device="cpu", dtype=torch.float32, requires_grad=True):
    """{name: tensor} from any module with reference key names; trainable leaves require grad.
    The tied decoder weight shares one leaf with the word embedding."""
    P, seen = {}, {}
    params = dict(model.named_parameters(remove_duplicate=False))
    for k, v in model.state_dict(keep_vars=True).items():
        key = v.data_ptr()
        if key in seen and v.is_floating_point() and v.numel() > 1:
            P[k] = P[seen[key]]
            continue
        t = v.detach().to(device=device).clone()
        if t.is_floating_point():
            t = t.to(dtype)
            if requires_grad and k in params and params[k].requires_grad:
                t.requires_grad_(True)
        P[k] = t
        seen[key] = k
    return P
