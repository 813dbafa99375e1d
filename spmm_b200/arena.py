"""Flat parameter / gradient / momentum arenas in HBM.

Data layout (one allocation each, fp32 unless noted):

    P   [ pe.word_emb | online EMA set ............ | itm_head, property_embed, mtr_head, cls, mask, temp ]
    G   same layout: gradients.  wgrad GEMMs accumulate into it, NCCL all-reduces it as one buffer,
        the fused clip+AdamW kernel walks it.
    M   [ momentum twins of the EMA range ]          (SPMM_models.py:46-62 model_pairs)
    P16 / M16   bf16 shadows of P / M: the operands the tcgen05 GEMMs actually read.

Every nn.Parameter of the model becomes a view into P (or M), `.grad` a view into G, so `state_dict()`,
`load_state_dict()`, `torch.optim` and the reference's key names keep working, while
  * the EMA (SPMM_models.py:265-269) is ONE vectorised kernel over [0, n_ema) that also refreshes both shadows,
  * clip_grad_norm_ + AdamW (SPMM_models.py:361-362) is one reduction + one update kernel.
Within each attention module the order is [q.w, k.w, v.w, q.b, k.b, v.b] so that fused QKV / KV projections are
plain views.  Tensor starts are aligned to 64 elements (TMA needs 16-byte aligned bases).
"""
from types import SimpleNamespace

import torch

from . import kernels as K

ALIGN = 64
SHARD_ALIGN = 8 * 64      # (n_total - adam_start) is a multiple of this: equal, aligned optimiser shards for 1/2/4/8 ranks
MODEL_PAIRS = (("property_encoder", "property_encoder_m"), ("property_proj", "property_proj_m"),
               ("text_encoder", "text_encoder_m"), ("text_proj", "text_proj_m"))
TAIL_MODULES = ("itm_head", "property_embed", "property_mtr_head")
TAIL_PARAMS = ("property_cls", "property_mask", "temp")
NEVER_GRAD = "property_encoder.embeddings.word_embeddings.weight"   # stays grad-free


def _qkv_order(names):
    out, done = [], set()
    for n in names:
        if n in done:
            continue
        if n.endswith("self.query.weight"):
            pre = n[:-len("query.weight")]
            grp = [pre + s for s in ("query.weight", "key.weight", "value.weight", "query.bias", "key.bias", "value.bias")]
            out += grp
            done.update(grp)
        else:
            out.append(n)
            done.add(n)
    return out


def _pad(n):
    return (n + ALIGN - 1) // ALIGN * ALIGN


class ParamArena:
    def __init__(self, model, device):
        self.device = torch.device(device)
        entries = []          # (qualified name, online param, momentum param or None)
        for on, mn in MODEL_PAIRS:
            mod, mod_m = getattr(model, on), getattr(model, mn)
            po, pm = dict(mod.named_parameters()), dict(mod_m.named_parameters())
            for n in _qkv_order(list(po.keys())):
                entries.append((on + "." + n, po[n], pm[n]))
        entries.sort(key=lambda e: 0 if e[0] == NEVER_GRAD else 1)      # stable: the grad-free tensor goes first
        n_ema_entries = len(entries)
        for mname in TAIL_MODULES:
            for n, p in getattr(model, mname).named_parameters():
                entries.append((mname + "." + n, p, None))
        for pname in TAIL_PARAMS:
            if getattr(model, pname, None) is not None:
                entries.append((pname, getattr(model, pname), None))

        self.offset, off = {}, 0
        for i, (name, p, pm) in enumerate(entries):
            if i == n_ema_entries:
                self.n_ema = off
            self.offset[name] = (off, p.numel(), tuple(p.shape))
            off += _pad(p.numel())
        if n_ema_entries == len(entries):
            self.n_ema = off
        first_pad = _pad(entries[0][1].numel()) if entries[0][0] == NEVER_GRAD else 0
        off += (-(off - first_pad)) % SHARD_ALIGN      # the optimiser range splits evenly over up to 8 ranks (16-byte shards)
        self.n_total = off
        first = entries[0]
        self.adam_start = _pad(first[1].numel()) if first[0] == NEVER_GRAD else 0
        self.n_params_ema = sum(e[1].numel() for e in entries[:n_ema_entries])

        f32 = dict(device=self.device, dtype=torch.float32)
        self.P = torch.zeros(self.n_total, **f32)
        self.G = torch.zeros(self.n_total, **f32)
        self.M = torch.zeros(self.n_ema, **f32)
        self.P16 = torch.zeros(self.n_total, device=self.device, dtype=torch.bfloat16)
        self.M16 = torch.zeros(self.n_ema, device=self.device, dtype=torch.bfloat16)
        self._params = []
        with torch.no_grad():
            for name, p, pm in entries:
                o, n, shape = self.offset[name]
                self.P[o:o + n].copy_(p.data.reshape(-1))
                p.data = self.P[o:o + n].view(shape)
                if p.requires_grad and name != NEVER_GRAD:     # stays None like in the reference (torch skips it)
                    p.grad = self.G[o:o + n].view(shape)
                if pm is not None:
                    self.M[o:o + n].copy_(pm.data.reshape(-1))
                    pm.data = self.M[o:o + n].view(shape)
                self._params.append((name, p, pm))
        self.sumsq = torch.zeros(1, **f32)
        self.refresh_shadows(ema=False)

    # ------------------------------------------------------------------ views
    def _slice(self, buf, name, rows=None, count=1):
        """view of `count` consecutive equally-shaped tensors starting at `name` as one 2-D/1-D tensor"""
        o, n, shape = self.offset[name]
        flat = buf[o:o + n * count]
        if len(shape) == 2:
            return flat.view(shape[0] * count, shape[1])
        return flat.view(n * count)

    def w16(self, name, momentum=False, count=1):
        return self._slice(self.M16 if momentum else self.P16, name, count=count)

    def f32(self, name, momentum=False, count=1):
        return self._slice(self.M if momentum else self.P, name, count=count)

    def grad(self, name, count=1):
        return self._slice(self.G, name, count=count)

    def prefix_range(self, prefix):
        """[lo, hi) of the arena occupied by the parameters whose name starts with `prefix` (e.g. one encoder layer):
        contiguous by construction (named_parameters order); used to all-reduce a layer's gradients as soon as its last
        backward contribution has been enqueued (trainer.GradOverlap)."""
        spans = [(o, o + _pad(n)) for name, (o, n, _) in self.offset.items() if name.startswith(prefix)]
        lo, hi = min(s[0] for s in spans), max(s[1] for s in spans)
        assert all(name.startswith(prefix) for name, (o, n, _) in self.offset.items() if lo <= o < hi), prefix
        return lo, hi

    # ------------------------------------------------------------------ maintenance
    def valid_for(self, model):
        name, p, _ = self._params[-1]
        o = self.offset[name][0]
        return p.data_ptr() == self.P.data_ptr() + 4 * o

    def ensure_grads(self):
        """Re-attach .grad views (torch's zero_grad(set_to_none=True) detaches them)."""
        for name, p, _ in self._params:
            if p.requires_grad and name != NEVER_GRAD and (p.grad is None or p.grad.data_ptr() != self.G.data_ptr() + 4 * self.offset[name][0]):
                o, n, shape = self.offset[name]
                view = self.G[o:o + n].view(shape)
                if p.grad is not None:
                    view.copy_(p.grad)
                else:
                    view.zero_()
                p.grad = view

    def zero_grad(self):
        self.G.zero_()

    def refresh_shadows(self, ema, momentum=0.995):
        """ema=True: p_m = p_m*m + p*(1-m) over the EMA range + both bf16 shadows in one kernel (K15);
        the tail (heads outside the EMA set) gets a plain fp32->bf16 cast."""
        if ema:
            K.ema(self.P[:self.n_ema], self.M, self.P16[:self.n_ema], self.M16, momentum)
        else:
            K.cast_bf16(self.P[:self.n_ema], self.P16[:self.n_ema])
            K.cast_bf16(self.M, self.M16)
        if self.n_total > self.n_ema:
            K.cast_bf16(self.P[self.n_ema:], self.P16[self.n_ema:])


# ---------------------------------------------------------------------------------------------- weight bundles
def _attn_bundle(A, prefix, cfg, momentum, cross):
    g = (lambda n, c=1: None) if momentum else A.grad
    W = SimpleNamespace(heads=cfg.num_attention_heads, eps=cfg.layer_norm_eps)
    s = prefix + ".self."
    if cross:
        W.wq, W.bq = A.w16(s + "query.weight", momentum), A.f32(s + "query.bias", momentum)
        W.wkv, W.bkv = A.w16(s + "key.weight", momentum, 2), A.f32(s + "key.bias", momentum, 2)
        W.g_wq, W.g_bq = g(s + "query.weight"), g(s + "query.bias")
        W.g_wkv, W.g_bkv = g(s + "key.weight", 2), g(s + "key.bias", 2)
    else:
        W.wqkv, W.bqkv = A.w16(s + "query.weight", momentum, 3), A.f32(s + "query.bias", momentum, 3)
        W.g_wqkv, W.g_bqkv = g(s + "query.weight", 3), g(s + "query.bias", 3)
    o = prefix + ".output."
    W.wo, W.bo = A.w16(o + "dense.weight", momentum), A.f32(o + "dense.bias", momentum)
    W.ln_g, W.ln_b = A.f32(o + "LayerNorm.weight", momentum), A.f32(o + "LayerNorm.bias", momentum)
    W.g_wo, W.g_bo = g(o + "dense.weight"), g(o + "dense.bias")
    W.g_ln_g, W.g_ln_b = g(o + "LayerNorm.weight"), g(o + "LayerNorm.bias")
    return W


def _ffn_bundle(A, prefix, cfg, momentum):
    g = (lambda n, c=1: None) if momentum else A.grad
    W = SimpleNamespace(eps=cfg.layer_norm_eps)
    W.w1, W.b1 = A.w16(prefix + ".intermediate.dense.weight", momentum), A.f32(prefix + ".intermediate.dense.bias", momentum)
    W.w2, W.b2 = A.w16(prefix + ".output.dense.weight", momentum), A.f32(prefix + ".output.dense.bias", momentum)
    W.ln_g, W.ln_b = A.f32(prefix + ".output.LayerNorm.weight", momentum), A.f32(prefix + ".output.LayerNorm.bias", momentum)
    W.g_w1, W.g_b1 = g(prefix + ".intermediate.dense.weight"), g(prefix + ".intermediate.dense.bias")
    W.g_w2, W.g_b2 = g(prefix + ".output.dense.weight"), g(prefix + ".output.dense.bias")
    W.g_ln_g, W.g_ln_b = g(prefix + ".output.LayerNorm.weight"), g(prefix + ".output.LayerNorm.bias")
    return W


def bert_bundles(A, name, bert, momentum, anchor, head_prefix=None, is_property=False):
    """Attaches the kernel-facing weight views of one BertModel (arena name prefix `name`, e.g. 'text_encoder.bert')."""
    cfg = bert.config
    # arena names are the ONLINE names; momentum twins live at the same offsets of the M arenas
    on = name.replace("_m.", ".", 1) if "_m." in name else (name[:-2] if name.endswith("_m") else name)
    g = (lambda n, c=1: None) if momentum else A.grad
    e = on + ".embeddings."
    emb = SimpleNamespace(H=cfg.hidden_size, eps=cfg.layer_norm_eps, pad_id=cfg.pad_token_id)
    emb.word = A.f32(e + "word_embeddings.weight", momentum)
    emb.pos = A.f32(e + "position_embeddings.weight", momentum)
    emb.type0 = A.f32(e + "token_type_embeddings.weight", momentum)[0]
    emb.ln_g, emb.ln_b = A.f32(e + "LayerNorm.weight", momentum), A.f32(e + "LayerNorm.bias", momentum)
    emb.g_word = None if (momentum or is_property) else g(e + "word_embeddings.weight")
    emb.g_pos = g(e + "position_embeddings.weight")
    gt = g(e + "token_type_embeddings.weight")
    emb.g_type0 = None if gt is None else gt[0]
    emb.g_ln_g, emb.g_ln_b = g(e + "LayerNorm.weight"), g(e + "LayerNorm.bias")
    layers = []
    for i, layer in enumerate(bert.encoder.layer):
        lp = "%s.encoder.layer.%d" % (on, i)
        layers.append(SimpleNamespace(
            attn=_attn_bundle(A, lp + ".attention", cfg, momentum, False),
            cross=_attn_bundle(A, lp + ".crossattention", cfg, momentum, True) if layer.has_cross_attention else None,
            ffn=_ffn_bundle(A, lp, cfg, momentum),
            grad_range=None if momentum else A.prefix_range(lp + ".")))
    head = None
    if head_prefix is not None:
        hp = head_prefix
        head = SimpleNamespace(eps=cfg.layer_norm_eps)
        head.wt, head.bt = A.w16(hp + ".transform.dense.weight", momentum), A.f32(hp + ".transform.dense.bias", momentum)
        head.ln_g, head.ln_b = A.f32(hp + ".transform.LayerNorm.weight", momentum), A.f32(hp + ".transform.LayerNorm.bias", momentum)
        head.wdec, head.bdec = A.w16(e + "word_embeddings.weight", momentum), A.f32(hp + ".bias", momentum)
        head.g_wt, head.g_bt = g(hp + ".transform.dense.weight"), g(hp + ".transform.dense.bias")
        head.g_ln_g, head.g_ln_b = g(hp + ".transform.LayerNorm.weight"), g(hp + ".transform.LayerNorm.bias")
        head.g_wdec = g(e + "word_embeddings.weight")
        gb = g(hp + ".bias")
        if gb is not None:   # colsum writes in multiples of 8 columns; the arena pads every tensor to 64
            o, n, _ = A.offset[hp + ".bias"]
            gb = A.G[o:o + _pad(n)]
        head.g_bdec = gb
    bert._spmm_bundles = SimpleNamespace(emb=emb, layers=layers, head=head, anchor=anchor)


def linear_bundle(A, name, momentum=False):
    on = name[:-2] if name.endswith("_m") else name
    g = (lambda n: None) if momentum else A.grad
    return SimpleNamespace(w=A.w16(on + ".weight", momentum), b=A.f32(on + ".bias", momentum),
                           g_w=g(on + ".weight"), g_b=g(on + ".bias"))
