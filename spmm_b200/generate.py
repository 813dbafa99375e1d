"""Inference drivers mirroring the reference's d_smiles2pv.py:14-52 (SMILES -> 53 property values, autoregressive over
the property tokens) and d_pv2smiles_batched.py:17-59 + d_pv2smiles_single.py:26-51 (property vector -> SMILES, beam
search with k beams over the fusion decoder).

Two layers:
  * the reference-shaped loops over the sub-module API (`smiles2pv`, `pv2smiles`: what the d_*.py scripts do, every
    pass through the same sm_100a kernels as pre-training, full prefix recomputed per token like the reference), and
  * `PvDecoder` / `pv2smiles_batched`: the B200-native decode path - KV-cached single-token steps for MANY molecules'
    beams at once, cross K/V of the 54 property tokens projected once per molecule, beam bookkeeping on the device
    (csrc/decode.cu), the whole step captured in ONE CUDA graph that is replayed per token with no host sync.
"""
import math

import torch

from . import kernels as K
from . import ops


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


# ====================================================================================================================
# KV-cached batched beam decode (B200-native path for d_pv2smiles_batched.py:24-59)
class BeamState:
    """Device-side beam-search state consumed by `spmm_beam_step` (csrc/decode.cu): k live beams per molecule."""

    def __init__(self, n_mol, k, tmax, device, trace_steps=0):
        self.n_mol, self.k, self.tmax, self.fin_cap = n_mol, k, tmax, k * k + k
        R = n_mol * k
        i32, i64, f32 = dict(device=device, dtype=torch.int32), dict(device=device, dtype=torch.int64), dict(device=device, dtype=torch.float32)
        self.t_dev = torch.zeros(1, **i32)
        self.scores = torch.zeros(n_mol, k, **f32)
        self.tokens = torch.zeros(R, tmax, **i64)
        self.anc = torch.zeros(R, tmax, **i32)
        self.next_ids = torch.zeros(R, **i64)
        self.fin_scores = torch.zeros(n_mol, self.fin_cap, **f32)
        self.fin_tokens = torch.zeros(n_mol, self.fin_cap, tmax, **i64)
        self.fin_len = torch.zeros(n_mol, self.fin_cap, **i32)
        self.fin_count = torch.zeros(n_mol, **i32)
        self.done = torch.zeros(n_mol, **i32)
        self.ticket = torch.zeros(1, **i32)
        self.trace_logp = torch.zeros(trace_steps, R, k, **f32) if trace_steps else None
        self.trace_tok = torch.zeros(trace_steps, R, k, **i32) if trace_steps else None

    def reset(self, cls_id):
        for t in (self.t_dev, self.scores, self.tokens, self.anc, self.fin_scores, self.fin_tokens, self.fin_len,
                  self.fin_count, self.done):
            t.zero_()
        self.tokens[:, 0] = cls_id
        self.next_ids.fill_(cls_id)


class PvDecoder:
    """k-beam PV -> SMILES decoding of `n_mol` molecules at a time (rows = n_mol * k live beams).

    Per generated token ONE graph replay: embedding of the new token at position t, 12 decoder layers on [rows, H]
    (fused QKV GEMM -> cached causal self-attention -> output GEMM + residual -> LayerNorm; for the 6 fusion layers a
    query GEMM -> cross-attention over the molecule's 54 cached property keys -> output GEMM + LayerNorm; FFN), LM head,
    device-side beam update.  The reference recomputes the whole prefix for every token (d_pv2smiles_single.py:29-36);
    the stack is causal, so the cached path computes the same last-position logits."""

    def __init__(self, model, n_mol, k=2, tmax=104, trace_steps=0, use_graph=True):
        self.model, self.n_mol, self.k, self.tmax = model, n_mol, k, tmax
        te = model.text_encoder
        self.bd = te.bert._bundles()
        cfg = te.config
        self.H, self.heads, self.V, self.ld = cfg.hidden_size, cfg.num_attention_heads, cfg.vocab_size, te.logit_ld()
        self.I = cfg.intermediate_size
        dev = model.arena().device
        self.dev = dev
        R = n_mol * k
        self.R = R
        self.state = BeamState(n_mol, k, tmax, dev, trace_steps)
        nl = len(self.bd.layers)
        self.cache_k = [torch.zeros(tmax, R, self.H, device=dev, dtype=torch.bfloat16) for _ in range(nl)]
        self.cache_v = [torch.zeros(tmax, R, self.H, device=dev, dtype=torch.bfloat16) for _ in range(nl)]
        self.n_pv = 54
        self.cross_kv = {i: torch.zeros(n_mol * self.n_pv, 2 * self.H, device=dev, dtype=torch.bfloat16)
                         for i, lw in enumerate(self.bd.layers) if lw.cross is not None}
        self.use_graph = use_graph
        self.graph = None
        self.scale = 1.0 / math.sqrt(self.H // self.heads)

    # ---- once per batch of molecules: property encoder + cross K/V of its 54 tokens for the 6 fusion layers
    @torch.no_grad()
    def encode(self, prop):
        from .xbert import raw_outputs
        m = self.model
        assert prop.shape[0] == self.n_mol
        tokens = torch.cat([m.property_cls.expand(prop.shape[0], -1, -1), m.property_embed(prop.unsqueeze(2))], dim=1)
        assert tokens.shape[1] == self.n_pv
        with raw_outputs():
            pe = m.property_encoder(inputs_embeds=tokens, return_dict=True).last_hidden_state      # [N, 54, H] bf16
        pe2 = pe.reshape(self.n_mol * self.n_pv, self.H)
        for i, buf in self.cross_kv.items():
            C = self.bd.layers[i].cross
            K.gemm(pe2, C.wkv, pe2.shape[0], 2 * self.H, self.H, bias=C.bkv, out=buf)

    def _ln(self, x, W):
        return K.layernorm_fwd(x, W.ln_g, W.ln_b, W.eps, save_stats=False)[0]

    # ---- one token for every live beam (the captured graph body)
    def _step(self):
        st, bd, R, H = self.state, self.bd, self.R, self.H
        emb = bd.emb
        x = K.decode_embed(st.next_ids, st.t_dev, emb.word, emb.pos, emb.type0, H)
        x = self._ln(x, emb)
        for i, lw in enumerate(bd.layers):
            W = lw.attn
            qkv = K.gemm(x, W.wqkv, R, 3 * H, H, bias=W.bqkv)
            o = torch.empty(R, H, device=self.dev, dtype=torch.bfloat16)
            K.decode_attn_self(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], self.cache_k[i], self.cache_v[i], st.anc, st.tokens,
                               st.t_dev, o, self.heads, self.scale)
            x = self._ln(K.gemm(o, W.wo, R, H, H, bias=W.bo, residual=x), W)
            if lw.cross is not None:
                C = lw.cross
                q = K.gemm(x, C.wq, R, H, H, bias=C.bq)
                kv = self.cross_kv[i]
                o2 = torch.empty(R, H, device=self.dev, dtype=torch.bfloat16)
                K.decode_attn_cross(q, kv[:, :H], kv[:, H:], self.n_pv, self.k, o2, self.heads, self.scale)
                x = self._ln(K.gemm(o2, C.wo, R, H, H, bias=C.bo, residual=x), C)
            F_ = lw.ffn
            act = K.gemm(x, F_.w1, R, self.I, H, bias=F_.b1, gelu=True)
            x = self._ln(K.gemm(act, F_.w2, R, H, self.I, bias=F_.b2, residual=x), F_)
        logits = ops.lm_logits(x, bd.head, self.V, self.ld)
        K.beam_step(logits, self.V, st, 0, self.sep_id)
        return logits

    @torch.no_grad()
    def generate(self, prop, cls_id=2, sep_id=3, max_steps=100, on_step=None):
        """Beam search for prop [n_mol, 53]; returns, per molecule, up to k (log-prob, token ids incl. [CLS] .. [SEP])
        pairs, best first (d_pv2smiles_batched.py:24-52).  `on_step` (eager mode only) is called as on_step(t, None) before and on_step(t, logits) after each step."""
        self.model.eval()
        self.sep_id = sep_id
        st = self.state
        st.reset(cls_id)
        self.encode(prop)
        n_steps = min(max_steps + 1, self.tmax - 2)
        if st.trace_logp is not None:
            n_steps = min(n_steps, st.trace_logp.shape[0])
        if self.use_graph and on_step is None:
            if self.graph is None:
                self._step()                                   # warm-up (lazy kernel attributes); state is reset below
                torch.cuda.synchronize()
                st.reset(cls_id)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._step()
            for t in range(n_steps):
                self.graph.replay()
                if t % 16 == 15 and bool(st.done.all()):       # one sync per 16 tokens
                    break
        else:
            for t in range(n_steps):
                if on_step is not None:
                    on_step(t, None)                           # state before the step (prefixes that are fed)
                logits = self._step()
                if on_step is not None:
                    on_step(t, logits)
                if t % 16 == 15 and bool(st.done.all()):
                    break
        cnt = st.fin_count.tolist()
        sc, ln, tk = st.fin_scores.cpu(), st.fin_len.cpu(), st.fin_tokens.cpu()
        out = []
        for m in range(self.n_mol):
            items = [(float(sc[m, e]), tk[m, e, :int(ln[m, e])].clone()) for e in range(min(cnt[m], st.fin_cap))]
            items.sort(key=lambda it: it[0], reverse=True)     # stable, like sorted(final_output, key=p, reverse=True)
            out.append(items[:self.k])
        return out


_DECODERS = {}


@torch.no_grad()
def pv2smiles_batched(model, prop, cls_id=2, sep_id=3, k=2, max_steps=100):
    """d_pv2smiles_batched.py:24-52 for a whole batch prop [N, 53] at once (deterministic beams): list of N result lists."""
    key = (id(model), prop.shape[0], k)
    dec = _DECODERS.get(key)
    if dec is None or dec.model is not model:
        dec = _DECODERS[key] = PvDecoder(model, prop.shape[0], k=k)
    return dec.generate(prop.to(model.arena().device).float(), cls_id, sep_id, max_steps)


@torch.no_grad()
def evaluate_pv2smiles(model, data_loader, tokenizer, k=2):
    """`evaluate` of d_pv2smiles_batched.py:17-59: (reference SMILES, best generated SMILES) per molecule; every loader
    batch is decoded at once."""
    reference, candidate = [], []
    for prop, text in data_loader:
        res = pv2smiles_batched(model, prop, tokenizer.cls_token_id, tokenizer.sep_token_id, k=k)
        for b, items in enumerate(res):
            reference.append(text[b].replace('[CLS]', ''))
            if items:
                sent = items[0][1]
                candidate.append(tokenizer.convert_tokens_to_string(tokenizer.convert_ids_to_tokens(sent[:-1])).replace('[CLS]', ''))
            else:
                candidate.append('')
    return reference, candidate


# ====================================================================================================================
# SMILES -> PV with cached text-side cross K/V (B200-native path for d_smiles2pv.py:14-52)
class Smiles2PvDecoder:
    """53 autoregressive property predictions for a batch of B tokenised SMILES.

    What can be cached is cached: the text encoder runs once, and the keys / values every fusion layer derives from its
    output (6 x [B*L, 2H]) are projected once instead of 53 times.  The property side cannot be cached - the PV encoder
    is bidirectional over the growing prefix, so every earlier position changes when a token is appended (the reference
    recomputes it too, d_smiles2pv.py:15) - but each step is ONE CUDA-graph replay: the prefix lives in a fixed-size
    buffer (three length buckets 16 / 32 / 56), its current length is a device scalar that feeds the attention
    kernels' key lengths, the last-position gather and the write of the next token, so no step syncs with the host."""

    BUCKETS = (16, 32, 56)
    TMAX = 56

    def __init__(self, model, B, L):
        self.model, self.B, self.L = model, B, L
        te, pe = model.text_encoder, model.property_encoder
        self.tb, self.pb = te.bert._bundles(), pe._bundles()
        cfg = te.config
        self.H, self.heads, self.I = cfg.hidden_size, cfg.num_attention_heads, cfg.intermediate_size
        self.fl = cfg.fusion_layer
        dev = model.arena().device
        self.dev = dev
        H = self.H
        bf = dict(device=dev, dtype=torch.bfloat16)
        self.prop_in = torch.zeros(B, self.TMAX, H, **bf)
        self.t_dev = torch.ones(1, device=dev, dtype=torch.int32)             # current prefix length T
        self.kv_len_pv = torch.ones(B, device=dev, dtype=torch.int32)
        self.text_len = torch.zeros(B, device=dev, dtype=torch.int32)
        self.cross_kv = [torch.zeros(B * L, 2 * H, **bf) for _ in range(self.fl, cfg.num_hidden_layers)]
        self.preds = torch.zeros(B, 53, device=dev, dtype=torch.float32)
        self.base = {Tp: torch.arange(B, device=dev, dtype=torch.int32) * Tp for Tp in self.BUCKETS}
        self.base_max = torch.arange(B, device=dev, dtype=torch.int64) * self.TMAX
        self.graphs = {}
        self.scale = 1.0 / math.sqrt(H // self.heads)
        W = model._W
        self.mtr, self.pvw = W["mtr"], W["pv"]

    def _ln(self, x, W):
        return K.layernorm_fwd(x, W.ln_g, W.ln_b, W.eps, save_stats=False)[0]

    def _self_block(self, x, W, Tp, causal):
        B, H = self.B, self.H
        M = B * Tp
        qkv = K.gemm(x, W.wqkv, M, 3 * H, H, bias=W.bqkv)
        o = torch.empty(M, H, device=self.dev, dtype=torch.bfloat16)
        K.attn_fwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], o, None, B, self.heads, Tp, Tp, self.kv_len_pv, causal, self.scale)
        return self._ln(K.gemm(o, W.wo, M, H, H, bias=W.bo, residual=x), W)

    def _ffn(self, x, F_, M):
        act = K.gemm(x, F_.w1, M, self.I, self.H, bias=F_.b1, gelu=True)
        return self._ln(K.gemm(act, F_.w2, M, self.H, self.I, bias=F_.b2, residual=x), F_)

    def _step(self, Tp):
        B, H = self.B, self.H
        M = B * Tp
        self.kv_len_pv.copy_(self.t_dev.expand(B))
        x_in = self.prop_in[:, :Tp].contiguous()
        emb = self.pb.emb
        x = self._ln(K.embed_inputs_fwd(x_in, emb.pos, emb.type0), emb)
        for lw in self.pb.layers:                                    # property encoder, bidirectional (d_smiles2pv.py:15)
            x = self._self_block(x, lw.attn, Tp, False)
            x = self._ffn(x, lw.ffn, M)
        for j, lw in enumerate(self.tb.layers[self.fl:]):            # fusion layers, causal over the prefix (:17-24)
            x = self._self_block(x, lw.attn, Tp, True)
            C = lw.cross
            q = K.gemm(x, C.wq, M, H, H, bias=C.bq)
            kv = self.cross_kv[j]
            o = torch.empty(M, H, device=self.dev, dtype=torch.bfloat16)
            K.attn_fwd(q, kv[:, :H], kv[:, H:], o, None, B, self.heads, Tp, self.L, self.text_len, False, self.scale)
            x = self._ln(K.gemm(o, C.wo, M, H, H, bias=C.bo, residual=x), C)
            x = self._ffn(x, lw.ffn, M)
        # regression head at the last position only (the reference applies it everywhere and reads [:, -1], :25-26)
        idx = self.base[Tp] + self.t_dev - 1
        last = K.gather_rows(x, idx, B)
        mt = self.mtr
        a = K.gemm(last, mt.w0, B, H, H, bias=mt.b0, gelu=True)
        t = K.layernorm_fwd(a, mt.ln_g, mt.ln_b, mt.eps, save_stats=False)[0]
        pred = t.float() @ mt.w3.view(-1) + mt.b3                   # [B]
        tl = self.t_dev.long()
        self.preds.index_copy_(1, tl - 1, pred[:, None])
        nxt = (pred[:, None] * self.pvw.w[None, :] + self.pvw.b[None, :]).to(torch.bfloat16)     # property_embed, :49
        self.prop_in.view(B * self.TMAX, H).index_copy_(0, self.base_max + tl, nxt)
        self.t_dev.add_(1)

    @torch.no_grad()
    def generate(self, input_ids, attention_mask, n_prop=53):
        from .xbert import raw_outputs
        m = self.model
        m.eval()
        B, L, H = self.B, self.L, self.H
        assert tuple(input_ids.shape) == (B, L) and n_prop <= 53
        with raw_outputs():
            text = m.text_encoder.bert(input_ids, attention_mask=attention_mask, return_dict=True, mode='text').last_hidden_state
        t2 = text.reshape(B * L, H)
        for j, lw in enumerate(self.tb.layers[self.fl:]):
            K.gemm(t2, lw.cross.wkv, B * L, 2 * H, H, bias=lw.cross.bkv, out=self.cross_kv[j])
        self.text_len.copy_(attention_mask.sum(1).to(torch.int32))
        self.prop_in.zero_()
        self.prop_in[:, 0] = m.property_cls.view(1, H).to(torch.bfloat16)
        self.t_dev.fill_(1)
        for s in range(n_prop):
            Tp = next(b for b in self.BUCKETS if s + 1 <= b)
            g = self.graphs.get(Tp)
            if g is None:
                snap = (self.prop_in.clone(), self.preds.clone(), self.t_dev.clone())
                self._step(Tp)                                       # warm-up outside capture, then undo it
                torch.cuda.synchronize()
                self.prop_in.copy_(snap[0]); self.preds.copy_(snap[1]); self.t_dev.copy_(snap[2])
                g = self.graphs[Tp] = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step(Tp)
            g.replay()
        return self.preds[:, :n_prop].clone()


_S2P = {}


@torch.no_grad()
def smiles2pv_fast(model, input_ids, attention_mask, n_prop=53):
    """d_smiles2pv.py:41-52 through `Smiles2PvDecoder` (cached text-side cross K/V, graph-replayed steps)."""
    key = (id(model), tuple(input_ids.shape))
    dec = _S2P.get(key)
    if dec is None or dec.model is not model:
        dec = _S2P[key] = Smiles2PvDecoder(model, input_ids.shape[0], input_ids.shape[1])
    dev = model.arena().device
    return dec.generate(input_ids.to(dev), attention_mask.to(dev), n_prop)
