"""Block-level autograd functions of the SPMM step.  Each one is a short sequence of C-ABI kernel launches
(spmm_b200/kernels.py); PyTorch only provides tensors, streams and the autograd tape.

Granularity follows the reference's modules so that the residual / LayerNorm / dropout / bias work can be fused
into GEMM epilogues and the LayerNorm-backward kernel:
  attn_block  = BertAttention        (xbert.py:376-422: BertSelfAttention + BertSelfOutput)
  ffn_block   = BertIntermediate + BertOutput                (xbert.py:425-451)
  embed_*     = BertEmbeddings                               (xbert.py:173-220)
  pv_tokens   = SPMM_models.py:82-88
  lm_head_loss, itm_loss, mtr_head_loss, itc = the four loss heads (SPMM_models.py:102-131,201-206,233-238,251-254)

Weight gradients are accumulated by the wgrad GEMMs straight into the flat fp32 gradient arena (spmm_b200/arena.py);
the functions therefore return gradients only for activations.  `anchor` is a dummy requires-grad scalar that keeps
a backward node alive for functions whose only differentiable inputs are parameters.

The product path has no fallback: every function here launches kernels from libspmm_b200.so.
"""
import math
from types import SimpleNamespace

import os

import torch

from . import kernels as K

BF16 = torch.bfloat16
_rng = SimpleNamespace(seed=0x5EED5EED, counter=0)


def manual_seed(seed):
    """Seeds the counter-based dropout / sampler streams (explicit generator, north star)."""
    _rng.seed = int(seed) & 0xFFFFFFFFFFFF
    _rng.counter = 0


def next_seed():
    _rng.counter += 1
    return ((_rng.seed * 0x9E3779B97F4A7C15) + _rng.counter * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF


class StepRng:
    """Per-training-step randomness that survives CUDA-graph replay.

    Per-op seeds (`next_seed()`) are host integers and get baked into a captured graph; the kernels therefore add a
    DEVICE scalar ("salt") to every seed.  `advance()` increments the salt ON THE DEVICE (a captured graph replays the
    increment), so a host that enqueues steps ahead of the GPU cannot make two steps share a salt; `host` mirrors it.
    """

    def __init__(self, device):
        if torch.device(device).type != "cuda":
            raise K._lib.SpmmKernelError("StepRng lives on the GPU the kernels run on, got %s" % (device,))
        self.host = torch.zeros(1, dtype=torch.int64).pin_memory()
        self.dev = torch.zeros(1, dtype=torch.int64, device=device)
        K.set_rng_salt(self.dev)

    def advance(self, bump_host=True):
        if bump_host:
            self.host += 1
        self.dev.add_(1)

    def reset(self, value=0):
        self.host.fill_(value)
        self.dev.fill_(value)


_step_rngs = {}


def step_rng(device):
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:      # "cuda" and "cuda:0" must name the same generator
        device = torch.device("cuda", torch.cuda.current_device())
    r = _step_rngs.get(device)
    if r is None:
        r = _step_rngs[device] = StepRng(device)
    return r


def _wgrad(dy, x, gw, n_out, k_in, m_tokens):
    """gw[n_out, k_in] += dy^T . x  (both operands MN-major: no transpose pass)."""
    if gw is not None:
        K.gemm(dy, x, n_out, k_in, m_tokens, a_mn=True, b_mn=True, out=gw, out_f32=True, accumulate=True)


# --------------------------------------------------------------------------------------------- backward markers
class _GradReady(torch.autograd.Function):
    """Identity whose backward runs `callback()` first.  Placed on the INPUT of a layer in the first forward pass that
    uses the layer's weights: autograd runs nodes in reverse creation order, so when this backward fires every kernel
    that adds to those weights' gradients (in this and all later passes) has been enqueued."""

    @staticmethod
    def forward(ctx, x, callback):
        ctx.callback = callback
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        ctx.callback()
        return g, None


def grad_ready(x, callback):
    return _GradReady.apply(x, callback)


# --------------------------------------------------------------------------------------------- attention block
class _AttnBlock(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, enc, anchor, W, g, p_attn, p_hid, grad_mode):
        # needs_input_grad ignores torch.no_grad() (and grad mode is always off inside forward, so the caller passes
        # it): the momentum passes must not write backward state (LSE, LayerNorm statistics, 2nd GELU-GEMM output)
        need = grad_mode and any(ctx.needs_input_grad)
        M, H = x.shape
        scale = 1.0 / math.sqrt(H // W.heads)
        if enc is None:
            qkv = K.gemm(x, W.wqkv, M, 3 * H, H, bias=W.bqkv)
            q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
            kvbuf = None
        else:
            qkv = K.gemm(x, W.wq, M, H, H, bias=W.bq)
            q = qkv
            kvbuf = K.gemm(enc, W.wkv, enc.shape[0], 2 * H, H, bias=W.bkv)
            k, v = kvbuf[:, :H], kvbuf[:, H:]
        o = torch.empty(M, H, device=x.device, dtype=BF16)
        lse = torch.empty(g.B * W.heads * g.Tq, device=x.device, dtype=torch.float32) if need else None
        seed_a = next_seed() if p_attn > 0 else 0
        # cross attention over SHARED encoder states: `enc` holds the distinct states, g.kv_index maps each query batch
        # element to one of them, so the K/V projection runs once per distinct state (SPMM_models.py:137-198 pairs
        # the same states with positives and hard negatives)
        kvi = getattr(g, 'kv_index', None)
        nkv = enc.shape[0] // g.Tk if kvi is not None else 0
        K.attn_fwd(q, k, v, o, lse, g.B, W.heads, g.Tq, g.Tk, g.kv_len, g.causal, scale, p_attn, seed_a,
                   kv_index=kvi, kv_batches=nkv)
        seed_h = next_seed() if p_hid > 0 else 0
        xsum = K.gemm(o, W.wo, M, H, H, bias=W.bo, residual=x, dropout_p=p_hid, seed=seed_h)
        y, mean, rstd = K.layernorm_fwd(xsum, W.ln_g, W.ln_b, W.eps, save_stats=need)
        if need:
            ctx.save_for_backward(x, enc, qkv, kvbuf, o, lse, xsum, mean, rstd)
            ctx.W, ctx.g, ctx.p = W, g, (p_attn, p_hid, seed_a, seed_h, scale)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, enc, qkv, kvbuf, o, lse, xsum, mean, rstd = ctx.saved_tensors
        W, g = ctx.W, ctx.g
        p_attn, p_hid, seed_a, seed_h, scale = ctx.p
        M, H = x.shape
        dy = dy.contiguous()
        dxs, dxb = K.layernorm_bwd(dy, xsum, mean, rstd, W.ln_g, W.g_ln_g, W.g_ln_b, dbias=W.g_bo, want_branch=True,
                                   branch_dropout_p=p_hid, branch_seed=seed_h)
        _wgrad(dxb, o, W.g_wo, H, H, M)
        do = K.gemm(dxb, W.wo, M, H, H, b_mn=True)
        if enc is None:
            q, k, v = qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:]
            dqkv = torch.empty_like(qkv)
            gb = W.g_bqkv                                 # [q.b | k.b | v.b] gradients: summed inside the attention kernel
            K.attn_bwd(do, q, k, v, o, lse, dqkv[:, :H], dqkv[:, H:2 * H], dqkv[:, 2 * H:], g.B, W.heads, g.Tq, g.Tk,
                       g.kv_len, g.causal, scale, p_attn, seed_a,
                       dbias=None if gb is None else (gb[:H], gb[H:2 * H], gb[2 * H:]))
            _wgrad(dqkv, x, W.g_wqkv, 3 * H, H, M)
            dx = K.gemm(dqkv, W.wqkv, M, H, 3 * H, b_mn=True, residual=dxs)
            return dx, None, None, None, None, None, None, None
        Mk = enc.shape[0]
        k, v = kvbuf[:, :H], kvbuf[:, H:]
        dq = torch.empty_like(qkv)
        kvi = getattr(g, 'kv_index', None)
        dkv = torch.empty_like(kvbuf) if kvi is None else torch.empty(g.B * g.Tk, 2 * H, device=x.device, dtype=BF16)
        K.attn_bwd(do, qkv, k, v, o, lse, dq, dkv[:, :H], dkv[:, H:], g.B, W.heads, g.Tq, g.Tk, g.kv_len, g.causal,
                   scale, p_attn, seed_a,
                   dbias=None if W.g_bq is None else (W.g_bq, W.g_bkv[:H], W.g_bkv[H:]),
                   kv_index=kvi, kv_batches=Mk // g.Tk if kvi is not None else 0)
        if kvi is not None:                              # per-pair dK/dV -> the distinct states they were projected from
            dkv = K.segment_sum_rows(dkv.view(g.B, g.Tk * 2 * H), kvi, Mk // g.Tk).view(Mk, 2 * H)
        _wgrad(dq, x, W.g_wq, H, H, M)
        dx = K.gemm(dq, W.wq, M, H, H, b_mn=True, residual=dxs)
        _wgrad(dkv, enc, W.g_wkv, 2 * H, H, Mk)
        denc = K.gemm(dkv, W.wkv, Mk, H, 2 * H, b_mn=True) if ctx.needs_input_grad[1] else None
        return dx, denc, None, None, None, None, None, None


def attn_block(x, enc, W, geom, p_attn, p_hid, anchor):
    """LN(dropout(dense(attention(x, enc or x))) + x) on [tokens, H] bf16."""
    return _AttnBlock.apply(x, enc, anchor, W, geom, p_attn, p_hid, torch.is_grad_enabled())


# --------------------------------------------------------------------------------------------- FFN block
_DGELU_STORED = os.environ.get("SPMM_DGELU_STORED", "1") != "0"


class _FfnBlock(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, W, p_hid, grad_mode):
        need = grad_mode and any(ctx.needs_input_grad)
        M, H = x.shape
        I = W.w1.shape[0]
        pre = torch.empty(M, I, device=x.device, dtype=BF16) if need else None
        # `pre` receives gelu'(x W1^T + b1) (same erf evaluation as the activation): backward only multiplies
        act = K.gemm(x, W.w1, M, I, H, bias=W.b1, gelu=True, pre_act_out=pre, dgelu_stored=_DGELU_STORED)
        seed_h = next_seed() if p_hid > 0 else 0
        xsum = K.gemm(act, W.w2, M, H, I, bias=W.b2, residual=x, dropout_p=p_hid, seed=seed_h)
        y, mean, rstd = K.layernorm_fwd(xsum, W.ln_g, W.ln_b, W.eps, save_stats=need)
        if need:
            ctx.save_for_backward(x, pre, act, xsum, mean, rstd)
            ctx.W, ctx.p = W, (p_hid, seed_h)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, pre, act, xsum, mean, rstd = ctx.saved_tensors
        W = ctx.W
        p_hid, seed_h = ctx.p
        M, H = x.shape
        I = W.w1.shape[0]
        dxs, dxb = K.layernorm_bwd(dy.contiguous(), xsum, mean, rstd, W.ln_g, W.g_ln_g, W.g_ln_b, dbias=W.g_b2,
                                   want_branch=True, branch_dropout_p=p_hid, branch_seed=seed_h)
        _wgrad(dxb, act, W.g_w2, H, I, M)
        # (dxb . W2) * gelu'(pre); the intermediate bias gradient (column sums of dpre) rides on the same epilogue
        dpre = K.gemm(dxb, W.w2, M, I, H, b_mn=True, dgelu_pre=pre, dgelu_stored=_DGELU_STORED, colsum_out=W.g_b1)
        _wgrad(dpre, x, W.g_w1, I, H, M)
        dx = K.gemm(dpre, W.w1, M, H, I, b_mn=True, residual=dxs)
        return dx, None, None, None, None


def ffn_block(x, W, p_hid, anchor):
    """LN(dropout(dense2(gelu(dense1(x)))) + x)."""
    return _FfnBlock.apply(x, anchor, W, p_hid, torch.is_grad_enabled())


# --------------------------------------------------------------------------------------------- embeddings
class _EmbedText(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, ids, W, p):
        x = K.embed_text_fwd(ids, W.word, W.pos, W.type0, W.H)
        seed = next_seed() if p > 0 else 0
        y, mean, rstd = K.layernorm_fwd(x, W.ln_g, W.ln_b, W.eps, save_stats=True, dropout_p=p, seed=seed)
        ctx.save_for_backward(x, mean, rstd, ids)
        ctx.W, ctx.p = W, (p, seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, ids = ctx.saved_tensors
        W = ctx.W
        dx, _ = K.layernorm_bwd(dy.contiguous(), x, mean, rstd, W.ln_g, W.g_ln_g, W.g_ln_b, out_dropout_p=ctx.p[0],
                                out_seed=ctx.p[1])
        K.embed_text_bwd(dx, ids, W.g_word, W.g_pos, W.g_type0, W.pad_id)
        return None, None, None, None


def embed_text(ids, W, p, anchor):
    """dropout(LN(word[ids] + type[0] + pos)) -> [B*T, H] bf16."""
    return _EmbedText.apply(anchor, ids, W, p)


class _EmbedInputs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, anchor, W, p):
        x = K.embed_inputs_fwd(inputs, W.pos, W.type0)
        seed = next_seed() if p > 0 else 0
        y, mean, rstd = K.layernorm_fwd(x, W.ln_g, W.ln_b, W.eps, save_stats=True, dropout_p=p, seed=seed)
        ctx.save_for_backward(x, mean, rstd)
        ctx.W, ctx.p, ctx.shape = W, (p, seed), inputs.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        W = ctx.W
        dx, _ = K.layernorm_bwd(dy.contiguous(), x, mean, rstd, W.ln_g, W.g_ln_g, W.g_ln_b, out_dropout_p=ctx.p[0],
                                out_seed=ctx.p[1])
        if W.g_pos is not None:
            K.embed_inputs_bwd(dx, ctx.shape[1], W.g_pos, W.g_type0)
        return dx.view(ctx.shape), None, None, None


def embed_inputs(inputs, W, p, anchor):
    """dropout(LN(inputs_embeds + type[0] + pos)); inputs [B,T,H] bf16 -> [B*T, H]."""
    return _EmbedInputs.apply(inputs, anchor, W, p)


class _PvTokens(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, pv, mpm, W):
        ctx.save_for_backward(pv, mpm)
        ctx.W = W
        return K.pv_tokens_fwd(pv, mpm, W.w, W.b, W.cls, W.mask)

    @staticmethod
    def backward(ctx, dprop):
        pv, mpm = ctx.saved_tensors
        W = ctx.W
        K.pv_tokens_bwd(dprop.contiguous(), pv, mpm, W.g_w, W.g_b, W.g_cls, W.g_mask)
        return None, None, None, None


def pv_tokens(pv, mpm_mask, W, anchor):
    """cat(cls, embed(pv) * (1-m) + mask_token * m) -> [B, 54, H] bf16 (SPMM_models.py:82-88)."""
    return _PvTokens.apply(anchor, pv, mpm_mask, W)


# --------------------------------------------------------------------------------------------- projection heads
class _ProjF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W):
        ctx.save_for_backward(x)
        ctx.W = W
        M, Kd = x.shape
        return K.gemm(x, W.w, M, W.w.shape[0], Kd, bias=W.b, out_f32=True)

    @staticmethod
    def backward(ctx, dz):
        (x,) = ctx.saved_tensors
        W = ctx.W
        M, Kd = x.shape
        N = W.w.shape[0]
        dzb = dz.to(BF16).contiguous()
        if W.g_b is not None:
            K.colsum(dzb, W.g_b)
        _wgrad(dzb, x, W.g_w, N, Kd, M)
        return K.gemm(dzb, W.w, M, Kd, N, b_mn=True), None


def proj_f32(x, W):
    """fp32 = x . W^T + b for the 768->256 projection heads (SPMM_models.py:92,95); x may be a strided CLS view."""
    return _ProjF32.apply(x, W)


# --------------------------------------------------------------------------------------------- ITC head
class _Itc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_prop, z_text, temp, z_prop_m, z_text_m, pq, tq, alpha, side):
        out = K.itc(z_prop.contiguous(), z_text.contiguous(), z_prop_m.contiguous(), z_text_m.contiguous(), pq, tq,
                    temp, alpha)
        side.update(out)
        ctx.save_for_backward(out["dz_prop"], out["dz_text"], out["dtemp"])
        return out["loss"].clone()

    @staticmethod
    def backward(ctx, g):
        dzp, dzt, dtemp = ctx.saved_tensors
        return dzp * g, dzt * g, dtemp * g, None, None, None, None, None, None


def itc(z_prop, z_text, temp, z_prop_m, z_text_m, prop_queue, text_queue, alpha, side):
    """loss_ita of SPMM_models.py:102-131 from raw projections; `side` receives sims / momentum feats / nan flag."""
    return _Itc.apply(z_prop, z_text, temp, z_prop_m, z_text_m, prop_queue, text_queue, alpha, side)


def sample_negatives(side, seed, step):
    return K.sample_negatives(side["sim_i2t"], side["sim_t2i"], seed, step)


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, idx):
        ctx.save_for_backward(idx)
        ctx.shape = src.shape
        return K.gather_rows(src.contiguous(), idx, idx.numel())

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        return K.segment_sum_rows(dout.contiguous(), idx, ctx.shape[0]), None


def gather_rows(src, idx):
    """src[idx] along dim 0 (hard-negative gather, SPMM_models.py:165-178), idx int32 on device."""
    return _GatherRows.apply(src, idx)


# --------------------------------------------------------------------------------------------- loss heads
def lm_logits(h, W, V, ld):
    """BertLMPredictionHead (xbert.py:693-696) without autograd: teacher logits [tokens, ld] bf16."""
    M, H = h.shape
    a = K.gemm(h, W.wt, M, H, H, bias=W.bt, gelu=True)
    t, _, _ = K.layernorm_fwd(a, W.ln_g, W.ln_b, W.eps, save_stats=False)
    logits = torch.empty(M, ld, device=h.device, dtype=BF16)
    K.gemm(t, W.wdec, M, V, H, bias=W.bdec, out=logits)
    return logits


class _LmHeadLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, logits_m, ids, W, alpha, V, valid_len):
        M, H = h.shape
        ld = logits_m.shape[1]
        pre = torch.empty(M, H, device=h.device, dtype=BF16)
        a = K.gemm(h, W.wt, M, H, H, bias=W.bt, gelu=True, pre_act_out=pre)
        t, mean, rstd = K.layernorm_fwd(a, W.ln_g, W.ln_b, W.eps)
        logits = torch.empty(M, ld, device=h.device, dtype=BF16)
        K.gemm(t, W.wdec, M, V, H, bias=W.bdec, out=logits)
        loss, dlogits = K.lm_loss(logits, logits_m, ids, V, alpha, valid_len)
        ctx.save_for_backward(h, pre, a, t, mean, rstd, dlogits)
        ctx.W, ctx.V = W, V
        return loss

    @staticmethod
    def backward(ctx, g):
        h, pre, a, t, mean, rstd, dlogits = ctx.saved_tensors
        W, V = ctx.W, ctx.V
        M, H = h.shape
        dl = dlogits * g.to(BF16)
        V8 = (V + 7) // 8 * 8
        K.colsum(dl[:, :V8], W.g_bdec)                   # padded columns are zero (arena padding absorbs them)
        _wgrad(dl[:, :V], t, W.g_wdec, V, H, M)          # tied decoder: accumulates into the word-embedding grad
        dt = K.gemm(dl[:, :V], W.wdec, M, H, V, b_mn=True)
        da, _ = K.layernorm_bwd(dt, a, mean, rstd, W.ln_g, W.g_ln_g, W.g_ln_b)
        dpre = K.dgelu(da, pre)
        K.colsum(dpre, W.g_bt)
        _wgrad(dpre, h, W.g_wt, H, H, M)
        return K.gemm(dpre, W.wt, M, H, H, b_mn=True), None, None, None, None, None, None


def lm_head_loss(h, logits_m, ids, W, alpha, V, valid_len=None):
    """LM head + (1-alpha) CE + alpha distillation (SPMM_models.py:224-238) on h [B*L, H].  `alpha` may be a device
    scalar; `valid_len` (device int32) marks bucket padding beyond the batch's own width."""
    return _LmHeadLoss.apply(h, logits_m, ids, W, alpha, V, valid_len)


class _ItmLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vl, W, n_pos):
        dw, db = torch.zeros_like(W.w), torch.zeros_like(W.b)
        loss, dx = K.itm_loss(vl.contiguous(), W.w, W.b, n_pos, dw, db)
        ctx.save_for_backward(dx, dw, db)
        ctx.W = W
        return loss

    @staticmethod
    def backward(ctx, g):
        dx, dw, db = ctx.saved_tensors
        W = ctx.W
        W.g_w.add_(dw * g)
        W.g_b.add_(db * g)
        return dx * g.to(BF16), None, None


def itm_loss(vl, W, n_pos):
    """CE(itm_head(vl), [1]*n_pos + [0]*rest) (SPMM_models.py:201-206); vl [3B, 2H] bf16."""
    return _ItmLoss.apply(vl, W, n_pos)


class _MtrHeadLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pv, mpm, W):
        M, H = x.shape
        pre = torch.empty(M, H, device=x.device, dtype=BF16)
        a = K.gemm(x, W.w0, M, H, H, bias=W.b0, gelu=True, pre_act_out=pre)
        t, mean, rstd = K.layernorm_fwd(a, W.ln_g, W.ln_b, W.eps)
        dw3, db3 = torch.zeros_like(W.w3), torch.zeros_like(W.b3)
        loss, dt = K.mpm_loss(t, W.w3, W.b3, pv, mpm, dw3, db3)
        ctx.save_for_backward(x, pre, a, mean, rstd, dt, dw3, db3)
        ctx.W = W
        return loss

    @staticmethod
    def backward(ctx, g):
        x, pre, a, mean, rstd, dt, dw3, db3 = ctx.saved_tensors
        W = ctx.W
        M, H = x.shape
        W.g_w3.add_(dw3 * g)
        W.g_b3.add_(db3 * g)
        da, _ = K.layernorm_bwd(dt * g.to(BF16), a, mean, rstd, W.ln_g, W.g_ln_g, W.g_ln_b)
        dpre = K.dgelu(da, pre)
        K.colsum(dpre, W.g_b0)
        _wgrad(dpre, x, W.g_w0, H, H, M)
        return K.gemm(dpre, W.w0, M, H, H, b_mn=True), None, None, None


def mtr_head_loss(x, pv, mpm_mask, W):
    """5 * MSE(property_mtr_head(x)[:, :-1][keep], pv[keep]) (SPMM_models.py:251-256); x [B*54, H] bf16."""
    return _MtrHeadLoss.apply(x, pv, mpm_mask, W)
