"""Thin tensor-level wrappers over the C ABI (include/spmm_b200.h): pointer extraction + stream only.

Nothing here computes with PyTorch; every function launches hand-written sm_100a kernels from
libspmm_b200.so on torch's current CUDA stream and raises if the library is unavailable.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import GEMM_ACCUMULATE, GEMM_DGELU, GEMM_DGELU_STORED, GEMM_GELU, GEMM_OUT_F32, GemmEpilogue, call

BF16 = torch.bfloat16


def _p(t):
    return None if t is None else t.data_ptr()


def _st():
    return torch.cuda.current_stream().cuda_stream


def _chk_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.SpmmKernelError("spmm_b200 kernels need CUDA tensors (no CPU fallback)")


def gemm(a, b, M, N, K, *, a_mn=False, b_mn=False, out=None, out_f32=False, accumulate=False, bias=None,
         residual=None, pre_act_out=None, dgelu_pre=None, gelu=False, dropout_p=0.0, seed=0, ldc=None,
         dgelu_stored=False, colsum_out=None):
    """C[M,N] (+)= epi(A.B^T).  a: [M,K] (K-major) or [K,M] (a_mn); b: [N,K] or [K,N] (b_mn); row stride = ld.
    `colsum_out` [N] fp32 += column sums of the bf16 result (bias gradient fused into the producing GEMM)."""
    _chk_cuda(a, b)
    assert a.dtype == BF16 and b.dtype == BF16 and a.stride(-1) == 1 and b.stride(-1) == 1
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if out_f32 else BF16)
    assert out.stride(-1) == 1
    flags = (GEMM_OUT_F32 if out_f32 else 0) | (GEMM_ACCUMULATE if accumulate else 0) | (GEMM_GELU if gelu else 0) | \
            (GEMM_DGELU if dgelu_pre is not None else 0) | (GEMM_DGELU_STORED if dgelu_stored else 0)
    epi = GemmEpilogue(_p(bias), _p(residual), residual.stride(0) if residual is not None else 0,
                       _p(pre_act_out), pre_act_out.stride(0) if pre_act_out is not None else 0,
                       _p(dgelu_pre), dgelu_pre.stride(0) if dgelu_pre is not None else 0,
                       flags, 1.0, float(dropout_p), int(seed), _p(colsum_out))
    call("spmm_gemm_bf16", a.data_ptr(), a.stride(0), int(a_mn), b.data_ptr(), b.stride(0), int(b_mn),
         out.data_ptr(), out.stride(0) if ldc is None else ldc, M, N, K, C.byref(epi), _st())
    return out


def attn_fwd(q, k, v, out, lse, batch, heads, Tq, Tk, kv_len, causal, scale, dropout_p=0.0, seed=0, kv_bstride=None,
             kv_index=None, kv_batches=0):
    """`kv_index` (int32 [batch]) + `kv_batches`: batch element b reads K/V of element kv_index[b] of a kv_batches-long buffer."""
    call("spmm_attn_fwd", q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
         out.data_ptr(), out.stride(0), _p(lse), batch, heads, Tq, Tk, _p(kv_len), int(causal),
         Tk if kv_bstride is None else kv_bstride, float(scale), float(dropout_p), int(seed), _p(kv_index), int(kv_batches),
         _st())
    return out


def attn_bwd(do, q, k, v, o, lse, dq, dk, dv, batch, heads, Tq, Tk, kv_len, causal, scale, dropout_p=0.0, seed=0,
             dbias=None, kv_index=None, kv_batches=0):
    """`dbias` = (dbq, dbk, dbv) fp32 [heads*64] views: += column sums of dq / dk / dv (projection bias gradients)."""
    dbq, dbk, dbv = dbias if dbias is not None else (None, None, None)
    call("spmm_attn_bwd", do.data_ptr(), do.stride(0), q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0),
         v.data_ptr(), v.stride(0), o.data_ptr(), o.stride(0), lse.data_ptr(), dq.data_ptr(), dq.stride(0),
         dk.data_ptr(), dk.stride(0), dv.data_ptr(), dv.stride(0), batch, heads, Tq, Tk, _p(kv_len), int(causal),
         float(scale), float(dropout_p), int(seed), _p(dbq), _p(dbk), _p(dbv), _p(kv_index), int(kv_batches), _st())


def layernorm_fwd(x, gamma, beta, eps, save_stats=True, dropout_p=0.0, seed=0):
    rows, H = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if save_stats else None
    call("spmm_layernorm_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), _p(mean), _p(rstd), rows,
         H, float(eps), float(dropout_p), int(seed), _st())
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, dbias=None, want_branch=False, out_dropout_p=0.0,
                  out_seed=0, branch_dropout_p=0.0, branch_seed=0):
    rows, H = x.shape
    dx = torch.empty_like(x)
    dxb = torch.empty_like(x) if (want_branch and branch_dropout_p > 0) else None
    call("spmm_layernorm_bwd", dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
         dx.data_ptr(), _p(dgamma), _p(dbeta), _p(dxb), _p(dbias), rows, H, float(out_dropout_p), int(out_seed),
         float(branch_dropout_p), int(branch_seed), _ln_workspace(x.device).data_ptr(), _st())
    return dx, (dxb if dxb is not None else dx)


_LN_WS = {}


def _ln_workspace(device):
    """Persistent zeroed scratch for the LayerNorm-backward column reduction (the kernel re-zeroes it)."""
    ws = _LN_WS.get(device)
    if ws is None:
        ws = _LN_WS[device] = torch.zeros(8 * 3 * 1024 + 8, device=device, dtype=torch.float32)
    return ws


def ema(p, p_m, p_bf16, p_m_bf16, momentum):
    n = p.numel()
    call("spmm_ema_multi", p.data_ptr(), p_m.data_ptr(), _p(p_bf16), _p(p_m_bf16), n, float(momentum),
         float(1.0 - momentum), _st())


def cast_bf16(src, dst):
    call("spmm_cast_f32_to_bf16", src.data_ptr(), dst.data_ptr(), src.numel(), _st())


def colsum(x, out):
    rows, cols = x.shape
    call("spmm_colsum_bf16", x.data_ptr(), x.stride(0), out.data_ptr(), rows, cols, _st())


def add_(dst, src):
    assert dst.is_contiguous() and src.is_contiguous() and dst.numel() == src.numel()
    call("spmm_add_bf16", dst.data_ptr(), src.data_ptr(), dst.numel(), _st())
    return dst


def dgelu(d_act, pre):
    out = torch.empty_like(pre)
    call("spmm_dgelu_bf16", d_act.data_ptr(), pre.data_ptr(), out.data_ptr(), pre.numel(), _st())
    return out


def gather_rows(src, idx, n_idx):
    row = src[0].numel()
    out = torch.empty((n_idx,) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
    call("spmm_gather_rows_bf16", src.data_ptr(), idx.data_ptr(), out.data_ptr(), n_idx, row, _st())
    return out


def segment_sum_rows(src, idx, n_dst):
    """out[t] = sum of src[r] over idx[r] == t; src [n_idx, ...] bf16, idx int32 on the device."""
    out = torch.empty((n_dst,) + tuple(src.shape[1:]), device=src.device, dtype=src.dtype)
    call("spmm_segment_sum_rows_bf16", out.data_ptr(), n_dst, idx.data_ptr(), src.data_ptr(), src.shape[0], src[0].numel(), _st())
    return out


def _scalar_or_dev(alpha):
    """(host float, device pointer or None): a CUDA tensor is read by the kernel itself (graph-replay safe)."""
    if torch.is_tensor(alpha):
        assert alpha.is_cuda and alpha.dtype == torch.float32 and alpha.numel() == 1
        return 0.0, alpha.data_ptr()
    return float(alpha), None


def itc(z_prop, z_text, z_prop_m, z_text_m, prop_queue, text_queue, temp, alpha):
    """`alpha`: python float, or a 1-element fp32 CUDA tensor read on the device."""
    B, E = z_prop.shape
    Q = prop_queue.shape[0]
    dev = z_prop.device
    f32 = dict(device=dev, dtype=torch.float32)
    ws_bytes = _lib.lib().spmm_itc_workspace_bytes(B, E, Q)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    out = dict(loss=torch.empty((), **f32), dz_prop=torch.empty(B, E, **f32), dz_text=torch.empty(B, E, **f32),
               dtemp=torch.empty((), **f32), sim_i2t=torch.empty(B, B, **f32), sim_t2i=torch.empty(B, B, **f32),
               feat_prop_m=torch.empty(B, E, **f32), feat_text_m=torch.empty(B, E, **f32),
               nan_flag=torch.empty((), **f32))
    call("spmm_itc_fwd_bwd", z_prop.data_ptr(), z_text.data_ptr(), z_prop_m.data_ptr(), z_text_m.data_ptr(),
         prop_queue.data_ptr(), text_queue.data_ptr(), temp.data_ptr(), *_scalar_or_dev(alpha), B, E, Q, out["loss"].data_ptr(),
         out["dz_prop"].data_ptr(), out["dz_text"].data_ptr(), out["dtemp"].data_ptr(), out["sim_i2t"].data_ptr(),
         out["sim_t2i"].data_ptr(), out["feat_prop_m"].data_ptr(), out["feat_text_m"].data_ptr(),
         out["nan_flag"].data_ptr(), ws.data_ptr(), ws_bytes, _st())
    return out


def sample_negatives(sim_i2t, sim_t2i, seed, step):
    B = sim_i2t.shape[0]
    t2i = torch.empty(B, device=sim_i2t.device, dtype=torch.int32)
    i2t = torch.empty(B, device=sim_i2t.device, dtype=torch.int32)
    call("spmm_sample_negatives", sim_i2t.data_ptr(), sim_t2i.data_ptr(), B, int(seed), int(step), t2i.data_ptr(),
         i2t.data_ptr(), _st())
    return t2i, i2t


def enqueue(prop_queue, text_queue, prop_feats, text_feats, queue_ptr, skip_flag=None):
    Q, E = prop_queue.shape
    call("spmm_enqueue", prop_queue.data_ptr(), text_queue.data_ptr(), prop_feats.data_ptr(), text_feats.data_ptr(),
         queue_ptr.data_ptr(), prop_feats.shape[0], E, Q, _p(skip_flag), _st())


def lm_loss(logits, teacher, ids, V, alpha, valid_len=None):
    """`alpha` float or 1-element fp32 CUDA tensor; `valid_len` optional int32 CUDA scalar: the batch's own padded width
    when `ids` sits in a longer length bucket (positions beyond it are ignored)."""
    B, L = ids.shape
    ld = logits.stride(0)
    loss = torch.empty((), device=logits.device, dtype=torch.float32)
    dlogits = torch.empty_like(logits)
    ws = torch.empty(4, device=logits.device, dtype=torch.float32)
    if valid_len is not None:
        assert valid_len.is_cuda and valid_len.dtype == torch.int32 and valid_len.numel() == 1
    call("spmm_lm_loss_fwd_bwd", logits.data_ptr(), teacher.data_ptr(), ld, ids.data_ptr(), B, L, V, *_scalar_or_dev(alpha),
         _p(valid_len), loss.data_ptr(), dlogits.data_ptr(), ws.data_ptr(), _st())
    return loss, dlogits


def itm_loss(x, w, b, n_pos, dw, db):
    n_rows, D = x.shape
    loss = torch.empty((), device=x.device, dtype=torch.float32)
    dx = torch.empty_like(x)
    ws = torch.empty(3 * n_rows, device=x.device, dtype=torch.float32)
    call("spmm_itm_loss_fwd_bwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), n_rows, n_pos, D, loss.data_ptr(),
         dx.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), _st())
    return loss, dx


def mpm_loss(t, w, b, pv, mpm_mask, dw, db):
    batch, n_prop = pv.shape
    H = t.shape[-1]
    loss = torch.empty((), device=t.device, dtype=torch.float32)
    dt = torch.empty_like(t)
    ws = torch.empty(4, device=t.device, dtype=torch.float32)
    call("spmm_mpm_loss_fwd_bwd", t.data_ptr(), w.data_ptr(), b.data_ptr(), pv.data_ptr(), mpm_mask.data_ptr(), batch,
         n_prop, H, loss.data_ptr(), dt.data_ptr(), dw.data_ptr(), db.data_ptr(), ws.data_ptr(), _st())
    return loss, dt


_SUMSQ_WS = {}


def grad_sumsq(g, out):
    """out[0] = sum g^2 with a fixed summation order (replicas get the bit-identical clip coefficient)."""
    ws = _SUMSQ_WS.get(g.device)
    if ws is None:
        ws = _SUMSQ_WS[g.device] = torch.zeros(1024, device=g.device, dtype=torch.float32)
    call("spmm_grad_sumsq", g.data_ptr(), g.numel(), out.data_ptr(), ws.data_ptr(), _st())


def adam_tick(t_dev, lr_dev, hyper_dev, beta1, beta2, skip_flag=None):
    call("spmm_adam_tick", t_dev.data_ptr(), lr_dev.data_ptr(), hyper_dev.data_ptr(), float(beta1), float(beta2),
         _p(skip_flag), _st())


def adamw(p, g, m1, m2, lr, beta1, beta2, eps, wd, step, sumsq=None, max_norm=0.0, grad_scale=1.0, skip_flag=None,
          hyper_dev=None):
    call("spmm_adamw_step", p.data_ptr(), g.data_ptr(), m1.data_ptr(), m2.data_ptr(), p.numel(), float(lr), float(beta1),
         float(beta2), float(eps), float(wd), int(step), _p(sumsq), float(max_norm), float(grad_scale), _p(skip_flag),
         _p(hyper_dev), _st())


def embed_text_fwd(ids, word, pos, type0, H):
    B, T = ids.shape
    x = torch.empty(B * T, H, device=ids.device, dtype=BF16)
    call("spmm_embed_text_fwd", ids.data_ptr(), word.data_ptr(), pos.data_ptr(), type0.data_ptr(), x.data_ptr(), B * T, T,
         H, _st())
    return x


def embed_text_bwd(dx, ids, dword, dpos, dtype0, pad_id):
    B, T = ids.shape
    call("spmm_embed_text_bwd", dx.data_ptr(), ids.data_ptr(), dword.data_ptr(), dpos.data_ptr(), dtype0.data_ptr(),
         B * T, T, dx.shape[-1], int(pad_id), _st())


def embed_inputs_fwd(inputs, pos, type0):
    B, T, H = inputs.shape
    x = torch.empty(B * T, H, device=inputs.device, dtype=BF16)
    call("spmm_embed_inputs_fwd", inputs.data_ptr(), pos.data_ptr(), type0.data_ptr(), x.data_ptr(), B * T, T, H, _st())
    return x


def embed_inputs_bwd(dx, T, dpos, dtype0):
    rows, H = dx.shape
    call("spmm_embed_inputs_bwd", dx.data_ptr(), dpos.data_ptr(), dtype0.data_ptr(), rows, T, H, _st())


def pv_tokens_fwd(pv, mpm_mask, w, b, cls_tok, mask_tok):
    B, n_prop = pv.shape
    H = w.numel()
    out = torch.empty(B, n_prop + 1, H, device=pv.device, dtype=BF16)
    call("spmm_pv_tokens_fwd", pv.data_ptr(), mpm_mask.data_ptr(), w.data_ptr(), b.data_ptr(), cls_tok.data_ptr(),
         mask_tok.data_ptr(), out.data_ptr(), B, n_prop, H, _st())
    return out


def pv_tokens_bwd(dprop, pv, mpm_mask, dw, db, dcls, dmask):
    B, n_prop = pv.shape
    call("spmm_pv_tokens_bwd", dprop.data_ptr(), pv.data_ptr(), mpm_mask.data_ptr(), dw.data_ptr(), db.data_ptr(),
         dcls.data_ptr(), dmask.data_ptr(), B, n_prop, dprop.shape[-1], _st())


def set_rng_salt(dev_tensor):
    """Registers the device int64 scalar the kernels add to every dropout / sampler seed (None clears it)."""
    call("spmm_set_rng_salt_ptr", _p(dev_tensor))


# ---------------------------------------------------------------------------------------------- single-token decode
def decode_embed(ids, t_dev, word, pos, type0, H):
    x = torch.empty(ids.numel(), H, device=ids.device, dtype=BF16)
    call("spmm_decode_embed", ids.data_ptr(), t_dev.data_ptr(), word.data_ptr(), pos.data_ptr(), type0.data_ptr(),
         x.data_ptr(), ids.numel(), H, _st())
    return x


def decode_attn_self(q, k_new, v_new, cache_k, cache_v, anc, tokens, t_dev, out, heads, scale):
    rows, tmax = tokens.shape
    assert k_new.stride(0) == v_new.stride(0) and cache_k.is_contiguous() and cache_v.is_contiguous()
    call("spmm_decode_attn_self", q.data_ptr(), q.stride(0), k_new.data_ptr(), v_new.data_ptr(), k_new.stride(0),
         cache_k.data_ptr(), cache_v.data_ptr(), anc.data_ptr(), tokens.data_ptr(), tmax, t_dev.data_ptr(), out.data_ptr(),
         out.stride(0), rows, heads, float(scale), _st())
    return out


def decode_attn_cross(q, k, v, Tk, group, out, heads, scale, kv_len=None):
    assert k.stride(0) == v.stride(0)
    call("spmm_decode_attn_cross", q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), Tk, group,
         _p(kv_len), out.data_ptr(), out.stride(0), q.shape[0], heads, float(scale), _st())
    return out


def beam_step(logits, V, st, cls_id, sep_id):
    """`st`: the device-side beam state (spmm_b200/generate.py:BeamState)."""
    call("spmm_beam_step", logits.data_ptr(), logits.stride(0), V, st.k, st.tmax, st.n_mol, st.fin_cap, cls_id, sep_id,
         st.t_dev.data_ptr(), st.scores.data_ptr(), st.tokens.data_ptr(), st.anc.data_ptr(), st.next_ids.data_ptr(),
         st.fin_scores.data_ptr(), st.fin_tokens.data_ptr(), st.fin_len.data_ptr(), st.fin_count.data_ptr(),
         st.done.data_ptr(), _p(st.trace_logp), _p(st.trace_tok), st.ticket.data_ptr(), _st())
