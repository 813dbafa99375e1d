"""ctypes binding of libspmm_b200.so (C ABI declared in include/spmm_b200.h).

There is no CPU fallback: if the library is missing, or a call returns non-zero, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspmm_b200.so")
_lib = None

vp, i32, i64, f32, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_ulonglong


class GemmEpilogue(C.Structure):
    _fields_ = [("bias", vp), ("residual", vp), ("ld_residual", i32), ("pre_act", vp), ("ld_pre_act", i32),
                ("dgelu_pre_act", vp), ("ld_dgelu_pre_act", i32), ("flags", i32), ("alpha", f32),
                ("dropout_p", f32), ("dropout_seed", u64), ("colsum", vp)]


GEMM_OUT_F32, GEMM_ACCUMULATE, GEMM_GELU, GEMM_DGELU, GEMM_DGELU_STORED = 1, 2, 4, 8, 16

SIGNATURES = {
    "spmm_version": (i32, []),
    "spmm_set_rng_salt_ptr": (i32, [vp]),
    "spmm_gemm_bf16": (i32, [vp, i32, i32, vp, i32, i32, vp, i32, i32, i32, i32, C.POINTER(GemmEpilogue), vp]),
    "spmm_gemm_debug_config": (i32, [i32, i32, i32, i32]),
    "spmm_gemm_debug_trace": (i32, [vp]),
    "spmm_gemm_debug_trace_ring": (i32, [vp, C.c_long]),
    "spmm_attn_debug_trace": (i32, [vp]),
    "spmm_attn_fwd": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, vp, i32, i32, f32, f32, u64, vp, i32,
                            vp]),
    "spmm_attn_bwd": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp, vp, i32, vp, i32, vp, i32, i32, i32, i32,
                            i32, vp, i32, f32, f32, u64, vp, vp, vp, vp, i32, vp]),
    "spmm_layernorm_fwd": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, f32, f32, u64, vp]),
    "spmm_layernorm_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, f32, u64, f32, u64, vp, vp]),
    "spmm_embed_text_fwd": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "spmm_embed_text_bwd": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "spmm_pv_tokens_fwd": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "spmm_pv_tokens_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp]),
    "spmm_embed_inputs_fwd": (i32, [vp, vp, vp, vp, i32, i32, i32, vp]),
    "spmm_embed_inputs_bwd": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "spmm_colsum_bf16": (i32, [vp, i32, vp, i32, i32, vp]),
    "spmm_cast_f32_to_bf16": (i32, [vp, vp, i64, vp]),
    "spmm_add_bf16": (i32, [vp, vp, i64, vp]),
    "spmm_dgelu_bf16": (i32, [vp, vp, vp, i64, vp]),
    "spmm_gather_rows_bf16": (i32, [vp, vp, vp, i32, i64, vp]),
    "spmm_segment_sum_rows_bf16": (i32, [vp, i32, vp, vp, i32, i64, vp]),
    "spmm_itc_fwd_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, f32, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                               i64, vp]),
    "spmm_itc_debug_trace": (i32, [vp]),
    "spmm_itc_workspace_bytes": (i64, [i32, i32, i32]),
    "spmm_sample_negatives": (i32, [vp, vp, i32, u64, u64, vp, vp, vp]),
    "spmm_enqueue": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp]),
    "spmm_lm_loss_fwd_bwd": (i32, [vp, vp, i32, vp, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp]),
    "spmm_itm_loss_fwd_bwd": (i32, [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
    "spmm_mpm_loss_fwd_bwd": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]),
    "spmm_decode_embed": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "spmm_decode_attn_self": (i32, [vp, i32, vp, vp, i32, vp, vp, vp, vp, i32, vp, vp, i32, i32, i32, f32, vp]),
    "spmm_decode_attn_cross": (i32, [vp, i32, vp, vp, i32, i32, i32, vp, vp, i32, i32, i32, f32, vp]),
    "spmm_beam_step": (i32, [vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "spmm_wordpiece_create": (vp, [C.POINTER(C.c_char_p), i32, i32, i32]),
    "spmm_wordpiece_destroy": (None, [vp]),
    "spmm_wordpiece_encode_batch": (i32, [vp, C.POINTER(C.c_char_p), i32, i32, i32, i32, i32, vp, vp, i32]),
    "spmm_ema_multi": (i32, [vp, vp, vp, vp, i64, f32, f32, vp]),
    "spmm_grad_sumsq": (i32, [vp, i64, vp, vp, vp]),
    "spmm_adam_tick": (i32, [vp, vp, vp, f32, f32, vp, vp]),
    "spmm_adamw_step": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, vp, f32, f32, vp, vp, vp]),
}


class SpmmKernelError(RuntimeError):
    pass


def lib():
    """Loads the C-ABI library.  Fails loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SpmmKernelError("%s is missing: run `python -m spmm_b200.build` (or __graft_entry__.build()). "
                                  "spmm_b200 has no CPU / PyTorch fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if os.environ.get("SPMM_GEMM_DEBUG_FLAGS"):          # debug / A-B runs: e.g. 0x40000 disables the 2-CTA GEMM kernel
            l.spmm_gemm_debug_config(0, 0, int(os.environ["SPMM_GEMM_DEBUG_FLAGS"], 0), 0)
        _lib = l
    return _lib


# kernels launched per C call (for bench.py's `gpu_launches` claim)
KERNELS_PER_CALL = {"spmm_itc_fwd_bwd": 6, "spmm_lm_loss_fwd_bwd": 3, "spmm_mpm_loss_fwd_bwd": 3, "spmm_enqueue": 2,
                    "spmm_itm_loss_fwd_bwd": 2}
_launches = 0


def reset_launch_count():
    global _launches
    _launches = 0


def launch_count():
    return _launches


# Measurement aid (tools/skip_table.sh): SPMM_DEBUG_SKIP="spmm_layernorm_fwd,spmm_colsum_bf16" makes the named entry points
# no-ops, so the step time WITHOUT a kernel family can be read from `bench.py --profile` (results are garbage, timing only).
_SKIP = frozenset(x for x in os.environ.get("SPMM_DEBUG_SKIP", "").split(",") if x)


def call(name, *args):
    global _launches
    if _SKIP and name in _SKIP:
        return 0
    _launches += KERNELS_PER_CALL.get(name, 1)
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise SpmmKernelError("%s returned %d (%s)" % (name, rc, "argument/setup error" if rc < 0 else "cudaError"))
    return rc
