"""ctypes binding of libvlk.so — the C ABI declared in include/vlk.h.

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError is raised
(the reference convention is plain Python exceptions, e.g. the assert at source/gpt2/train_gpt2.py:113).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvlk.so")

c_void_p, c_int, c_float, c_ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong

# name -> argtypes; every function returns int unless listed in _RESTYPES.
SIGNATURES = {
    "vlk_version": [],
    "vlk_last_error_string": [],
    "vlk_num_sms": [],
    "vlk_launch_count": [],
    "vlk_gemm_bf16": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                      c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_float, c_int,
                      c_int, c_void_p],
    "vlk_gemm_bf16_tile": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                           c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "vlk_gemm_bf16_splitk": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                             c_int, c_float, c_int, c_int, c_void_p],
    "vlk_row_stats": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p],
    "vlk_gemm_bf16_lnfold": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_int, c_void_p],
    "vlk_gemm_bf16_lnfold_sums": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_float, c_void_p, c_int, c_void_p],
    "vlk_gemm_bf16_stats": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                            c_int, c_void_p, c_void_p],
    "vlk_colsum_bf16": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "vlk_colsum_bf16_acc": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "vlk_transpose_bf16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "vlk_layernorm_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float,
                          c_void_p],
    "vlk_layernorm_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                          c_int, c_int, c_int, c_void_p],
    "vlk_layernorm_bwd_acc": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                              c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "vlk_sum_copies": [c_void_p, c_int, c_ll, c_void_p, c_int, c_int, c_int, c_void_p],
    "vlk_attn_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                     c_ll, c_int, c_ll, c_int, c_ll, c_int, c_ll, c_int, c_int, c_float,
                     c_float, c_void_p, ctypes.c_uint, c_void_p],
    "vlk_attn_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                     c_int, c_int, c_int, c_int,
                     c_ll, c_int, c_ll, c_int, c_ll, c_int, c_ll, c_int,
                     c_ll, c_int, c_ll, c_int, c_ll, c_int, c_int, c_float, c_void_p,
                     c_float, c_void_p, ctypes.c_uint, c_void_p],
    "vlk_pool33_l2norm": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "vlk_embed_concat_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                             c_int, c_void_p],
    "vlk_scalar_pack_digits": [c_void_p, c_void_p, c_int, c_void_p],
    "vlk_scalar_unpack_digits": [c_void_p, c_void_p, c_int, c_void_p],
    "vlk_embed_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "vlk_embed_bwd_acc": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                          c_void_p],
    "vlk_softmax_ce_rows": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "vlk_lmhead_ce_workspace_bytes": [c_int, c_int, c_int, c_int, c_int, c_int],
    "vlk_lmhead_ce_fwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                          c_int, c_void_p, c_ll, c_void_p],
    "vlk_lmhead_ce_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                          c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p],
    "vlk_ce_count": [c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "vlk_ce_finalize": [c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "vlk_grad_sumsq_workspace_floats": [c_int, c_ll],
    "vlk_grad_sumsq": [c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p, c_void_p],
    "vlk_adamw_step": [c_void_p, c_int, c_ll, c_int, c_void_p, c_float, c_void_p, c_float, c_float, c_float,
                       c_void_p, c_void_p],
    "vlk_im2col_patch14": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "vlk_clip_assemble": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p],
    "vlk_add_bf16": [c_void_p, c_void_p, c_void_p, c_ll, c_void_p],
    "vlk_cast_f32_to_bf16": [c_void_p, c_void_p, c_ll, c_void_p],
    "vlk_cast_bf16_to_f32": [c_void_p, c_void_p, c_ll, c_void_p],
    "vlk_dropout_add_bf16": [c_void_p, c_void_p, c_void_p, c_ll, c_float, c_void_p, ctypes.c_uint, c_void_p],
    "vlk_gate_grad": [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_void_p],
    "vlk_argmax_rows": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
}
_RESTYPES = {"vlk_last_error_string": ctypes.c_char_p, "vlk_launch_count": c_ll, "vlk_lmhead_ce_workspace_bytes": c_ll,
             "vlk_grad_sumsq_workspace_floats": c_ll}


class TensorDesc(ctypes.Structure):
    """Mirror of `vlk_tensor_desc` (include/vlk.h)."""
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("numel", c_ll), ("weight_decay", c_float), ("pad_", c_int)]


_lib = None


def load():
    """Load libvlk.so and bind every symbol of include/vlk.h. Raises RuntimeError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"libvlk.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


def check(rc, what=""):
    """Turn a non-zero libvlk return code into RuntimeError carrying vlk_last_error_string()."""
    if rc != 0:
        msg = load().vlk_last_error_string()
        raise RuntimeError(f"libvlk {what} failed (rc={rc}): {msg.decode() if msg else ''}")


def launch_count():
    return int(load().vlk_launch_count())
