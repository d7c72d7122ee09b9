"""The C-ABI library loads and exports every symbol include/vlk.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "vlk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vlk_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for must in ("vlk_gemm_bf16", "vlk_layernorm_fwd", "vlk_layernorm_bwd", "vlk_attn_fwd", "vlk_attn_bwd",
                 "vlk_pool33_l2norm", "vlk_embed_concat_fwd", "vlk_softmax_ce_rows", "vlk_adamw_step",
                 "vlk_grad_sumsq", "vlk_version", "vlk_last_error_string"):
        assert must in syms


def test_library_exports_every_header_symbol():
    from gpt2_vision_language_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert set(_lib.SIGNATURES) == set(header_symbols()), set(_lib.SIGNATURES) ^ set(header_symbols())
    bound = _lib.load()
    assert bound.vlk_version() == 100


def test_product_path_fails_loudly_without_gpu():
    import torch
    from gpt2_vision_language_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    from gpt2_vision_language_b200.caption import pool_clip_197_to_33_avg_with_cls
    with pytest.raises(RuntimeError):
        pool_clip_197_to_33_avg_with_cls(torch.zeros(1, 257, 64))
