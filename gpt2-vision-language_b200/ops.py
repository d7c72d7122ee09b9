"""Tensor-level wrappers over the libvlk C ABI (raw device pointers in, nothing allocated natively).

Everything here runs on the CURRENT torch CUDA stream, is asynchronous and CUDA-graph capturable.
PyTorch owns every buffer; the native side borrows pointers for the duration of a call.
"""
import torch

from . import _lib
from ._lib import check

ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_QUICK_GELU = 0, 1, 2, 3
_ACT = {None: 0, "none": 0, "gelu_tanh": 1, "gelu_erf": 2, "quick_gelu": 3}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("libvlk ops need CUDA tensors: there is no CPU path for the B200 kernels")


def _bf16c(t):
    """bf16, 2-D view requirements are the caller's; ensure dtype + inner contiguity."""
    if t.dtype != torch.bfloat16:
        t = t.to(torch.bfloat16)
    if t.stride(-1) != 1:
        t = t.contiguous()
    return t


def gemm(a, b, *, trans_a=False, trans_b=False, bias=None, residual=None, aux_in=None, aux_out=False,
         scale=None, act=None, dact=False, alpha=1.0, out=None, out_fp32=False):
    """D = epi(alpha * op(a) @ op(b)) with the epilogue of vlk_gemm_bf16 (include/vlk.h).

    a: [M,K] (or [K,M] if trans_a); b: [N,K] — nn.Linear weight layout — (or [K,N] if trans_b).
    Returns D, or (D, aux) when aux_out=True (aux = pre-activation, bf16).
    """
    _need_cuda(a, b)
    lib = _lib.load()
    assert a.dim() == 2 and b.dim() == 2
    a = _bf16c(a)
    b = _bf16c(b)
    M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if trans_b else (b.shape[0], b.shape[1])
    if K != Kb:
        raise RuntimeError(f"gemm: contraction mismatch {K} vs {Kb}")
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if out_fp32 else torch.bfloat16)
    aux = None
    if aux_out:
        aux = torch.empty((M, N), device=a.device, dtype=torch.bfloat16)
    elif aux_in is not None:
        aux = aux_in
    if bias is not None:
        bias = _bf16c(bias)
    if residual is not None:
        residual = _bf16c(residual)
    rc = lib.vlk_gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a.stride(0), b.stride(0),
                           out.stride(0), int(trans_a), int(trans_b), _p(bias), _p(residual),
                           residual.stride(0) if residual is not None else 0,
                           _p(aux) if (dact or aux_in is not None) else 0,
                           _p(aux) if aux_out else 0, aux.stride(0) if aux is not None else 0,
                           _p(scale), _ACT[act] if not isinstance(act, int) else act, int(dact), float(alpha),
                           int(out_fp32), _stream())
    check(rc, "vlk_gemm_bf16")
    return (out, aux) if aux_out else out
