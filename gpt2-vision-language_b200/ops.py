"""Tensor-level wrappers and autograd Functions over the libvlk C ABI (include/vlk.h).

Everything here runs on the CURRENT torch CUDA stream, is asynchronous and CUDA-graph capturable.
PyTorch owns every buffer; the native side borrows raw device pointers for the duration of a call
(SURVEY 8b "Ownership").  There is no CPU or eager-PyTorch fallback: CPU tensors raise RuntimeError.

Numerics contract (same as the reference's CUDA path, train_gpt2.py:264 + autocast): bf16 parameters,
bf16 activations and gradients, fp32 accumulation / statistics / softmax / loss inside the kernels.
"""
import math
import os

import torch

from . import _lib
from ._lib import check

ACT = {None: 0, "none": 0, "gelu_tanh": 1, "gelu_erf": 2, "quick_gelu": 3}
BF16 = torch.bfloat16


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("libvlk ops need CUDA tensors: there is no CPU path for the B200 kernels")


def _bf16c(t):
    """bf16 with a contiguous innermost dim (row stride may be arbitrary)."""
    if t.dtype != BF16:
        t = t.to(BF16)
    if t.stride(-1) != 1:
        t = t.contiguous()
    return t


def _rows(t):
    """[..., C] -> contiguous 2-D [rows, C] view (copy only if needed)."""
    t = _bf16c(t)
    t2 = t.reshape(-1, t.shape[-1])
    if t2.stride(-1) != 1 or (t2.shape[0] > 1 and t2.stride(0) % 8 != 0):
        t2 = t2.contiguous()
    return t2


# ======================================================================================================
# raw ops
# ======================================================================================================
def gemm(a, b, *, trans_a=False, trans_b=False, bias=None, residual=None, aux_in=None, aux_out=False,
         scale=None, act=None, dact=False, alpha=1.0, out=None, out_fp32=False, split_k=1):
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
    if split_k > 1:
        # deterministic split-K: fp32 partial slabs + one fixed-order reduction (vlk_gemm_bf16_splitk)
        if out_fp32 or bias is not None or residual is not None or aux_in is not None or aux_out or scale is not None \
                or act not in (None, "none", 0) or dact:
            raise RuntimeError("gemm: split_k > 1 supports a plain bf16 product only")
        ws = torch.empty((split_k, M, N), device=a.device, dtype=torch.float32)
        if out is None:
            out = torch.empty((M, N), device=a.device, dtype=BF16)
        check(lib.vlk_gemm_bf16_splitk(a.data_ptr(), b.data_ptr(), out.data_ptr(), ws.data_ptr(), M, N, K, a.stride(0),
                                       b.stride(0), out.stride(0), int(trans_a), int(trans_b), float(alpha),
                                       int(split_k), 0, _stream()), "vlk_gemm_bf16_splitk")
        return out
    split_k = -split_k if split_k < 0 else 1   # negative: raw atomic split-K into a caller-zeroed fp32 `out`
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if out_fp32 else BF16)
    aux = None
    if aux_out:
        aux = torch.empty((M, N), device=a.device, dtype=BF16)
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
                           _p(scale), ACT[act] if not isinstance(act, int) else act, int(dact), float(alpha),
                           int(out_fp32), int(split_k), _stream())
    check(rc, "vlk_gemm_bf16")
    return (out, aux) if aux_out else out


def gemm_tile(a, b, *, tile_n=0, cta_pair=-1, trans_a=False, trans_b=False, bias=None, residual=None, act=None):
    """ops.gemm with an EXPLICIT tile shape (vlk_gemm_bf16_tile): tuning sweeps and per-instantiation regression tests."""
    _need_cuda(a, b)
    a, b = _bf16c(a), _bf16c(b)
    M, K = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    N = b.shape[1] if trans_b else b.shape[0]
    out = torch.empty((M, N), device=a.device, dtype=BF16)
    if residual is not None:
        residual = _bf16c(residual)
    check(_lib.load().vlk_gemm_bf16_tile(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, a.stride(0), b.stride(0),
                                         out.stride(0), int(trans_a), int(trans_b), _p(bias), _p(residual),
                                         residual.stride(0) if residual is not None else 0,
                                         ACT[act] if not isinstance(act, int) else act, int(tile_n), int(cta_pair),
                                         _stream()), "vlk_gemm_bf16_tile")
    return out


_SM_COUNT = None


def auto_split_k(M, N, K):
    """Slices of the contraction for a product whose [M,N] output has too few 256x256 tiles to fill the GPU.
    Cost model: waves of 2-CTA tile slots x (k-blocks per slice + ~6 k-blocks of per-tile prologue/epilogue)."""
    global _SM_COUNT
    if _SM_COUNT is None:
        _SM_COUNT = max(2, _lib.load().vlk_num_sms())
    slots = _SM_COUNT // 2
    units = -(-M // 256) * -(-N // 256)
    kb = -(-K // 64)
    if units * 2 > slots or kb < 32:
        return 1
    best, best_cost = 1, -(-units // slots) * (kb + 6)
    for s in range(2, 17):
        cost = -(-units * s // slots) * (-(-kb // s) + 6)
        if cost < best_cost:
            best, best_cost = s, cost
    return best


def wgrad(dy2, x2, param=None):
    """dW[N_out, K_in] = dy^T . x  — the weight gradient of nn.Linear, contracted over the rows (tokens).

    When ``param`` already has a dense bf16 ``.grad`` (gradient accumulation over micro-batches, train_gpt2.py:458-469;
    the flat bucket of dp.FlatGradBucket), the product is ACCUMULATED into it inside the GEMM — in the split-K
    reduction or as the epilogue's residual operand, fp32 sum rounded once — and None is returned, so autograd does
    not launch a separate bf16 add per parameter.  Otherwise the gradient tensor is returned as usual."""
    split = auto_split_k(dy2.shape[1], x2.shape[1], dy2.shape[0])
    g = param.grad if (param is not None and param.is_leaf) else None   # slices of packed weights are not leaves
    if g is not None and g.dtype == BF16 and g.is_cuda and g.is_contiguous() and g.dim() == 2 \
            and g.data_ptr() % 16 == 0 and not torch.is_grad_enabled():
        if split > 1:
            dy2, x2 = _bf16c(dy2), _bf16c(x2)
            M, N, K = dy2.shape[1], x2.shape[1], dy2.shape[0]
            ws = torch.empty((split, M, N), device=g.device, dtype=torch.float32)
            check(_lib.load().vlk_gemm_bf16_splitk(dy2.data_ptr(), x2.data_ptr(), g.data_ptr(), ws.data_ptr(), M, N, K,
                                                   dy2.stride(0), x2.stride(0), g.stride(0), 1, 1, 1.0, int(split), 1,
                                                   _stream()), "vlk_gemm_bf16_splitk")
        else:
            gemm(dy2, x2, trans_a=True, trans_b=True, out=g, residual=g)
        return None
    return gemm(dy2, x2, trans_a=True, trans_b=True, split_k=split)


def colsum(x2d):
    """fp32 [cols] column sums of a bf16 [rows, cols] matrix (bias gradients)."""
    _need_cuda(x2d)
    out = torch.empty(x2d.shape[1], device=x2d.device, dtype=torch.float32)
    check(_lib.load().vlk_colsum_bf16(x2d.data_ptr(), out.data_ptr(), x2d.shape[0], x2d.shape[1], x2d.stride(0),
                                      _stream()), "vlk_colsum_bf16")
    return out


def grad_slot(param):
    """``param.grad`` when a libvlk kernel may ACCUMULATE this parameter's gradient into it in place: a dense bf16 CUDA
    tensor already attached (gradient accumulation over micro-batches, train_gpt2.py:458-469; the views of
    dp.FlatGradBucket), and autograd not recording (we are inside backward).  The Function then returns None for that
    input, so autograd launches no ``grad += new`` kernel of its own.  Otherwise None: the gradient is returned."""
    if param is None or not param.is_leaf:
        return None
    g = param.grad
    if g is not None and g.dtype == BF16 and g.is_cuda and g.is_contiguous() and g.data_ptr() % 16 == 0 \
            and not torch.is_grad_enabled():
        return g
    return None


def sum_copies(src, copies, n, dst, accumulate=False, clear_src=False):
    """dst[i] (= or +=) sum_c src[c*n + i]; dst fp32 or bf16 (vlk_sum_copies)."""
    check(_lib.load().vlk_sum_copies(src.data_ptr(), int(copies), int(n), dst.data_ptr(), int(dst.dtype == BF16),
                                     int(accumulate), int(clear_src), _stream()), "vlk_sum_copies")


def bias_grad(dy2, bias):
    """Bias gradient of nn.Linear = column sums of dy.  Accumulated straight into ``bias.grad`` when possible (returns
    None), else returned as a bf16 tensor."""
    g = grad_slot(bias)
    if g is not None:
        cols = dy2.shape[1]
        ws = persistent_workspace(("colsum_ws", dy2.device.index),
                                  lambda: (torch.zeros(65536, device=dy2.device, dtype=torch.float32),
                                           torch.zeros(256, device=dy2.device, dtype=torch.int32)))
        if ws is not None and cols <= 65536 and dy2.dtype == BF16 and dy2.stride(1) == 1:
            # one launch: the last block of each column group adds the sums into bias.grad and re-zeroes the workspace
            check(_lib.load().vlk_colsum_bf16_acc(dy2.data_ptr(), ws[0].data_ptr(), ws[1].data_ptr(), g.data_ptr(),
                                                  dy2.shape[0], cols, dy2.stride(0), _stream()), "vlk_colsum_bf16_acc")
            return None
        s32 = colsum(dy2)
        sum_copies(s32, 1, s32.numel(), g, accumulate=True)
        return None
    s32 = colsum(dy2)
    out = torch.empty(s32.numel(), device=s32.device, dtype=BF16)
    sum_copies(s32, 1, s32.numel(), out)
    return out


_WORKSPACES = {}


def persistent_workspace(key, make):
    """A process-lifetime device workspace (static address: CUDA-graph friendly) that its users leave in its initial
    state.  Never created while a stream is capturing — callers fall back to a per-call temporary then."""
    ws = _WORKSPACES.get(key)
    if ws is None and not torch.cuda.is_current_stream_capturing():
        ws = _WORKSPACES[key] = make()
    return ws


def transpose(x2d):
    _need_cuda(x2d)
    x2d = _bf16c(x2d)
    out = torch.empty((x2d.shape[1], x2d.shape[0]), device=x2d.device, dtype=BF16)
    check(_lib.load().vlk_transpose_bf16(x2d.data_ptr(), out.data_ptr(), x2d.shape[0], x2d.shape[1], x2d.stride(0),
                                         out.stride(0), _stream()), "vlk_transpose_bf16")
    return out


def layernorm_fwd(x2d, weight, bias, eps=1e-5, save_stats=True):
    _need_cuda(x2d, weight, bias)
    rows, cols = x2d.shape
    y = torch.empty_like(x2d)
    mean = rstd = None
    if save_stats:
        mean = torch.empty(rows, device=x2d.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x2d.device, dtype=torch.float32)
    check(_lib.load().vlk_layernorm_fwd(x2d.data_ptr(), weight.data_ptr(), bias.data_ptr(), y.data_ptr(), _p(mean),
                                        _p(rstd), rows, cols, float(eps), _stream()), "vlk_layernorm_fwd")
    return y, mean, rstd


def row_stats(x2d, eps=1e-5):
    """(mean, rstd) fp32 [rows] of a bf16 [rows, cols] matrix — the LayerNorm statistics without the normalised copy."""
    _need_cuda(x2d)
    rows, cols = x2d.shape
    mean = torch.empty(rows, device=x2d.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=x2d.device, dtype=torch.float32)
    check(_lib.load().vlk_row_stats(x2d.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, cols, float(eps), _stream()),
          "vlk_row_stats")
    return mean, rstd


def fold_layernorm(weight, bias, gamma, beta):
    """Fold LayerNorm(gamma, beta) into the Linear(weight [N,K], bias [N] or None) that consumes it (frozen weights):
    returns (Wf bf16 [N,K], colsum fp32 [N] of the ROUNDED Wf, biasf bf16 [N])."""
    w32, g32, b32 = weight.float(), gamma.float(), beta.float()
    wf = (w32 * g32[None, :]).to(BF16)
    colsum = wf.float().sum(dim=1).contiguous()
    biasf = w32 @ b32
    if bias is not None:
        biasf = biasf + bias.float()
    return wf.contiguous(), colsum, biasf.to(BF16).contiguous()


def gemm_stats(a, w, bias, residual, stats_out):
    """bf16(a @ w^T + bias + residual), with the row sums (sum, sum of squares) of the rounded result accumulated into
    ``stats_out`` [M, 2] fp32 (zeroed by the caller) by the same epilogue (vlk_gemm_bf16_stats)."""
    _need_cuda(a, w, residual)
    a, w, residual = _bf16c(a), _bf16c(w), _bf16c(residual)
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty((M, N), device=a.device, dtype=BF16)
    check(_lib.load().vlk_gemm_bf16_stats(a.data_ptr(), w.data_ptr(), out.data_ptr(), M, N, K, a.stride(0), w.stride(0),
                                          out.stride(0), _p(bias), residual.data_ptr(), residual.stride(0),
                                          stats_out.data_ptr(), _stream()), "vlk_gemm_bf16_stats")
    return out


def gemm_lnfold(x2d, wf, biasf, colsum, eps=1e-5, act=None, stats=None, sums=None):
    """act(LayerNorm(x) @ W^T + b) computed as rstd * (x @ Wf^T - mean * colsum) + biasf on the RAW x
    (vlk_gemm_bf16_lnfold).  The statistics are either ``stats`` = (mean, rstd) from row_stats (computed here when
    omitted) or ``sums`` = [M, 2] row sums accumulated by the GEMM that produced x (gemm_stats)."""
    _need_cuda(x2d, wf)
    x2d = _bf16c(x2d)
    M, K = x2d.shape
    N = wf.shape[0]
    out = torch.empty((M, N), device=x2d.device, dtype=BF16)
    if sums is not None:
        check(_lib.load().vlk_gemm_bf16_lnfold_sums(x2d.data_ptr(), wf.data_ptr(), out.data_ptr(), M, N, K, x2d.stride(0),
                                                    wf.stride(0), out.stride(0), _p(biasf), sums.data_ptr(), float(eps),
                                                    colsum.data_ptr(), ACT[act] if not isinstance(act, int) else act,
                                                    _stream()), "vlk_gemm_bf16_lnfold_sums")
        return out
    mean, rstd = stats if stats is not None else row_stats(x2d, eps)
    check(_lib.load().vlk_gemm_bf16_lnfold(x2d.data_ptr(), wf.data_ptr(), out.data_ptr(), M, N, K, x2d.stride(0),
                                           wf.stride(0), out.stride(0), _p(biasf), mean.data_ptr(), rstd.data_ptr(),
                                           colsum.data_ptr(), ACT[act] if not isinstance(act, int) else act, _stream()),
          "vlk_gemm_bf16_lnfold")
    return out


LN_GRAD_COPIES = 8


def layernorm_bwd(dy2d, x2d, weight, mean, rstd, param_grads=False, dx=None, accumulate=False, grads_bf16=False,
                  bias=None):
    """-> dx, dgamma, dbeta.  The parameter gradients are accumulated by ~300 blocks into LN_GRAD_COPIES replicas
    (fewer same-address atomics) that one small kernel sums — into fp32, or bf16 when ``grads_bf16``.  When ``bias`` (the
    LayerNorm's bias Parameter) is given and both ``weight.grad`` / ``bias.grad`` can take an in-place accumulation
    (grad_slot), the sums are ADDED into them and None is returned for dgamma / dbeta."""
    rows, cols = x2d.shape
    lib = _lib.load()
    if dx is None:
        assert not accumulate
        dx = torch.empty_like(x2d)
    acc = None
    persistent = False
    if param_grads:
        acc = persistent_workspace(("ln_acc", x2d.device.index, cols),
                                   lambda: torch.zeros((2, LN_GRAD_COPIES, cols), device=x2d.device, dtype=torch.float32))
        persistent = acc is not None
        if acc is None:
            acc = torch.zeros((2, LN_GRAD_COPIES, cols), device=x2d.device, dtype=torch.float32)
        gw, gb = grad_slot(weight), grad_slot(bias) if bias is not None else None
        counter = persistent_workspace(("ln_counter", x2d.device.index),
                                       lambda: torch.zeros(1, device=x2d.device, dtype=torch.int32))
        if persistent and counter is not None and gw is not None and gb is not None:
            # one launch: the last block folds the replicas into weight.grad / bias.grad and re-zeroes the workspace
            check(lib.vlk_layernorm_bwd_acc(dy2d.data_ptr(), x2d.data_ptr(), weight.data_ptr(), mean.data_ptr(),
                                            rstd.data_ptr(), dx.data_ptr(), acc[0].data_ptr(), acc[1].data_ptr(), rows, cols,
                                            int(accumulate), LN_GRAD_COPIES, gw.data_ptr(), gb.data_ptr(),
                                            counter.data_ptr(), _stream()), "vlk_layernorm_bwd_acc")
            return dx, None, None
    check(lib.vlk_layernorm_bwd(dy2d.data_ptr(), x2d.data_ptr(), weight.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                dx.data_ptr(), _p(acc), acc[1].data_ptr() if acc is not None else 0, rows, cols,
                                int(accumulate), LN_GRAD_COPIES, _stream()), "vlk_layernorm_bwd")
    dg = db = None
    if param_grads:
        gw, gb = grad_slot(weight), grad_slot(bias) if bias is not None else None
        if gw is not None and gb is not None:
            sum_copies(acc[0], LN_GRAD_COPIES, cols, gw, accumulate=True, clear_src=persistent)
            sum_copies(acc[1], LN_GRAD_COPIES, cols, gb, accumulate=True, clear_src=persistent)
        else:
            out = torch.empty((2, cols), device=x2d.device, dtype=BF16 if grads_bf16 else torch.float32)
            for k in range(2):
                sum_copies(acc[k], LN_GRAD_COPIES, cols, out[k], clear_src=persistent)
            dg, db = out[0], out[1]
    return dx, dg, db


def _bt_strides(t):
    """(batch stride, row stride) in elements of a [B,T,W] view whose last dim is contiguous."""
    assert t.dim() == 3 and t.stride(2) == 1
    return t.stride(0), t.stride(1)


class DropoutState:
    """Device-resident Philox key for the dropout kernels: int64 [seed, step].

    Every dropout call site draws a fresh ``stream id`` from a host counter (so eager calls never repeat a mask)
    and the device ``step`` is bumped by ``advance()`` once per optimizer step — that is what makes masks change
    between replays of a captured CUDA graph, where the stream ids are baked in.  Forward and backward of one site
    use the same (seed, step, stream id) triple and therefore regenerate the same mask."""

    _per_device = {}

    def __init__(self, device, seed=None):
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        self.state = torch.tensor([seed, 0], dtype=torch.int64, device=device)
        self._next = 0

    def new_stream(self):
        self._next = (self._next + 1) & 0x7FFFFFFF
        return self._next

    def advance(self):
        self.state[1:2] += 1

    @classmethod
    def default(cls, device):
        key = str(device)
        if key not in cls._per_device:
            cls._per_device[key] = cls(device)
        return cls._per_device[key]


def attention_fwd(q, k, v, n_head, causal, scale=None, need_lse=True, dropout_p=0.0, rng=None, stream_id=0):
    """q: [B,Tq,H*64] view, k/v: [B,Tk,H*64] views (slices of packed projections are fine). Returns o [B,Tq,H*64], lse."""
    _need_cuda(q, k, v)
    B, Tq, W = q.shape
    Tk = k.shape[1]
    assert W == n_head * 64, "libvlk attention is specialised for head_dim 64"
    scale = 1.0 / math.sqrt(64) if scale is None else scale
    o = torch.empty((B, Tq, W), device=q.device, dtype=BF16)
    lse = torch.empty((B, n_head, Tq), device=q.device, dtype=torch.float32) if need_lse else None
    (qb, qr), (kb, kr), (vb, vr), (ob, orr) = _bt_strides(q), _bt_strides(k), _bt_strides(v), _bt_strides(o)
    check(_lib.load().vlk_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), _p(lse), B, n_head, Tq,
                                   Tk, qb, qr, kb, kr, vb, vr, ob, orr, int(causal), float(scale), float(dropout_p),
                                   rng.state.data_ptr() if rng is not None else 0, int(stream_id), _stream()),
          "vlk_attn_fwd")
    return o, lse


def attention_bwd(q, k, v, o, d_o, lse, dq, dk, dv, n_head, causal, scale=None, dropout_p=0.0, rng=None, stream_id=0):
    """Writes dq/dk/dv (pre-allocated views with the primal shapes)."""
    B, Tq, W = q.shape
    Tk = k.shape[1]
    scale = 1.0 / math.sqrt(64) if scale is None else scale
    d_o = _bf16c(d_o)
    if d_o.stride() != o.stride():
        d_o = d_o.contiguous()
        assert d_o.stride() == o.stride()
    delta = torch.empty((B, n_head, Tq), device=q.device, dtype=torch.float32)
    (qb, qr), (kb, kr), (vb, vr), (ob, orr) = _bt_strides(q), _bt_strides(k), _bt_strides(v), _bt_strides(o)
    (dqb, dqr), (dkb, dkr), (dvb, dvr) = _bt_strides(dq), _bt_strides(dk), _bt_strides(dv)
    check(_lib.load().vlk_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), d_o.data_ptr(),
                                   lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), B, n_head, Tq, Tk,
                                   qb, qr, kb, kr, vb, vr, ob, orr, dqb, dqr, dkb, dkr, dvb, dvr, int(causal),
                                   float(scale), delta.data_ptr(), float(dropout_p),
                                   rng.state.data_ptr() if rng is not None else 0, int(stream_id), _stream()),
          "vlk_attn_bwd")


def dropout_add_raw(x, residual, p, rng, stream_id):
    """y = residual + x * keep/(1-p) (residual may be None); mask = Philox(rng.state, stream_id)."""
    _need_cuda(x)
    x = _bf16c(x).contiguous()
    if residual is not None:
        residual = _bf16c(residual).contiguous()
    y = torch.empty_like(x)
    check(_lib.load().vlk_dropout_add_bf16(x.data_ptr(), _p(residual), y.data_ptr(), x.numel(), float(p),
                                           rng.state.data_ptr(), int(stream_id), _stream()), "vlk_dropout_add_bf16")
    return y


def pool33(tokens, normalize=True):
    """[B,257,D] -> [B,33,D] (CLS + 4x8 average bins, optional per-token L2 normalise); bf16 or fp32."""
    _need_cuda(tokens)
    assert tokens.dim() == 3 and tokens.shape[1] == 257, "expects CLS + 16x16 patch tokens"
    fp32 = tokens.dtype == torch.float32
    if not fp32:
        tokens = _bf16c(tokens)
    tokens = tokens.contiguous()
    B, _, D = tokens.shape
    out = torch.empty((B, 33, D), device=tokens.device, dtype=tokens.dtype)
    check(_lib.load().vlk_pool33_l2norm(tokens.data_ptr(), out.data_ptr(), B, D, int(fp32), int(normalize),
                                        _stream()), "vlk_pool33_l2norm")
    return out


def embed_concat(ids, wte, wpe, prefix=None, pos0=0):
    _need_cuda(ids, wte, wpe)
    ids = ids.contiguous()
    assert ids.dtype == torch.int64
    B, T = ids.shape
    C = wte.shape[1]
    P = 0 if prefix is None else prefix.shape[1]
    if prefix is not None:
        prefix = _bf16c(prefix).contiguous()
    out = torch.empty((B, P + T, C), device=ids.device, dtype=BF16)
    check(_lib.load().vlk_embed_concat_fwd(ids.data_ptr(), wte.data_ptr(), wpe.data_ptr(), _p(prefix),
                                           out.data_ptr(), B, T, P, C, int(pos0), int(wte.shape[0]), _stream()),
          "vlk_embed_concat_fwd")
    return out


def add(a, b):
    _need_cuda(a, b)
    a, b = _bf16c(a).contiguous(), _bf16c(b).contiguous()
    y = torch.empty_like(a)
    check(_lib.load().vlk_add_bf16(a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), _stream()), "vlk_add_bf16")
    return y


def argmax_rows(logits2d):
    _need_cuda(logits2d)
    rows, V = logits2d.shape
    out = torch.empty(rows, device=logits2d.device, dtype=torch.int64)
    check(_lib.load().vlk_argmax_rows(logits2d.data_ptr(), out.data_ptr(), rows, V, logits2d.stride(0), _stream()),
          "vlk_argmax_rows")
    return out


def im2col_patch14(pixels, kpad=640):
    _need_cuda(pixels)
    assert pixels.shape[1:] == (3, 224, 224)
    fp32 = pixels.dtype == torch.float32
    if not fp32:
        pixels = pixels.to(BF16)
    pixels = pixels.contiguous()
    B = pixels.shape[0]
    out = torch.empty((B * 256, kpad), device=pixels.device, dtype=BF16)
    check(_lib.load().vlk_im2col_patch14(pixels.data_ptr(), out.data_ptr(), B, kpad, int(fp32), _stream()),
          "vlk_im2col_patch14")
    return out


def clip_assemble(patch2d, cls, pos, B):
    D = patch2d.shape[1]
    out = torch.empty((B, 257, D), device=patch2d.device, dtype=BF16)
    check(_lib.load().vlk_clip_assemble(patch2d.data_ptr(), cls.data_ptr(), pos.data_ptr(), out.data_ptr(), B, D,
                                        _stream()), "vlk_clip_assemble")
    return out


# ======================================================================================================
# autograd Functions (what the drop-in nn.Modules are made of)
# ======================================================================================================
def _param_ok(*ps):
    for p in ps:
        if p is not None and (p.dtype != BF16 or not p.is_cuda):
            raise RuntimeError(
                "the B200 path needs bf16 CUDA parameters (the reference does model.to(device).to(torch.bfloat16), "
                "source/gpt2/train_gpt2.py:263-264); got dtype=%s device=%s" % (p.dtype, p.device))


class LinearFn(torch.autograd.Function):
    """y = x @ W^T + b (+ residual).  nn.Linear forward + dgrad/wgrad, all on vlk_gemm_bf16."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual):
        _param_ok(weight, bias)
        x2 = _rows(x)
        res2 = _rows(residual) if residual is not None else None
        y = gemm(x2, weight, bias=bias, residual=res2)
        ctx.save_for_backward(x2, weight)
        ctx.has_bias = bias is not None
        ctx.bias_param = bias
        ctx.has_res = residual is not None
        ctx.x_shape = x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        dy2 = _rows(dy)
        dx = dw = db = dres = None
        if ctx.needs_input_grad[0]:
            dx = gemm(dy2, weight, trans_b=True).view(ctx.x_shape)          # [M,N] x [N,K]
        if ctx.needs_input_grad[1]:
            dw = wgrad(dy2, x2, weight)                                     # dy^T [N,M] x x [M,K]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = bias_grad(dy2, ctx.bias_param)
        if ctx.has_res and ctx.needs_input_grad[3]:
            dres = dy
        return dx, dw, db, dres


def linear(x, weight, bias=None, residual=None):
    return LinearFn.apply(x, weight, bias, residual)


class MLPFn(torch.autograd.Function):
    """y = residual + act(x @ Wfc^T + bfc) @ Wproj^T + bproj with the activation and its gradient fused in the
    GEMM epilogues (MLP at train_gpt2.py:46-59; Q-Former mlp at gpt2_q_former/model.py:126-130)."""

    @staticmethod
    def forward(ctx, x, w_fc, b_fc, w_proj, b_proj, residual, act):
        _param_ok(w_fc, b_fc, w_proj, b_proj)
        x2 = _rows(x)
        res2 = _rows(residual) if residual is not None else None
        need_bwd = any(ctx.needs_input_grad)
        if need_bwd:
            h, u = gemm(x2, w_fc, bias=b_fc, act=act, aux_out=True)
        else:
            h, u = gemm(x2, w_fc, bias=b_fc, act=act), None
        y = gemm(h, w_proj, bias=b_proj, residual=res2)
        if need_bwd:
            wgrad = any(ctx.needs_input_grad[1:5])
            ctx.save_for_backward(x2 if wgrad else None, u, h if wgrad else None, w_fc, w_proj)
        ctx.act = act
        ctx.x_shape = x.shape
        ctx.has_res = residual is not None
        ctx.bias_params = (b_fc, b_proj)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, u, h, w_fc, w_proj = ctx.saved_tensors
        dy2 = _rows(dy)
        du = gemm(dy2, w_proj, trans_b=True, aux_in=u, act=ctx.act, dact=True)   # (dy @ Wproj) * act'(u)
        dx = dwfc = dbfc = dwp = dbp = dres = None
        if ctx.needs_input_grad[0]:
            dx = gemm(du, w_fc, trans_b=True).view(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            dwfc = wgrad(du, x2, w_fc)
        if ctx.needs_input_grad[2]:
            dbfc = bias_grad(du, ctx.bias_params[0])
        if ctx.needs_input_grad[3]:
            dwp = wgrad(dy2, h, w_proj)
        if ctx.needs_input_grad[4]:
            dbp = bias_grad(dy2, ctx.bias_params[1])
        if ctx.has_res and ctx.needs_input_grad[5]:
            dres = dy
        return dx, dwfc, dbfc, dwp, dbp, dres, None


def mlp(x, w_fc, b_fc, w_proj, b_proj, residual=None, act="gelu_tanh"):
    return MLPFn.apply(x, w_fc, b_fc, w_proj, b_proj, residual, act)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        _param_ok(weight, bias)
        x2 = _rows(x)
        need_bwd = any(ctx.needs_input_grad)
        y, mean, rstd = layernorm_fwd(x2, weight, bias, eps, save_stats=need_bwd)
        if need_bwd:
            ctx.save_for_backward(x2, weight, mean, rstd)
        ctx.bias_param = bias
        ctx.x_shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, weight, mean, rstd = ctx.saved_tensors
        pg = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx, dg, db = layernorm_bwd(_rows(dy).contiguous(), x2, weight, mean, rstd, param_grads=pg, grads_bf16=True,
                                   bias=ctx.bias_param if (ctx.needs_input_grad[1] and ctx.needs_input_grad[2]) else None)
        return (dx.view(ctx.x_shape) if ctx.needs_input_grad[0] else None,
                dg if (pg and ctx.needs_input_grad[1]) else None,
                db if (pg and ctx.needs_input_grad[2]) else None, None)


def layernorm(x, weight, bias, eps=1e-5):
    return LayerNormFn.apply(x, weight, bias, eps)


_RESIDUAL_GRAD_INPLACE = False


class residual_grad_inplace:
    """Context manager: inside it, ``ResidualLayerNormFn.backward`` accumulates the LayerNorm input gradient IN PLACE
    into the incoming residual-stream gradient (one read-modify-write inside the LayerNorm-backward kernel instead of
    a separate elementwise add per norm).

    Aliasing contract: LinearFn / MLPFn / GatedProjFn hand the residual gradient through unchanged (``dres = dy``, the
    same tensor), so with the in-place path ONE buffer is mutated all the way down the residual stream.  That is only
    safe when nobody else reads those gradient tensors: no ``retain_grad()`` / tensor hooks on the residual stream, no
    ``backward(retain_graph=True)`` run twice, and a tensor passed as ``backward(gradient=G)`` is clobbered.  The step
    classes (step.CaptionTrainStep / PretrainStep) own every tensor involved and opt in; everywhere else the default
    out-of-place path keeps autograd's value semantics."""

    def __init__(self, enabled=True):
        self.enabled = enabled

    def __enter__(self):
        global _RESIDUAL_GRAD_INPLACE
        self.prev, _RESIDUAL_GRAD_INPLACE = _RESIDUAL_GRAD_INPLACE, self.enabled

    def __exit__(self, *exc):
        global _RESIDUAL_GRAD_INPLACE
        _RESIDUAL_GRAD_INPLACE = self.prev


class ResidualLayerNormFn(torch.autograd.Function):
    """(x, LayerNorm(x)) for a pre-LN residual block (train_gpt2.py:72-73: ``x = x + f(ln(x))``).  Returning the
    residual stream through the same node lets backward ADD the LayerNorm input gradient straight into the
    incoming residual gradient (in place when ``residual_grad_inplace`` is active, see its aliasing contract)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        _param_ok(weight, bias)
        x2 = _rows(x)
        need_bwd = any(ctx.needs_input_grad)
        y, mean, rstd = layernorm_fwd(x2, weight, bias, eps, save_stats=need_bwd)
        if need_bwd:
            ctx.save_for_backward(x2, weight, mean, rstd)
        ctx.bias_param = bias
        ctx.x_shape = x.shape
        return x.view_as(x), y.view(x.shape)

    @staticmethod
    def backward(ctx, g_x, g_y):
        x2, weight, mean, rstd = ctx.saved_tensors
        pg = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dx = dg = db = None
        if g_y is None:
            return g_x, None, None, None
        dy2 = _rows(g_y).contiguous()
        bias_p = ctx.bias_param if (ctx.needs_input_grad[1] and ctx.needs_input_grad[2]) else None
        if _RESIDUAL_GRAD_INPLACE and g_x is not None and g_x.dtype == BF16 and g_x.is_contiguous():
            # opted in: the residual gradient has no reader left but this node (its producers ran earlier in backward)
            acc = g_x.view(-1, g_x.shape[-1])
            _, dg, db = layernorm_bwd(dy2, x2, weight, mean, rstd, param_grads=pg, dx=acc, accumulate=True, grads_bf16=True,
                                      bias=bias_p)
            dx = g_x
        else:
            dxl, dg, db = layernorm_bwd(dy2, x2, weight, mean, rstd, param_grads=pg, grads_bf16=True, bias=bias_p)
            dx = dxl.view(ctx.x_shape) if g_x is None else add(g_x, dxl.view(ctx.x_shape)).view(ctx.x_shape)
        return (dx if ctx.needs_input_grad[0] else None,
                dg if (pg and ctx.needs_input_grad[1]) else None,
                db if (pg and ctx.needs_input_grad[2]) else None, None)


def residual_layernorm(x, weight, bias, eps=1e-5):
    """-> (x, LayerNorm(x)); use the returned x as the residual operand of the branch (see ResidualLayerNormFn)."""
    return ResidualLayerNormFn.apply(x, weight, bias, eps)


class SelfAttnFn(torch.autograd.Function):
    """Attention over a packed projection qkv [B,T,3C] (c_attn output, train_gpt2.py:35-41; CLIP fused q/k/v;
    nn.MultiheadAttention in_proj with q=k=v).  The gradient is written straight into a packed [B,T,3C] buffer."""

    @staticmethod
    def forward(ctx, qkv, n_head, causal, dropout_p=0.0, rng=None, stream_id=0):
        qkv = _bf16c(qkv)
        C = qkv.shape[-1] // 3
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        need_bwd = ctx.needs_input_grad[0]
        o, lse = attention_fwd(q, k, v, n_head, causal, need_lse=need_bwd, dropout_p=dropout_p, rng=rng,
                               stream_id=stream_id)
        if need_bwd:
            ctx.save_for_backward(qkv, o, lse)
        ctx.n_head, ctx.causal = n_head, causal
        ctx.drop = (dropout_p, rng, stream_id)
        return o

    @staticmethod
    def backward(ctx, d_o):
        qkv, o, lse = ctx.saved_tensors
        C = qkv.shape[-1] // 3
        dqkv = torch.empty_like(qkv)
        p, rng, sid = ctx.drop
        attention_bwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], o, d_o, lse, dqkv[..., :C],
                      dqkv[..., C:2 * C], dqkv[..., 2 * C:], ctx.n_head, ctx.causal, dropout_p=p, rng=rng,
                      stream_id=sid)
        return dqkv, None, None, None, None, None


def self_attention(qkv, n_head, causal, dropout_p=0.0, rng=None):
    sid = rng.new_stream() if (dropout_p > 0 and rng is not None) else 0
    return SelfAttnFn.apply(qkv, n_head, causal, dropout_p, rng, sid)


class CrossAttnFn(torch.autograd.Function):
    """q [B,T,C] attends to a packed kv [B,S,2C] (CrossAttention at gpt2_cross-att/model.py:45-57; the cross
    nn.MultiheadAttention of the Q-Former, gpt2_q_former/model.py:140).  Non-causal."""

    @staticmethod
    def forward(ctx, q, kv, n_head, dropout_p=0.0, rng=None, stream_id=0):
        q, kv = _bf16c(q), _bf16c(kv)
        C = q.shape[-1]
        need_bwd = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        o, lse = attention_fwd(q, kv[..., :C], kv[..., C:], n_head, False, need_lse=need_bwd, dropout_p=dropout_p,
                               rng=rng, stream_id=stream_id)
        if need_bwd:
            ctx.save_for_backward(q, kv, o, lse)
        ctx.n_head = n_head
        ctx.drop = (dropout_p, rng, stream_id)
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, kv, o, lse = ctx.saved_tensors
        C = q.shape[-1]
        dq, dkv = torch.empty_like(q), torch.empty_like(kv)
        p, rng, sid = ctx.drop
        attention_bwd(q, kv[..., :C], kv[..., C:], o, d_o, lse, dq, dkv[..., :C], dkv[..., C:], ctx.n_head, False,
                      dropout_p=p, rng=rng, stream_id=sid)
        return dq, dkv, None, None, None, None


def cross_attention(q, kv, n_head, dropout_p=0.0, rng=None):
    sid = rng.new_stream() if (dropout_p > 0 and rng is not None) else 0
    return CrossAttnFn.apply(q, kv, n_head, dropout_p, rng, sid)


class DropoutAddFn(torch.autograd.Function):
    """residual + dropout(x): nn.Dropout on a residual branch (gpt2_q_former/model.py:136,141,144)."""

    @staticmethod
    def forward(ctx, x, residual, p, rng, stream_id):
        ctx.args = (p, rng, stream_id)
        return dropout_add_raw(x, residual, p, rng, stream_id).view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        p, rng, sid = ctx.args
        dx = dropout_add_raw(dy, None, p, rng, sid).view(dy.shape) if ctx.needs_input_grad[0] else None
        return dx, dy if ctx.needs_input_grad[1] else None, None, None, None


def dropout_add(x, residual, p, rng):
    return DropoutAddFn.apply(x, residual, p, rng, rng.new_stream())


class GatedProjFn(torch.autograd.Function):
    """x + tanh(gate) * (y @ W^T + b): the gated cross-attention residual of gpt2_cross-att/model.py:101.
    tanh(gate) is a device scalar applied in the GEMM epilogue; the un-gated projection is kept (aux) for
    the gate gradient  d gate = (1 - tanh^2) * sum(dout * proj)."""

    @staticmethod
    def forward(ctx, y, weight, bias, gate, residual):
        _param_ok(weight, bias)
        y2, res2 = _rows(y), _rows(residual)
        tg = torch.tanh(gate.detach().float()).reshape(1)
        need_gate = ctx.needs_input_grad[3]
        if need_gate:
            out, proj = gemm(y2, weight, bias=bias, scale=tg, residual=res2, aux_out=True)
        else:
            out, proj = gemm(y2, weight, bias=bias, scale=tg, residual=res2), None
        ctx.save_for_backward(y2, weight, tg, proj, gate)
        ctx.shape = y.shape
        ctx.has_bias = bias is not None
        return out.view(residual.shape)

    @staticmethod
    def backward(ctx, dout):
        y2, weight, tg, proj, gate = ctx.saved_tensors
        d2 = _rows(dout).contiguous()
        dy = dw = db = dgate = None
        if ctx.needs_input_grad[3]:
            acc = torch.zeros(1, device=d2.device, dtype=torch.float32)
            g32 = gate.detach().float().reshape(1)
            check(_lib.load().vlk_gate_grad(d2.data_ptr(), proj.data_ptr(), g32.data_ptr(), acc.data_ptr(),
                                            d2.numel(), _stream()), "vlk_gate_grad")
            dgate = acc.reshape(gate.shape).to(gate.dtype)
        if ctx.needs_input_grad[0]:
            dy = gemm(d2, weight, trans_b=True, scale=tg).view(ctx.shape)     # tanh(g) * dout @ W
        if ctx.needs_input_grad[1]:
            dw = gemm(d2, y2, trans_a=True, trans_b=True, scale=tg)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = (colsum(d2) * tg).to(BF16)
        return dy, dw, db, dgate, dout if ctx.needs_input_grad[4] else None


def gated_proj_residual(y, weight, bias, gate, residual):
    return GatedProjFn.apply(y, weight, bias, gate, residual)


class EmbedConcatFn(torch.autograd.Function):
    """[prefix ; wte[ids] + wpe[pos]] (GPT.forward train_gpt2.py:114-117; caption variant
    gpt2_linear/model.py:187-200: positions restart at 0 for the text, the image prefix has none)."""

    @staticmethod
    def forward(ctx, ids, wte, wpe, prefix):
        _param_ok(wte, wpe)
        out = embed_concat(ids, wte, wpe, prefix)
        ctx.save_for_backward(ids)
        ctx.P = 0 if prefix is None else prefix.shape[1]
        ctx.wte_shape, ctx.wpe_shape = wte.shape, wpe.shape
        ctx.params = (wte, wpe)
        return out

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        dout = _bf16c(dout).contiguous()
        dwte = dwpe = dprefix = None
        need_wte, need_wpe = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if need_wte or need_wpe:
            B, T = ids.shape
            C = dout.shape[-1]
            wte, wpe = ctx.params
            V = ctx.wte_shape[0]
            g_wte = grad_slot(wte) if need_wte else None
            g_wpe = grad_slot(wpe) if need_wpe else None
            dev = dout.device
            first = scratch = None
            if (not need_wte or g_wte is not None) and (not need_wpe or g_wpe is not None):
                if need_wte:
                    first = persistent_workspace(("embed_first", dev.index, V), lambda: torch.full(
                        (V,), 2 ** 31 - 1, device=dev, dtype=torch.int32))
                    scratch = persistent_workspace(("embed_scratch", dev.index, B * T, C), lambda: torch.zeros(
                        (B * T, C), device=dev, dtype=torch.float32))
                if not need_wte or (first is not None and scratch is not None):
                    # accumulate straight into the bf16 gradients (the flat bucket): no dense [V, C] temporaries
                    check(_lib.load().vlk_embed_bwd_acc(ids.data_ptr(), dout.data_ptr(), _p(g_wte), _p(g_wpe[:T] if g_wpe is not None else None),
                                                        _p(first), _p(scratch), B, T, ctx.P, C, V, _stream()),
                          "vlk_embed_bwd_acc")
                    dprefix = dout[:, :ctx.P] if (ctx.P and ctx.needs_input_grad[3]) else None
                    return None, None, None, dprefix
            f_wte = torch.zeros(ctx.wte_shape, device=dev, dtype=torch.float32) if need_wte else None
            f_wpe = torch.zeros(ctx.wpe_shape, device=dev, dtype=torch.float32) if need_wpe else None
            check(_lib.load().vlk_embed_bwd(ids.data_ptr(), dout.data_ptr(), _p(f_wte), _p(f_wpe), B, T, ctx.P, C,
                                            _stream()), "vlk_embed_bwd")
            dwte = f_wte.to(BF16) if need_wte else None
            dwpe = f_wpe.to(BF16) if need_wpe else None
        if ctx.P and ctx.needs_input_grad[3]:
            dprefix = dout[:, :ctx.P]
        return None, dwte, dwpe, dprefix


def embed(ids, wte, wpe, prefix=None):
    return EmbedConcatFn.apply(ids, wte, wpe, prefix)


def _lmhead_ce_forward(h2, weight, labels, rw):
    """vlk_lmhead_ce_fwd -> (stats [loss, 1/count], loss_row [rows], lse [rows]); no [rows, V] buffer exists."""
    lib = _lib.load()
    rows, C = h2.shape
    V = weight.shape[0]
    dev = h2.device
    stats = torch.empty(2, device=dev, dtype=torch.float32)
    loss_row = torch.empty(rows, device=dev, dtype=torch.float32)
    lse = torch.empty(rows, device=dev, dtype=torch.float32)
    nbytes = int(lib.vlk_lmhead_ce_workspace_bytes(rows, C, V, 0, 0, 0))
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    check(lib.vlk_lmhead_ce_fwd(h2.data_ptr(), weight.data_ptr(), labels.data_ptr(), _p(rw), stats.data_ptr(),
                                loss_row.data_ptr(), lse.data_ptr(), rows, C, V, h2.stride(0), weight.stride(0),
                                ws.data_ptr(), nbytes, _stream()), "vlk_lmhead_ce_fwd")
    return stats, loss_row, lse


class LMHeadCEFn(torch.autograd.Function):
    """Mean cross-entropy of (h @ W^T) against labels with the lm_head fused into the loss (vlk_lmhead_ce_fwd / _bwd):
    the forward GEMM's epilogue keeps a running (max, sum exp) per row on chip and never stores logits; the backward
    recomputes the logit tiles of L2-sized vocabulary chunks, turns them into softmax - onehot in the GEMM epilogue and
    feeds the chunk straight into the d h (and, for a trainable head, d W) products.  No [rows, V] buffer in HBM in
    either direction.  Replaces lm_head + F.cross_entropy at train_gpt2.py:121-124, gpt2_linear/model.py:172,204-210
    (ignore_index=-100) and the masked mean of gpt2_cross-att/model.py:176-185 (row_weight = mask).  The incoming
    scalar gradient (e.g. the 1/grad_accum of train_gpt2.py:464) is folded into the same epilogue."""

    # Backward geometry (row_block, chunk_cols); 0 = the library's automatic choice: 4,096-row blocks x vocabulary chunks
    # whose d-logits stay L2-resident.  Measured on B200 (scripts/lmhead_ce_ab.py, profiles/r02_lmhead_ce_ab.log):
    #   frozen head, 1,984 rows:  automatic 473 us | one chunk 383 us | round-1 logits-in-HBM sequence 340 us
    #   trainable head, 16,384 rows:  automatic 5.50 ms | one block x one chunk 4.37 ms | round-1 sequence 3.88 ms
    # Recomputing the logit tiles costs one extra 2*rows*V*C product; with a trainable head (3 products become 4) that is
    # more than the logits' HBM round trip ever cost, so pretraining takes the widest chunks (the d-logits of one
    # micro-batch exist only inside the backward call, never across the trunk's backward); captioning, where the extra
    # product is 0.9 % of the step, keeps everything L2-resident.
    GEOMETRY = (0, 0)                      # frozen head (captioning)
    GEOMETRY_DW = (1 << 20, 1 << 20)       # trainable head (pretraining): clamped to (rows, V) by the library

    @staticmethod
    def forward(ctx, h, weight, labels, row_weight):
        _param_ok(weight)
        h2 = _rows(h)
        labels = labels.reshape(-1).contiguous()
        assert labels.dtype == torch.int64 and labels.numel() == h2.shape[0]
        rw = row_weight.reshape(-1).float().contiguous() if row_weight is not None else None
        stats, _, lse = _lmhead_ce_forward(h2, weight, labels, rw)
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            ctx.save_for_backward(h2, weight, labels, rw, lse, stats)
        ctx.h_shape = h.shape
        return stats[0]

    @staticmethod
    def backward(ctx, dloss):
        h2, weight, labels, rw, lse, stats = ctx.saved_tensors
        lib = _lib.load()
        rows, C = h2.shape
        V = weight.shape[0]
        need_dh, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dh = torch.empty_like(h2) if need_dh else None
        dw = dw_ret = None
        accumulate = 0
        if need_dw:
            g = weight.grad if weight.is_leaf else None
            if g is not None and g.dtype == BF16 and g.is_cuda and g.is_contiguous() and not torch.is_grad_enabled():
                dw, accumulate = g, 1          # gradient accumulation / flat bucket: summed inside the GEMM (see wgrad)
            else:
                dw = dw_ret = torch.empty_like(weight)
        dl = dloss.detach().reshape(1)
        if dl.dtype != torch.float32:
            dl = dl.float()
        rb, vc = LMHeadCEFn.GEOMETRY_DW if need_dw else LMHeadCEFn.GEOMETRY
        nbytes = int(lib.vlk_lmhead_ce_workspace_bytes(rows, C, V, 1, rb, vc))
        ws = torch.empty(nbytes, device=h2.device, dtype=torch.uint8)
        check(lib.vlk_lmhead_ce_bwd(h2.data_ptr(), weight.data_ptr(), labels.data_ptr(), _p(rw), lse.data_ptr(),
                                    stats[1:].data_ptr(), dl.data_ptr(), _p(dh), _p(dw), accumulate, rows, C, V,
                                    h2.stride(0), weight.stride(0), C, weight.stride(0), rb, vc, ws.data_ptr(), nbytes,
                                    _stream()), "vlk_lmhead_ce_bwd")
        return dh.view(ctx.h_shape) if dh is not None else None, dw_ret, None, None


def lmhead_ce(h, weight, labels, row_weight=None):
    return LMHeadCEFn.apply(h, weight, labels, row_weight)


@torch.no_grad()
def cross_entropy_rows(logits2d, labels):
    """Per-row cross-entropy (F.cross_entropy(..., reduction='none'), ignore_index rows -> 0) of bf16 logits
    [rows, V]; the logits are only read."""
    _need_cuda(logits2d, labels)
    lg = _rows(logits2d)
    labels = labels.reshape(-1).contiguous()
    rows, V = lg.shape
    out = torch.empty(rows, device=lg.device, dtype=torch.float32)
    one = torch.ones(1, device=lg.device, dtype=torch.float32)
    check(_lib.load().vlk_softmax_ce_rows(lg.data_ptr(), labels.data_ptr(), 0, out.data_ptr(), one.data_ptr(), rows, V,
                                          lg.stride(0), 0, _stream()), "vlk_softmax_ce_rows")
    return out


@torch.no_grad()
def lmhead_ce_rows(h, weight, labels):
    """Per-token losses of (h @ W^T) against labels without ever holding logits (evaluation paths:
    get_most_likely_row at train_gpt2.py:190-202, validation at gpt2_linear/train.py:218-252): the forward half of
    the fused lm_head + cross-entropy; ignored rows (-100) come back as 0."""
    _param_ok(weight)
    h2 = _rows(h)
    _, loss_row, _ = _lmhead_ce_forward(h2, weight, labels.reshape(-1).contiguous(), None)
    return loss_row
