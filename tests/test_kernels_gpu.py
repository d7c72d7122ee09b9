"""libvlk memory-bound + attention kernels through the C ABI vs torch fp32 on the same (bf16-rounded) inputs."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relerr(a, b):
    # relative to the reference's magnitude, with an absolute floor for references that are exactly zero
    return (a.float() - b.float()).abs().max().item() / max(b.float().abs().max().item(), 1e-3)


@pytest.mark.parametrize("rows,cols", [(4096, 768), (16448, 1024), (37, 768), (5, 128), (2112, 2048)])
def test_layernorm_fwd_bwd(cuda, rows, cols):
    from gpt2_vision_language_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + cols)
    x = (torch.randn(rows, cols, device=cuda, generator=g) * 2 + 0.3).bfloat16()
    w = (1 + 0.1 * torch.randn(cols, device=cuda, generator=g)).bfloat16()
    b = (0.1 * torch.randn(cols, device=cuda, generator=g)).bfloat16()
    dy = torch.randn(rows, cols, device=cuda, generator=g).bfloat16()
    y, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-5)
    xr = x.float().requires_grad_(True)
    wr, br = w.float().requires_grad_(True), b.float().requires_grad_(True)
    yr = F.layer_norm(xr, (cols,), wr, br, 1e-5)
    assert relerr(y, yr) < 1e-2
    assert relerr(mean, x.float().mean(-1)) < 1e-4
    yr.backward(dy.float())
    dx, dg, db = ops.layernorm_bwd(dy, x, w, mean, rstd, param_grads=True)
    assert relerr(dx, xr.grad) < 1e-2
    assert relerr(dg, wr.grad) < 2e-3
    assert relerr(db, br.grad) < 2e-3
    # frozen-norm variant + accumulation into an existing gradient
    base = torch.randn(rows, cols, device=cuda, generator=g).bfloat16()
    acc = base.clone()
    ops.layernorm_bwd(dy, x, w, mean, rstd, param_grads=False, dx=acc, accumulate=True)
    assert relerr(acc, base.float() + xr.grad) < 1.5e-2


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_pool33_vs_reference_semantics(cuda, dtype):
    from gpt2_vision_language_b200.caption import pool_clip_197_to_33_avg_with_cls
    from oracle import torch_oracle as O
    g = torch.load(os.path.join(GOLD, "caption_linear_tiny.pt"), weights_only=False)
    out = pool_clip_197_to_33_avg_with_cls(g["raw_tokens"].to(cuda).to(dtype))
    assert out.shape == (3, 33, 64) and out.dtype == dtype
    ref = O.pool33(g["raw_tokens"].to(dtype).float())
    assert relerr(out.cpu(), ref) < (1e-5 if dtype == torch.float32 else 1e-2)
    if dtype == torch.float32:  # golden vector from the reference function itself
        assert relerr(out.cpu(), g["pooled"]) < 1e-5
    big = torch.randn(64, 257, 768, device=cuda).to(dtype)
    ob = pool_clip_197_to_33_avg_with_cls(big)
    assert relerr(ob.float().norm(dim=-1), torch.ones(64, 33, device=cuda)) < 1e-2
    assert relerr(ob.cpu(), O.pool33(big.float().cpu())) < 1e-2


def test_embed_concat(cuda):
    from gpt2_vision_language_b200 import ops
    wte = torch.randn(512, 768, device=cuda).bfloat16()
    wpe = torch.randn(64, 768, device=cuda).bfloat16()
    ids = torch.randint(0, 512, (5, 31), device=cuda)
    prefix = torch.randn(5, 33, 768, device=cuda).bfloat16()
    out = ops.embed_concat(ids, wte, wpe, prefix)
    ref = torch.cat([prefix.float(), wte.float()[ids] + wpe.float()[:31]], 1)
    assert relerr(out, ref) < 1e-2
    out2 = ops.embed_concat(ids, wte, wpe)
    assert relerr(out2, wte.float()[ids] + wpe.float()[:31]) < 1e-2


_GUARD = 8192


def _guarded(like):
    """A contiguous bf16 tensor shaped like `like` inside a flat buffer with _GUARD sentinel elements on either side."""
    n = like.numel()
    buf = torch.full((n + 2 * _GUARD,), -77.0, device=like.device, dtype=torch.bfloat16)
    return buf, buf[_GUARD:_GUARD + n].view(like.shape)


def _attn_ref(q, k, v, H, causal):
    from oracle import torch_oracle as O
    return O.sdpa(q, k, v, H, causal)


@pytest.mark.parametrize("B,H,Tq,Tk,causal", [(2, 12, 64, 64, True), (3, 12, 63, 63, True), (2, 12, 31, 33, False),
                                              (2, 12, 32, 32, False), (2, 16, 257, 257, False), (1, 2, 300, 300, True),
                                              (2, 12, 1, 1, True), (1, 12, 130, 130, True),
                                              # head-pair tcgen05 kernels: odd number of (batch, head) problems, ragged
                                              (1, 3, 40, 40, True), (3, 1, 17, 64, False), (5, 3, 64, 9, False),
                                              (1, 1, 8, 8, True), (7, 12, 33, 33, True)])
def test_attention_fwd_bwd(cuda, B, H, Tq, Tk, causal):
    from gpt2_vision_language_b200 import ops
    C = H * 64
    g = torch.Generator(device="cuda").manual_seed(Tq * 13 + Tk)
    if Tq == Tk:   # packed qkv, consumed in place
        qkv = torch.randn(B, Tq, 3 * C, device=cuda, generator=g).bfloat16()
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    else:
        q = torch.randn(B, Tq, C, device=cuda, generator=g).bfloat16()
        kv = torch.randn(B, Tk, 2 * C, device=cuda, generator=g).bfloat16()
        k, v = kv[..., :C], kv[..., C:]
    o, lse = ops.attention_fwd(q, k, v, H, causal)
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    ref = _attn_ref(qr, kr, vr, H, causal)
    assert relerr(o, ref) < 1.5e-2
    d_o = torch.randn(B, Tq, C, device=cuda, generator=g).bfloat16()
    ref.backward(d_o.float())
    # the outputs sit between guard zones of a sentinel value: the kernels store rows through pointer arithmetic relative
    # to rows that may lie past the sequence (and, streaming kernels, through transposed lane quads) — nothing may land
    # outside the three gradients
    (bq, dq), (bk, dk), (bv, dv) = _guarded(q), _guarded(k), _guarded(v)
    ops.attention_bwd(q, k, v, o, d_o, lse, dq, dk, dv, H, causal)
    for buf in (bq, bk, bv):
        assert bool((buf[:_GUARD] == -77.0).all()) and bool((buf[-_GUARD:] == -77.0).all())
    assert relerr(dq, qr.grad) < 2e-2
    assert relerr(dk, kr.grad) < 2e-2
    assert relerr(dv, vr.grad) < 2e-2


@pytest.mark.parametrize("B,H,Tq,Tk,causal,p", [(4, 12, 64, 64, True, 0.0), (3, 3, 32, 33, False, 0.0),
                                                (1, 5, 31, 33, False, 0.0), (4, 12, 32, 32, False, 0.1),
                                                (3, 5, 32, 33, False, 0.25)])
def test_pair_attention_odd_head_counts_and_dropout_determinism(cuda, B, H, Tq, Tk, causal, p):
    """The two-heads-per-CTA tcgen05 kernels (attention_pair.cu) on packed K/V views with an ODD number of (batch, head)
    problems and Tq != Tk: against torch fp32 without dropout; with dropout the forward/backward pair is
    reproducible for one (seed, step, stream id) triple and changes with the stream id (the exact-mask comparison
    with torch lives in tests/test_dropout_gpu.py)."""
    from gpt2_vision_language_b200 import ops
    C = H * 64
    g = torch.Generator(device="cuda").manual_seed(B * 100 + Tq)
    q = torch.randn(B, Tq, C, device=cuda, generator=g).bfloat16()
    kv = torch.randn(B, Tk, 2 * C, device=cuda, generator=g).bfloat16()
    k, v = kv[..., :C], kv[..., C:]
    d_o = torch.randn(B, Tq, C, device=cuda, generator=g).bfloat16()
    rng = ops.DropoutState(cuda, seed=1234) if p > 0 else None

    def run(stream_id):
        o, lse = ops.attention_fwd(q, k, v, H, causal, dropout_p=p, rng=rng, stream_id=stream_id)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k.contiguous()), torch.empty_like(v.contiguous())
        ops.attention_bwd(q, k, v, o, d_o, lse, dq, dk, dv, H, causal, dropout_p=p, rng=rng, stream_id=stream_id)
        return o, lse, dq, dk, dv
    a = run(7)
    for t, name in zip(a, ("o", "lse", "dq", "dk", "dv")):
        assert torch.isfinite(t.float()).all(), name
    if p == 0:
        qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
        ref = _attn_ref(qr, kr, vr, H, causal)
        ref.backward(d_o.float())
        assert relerr(a[0], ref) < 2e-2
        for got, want, name in zip(a[2:], (qr.grad, kr.grad, vr.grad), ("dq", "dk", "dv")):
            assert relerr(got, want) < 2e-2, name
        return
    b = run(7)
    for x, y, name in zip(a, b, ("o", "lse", "dq", "dk", "dv")):
        assert torch.equal(x, y), name                       # same (seed, step, stream id) -> same mask, bit-identical
    c = run(8)
    assert not torch.equal(a[0], c[0])                       # another call site draws another mask


@pytest.mark.parametrize("B,H,Tq,Tk,causal", [(2, 12, 1024, 1024, True), (1, 4, 384, 384, True), (1, 2, 500, 500, False),
                                              (2, 3, 129, 300, False), (1, 2, 640, 640, True),
                                              # ragged causal (a half-empty last 128-block), causal with more keys than
                                              # queries (the diagonal is shifted by Tk - Tq), few queries / many keys
                                              (1, 2, 1000, 1000, True), (1, 2, 200, 456, True), (2, 2, 65, 1000, False),
                                              (1, 3, 321, 321, True)])
def test_flash_attention_long_sequences(cuda, B, H, Tq, Tk, causal):
    """Streaming tcgen05 forward/backward (GPT-2 pretraining shape T=1024 and ragged lengths) vs torch fp32."""
    from gpt2_vision_language_b200 import ops
    C = H * 64
    g = torch.Generator(device="cuda").manual_seed(Tq + 7 * Tk)
    if Tq == Tk:
        qkv = torch.randn(B, Tq, 3 * C, device=cuda, generator=g).bfloat16()
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    else:
        q = torch.randn(B, Tq, C, device=cuda, generator=g).bfloat16()
        kv = torch.randn(B, Tk, 2 * C, device=cuda, generator=g).bfloat16()
        k, v = kv[..., :C], kv[..., C:]
    o, lse = ops.attention_fwd(q, k, v, H, causal)
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    ref = _attn_ref(qr, kr, vr, H, causal)
    assert relerr(o, ref) < 1.5e-2
    d_o = torch.randn(B, Tq, C, device=cuda, generator=g).bfloat16()
    ref.backward(d_o.float())
    # the outputs sit between guard zones of a sentinel value: the kernels store rows through pointer arithmetic relative
    # to rows that may lie past the sequence (and, streaming kernels, through transposed lane quads) — nothing may land
    # outside the three gradients
    (bq, dq), (bk, dk), (bv, dv) = _guarded(q), _guarded(k), _guarded(v)
    ops.attention_bwd(q, k, v, o, d_o, lse, dq, dk, dv, H, causal)
    for buf in (bq, bk, bv):
        assert bool((buf[:_GUARD] == -77.0).all()) and bool((buf[-_GUARD:] == -77.0).all())
    assert relerr(dq, qr.grad) < 2e-2
    assert relerr(dk, kr.grad) < 2e-2
    assert relerr(dv, vr.grad) < 2e-2


@pytest.mark.parametrize("rows,V", [(64, 50304), (7, 512), (300, 1024), (1984, 50304), (4500, 50304), (130, 264)])
def test_lmhead_ce(cuda, rows, V):
    """Fused lm_head + cross-entropy (vlk_lmhead_ce_fwd / _bwd) vs torch fp32: loss, d h, d W — one vocabulary chunk and
    several, one row block and two (4500 rows > 4096), a ragged last tile (V = 264), ignored rows, an upstream scale,
    accumulation into an existing .grad, and the masked-mean (row_weight) variant."""
    from gpt2_vision_language_b200 import ops
    C = 768 if V > 1024 else 128
    g = torch.Generator(device="cuda").manual_seed(V + rows)
    h = torch.randn(rows, C, device=cuda, generator=g).bfloat16().requires_grad_(True)
    w = (torch.randn(V, C, device=cuda, generator=g) * 0.05).bfloat16().requires_grad_(True)
    labels = torch.randint(0, V, (rows,), device=cuda, generator=g)
    labels[::5] = -100
    loss = ops.lmhead_ce(h, w, labels)
    (loss * 0.5).backward()
    hr, wr = h.detach().float().requires_grad_(True), w.detach().float().requires_grad_(True)
    lr = F.cross_entropy(hr @ wr.t(), labels, ignore_index=-100)
    (lr * 0.5).backward()
    assert abs(loss.item() - lr.item()) / lr.item() < 1e-3
    assert F.cosine_similarity(h.grad.float().flatten(), hr.grad.flatten(), dim=0) > 0.9995
    assert F.cosine_similarity(w.grad.float().flatten(), wr.grad.flatten(), dim=0) > 0.9995
    assert relerr(h.grad, hr.grad) < 3e-2 and relerr(w.grad, wr.grad) < 3e-2
    assert h.grad[::5].float().abs().max().item() == 0.0              # ignored rows get exactly zero gradient
    # second backward ACCUMULATES into the existing .grad inside the GEMM (gradient accumulation, train_gpt2.py:458-469)
    g1 = w.grad.detach().float().clone()
    h.grad = None
    (ops.lmhead_ce(h, w, labels) * 0.5).backward()
    assert relerr(w.grad, 2.0 * g1) < 3e-2
    # masked-mean variant (x-attn loss, gpt2_cross-att/model.py:176-185): row_weight = mask, labels all valid
    mask = (labels >= 0)
    h2 = h.detach().clone().requires_grad_(True)
    loss2 = ops.lmhead_ce(h2, w.detach(), labels.clamp_min(0), mask)
    loss2.backward()
    assert abs(loss2.item() - lr.item()) / lr.item() < 1e-3
    assert F.cosine_similarity(h2.grad.float().flatten(), hr.grad.flatten(), dim=0) > 0.9995
    assert relerr(h2.grad, 2.0 * hr.grad) < 3e-2
    # per-token losses of the evaluation path come from the same forward
    per = ops.lmhead_ce_rows(h.detach(), w.detach(), labels)
    ref_rows = F.cross_entropy(hr.detach() @ wr.detach().t(), labels, ignore_index=-100, reduction="none")
    assert relerr(per, ref_rows) < 2e-3


def test_fused_clip_adamw_matches_torch_golden(cuda):
    from gpt2_vision_language_b200.optim import FusedAdamW
    g = torch.load(os.path.join(GOLD, "adamw_steps.pt"), weights_only=False)
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 3e-2)):
        ps = [torch.nn.Parameter(p.to(cuda).to(dtype)) for p in g["p0"]]
        opt = FusedAdamW([{"params": [ps[0]], "weight_decay": 0.1}, {"params": [ps[1]], "weight_decay": 0.0}],
                         lr=g["lr"], betas=(0.9, 0.95), eps=1e-8)
        for step, grads in enumerate(g["grads"]):
            for p, gr in zip(ps, grads):
                p.grad = gr.to(cuda).to(dtype)
            norm = opt.clip_grad_norm(1.0)
            opt.step()
            assert abs(norm.item() - g["norms"][step].item()) / g["norms"][step].item() < (1e-4 if dtype == torch.float32 else 1e-2)
        for p, ref in zip(ps, g["p_final"]):
            assert relerr(p.detach().cpu(), ref) < tol


def test_colsum_argmax_add(cuda):
    from gpt2_vision_language_b200 import ops
    x = torch.randn(2112, 768, device=cuda).bfloat16()
    assert relerr(ops.colsum(x), x.float().sum(0)) < 1e-3
    lg = torch.randn(9, 50304, device=cuda).bfloat16()
    assert torch.equal(ops.argmax_rows(lg), lg.float().argmax(-1))
    y = torch.randn(2112, 768, device=cuda).bfloat16()
    assert relerr(ops.add(x, y), x.float() + y.float()) < 1e-2
    assert relerr(ops.transpose(x), x.float().t()) == 0.0
