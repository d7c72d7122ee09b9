"""Dropout on the B200 path (Q-Former: nn.Dropout x3 + attention-probability dropout x2, p = 0.1,
source/gpt2_q_former/model.py:114-145).  Philox masks cannot be bit-matched with torch's, so the masks are
EXTRACTED from the kernels and the math is then checked exactly against torch with that explicit mask; plus
keep-rate statistics and determinism."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def test_dropout_add_mask_statistics_and_backward(cuda):
    from gpt2_vision_language_b200 import ops
    rng = ops.DropoutState(cuda, seed=1234)
    p = 0.1
    n = 1 << 20
    ones = torch.ones(n, device=cuda, dtype=torch.bfloat16)
    m1 = ops.dropout_add_raw(ones, None, p, rng, 7).float()
    vals = torch.unique(m1)
    assert vals.numel() == 2 and vals[0] == 0 and abs(vals[1].item() - 1 / (1 - p)) < 1e-2
    keep = (m1 > 0).float().mean().item()
    assert abs(keep - (1 - p)) < 4 * math.sqrt(p * (1 - p) / n)
    assert torch.equal(m1, ops.dropout_add_raw(ones, None, p, rng, 7).float())          # deterministic per stream id
    assert not torch.equal(m1, ops.dropout_add_raw(ones, None, p, rng, 8).float())      # new call site -> new mask
    rng.advance()
    assert not torch.equal(m1, ops.dropout_add_raw(ones, None, p, rng, 7).float())      # new step -> new mask
    # autograd: y = res + x*mask ; dx = dy*mask ; dres = dy
    x = torch.randn(64, 768, device=cuda).bfloat16().requires_grad_(True)
    res = torch.randn(64, 768, device=cuda).bfloat16().requires_grad_(True)
    y = ops.DropoutAddFn.apply(x, res, p, rng, 99)
    mask = (ops.dropout_add_raw(torch.ones_like(x), None, p, rng, 99).float() > 0).float() / (1 - p)
    assert torch.allclose(y.float(), res.float() + x.float() * mask, rtol=1e-2, atol=2e-2)
    dy = torch.randn_like(y)
    y.backward(dy)
    assert torch.allclose(x.grad.float(), dy.float() * mask, rtol=1e-2, atol=1e-3)
    assert torch.equal(res.grad, dy)


@pytest.mark.parametrize("Tq,Tk", [(32, 32), (32, 33), (64, 64)])
def test_attention_dropout_matches_torch_with_extracted_mask(cuda, Tq, Tk):
    from gpt2_vision_language_b200 import ops
    B, H, p = 3, 12, 0.1
    C = H * 64
    rng = ops.DropoutState(cuda, seed=77)
    sid = 5
    # mask extraction: q = k = 0 -> uniform probabilities 1/Tk; V = [I | 0] -> O[i, j] = mask_ij / Tk
    zq = torch.zeros(B, Tq, C, device=cuda, dtype=torch.bfloat16)
    zk = torch.zeros(B, Tk, C, device=cuda, dtype=torch.bfloat16)
    eye = torch.zeros(B, Tk, H, 64, device=cuda)
    eye[:, torch.arange(Tk), :, torch.arange(Tk)] = 1.0
    o, _ = ops.attention_fwd(zq, zk, eye.view(B, Tk, C).bfloat16(), H, False, dropout_p=p, rng=rng, stream_id=sid)
    mask = (o.float().view(B, Tq, H, 64)[..., :Tk] * Tk).permute(0, 2, 1, 3)          # [B,H,Tq,Tk], 0 or 1/(1-p)
    assert abs((mask > 0).float().mean().item() - (1 - p)) < 0.02
    assert (mask[mask > 0] - 1 / (1 - p)).abs().max() < 2e-2
    mask = (mask > 0).float() / (1 - p)
    # real inputs, same (seed, step, stream id) -> same mask
    g = torch.Generator(device="cuda").manual_seed(Tq * 100 + Tk)
    q = torch.randn(B, Tq, C, device=cuda, generator=g).bfloat16()
    kv = torch.randn(B, Tk, 2 * C, device=cuda, generator=g).bfloat16()
    k, v = kv[..., :C], kv[..., C:]
    o, lse = ops.attention_fwd(q, k, v, H, False, dropout_p=p, rng=rng, stream_id=sid)
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    qh = qr.view(B, Tq, H, 64).transpose(1, 2)
    kh = kr.view(B, Tk, H, 64).transpose(1, 2)
    vh = vr.view(B, Tk, H, 64).transpose(1, 2)
    P = torch.softmax(qh @ kh.transpose(-1, -2) / 8.0, dim=-1) * mask
    ref = (P @ vh).transpose(1, 2).reshape(B, Tq, C)
    assert (o.float() - ref).abs().max() / ref.abs().max() < 2e-2
    d_o = torch.randn(B, Tq, C, device=cuda, generator=g).bfloat16()
    ref.backward(d_o.float())
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k.contiguous()), torch.empty_like(v.contiguous())
    ops.attention_bwd(q, k, v, o, d_o, lse, dq, dk, dv, H, False, dropout_p=p, rng=rng, stream_id=sid)
    for got, want in ((dq, qr.grad), (dk, kr.grad), (dv, vr.grad)):
        assert (got.float() - want).abs().max() / want.abs().max() < 2.5e-2


def test_qformer_train_mode_dropout_runs_and_is_unbiased(cuda):
    from gpt2_vision_language_b200 import gpt2, gpt2_q_former, ops
    torch.manual_seed(0)
    cfg = gpt2.GPTConfig(block_size=64, vocab_size=256, n_layer=1, n_head=2, n_embd=128)
    m = gpt2_q_former.GPT_Caption(enc_dim=64, lm=gpt2.GPT_previous(cfg), m_vis_tokens=8).to(cuda).to(torch.bfloat16)
    z = F.normalize(torch.randn(4, 33, 64, device=cuda), dim=-1).bfloat16()
    x = torch.randint(0, 256, (4, 15), device=cuda)
    y = torch.randint(0, 256, (4, 15), device=cuda)
    m.eval()
    with torch.no_grad():
        ref = m.bridge(z).float()
    m.train()
    _, loss = m(z, x, labels=y)
    loss.backward()
    assert torch.isfinite(loss)
    assert all(torch.isfinite(p.grad).all() for p in m.bridge.parameters() if p.grad is not None)
    outs = []
    with torch.no_grad():
        for _ in range(48):
            outs.append(m.bridge(z).float())
    outs = torch.stack(outs)
    assert outs.std(0).mean() > 1e-3                                   # masks really differ from call to call
    # inverted dropout keeps the bridge output unbiased to first order: the mean over masks tracks eval mode
    err = (outs.mean(0) - ref).abs().mean() / ref.abs().mean()
    assert err < 0.1, err
