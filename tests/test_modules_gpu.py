"""Drop-in modules on the B200 kernels vs the fp32 CPU oracle (pinned to the reference by test_oracle_cpu.py)
on identical weights and inputs.  Tolerances are BASELINE.json's: loss within 2e-3 relative, gradient cosine
>= 0.999 against fp32 (weights are rounded to bf16 first, SURVEY 8c pitfall 3)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name), map_location="cpu", weights_only=False)


def bf16_round(sd):
    return {k: v.to(torch.bfloat16).float() if v.is_floating_point() else v for k, v in sd.items()}


def cos(a, b):
    return F.cosine_similarity(a.float().flatten().cpu(), b.float().flatten().cpu(), dim=0).item()


def check_grads(model, ref_grads, min_cos=0.999):
    for n, p in model.named_parameters():
        if p.requires_grad and n in ref_grads:
            assert p.grad is not None, n
            if ref_grads[n].abs().max() < 1e-9:
                continue
            c = cos(p.grad, ref_grads[n])
            assert c > min_cos, (n, c)


def oracle_grads(sd, names, fn):
    sd = {k: v.clone() for k, v in sd.items()}
    for n in names:
        sd[n].requires_grad_(True)
    logits, loss = fn(sd)
    loss.backward()
    return logits.detach(), loss.detach(), {n: sd[n].grad for n in names}


def test_gpt2_tiny_vs_oracle(cuda):
    from gpt2_vision_language_b200 import gpt2
    from oracle import torch_oracle as O
    g = load("gpt2_tiny.pt")
    m = gpt2.GPT(gpt2.GPTConfig(**g["cfg"]))
    m.load_state_dict(g["sd"])               # reference state_dict keys load strictly
    m = m.to(cuda).to(torch.bfloat16)
    m.return_logits_with_loss = True
    sd = bf16_round(g["sd"])
    sd["transformer.wte.weight"] = sd["lm_head.weight"]
    names = [n for n, _ in m.named_parameters()]
    sd_o = dict(sd)
    lo, ls, gr = oracle_grads(sd_o, [n for n in names if n != "transformer.wte.weight"],
                              lambda s: O.gpt2_forward({**s, "transformer.wte.weight": s["lm_head.weight"]}, g["idx"],
                                                       g["targets"], g["cfg"]["n_layer"], g["cfg"]["n_head"]))
    logits, loss = m(g["idx"].to(cuda), g["targets"].to(cuda))
    loss.backward()
    assert abs(loss.item() - ls.item()) / ls.item() < 2e-3
    assert (logits.float().cpu() - lo).abs().max() < 0.05
    check_grads(m, gr, 0.995)   # full-model grads incl. tiny LN/bias grads; bridge-grad bar (0.999) is below
    big = [n for n in gr if gr[n].numel() > 10000]
    for n in big:
        assert cos(dict(m.named_parameters())[n].grad, gr[n]) > 0.999, n


def test_caption_linear_tiny_vs_oracle_and_golden(cuda):
    from gpt2_vision_language_b200 import gpt2, gpt2_linear
    from oracle import torch_oracle as O
    g = load("caption_linear_tiny.pt")
    m = gpt2_linear.GPT_Caption(enc_dim=64, lm=gpt2.GPT_previous(gpt2.GPTConfig(**g["cfg"])), m_vis_tokens=32)
    m.load_state_dict(g["sd"])
    m = m.to(cuda).to(torch.bfloat16)
    m.return_logits_with_loss = True
    sd = bf16_round(g["sd"])
    z = g["pooled"].to(torch.bfloat16)
    lo, ls, gr = oracle_grads(sd, list(g["grads"].keys()),
                              lambda s: O.caption_linear_forward(s, z.float(), g["input_ids"], g["labels"],
                                                                 g["cfg"]["n_layer"], g["cfg"]["n_head"]))
    logits, loss = m(z.to(cuda), g["input_ids"].to(cuda), labels=g["labels"].to(cuda))
    loss.backward()
    assert abs(loss.item() - ls.item()) / ls.item() < 2e-3
    assert abs(loss.item() - g["loss"].item()) / g["loss"].item() < 5e-3     # vs the reference's own fp32 run
    assert logits.shape == lo.shape
    check_grads(m, gr)
    assert all(p.grad is None for p in m.gpt.parameters())                   # frozen LM gets no gradients


def test_caption_qformer_tiny_vs_oracle(cuda):
    from gpt2_vision_language_b200 import gpt2, gpt2_q_former
    from oracle import torch_oracle as O
    g = load("caption_qformer_tiny.pt")
    m = gpt2_q_former.GPT_Caption(enc_dim=64, lm=gpt2.GPT_previous(gpt2.GPTConfig(**g["cfg"])), m_vis_tokens=8)
    m.load_state_dict(g["sd"])
    m = m.to(cuda).to(torch.bfloat16).eval()
    sd = bf16_round(g["sd"])
    z = g["pooled"].to(torch.bfloat16)
    lo, ls, gr = oracle_grads(sd, list(g["grads"].keys()),
                              lambda s: O.caption_qformer_forward(s, z.float(), g["input_ids"], g["labels"],
                                                                  g["cfg"]["n_layer"], g["cfg"]["n_head"]))
    _, loss = m(z.to(cuda), g["input_ids"].to(cuda), labels=g["labels"].to(cuda))
    loss.backward()
    assert abs(loss.item() - ls.item()) / ls.item() < 2e-3
    check_grads(m, gr, 0.998)
    flat = torch.cat([p.grad.float().flatten().cpu() for n, p in m.named_parameters() if p.requires_grad])
    ref = torch.cat([gr[n].flatten() for n, p in m.named_parameters() if p.requires_grad])
    assert cos(flat, ref) > 0.999


def test_caption_xattn_tiny_vs_oracle(cuda):
    from gpt2_vision_language_b200 import gpt2_cross_att as xa
    from oracle import torch_oracle as O
    g = load("caption_xattn_tiny.pt")
    m = xa.GPT(xa.GPTConfig(**g["cfg"]))
    m.load_state_dict(g["sd"])
    m = m.to(cuda).to(torch.bfloat16)
    sd = bf16_round(g["sd"])
    z = g["pooled"].to(torch.bfloat16)
    lo, ls, gr = oracle_grads(sd, list(g["grads"].keys()),
                              lambda s: O.xattn_forward(s, g["idx"], z.float(), g["targets"], g["mask"],
                                                        g["cfg"]["n_layer"], g["cfg"]["n_head"]))
    _, loss = m(g["idx"].to(cuda), z=z.to(cuda), targets=g["targets"].to(cuda), target_mask=g["mask"].to(cuda))
    loss.backward()
    assert abs(loss.item() - ls.item()) / ls.item() < 2e-3
    check_grads(m, gr, 0.998)
    flat = torch.cat([p.grad.float().flatten().cpu() for n, p in m.named_parameters() if p.requires_grad])
    ref = torch.cat([gr[n].flatten() for n, p in m.named_parameters() if p.requires_grad])
    assert cos(flat, ref) > 0.999


def test_clip_tower_vs_hf_golden(cuda):
    """Tiny-width CLIP with the real 224px/patch-14 geometry against HF's own forward (golden)."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    from gpt2_vision_language_b200.clip import ClipVisionTower
    g = load("clip_tiny.pt")
    cfg = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                           patch_size=14, image_size=224, projection_dim=64, hidden_act="quick_gelu",
                           layer_norm_eps=1e-5)
    hf = CLIPVisionModelWithProjection(cfg)
    hf.load_state_dict(g["sd"])
    tower = ClipVisionTower.from_hf(hf, device=cuda)
    feats = tower(g["pixels"].float().to(cuda))
    ref = g["feats"]
    err = (feats.float().cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 3e-2, err
    assert cos(feats, ref) > 0.9995


def test_greedy_decode_ids_match_oracle(cuda):
    """north_star: greedy-decoded caption token ids identical at >= 99 % of positions.  Random-init logits are
    nearly flat, so positions whose fp32 top-2 margin is below bf16 resolution are reported separately (SURVEY 8c
    pitfall 4); the bar applies to the decisive ones, and a sequence is followed only while it still agrees."""
    from gpt2_vision_language_b200 import gpt2, gpt2_linear
    from gpt2_vision_language_b200.decode import greedy_decode
    from oracle import torch_oracle as O
    g = load("caption_linear_tiny.pt")
    m = gpt2_linear.GPT_Caption(enc_dim=64, lm=gpt2.GPT_previous(gpt2.GPTConfig(**g["cfg"])), m_vis_tokens=32)
    m.load_state_dict(g["sd"])
    # sharpen the tiny model so that argmax is not a coin flip on flat logits
    with torch.no_grad():
        m.gpt.lm_head.weight.mul_(8.0)
    sd = bf16_round({k: v.detach().clone() for k, v in m.state_dict().items()})
    m = m.to(cuda).to(torch.bfloat16).eval()
    z = g["pooled"].to(torch.bfloat16)
    prompt = g["input_ids"][:, :3]
    ids = greedy_decode(m, z.to(cuda), prompt.to(cuda), max_new_tokens=24).cpu()
    ref, margins = O.greedy_decode(
        lambda x: O.caption_linear_forward(sd, z.float(), x, None, g["cfg"]["n_layer"], g["cfg"]["n_head"])[0], prompt, 24)
    assert ids.shape == ref.shape == (3, 27)
    new, new_ref = ids[:, 3:], ref[:, 3:]
    agree_prefix = (new == new_ref).long().cumprod(dim=1).bool()          # still on the oracle's trajectory
    comparable = torch.cat([torch.ones(3, 1, dtype=torch.bool), agree_prefix[:, :-1]], dim=1)
    decisive = comparable & (margins > 0.05)
    assert decisive.sum() >= 30, int(decisive.sum())
    rate = (new == new_ref)[decisive].float().mean().item()
    assert rate >= 0.99, rate


@pytest.mark.parametrize("kind", ["linear", "qformer", "xattn"])
def test_kv_cached_decode_matches_recompute(cuda, kind):
    """The KV-cached greedy decode (prefill + one row per step) against the reference-shaped loop that re-runs the
    whole captioner per token: same ids while the two trajectories agree, on decisive (non-tied) positions; and the
    hidden state of the first step agrees to bf16 round-off."""
    from gpt2_vision_language_b200 import gpt2, gpt2_linear, gpt2_q_former, gpt2_cross_att as xa
    from gpt2_vision_language_b200 import decode as D
    from gpt2_vision_language_b200 import ops
    if kind == "xattn":
        g = load("caption_xattn_tiny.pt")
        m = xa.GPT(xa.GPTConfig(**g["cfg"]))
        m.load_state_dict(g["sd"])
        with torch.no_grad():
            for i, blk in enumerate(m.transformer.h):
                blk.cross_gate.fill_(0.5 - 0.2 * i)      # gates are 0 at init: open them so the cached K/V matter
            m.lm_head.weight.mul_(8.0)
        prompt, dkind = g["idx"][:, :3], "xattn"
    else:
        g = load("caption_linear_tiny.pt" if kind == "linear" else "caption_qformer_tiny.pt")
        mod = gpt2_linear if kind == "linear" else gpt2_q_former
        m = mod.GPT_Caption(enc_dim=64, lm=gpt2.GPT_previous(gpt2.GPTConfig(**g["cfg"])),
                            m_vis_tokens=32 if kind == "linear" else 8)
        m.load_state_dict(g["sd"])
        with torch.no_grad():
            m.gpt.lm_head.weight.mul_(8.0)
        prompt, dkind = g["input_ids"][:, :3], "prefix"
    m = m.to(cuda).to(torch.bfloat16).eval()
    z = g["pooled"].to(torch.bfloat16).to(cuda)
    prompt = prompt.to(cuda)
    with torch.no_grad():      # first step: cached prefill hidden == recompute hidden (last row)
        dec = D._Decoder(m, dkind, z, n_text_max=8)
        h_c = dec.prefill(prompt.contiguous())
        h_r = (D._xattn_hidden(m, prompt, z) if dkind == "xattn" else D._prefix_hidden(m, prompt, z))[:, -1, :]
        assert cos(h_c, h_r) > 0.9995
        w = m.lm_head.weight if dkind == "xattn" else m.gpt.lm_head.weight
        # second step through the cache vs a recompute on the extended sequence
        nxt = dec.next_token(h_c)
        h_c2 = dec.step(nxt)
        ext = torch.cat([prompt, nxt[:, None]], dim=1)
        h_r2 = (D._xattn_hidden(m, ext, z) if dkind == "xattn" else D._prefix_hidden(m, ext, z))[:, -1, :]
        assert cos(h_c2, h_r2) > 0.9995
        margins = ops.gemm(h_r2.contiguous(), w).float().topk(2, dim=-1).values
    a = D.greedy_decode(m, z, prompt, max_new_tokens=24, kind=dkind).cpu()
    b = D.greedy_decode_recompute(m, z, prompt, max_new_tokens=24, kind=dkind).cpu()
    assert a.shape == b.shape == (prompt.shape[0], 27)
    assert torch.equal(a[:, :4], b[:, :4]) or (margins[:, 0] - margins[:, 1]).min() < 0.05
    agree = (a[:, 3:] == b[:, 3:]).long().cumprod(dim=1)
    # bf16 round-off may flip a near-tie; once a sequence diverges it is no longer comparable
    assert agree.float().mean().item() >= 0.9, agree


def test_plain_gpt_cached_decode_and_sampling(cuda):
    """kind='gpt': the KV-cached decode of the plain GPT-2 against argmax of a full forward on the same prefix, and the
    top-k sampler (train_gpt2.py:444-449) staying inside the top-k set of the full-forward logits."""
    from gpt2_vision_language_b200 import gpt2
    from gpt2_vision_language_b200 import decode as D
    g = load("gpt2_tiny.pt")
    m = gpt2.GPT(gpt2.GPTConfig(**g["cfg"]))
    m.load_state_dict(g["sd"])
    with torch.no_grad():
        m.lm_head.weight.mul_(8.0)
    m = m.to(cuda).to(torch.bfloat16).eval()
    prompt = g["idx"][:, :5].to(cuda)
    ids = D.greedy_decode(m, None, prompt, max_new_tokens=6, kind="gpt")
    assert ids.shape == (prompt.shape[0], 11) and torch.equal(ids[:, :5], prompt)
    with torch.no_grad():
        for t in range(5, 11):                       # teacher-forced check: each cached step = full forward argmax
            logits, _ = m(ids[:, :t])
            top2 = logits[:, -1].float().topk(2, dim=-1)
            decisive = (top2.values[:, 0] - top2.values[:, 1]) > 0.05
            assert torch.equal(ids[decisive, t], top2.indices[decisive, 0])
    gen = torch.Generator(device=cuda).manual_seed(42)
    smp = D.sample_decode(m, None, prompt, max_new_tokens=4, kind="gpt", top_k=5, generator=gen)
    with torch.no_grad():
        for t in range(5, 9):
            logits, _ = m(smp[:, :t])
            top = logits[:, -1].float().topk(8, dim=-1).indices       # slack of 3 for bf16 near-ties at the k-th place
            assert all(smp[b, t].item() in top[b].tolist() for b in range(smp.shape[0]))


def test_eval_helpers_match_torch(cuda):
    """get_most_likely_row (train_gpt2.py:190-202) with given logits and in its logits-free form, and per-row CE."""
    from gpt2_vision_language_b200 import gpt2, evaluate, ops
    g = load("gpt2_tiny.pt")
    m = gpt2.GPT(gpt2.GPTConfig(**g["cfg"]))
    m.load_state_dict(g["sd"])
    m = m.to(cuda).to(torch.bfloat16).eval()
    gen = torch.Generator().manual_seed(5)
    V = g["cfg"]["vocab_size"]
    tokens = torch.randint(0, V, (4, 24), generator=gen).to(cuda)
    mask = torch.zeros(4, 24, dtype=torch.long)
    mask[:, 10:20] = 1
    mask[2, 18:] = 0
    mask = mask.to(cuda)
    with torch.no_grad():
        logits, _ = m(tokens)
    lf = logits.float()
    ref_losses = F.cross_entropy(lf[:, :-1].reshape(-1, V), tokens[:, 1:].reshape(-1), reduction="none").view(4, -1)
    sm = mask[:, 1:].float()
    ref_avg = (ref_losses * sm).sum(1) / sm.sum(1)
    got = evaluate.masked_row_losses(tokens, mask, logits=logits)
    assert (got - ref_avg).abs().max().item() < 2e-3 * ref_avg.abs().max().item()
    assert evaluate.get_most_likely_row(tokens, mask, logits) == int(ref_avg.argmin())
    fused = evaluate.masked_row_losses(tokens, mask, hidden=m.trunk(ops.embed(tokens, m.transformer.wte.weight,
                                                                             m.transformer.wpe.weight)).detach(),
                                       lm_weight=m.lm_head.weight)
    assert (fused - ref_avg).abs().max().item() < 2e-3 * ref_avg.abs().max().item()
    assert evaluate.most_likely_row(m, tokens, mask) == int(ref_avg.argmin())
    lab = tokens[:, 1:].reshape(-1).clone()
    lab[::5] = -100
    rows = ops.cross_entropy_rows(logits[:, :-1].reshape(-1, V), lab)
    ref_rows = F.cross_entropy(lf[:, :-1].reshape(-1, V), lab, reduction="none", ignore_index=-100)
    assert (rows - ref_rows).abs().max().item() < 2e-3 * ref_rows.abs().max().item()


def test_smoke_entry(cuda):
    import __graft_entry__
    __graft_entry__.smoke()
