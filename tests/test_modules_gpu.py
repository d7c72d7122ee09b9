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


def test_smoke_entry(cuda):
    import __graft_entry__
    __graft_entry__.smoke()
