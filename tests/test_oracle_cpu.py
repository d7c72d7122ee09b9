"""Pins the CPU oracle (oracle/torch_oracle.py) against golden vectors produced by the REAL reference modules
and HF CLIP (tests/golden/*.pt, see oracle/make_golden.py).  fp32 on CPU; tolerances are fp32 round-off."""
import os

import pytest
import torch

from oracle import torch_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name), map_location="cpu", weights_only=False)


def leafify(sd, names):
    sd = {k: v.clone() for k, v in sd.items()}
    for n in names:
        sd[n].requires_grad_(True)
    return sd


def close(a, b, tol=2e-4):
    denom = b.abs().max().item() + 1e-12
    assert (a - b).abs().max().item() / denom < tol, ((a - b).abs().max().item(), denom)


def test_gpt2_matches_reference_golden():
    g = load("gpt2_tiny.pt")
    names = [n for n in g["grads"] if n != "lm_head.weight"] + ["lm_head.weight"]
    sd = leafify(g["sd"], set(names))
    sd["transformer.wte.weight"] = sd["lm_head.weight"]  # tied
    logits, loss = O.gpt2_forward(sd, g["idx"], g["targets"], g["cfg"]["n_layer"], g["cfg"]["n_head"])
    close(logits.detach(), g["logits"])
    assert abs(loss.item() - g["loss"].item()) < 1e-5
    loss.backward()
    for n, ref in g["grads"].items():
        if n == "transformer.wte.weight":
            n = "lm_head.weight"
        close(sd[n].grad, ref, 5e-4)


def test_pool33_matches_reference_golden():
    g = load("caption_linear_tiny.pt")
    close(O.pool33(g["raw_tokens"]), g["pooled"], 1e-5)


def test_caption_linear_matches_reference_golden():
    g = load("caption_linear_tiny.pt")
    sd = leafify(g["sd"], g["grads"].keys())
    logits, loss = O.caption_linear_forward(sd, g["pooled"], g["input_ids"], g["labels"], g["cfg"]["n_layer"],
                                            g["cfg"]["n_head"])
    close(logits.detach(), g["logits"])
    assert abs(loss.item() - g["loss"].item()) < 1e-5
    loss.backward()
    for n, ref in g["grads"].items():
        close(sd[n].grad, ref, 5e-4)


def test_caption_qformer_matches_reference_golden():
    g = load("caption_qformer_tiny.pt")
    sd = leafify(g["sd"], g["grads"].keys())
    logits, loss = O.caption_qformer_forward(sd, g["pooled"], g["input_ids"], g["labels"], g["cfg"]["n_layer"],
                                             g["cfg"]["n_head"])
    close(logits.detach(), g["logits"])
    assert abs(loss.item() - g["loss"].item()) < 1e-5
    loss.backward()
    for n, ref in g["grads"].items():
        close(sd[n].grad, ref, 5e-4)


def test_caption_xattn_matches_reference_golden():
    g = load("caption_xattn_tiny.pt")
    sd = leafify(g["sd"], g["grads"].keys())
    logits, loss = O.xattn_forward(sd, g["idx"], g["pooled"], g["targets"], g["mask"], g["cfg"]["n_layer"],
                                   g["cfg"]["n_head"])
    close(logits.detach(), g["logits"])
    assert abs(loss.item() - g["loss"].item()) < 1e-5
    loss.backward()
    for n, ref in g["grads"].items():
        close(sd[n].grad, ref, 5e-4)


@pytest.mark.parametrize("case", ["cls_only", "patch_tokens_2d", "truncate_text", "ignored_sample"])
def test_caption_edge_cases_match_reference_golden(case):
    """use_cls_only, 2-D patch_tokens, text truncation at block_size and a fully ignored sample
    (source/gpt2_linear/model.py:178-196, 206-210) — outputs of the reference module itself."""
    g = load("caption_edge_tiny.pt")[case]
    sd = leafify(g["sd"], g["grads"].keys())
    logits, loss = O.caption_linear_forward(sd, g["z"], g["input_ids"], g["labels"], g["cfg"]["n_layer"],
                                            g["cfg"]["n_head"], use_cls_only=g["kw"].get("use_cls_only", False),
                                            block_size=g["cfg"]["block_size"])
    assert logits.shape == g["logits"].shape
    close(logits.detach(), g["logits"])
    assert abs(loss.item() - g["loss"].item()) < 1e-5
    loss.backward()
    for n, ref in g["grads"].items():
        close(sd[n].grad, ref, 5e-4)


@pytest.mark.parametrize("case", ["xattn_masked_sample", "xattn_all_masked"])
def test_xattn_masked_loss_edge_cases_match_reference_golden(case):
    """Masked-mean CE with a fully masked sample, and with every token masked: 0 / clamp_min(1) = 0 with zero
    gradients (source/gpt2_cross-att/model.py:176-185)."""
    g = load("caption_edge_tiny.pt")[case]
    sd = leafify(g["sd"], g["grads"].keys())
    logits, loss = O.xattn_forward(sd, g["idx"], g["z"], g["targets"], g["mask"], g["cfg"]["n_layer"], g["cfg"]["n_head"])
    close(logits.detach(), g["logits"])
    assert abs(loss.item() - g["loss"].item()) < 1e-5
    loss.backward()
    for n, ref in g["grads"].items():
        if ref.abs().max() == 0:
            assert sd[n].grad is None or sd[n].grad.abs().max() == 0
        else:
            close(sd[n].grad, ref, 5e-4)


def test_clip_matches_hf_golden():
    g = load("clip_tiny.pt")
    px = g["pixels"].float()
    hidden = O.clip_features(g["sd"], px, g["cfg"]["n_layer"], g["cfg"]["n_head"], return_hidden=True)
    close(hidden, g["hidden"], 5e-4)
    close(O.clip_features(g["sd"], px, g["cfg"]["n_layer"], g["cfg"]["n_head"]), g["feats"], 5e-4)


def test_adamw_clip_matches_torch_golden():
    g = load("adamw_steps.pt")
    ps = [p.clone() for p in g["p0"]]
    ms = [torch.zeros_like(p) for p in ps]
    vs = [torch.zeros_like(p) for p in ps]
    for step, grads in enumerate(g["grads"], start=1):
        norm = O.clip_and_adamw(ps, [x.clone() for x in grads], ms, vs, step, g["lr"], g["wds"])
        assert abs(norm.item() - g["norms"][step - 1].item()) < 1e-4 * g["norms"][step - 1].item()
    for p, ref in zip(ps, g["p_final"]):
        close(p, ref, 1e-5)


@pytest.mark.skipif(not os.path.isdir("/root/reference/source"), reason="reference only exists in the build container")
def test_oracle_against_live_reference_full_width():
    """Same check at the real width (768/12 heads, 2 layers to keep it quick) against the live reference module."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_linear_live", "/root/reference/source/gpt2_linear/model.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(7)
    lm = ref.GPT_previous(ref.GPTConfig(block_size=128, vocab_size=1024, n_layer=2, n_head=12, n_embd=768))
    cap = ref.GPT_Caption(enc_dim=768, lm=lm, m_vis_tokens=32)
    z = ref.pool_clip_197_to_33_avg_with_cls(torch.randn(2, 257, 768))
    x, _, _, labels = O.synthetic_caption_batch(2, seed=3, vocab=1024, eot=1023)
    logits, loss = cap(z, x, labels=labels)
    sd = {k: v.detach() for k, v in cap.state_dict().items()}
    lo, ls = O.caption_linear_forward(sd, z, x, labels, 2, 12)
    close(lo, logits.detach())
    assert abs(ls.item() - loss.item()) < 1e-5


def test_host_caption_encoding_matches_oracle_and_reference_rule():
    """gpt2_vision_language_b200.data (host input formatting) vs the oracle's generator and the reference rule
    (source/gpt2_linear/data.py:35-49): empty caption, short caption, caption longer than max_len."""
    from gpt2_vision_language_b200 import data
    a = data.synthetic_caption_batch(16, seed=3)
    b = O.synthetic_caption_batch(16, seed=3)
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    x, y, m = data.encode_caption_ids([], max_len=8)
    assert x.tolist() == [data.EOT] * 7 and y.tolist() == [data.EOT] * 7 and m.tolist() == [True] + [False] * 6
    x, y, m = data.encode_caption_ids([5, 6, 7], max_len=8)
    assert x.tolist() == [5, 6, 7] + [data.EOT] * 4 and y.tolist() == [6, 7] + [data.EOT] * 5
    assert m.tolist() == [True] * 3 + [False] * 4
    x, y, m = data.encode_caption_ids(list(range(100, 120)), max_len=8)
    assert x.tolist() == list(range(100, 107)) and y.tolist() == list(range(101, 107)) + [data.EOT] and m.all()
