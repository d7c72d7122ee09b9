"""Generate tests/golden/*.pt from the REAL reference modules (/root/reference) and HF CLIP.

Run in the build container only (the reference is mounted read-only there; it does not exist on the GPU box):
    python oracle/make_golden.py
Each fixture holds: the (tiny) state_dict, the seeded inputs, and the reference's outputs — loss, logits and the
gradients of every trainable tensor — all fp32 CPU.  Tiny shapes keep head_dim = 64 (the kernels' specialisation).
TEST INFRASTRUCTURE ONLY.
"""
import importlib.util
import os
import sys
import types

import torch

REF = "/root/reference/source"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_train_gpt2_model_part():
    """train_gpt2.py is a module-level training script: exec only its model section (lines 1-145) with stubs for
    the un-vendored `hellaswag` import."""
    src = open(os.path.join(REF, "gpt2", "train_gpt2.py")).read().split("\n")[:145]
    for name in ("hellaswag", "tiktoken"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                stub = types.ModuleType(name)
                stub.render_example = stub.iterate_examples = None
                sys.modules[name] = stub
    ns = {"__name__": "ref_train_gpt2"}
    exec(compile("\n".join(src), "train_gpt2_model_part", "exec"), ns)
    return types.SimpleNamespace(**ns)


def grads_of(model):
    return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.requires_grad and p.grad is not None}


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(True)
    tiny = dict(block_size=64, vocab_size=256, n_layer=2, n_head=2, n_embd=128)

    # ---- plain GPT-2 (source/gpt2/train_gpt2.py) ----------------------------------------------------------
    ref = load_train_gpt2_model_part()
    torch.manual_seed(1337)
    m = ref.GPT(ref.GPTConfig(**tiny))
    idx = torch.randint(0, 256, (3, 24))
    tgt = torch.randint(0, 256, (3, 24))
    logits, loss = m(idx, tgt)
    loss.backward()
    torch.save(dict(cfg=tiny, sd={k: v.detach().clone() for k, v in m.state_dict().items()}, idx=idx, targets=tgt,
                    logits=logits.detach(), loss=loss.detach(), grads=grads_of(m)), os.path.join(OUT, "gpt2_tiny.pt"))
    print("gpt2_tiny loss", float(loss))

    # ---- linear captioner (source/gpt2_linear/model.py) ---------------------------------------------------
    lin = load_by_path("ref_linear", os.path.join(REF, "gpt2_linear", "model.py"))
    torch.manual_seed(1338)
    lm = lin.GPT_previous(lin.GPTConfig(**tiny))
    cap = lin.GPT_Caption(enc_dim=64, lm=lm, m_vis_tokens=32)
    raw = torch.randn(3, 257, 64)
    z = lin.pool_clip_197_to_33_avg_with_cls(raw)
    x = torch.randint(0, 256, (3, 15))
    labels = torch.randint(0, 256, (3, 15))
    labels[0, 9:] = -100
    labels[2, 4:] = -100
    logits, loss = cap(z, x, labels=labels)
    loss.backward()
    torch.save(dict(cfg=tiny, sd={k: v.detach().clone() for k, v in cap.state_dict().items()}, raw_tokens=raw,
                    pooled=z, input_ids=x, labels=labels, logits=logits.detach(), loss=loss.detach(),
                    grads=grads_of(cap)), os.path.join(OUT, "caption_linear_tiny.pt"))
    print("caption_linear_tiny loss", float(loss))

    # ---- Q-Former captioner (source/gpt2_q_former/model.py), eval mode = dropout off ----------------------
    qf = load_by_path("ref_qformer", os.path.join(REF, "gpt2_q_former", "model.py"))
    torch.manual_seed(1339)
    lm = qf.GPT_previous(qf.GPTConfig(**tiny))
    cap = qf.GPT_Caption(enc_dim=64, lm=lm, m_vis_tokens=8)
    cap.eval()
    logits, loss = cap(z, x, labels=labels)
    loss.backward()
    torch.save(dict(cfg=tiny, sd={k: v.detach().clone() for k, v in cap.state_dict().items()}, pooled=z,
                    input_ids=x, labels=labels, logits=logits.detach(), loss=loss.detach(), grads=grads_of(cap)),
               os.path.join(OUT, "caption_qformer_tiny.pt"))
    print("caption_qformer_tiny loss", float(loss))

    # ---- cross-attention captioner (source/gpt2_cross-att/model.py), gates made non-zero -------------------
    xa = load_by_path("ref_xattn", os.path.join(REF, "gpt2_cross-att", "model.py"))
    torch.manual_seed(1340)
    xm = xa.GPT(xa.GPTConfig(img_embd=64, **tiny))
    with torch.no_grad():
        for blk in xm.transformer.h:
            blk.cross_gate.copy_(torch.randn(()) * 0.5)
    mask = labels != -100
    tg = labels.clamp_min(0)
    logits, loss = xm(x, z=z, targets=tg, target_mask=mask)
    loss.backward()
    torch.save(dict(cfg=dict(img_embd=64, **tiny), sd={k: v.detach().clone() for k, v in xm.state_dict().items()},
                    pooled=z, idx=x, targets=tg, mask=mask, logits=logits.detach(), loss=loss.detach(),
                    grads=grads_of(xm)), os.path.join(OUT, "caption_xattn_tiny.pt"))
    print("caption_xattn_tiny loss", float(loss))

    # ---- CLIP vision tower (HF transformers), tiny width, real 224px / patch-14 geometry -------------------
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    torch.manual_seed(1341)
    cfg = CLIPVisionConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                           patch_size=14, image_size=224, projection_dim=64, hidden_act="quick_gelu",
                           layer_norm_eps=1e-5)
    hf = CLIPVisionModelWithProjection(cfg).eval()
    px = torch.randn(2, 3, 224, 224)
    with torch.no_grad():
        out = hf.vision_model(pixel_values=px)
        hidden = out.last_hidden_state
        feats = hf.visual_projection(hf.vision_model.post_layernorm(hidden))
    torch.save(dict(cfg=dict(n_layer=2, n_head=2), sd={k: v.detach().clone() for k, v in hf.state_dict().items()},
                    pixels=px.half(), hidden=hidden, feats=feats), os.path.join(OUT, "clip_tiny.pt"))
    print("clip_tiny feats", feats.shape, float(feats.abs().mean()))

    # ---- one clip_grad_norm_ + AdamW step (torch reference implementation) ---------------------------------
    torch.manual_seed(1342)
    ps = [torch.nn.Parameter(torch.randn(37, 16)), torch.nn.Parameter(torch.randn(50))]
    opt = torch.optim.AdamW([{"params": [ps[0]], "weight_decay": 0.1}, {"params": [ps[1]], "weight_decay": 0.0}],
                            lr=1e-2, betas=(0.9, 0.95), eps=1e-8)
    p0 = [p.detach().clone() for p in ps]
    gs, norms = [], []
    for step in range(3):
        for p in ps:
            p.grad = torch.randn_like(p) * (3.0 if step == 0 else 0.05)
        gs.append([p.grad.clone() for p in ps])
        norms.append(torch.nn.utils.clip_grad_norm_(ps, 1.0).clone())
        opt.step()
    torch.save(dict(p0=p0, grads=gs, norms=norms, p_final=[p.detach().clone() for p in ps], lr=1e-2,
                    wds=[0.1, 0.0]), os.path.join(OUT, "adamw_steps.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
