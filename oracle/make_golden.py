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


def edge_cases_and_loader():
    """Round-2 fixtures: the edge cases the reference forward handles (use_cls_only, 2-D patch_tokens, text
    truncation at block_size, a sample whose labels are all ignored, an all-masked batch in the cross-attention
    loss) and the first batches of the reference's own ``DataLoaderLite`` over a seeded shard directory."""
    import tempfile
    import numpy as np
    tiny = dict(block_size=64, vocab_size=256, n_layer=2, n_head=2, n_embd=128)
    lin = load_by_path("ref_linear_edge", os.path.join(REF, "gpt2_linear", "model.py"))
    cases = {}

    shared_sd = {}

    def snapshot(module):
        """state_dict copy that keeps tied tensors tied (one clone per storage), so torch.save stores them once."""
        memo, out = {}, {}
        for k, v in module.state_dict().items():
            if v.data_ptr() not in memo:
                memo[v.data_ptr()] = v.detach().clone()
            out[k] = memo[v.data_ptr()]
        return out

    def run_caption(name, cfg, z, x, labels, **kw):
        torch.manual_seed(2001)          # same weights for every case of one config: stored once
        lm = lin.GPT_previous(lin.GPTConfig(**cfg))
        cap = lin.GPT_Caption(enc_dim=64, lm=lm, m_vis_tokens=32, **kw)
        logits, loss = cap(z, x, labels=labels)
        loss.backward()
        sd = shared_sd.setdefault(cfg["block_size"], snapshot(cap))
        cases[name] = dict(cfg=cfg, kw=kw, sd=sd, z=z,
                           input_ids=x, labels=labels, logits=logits.detach(), loss=loss.detach(), grads=grads_of(cap))
        print(name, "loss", float(loss), "logits", tuple(logits.shape))

    g = torch.Generator().manual_seed(77)
    z = lin.pool_clip_197_to_33_avg_with_cls(torch.randn(3, 257, 64, generator=g))
    x = torch.randint(0, 256, (3, 15), generator=g)
    labels = torch.randint(0, 256, (3, 15), generator=g)
    labels[1, 6:] = -100
    # (1) use_cls_only=True: only the CLS token goes through the bridge (model.py:183-185) -> prefix of length 1
    run_caption("cls_only", tiny, z, x, labels, use_cls_only=True)
    # (2) 2-D patch_tokens [B, D] are unsqueezed to one visual token (model.py:178-179)
    run_caption("patch_tokens_2d", tiny, z[:, 0, :].contiguous(), x, labels)
    # (3) M + T > block_size: the TEXT is truncated to block_size - M tokens (model.py:189-196): 33 + 15 > 40 -> 7
    run_caption("truncate_text", dict(tiny, block_size=40), z, x, labels)
    # (4) one sample with every label ignored: it contributes nothing to the mean (ignore_index=-100, model.py:206-210)
    lab4 = labels.clone()
    lab4[2, :] = -100
    run_caption("ignored_sample", tiny, z, x, lab4)

    # (5) cross-attention loss with a fully masked sample, and (6) with EVERYTHING masked: loss = 0 / clamp_min(1)
    xa = load_by_path("ref_xattn_edge", os.path.join(REF, "gpt2_cross-att", "model.py"))
    for name, mask in (("xattn_masked_sample", None), ("xattn_all_masked", torch.zeros(3, 15, dtype=torch.bool))):
        torch.manual_seed(2100)
        xm = xa.GPT(xa.GPTConfig(img_embd=64, **tiny))
        with torch.no_grad():
            for blk in xm.transformer.h:
                blk.cross_gate.copy_(torch.randn(()) * 0.5)
        if mask is None:
            mask = labels != -100
            mask[0, :] = False
        tg = labels.clamp_min(0)
        logits, loss = xm(x, z=z, targets=tg, target_mask=mask)
        loss.backward()
        cases[name] = dict(cfg=dict(img_embd=64, **tiny), sd=shared_sd.setdefault("xattn", snapshot(xm)),
                           z=z, idx=x, targets=tg, mask=mask, logits=logits.detach(), loss=loss.detach(),
                           grads=grads_of(xm))
        print(name, "loss", float(loss))
    # logits are [3, <=48, 256] fp32: small; keep only what the tests read
    torch.save(cases, os.path.join(OUT, "caption_edge_tiny.pt"))

    # ---- DataLoaderLite (source/gpt2/train_gpt2.py:148-187), exec'd from the reference file itself ----------
    src = open(os.path.join(REF, "gpt2", "train_gpt2.py")).read().split("\n")[147:187]
    ns = {"np": np, "torch": torch, "os": os, "master_process": False}
    exec(compile("\n".join(src), "train_gpt2_loader_part", "exec"), ns)
    rng = np.random.default_rng(2024)
    shard_tokens = [rng.integers(0, 50257, size=n, dtype=np.uint16) for n in (1000, 777, 1200)]
    names = ["edufineweb_train_000001.npy", "edufineweb_train_000002.npy", "edufineweb_val_000000.npy"]
    with tempfile.TemporaryDirectory() as d:
        for n, t in zip(names, shard_tokens):
            np.save(os.path.join(d, n), t)
        old = os.environ.get("FW_OUT_DIR")
        os.environ["FW_OUT_DIR"] = d
        B, T, world, nb = 2, 8, 2, 70          # 70 batches: rolls over both train shards and wraps around
        out = {}
        for split in ("train", "val"):
            for rank in range(world):
                ld = ns["DataLoaderLite"](B, T, rank, world, split)
                xs, ys, shard_idx = [], [], []
                for _ in range(nb):
                    xb, yb = ld.next_batch()
                    xs.append(xb.clone()); ys.append(yb.clone()); shard_idx.append(ld.current_shard)
                out[(split, rank)] = dict(x=torch.stack(xs).to(torch.int32), y=torch.stack(ys).to(torch.int32),
                                          shard_after=torch.tensor(shard_idx))
        if old is None:
            del os.environ["FW_OUT_DIR"]
        else:
            os.environ["FW_OUT_DIR"] = old
    torch.save(dict(names=names, shards=[torch.from_numpy(t.astype(np.int32)) for t in shard_tokens], B=B, T=T,
                    world=world, batches=out), os.path.join(OUT, "dataloader_lite.pt"))
    for f in ("caption_edge_tiny.pt", "dataloader_lite.pt"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    if "--only-edge" not in sys.argv:
        main()
    edge_cases_and_loader()
