"""Frozen CLIP ViT-L/14 vision tower (forward only) on libvlk kernels.

The reference trains from PRE-COMPUTED CLIP token features ([257, 768] per image, source/gpt2_linear/data.py:55-63)
and does not ship the extractor ("imported directly from Hugging Face", README.md:21).  BASELINE.json's north_star
puts that forward on the hot path, so this module reproduces HF ``CLIPVisionModelWithProjection``
(transformers/models/clip/modeling_clip.py: embeddings :138-219, encoder layer :282-385, vision transformer
:647-696, visual_projection :1026) with per-token features = visual_projection(post_layernorm(last_hidden_state)).

Layout choices for B200: q/k/v weights are fused into one [3072,1024] GEMM whose packed output feeds the
attention kernel in place; the 14x14/14 patch convolution is an im2col + GEMM with K padded 588 -> 640
(16-byte TMA rows); bias, quick-GELU and both residual adds live in GEMM epilogues; weights are bf16.

The tower is frozen, so every LayerNorm that feeds a Linear (layer_norm1 -> q/k/v, layer_norm2 -> fc1,
post_layernorm -> visual_projection: 49 of the 50 norms) is FOLDED into that Linear's weights once
(``ops.fold_layernorm``): the GEMM runs on the raw residual stream, the per-row mean / rstd come from a
statistics-only kernel — or, from the second norm on, from row sums that the residual GEMM which wrote the
stream accumulated in its own epilogue — and are applied in the GEMM epilogue.  No normalised copy of the
[B*257, 1024] activations is ever written or re-read (67 MB of HBM traffic per norm at B=64).
``VLK_CLIP_NO_LNFOLD=1`` restores the LayerNorm kernel + plain GEMM pair, ``VLK_CLIP_NO_FUSED_STATS=1`` the
statistics kernel in front of every folded product (cross-checks).
"""
import os

import torch
import torch.nn as nn

from . import ops

KPAD = 640


class ClipVisionTower(nn.Module):
    def __init__(self, hidden=1024, layers=24, heads=16, intermediate=4096, proj_dim=768, eps=1e-5):
        super().__init__()
        self.hidden, self.n_layers, self.heads, self.inter, self.proj_dim, self.eps = hidden, layers, heads, intermediate, proj_dim, eps
        bf = torch.bfloat16

        def buf(name, *shape):
            self.register_buffer(name, torch.zeros(*shape, dtype=bf))

        buf("patch_w", hidden, KPAD)
        buf("cls", hidden)
        buf("pos", 257, hidden)
        for n in ("pre_ln", "post_ln"):
            buf(n + "_w", hidden)
            buf(n + "_b", hidden)
        buf("proj_w", proj_dim, hidden)
        L = layers
        buf("ln1_w", L, hidden); buf("ln1_b", L, hidden); buf("ln2_w", L, hidden); buf("ln2_b", L, hidden)
        buf("qkv_w", L, 3 * hidden, hidden); buf("qkv_b", L, 3 * hidden)
        buf("out_w", L, hidden, hidden); buf("out_b", L, hidden)
        buf("fc1_w", L, intermediate, hidden); buf("fc1_b", L, intermediate)
        buf("fc2_w", L, hidden, intermediate); buf("fc2_b", L, hidden)

    @classmethod
    def from_hf(cls, hf_model, device="cuda"):
        """Build from a ``transformers.CLIPVisionModelWithProjection`` (random-init or pretrained)."""
        cfg = hf_model.config
        assert cfg.patch_size == 14 and cfg.image_size == 224, "specialised for ViT-L/14 @ 224px"
        assert cfg.hidden_act == "quick_gelu"
        return cls.from_state_dict(hf_model.state_dict(), cfg.num_hidden_layers, cfg.num_attention_heads,
                                   cfg.layer_norm_eps, device)

    @classmethod
    def from_state_dict(cls, sd, layers=24, heads=16, eps=1e-5, device="cuda"):
        """Build from an HF-keyed CLIP vision state_dict (``vision_model.*`` + ``visual_projection.weight``)."""
        vm = "vision_model."
        hidden = sd[vm + "embeddings.class_embedding"].shape[0]
        inter = sd[vm + "encoder.layers.0.mlp.fc1.weight"].shape[0]
        proj = sd["visual_projection.weight"].shape[0]
        self = cls(hidden, layers, heads, inter, proj, eps).to(device)
        sd = {k: v.detach().to(device) for k, v in sd.items() if torch.is_tensor(v)}
        bf = torch.bfloat16
        with torch.no_grad():
            self.patch_w[:, :588] = sd[vm + "embeddings.patch_embedding.weight"].reshape(hidden, 588).to(bf)
            self.cls.copy_(sd[vm + "embeddings.class_embedding"])
            self.pos.copy_(sd[vm + "embeddings.position_embedding.weight"])
            self.pre_ln_w.copy_(sd[vm + "pre_layrnorm.weight"]); self.pre_ln_b.copy_(sd[vm + "pre_layrnorm.bias"])
            self.post_ln_w.copy_(sd[vm + "post_layernorm.weight"]); self.post_ln_b.copy_(sd[vm + "post_layernorm.bias"])
            self.proj_w.copy_(sd["visual_projection.weight"])
            for i in range(layers):
                p = f"{vm}encoder.layers.{i}."
                self.ln1_w[i] = sd[p + "layer_norm1.weight"]; self.ln1_b[i] = sd[p + "layer_norm1.bias"]
                self.ln2_w[i] = sd[p + "layer_norm2.weight"]; self.ln2_b[i] = sd[p + "layer_norm2.bias"]
                self.qkv_w[i] = torch.cat([sd[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], 0)
                self.qkv_b[i] = torch.cat([sd[p + f"self_attn.{n}_proj.bias"] for n in "qkv"], 0)
                self.out_w[i] = sd[p + "self_attn.out_proj.weight"]; self.out_b[i] = sd[p + "self_attn.out_proj.bias"]
                self.fc1_w[i] = sd[p + "mlp.fc1.weight"]; self.fc1_b[i] = sd[p + "mlp.fc1.bias"]
                self.fc2_w[i] = sd[p + "mlp.fc2.weight"]; self.fc2_b[i] = sd[p + "mlp.fc2.bias"]
        return self

    @staticmethod
    def random_state_dict(seed=0, hidden=1024, layers=24, intermediate=4096, proj_dim=768, device="cpu",
                          dtype=torch.float32):
        """HF-keyed random-init ViT-L/14 vision weights (there is no network for checkpoints): N(0, 0.02)
        matrices and embeddings, LayerNorm gains 1 + N(0, 0.02).  Both this tower and the CPU oracle
        (oracle.torch_oracle.clip_features) consume the same dict."""
        g = torch.Generator(device=device).manual_seed(seed)

        def rn(*shape, std=0.02, mean=0.0):
            return (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * std + mean).to(dtype)

        vm = "vision_model."
        sd = {vm + "embeddings.class_embedding": rn(hidden),
              vm + "embeddings.patch_embedding.weight": rn(hidden, 3, 14, 14),
              vm + "embeddings.position_embedding.weight": rn(257, hidden),
              vm + "pre_layrnorm.weight": rn(hidden, mean=1.0), vm + "pre_layrnorm.bias": rn(hidden),
              vm + "post_layernorm.weight": rn(hidden, mean=1.0), vm + "post_layernorm.bias": rn(hidden),
              "visual_projection.weight": rn(proj_dim, hidden)}
        for i in range(layers):
            p = f"{vm}encoder.layers.{i}."
            for n in "qkv":
                sd[p + f"self_attn.{n}_proj.weight"] = rn(hidden, hidden)
                sd[p + f"self_attn.{n}_proj.bias"] = rn(hidden)
            sd[p + "self_attn.out_proj.weight"] = rn(hidden, hidden); sd[p + "self_attn.out_proj.bias"] = rn(hidden)
            sd[p + "layer_norm1.weight"] = rn(hidden, mean=1.0); sd[p + "layer_norm1.bias"] = rn(hidden)
            sd[p + "layer_norm2.weight"] = rn(hidden, mean=1.0); sd[p + "layer_norm2.bias"] = rn(hidden)
            sd[p + "mlp.fc1.weight"] = rn(intermediate, hidden); sd[p + "mlp.fc1.bias"] = rn(intermediate)
            sd[p + "mlp.fc2.weight"] = rn(hidden, intermediate); sd[p + "mlp.fc2.bias"] = rn(hidden)
        return sd

    @torch.no_grad()
    def _fold(self):
        """(Re)build the LayerNorm-folded weights from the current buffers."""
        q, f = [], []
        for i in range(self.n_layers):
            q.append(ops.fold_layernorm(self.qkv_w[i], self.qkv_b[i], self.ln1_w[i], self.ln1_b[i]))
            f.append(ops.fold_layernorm(self.fc1_w[i], self.fc1_b[i], self.ln2_w[i], self.ln2_b[i]))
        self._fq, self._ff = q, f
        self._fp = ops.fold_layernorm(self.proj_w, None, self.post_ln_w, self.post_ln_b)
        self._fold_key = self.qkv_w._version

    def _folded(self):
        if os.environ.get("VLK_CLIP_NO_LNFOLD"):
            return False
        if getattr(self, "_fq", None) is None or self._fold_key != self.qkv_w._version \
                or self._fq[0][0].device != self.qkv_w.device:
            self._fold()
        return True

    @torch.no_grad()
    def hidden_states(self, pixel_values):
        """last_hidden_state [B,257,hidden] (before post_layernorm)."""
        B = pixel_values.shape[0]
        H = self.hidden
        cols = ops.im2col_patch14(pixel_values, KPAD)
        patch = ops.gemm(cols, self.patch_w)
        x = ops.clip_assemble(patch, self.cls, self.pos, B).view(B * 257, H)
        x, _, _ = ops.layernorm_fwd(x, self.pre_ln_w, self.pre_ln_b, self.eps, save_stats=False)
        if self._folded():
            # the residual GEMMs (out_proj, fc2) also accumulate the row sums of what they write, so the next folded
            # product needs no pass over x at all; only the very first norm reads its input (vlk_row_stats)
            fused = H >= 96 and not os.environ.get("VLK_CLIP_NO_FUSED_STATS")
            sums = torch.zeros((2 * self.n_layers, B * 257, 2), device=x.device, dtype=torch.float32) if fused else None
            nxt = None                                   # statistics of the current x, when a GEMM produced them
            for i in range(self.n_layers):
                wq, cq, bq = self._fq[i]
                if nxt is None:
                    qkv = ops.gemm_lnfold(x, wq, bq, cq, self.eps, stats=ops.row_stats(x, self.eps))
                else:
                    qkv = ops.gemm_lnfold(x, wq, bq, cq, self.eps, sums=nxt)
                qkv = qkv.view(B, 257, 3 * H)
                a, _ = ops.attention_fwd(qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:], self.heads, False,
                                         need_lse=False)
                wf, cf, bf_ = self._ff[i]
                if fused:
                    x = ops.gemm_stats(a.view(B * 257, H), self.out_w[i], self.out_b[i], x, sums[2 * i])
                    f = ops.gemm_lnfold(x, wf, bf_, cf, self.eps, act="quick_gelu", sums=sums[2 * i])
                    x = ops.gemm_stats(f, self.fc2_w[i], self.fc2_b[i], x, sums[2 * i + 1])
                    nxt = sums[2 * i + 1]
                else:
                    x = ops.gemm(a.view(B * 257, H), self.out_w[i], bias=self.out_b[i], residual=x)
                    f = ops.gemm_lnfold(x, wf, bf_, cf, self.eps, act="quick_gelu", stats=ops.row_stats(x, self.eps))
                    x = ops.gemm(f, self.fc2_w[i], bias=self.fc2_b[i], residual=x)
            self._last_sums = nxt
            return x.view(B, 257, H)
        for i in range(self.n_layers):
            h, _, _ = ops.layernorm_fwd(x, self.ln1_w[i], self.ln1_b[i], self.eps, save_stats=False)
            qkv = ops.gemm(h, self.qkv_w[i], bias=self.qkv_b[i]).view(B, 257, 3 * H)
            a, _ = ops.attention_fwd(qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:], self.heads, False,
                                     need_lse=False)
            x = ops.gemm(a.view(B * 257, H), self.out_w[i], bias=self.out_b[i], residual=x)
            h, _, _ = ops.layernorm_fwd(x, self.ln2_w[i], self.ln2_b[i], self.eps, save_stats=False)
            f = ops.gemm(h, self.fc1_w[i], bias=self.fc1_b[i], act="quick_gelu")
            x = ops.gemm(f, self.fc2_w[i], bias=self.fc2_b[i], residual=x)
        return x.view(B, 257, H)

    @torch.no_grad()
    def forward(self, pixel_values):
        """pixel_values [B,3,224,224] (fp32 or bf16) -> per-token features [B,257,proj_dim] (bf16)."""
        B = pixel_values.shape[0]
        x = self.hidden_states(pixel_values).view(B * 257, self.hidden)
        if self._folded():
            wp, cp, bp = self._fp
            if getattr(self, "_last_sums", None) is not None:
                y = ops.gemm_lnfold(x, wp, bp, cp, self.eps, sums=self._last_sums)
                self._last_sums = None
                return y.view(B, 257, self.proj_dim)
            return ops.gemm_lnfold(x, wp, bp, cp, self.eps, stats=ops.row_stats(x, self.eps)).view(B, 257, self.proj_dim)
        y, _, _ = ops.layernorm_fwd(x, self.post_ln_w, self.post_ln_b, self.eps, save_stats=False)
        return ops.gemm(y, self.proj_w).view(B, 257, self.proj_dim)
