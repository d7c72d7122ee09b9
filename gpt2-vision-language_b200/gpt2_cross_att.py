"""Cross-attention captioner — drop-in for source/gpt2_cross-att/model.py (same public names).

Every GPT-2 block gets a trainable cross-attention to the 33 pooled CLIP tokens, gated by tanh(cross_gate)
(initialised to 0) and applied BEFORE the frozen self-attention (model.py:87-104).  Only vis_proj, the xattn
projections and the gates train (model.py:131-139).
"""
from dataclasses import dataclass

import torch
import torch.nn as nn

from . import ops
from .caption import pool_clip_197_to_33_avg_with_cls  # noqa: F401  (re-exported)
from .gpt2 import MLP, CausalSelfAttention, _check_config, _init_gpt_weights, build_adamw


@dataclass
class GPTConfig:
    block_size: int = 1024
    vocab_size: int = 50257
    n_layer: int = 12
    n_head: int = 12
    n_embd: int = 768
    img_embd: int = 768


class CrossAttention(nn.Module):
    """q_proj(x) attends (non-causally) to kv_proj(z); c_proj has the scaled residual init (model.py:34-58)."""

    def __init__(self, config):
        super().__init__()
        if config.n_embd % config.n_head != 0:
            raise AssertionError("n_embd must be divisible by n_head")
        self.n_head, self.n_embd = config.n_head, config.n_embd
        self.q_proj = nn.Linear(config.n_embd, config.n_embd)
        self.kv_proj = nn.Linear(config.n_embd, 2 * config.n_embd)
        self.c_proj = nn.Linear(config.n_embd, config.n_embd)
        self.c_proj.NANOGPT_SCALE_INIT = 1

    def attend(self, x, z):
        q = ops.linear(x, self.q_proj.weight, self.q_proj.bias)
        kv = ops.linear(z, self.kv_proj.weight, self.kv_proj.bias)
        return ops.cross_attention(q, kv, self.n_head)

    def forward(self, x, z):
        return ops.linear(self.attend(x, z), self.c_proj.weight, self.c_proj.bias)


class Vision_projector(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.z_proj = nn.Linear(config.img_embd, config.n_embd)

    def forward(self, z):
        return ops.linear(z, self.z_proj.weight, self.z_proj.bias)


class Block(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.ln_x = nn.LayerNorm(config.n_embd)
        self.xattn = CrossAttention(config)
        self.ln_1 = nn.LayerNorm(config.n_embd)
        self.attn = CausalSelfAttention(config)
        self.ln_2 = nn.LayerNorm(config.n_embd)
        self.mlp = MLP(config)
        self.cross_gate = nn.Parameter(torch.tensor(0.0))

    def forward(self, x, z):
        if z is not None:
            x, hx = ops.residual_layernorm(x, self.ln_x.weight, self.ln_x.bias, self.ln_x.eps)
            y = self.xattn.attend(hx, z)
            # x + tanh(gate) * c_proj(y): gate scale and residual fused into the projection GEMM epilogue
            x = ops.gated_proj_residual(y, self.xattn.c_proj.weight, self.xattn.c_proj.bias, self.cross_gate, x)
        x, h = ops.residual_layernorm(x, self.ln_1.weight, self.ln_1.bias, self.ln_1.eps)
        x = self.attn.attend(h, x)
        x, h = ops.residual_layernorm(x, self.ln_2.weight, self.ln_2.bias, self.ln_2.eps)
        return self.mlp.transform(h, x)


class GPT(nn.Module):
    """forward(idx, z=None, targets=None, target_mask=None) -> (logits, loss) (model.py:116-186).
    With a target_mask the loss is sum(ce * mask) / max(mask.sum(), 1)."""

    def __init__(self, config):
        super().__init__()
        _check_config(config)
        self.config = config
        self.transformer = nn.ModuleDict(dict(
            wte=nn.Embedding(config.vocab_size, config.n_embd),
            wpe=nn.Embedding(config.block_size, config.n_embd),
            vis_proj=Vision_projector(config),
            h=nn.ModuleList([Block(config) for _ in range(config.n_layer)]),
            ln_f=nn.LayerNorm(config.n_embd),
        ))
        self.lm_head = nn.Linear(config.n_embd, config.vocab_size, bias=False)
        self.transformer.wte.weight = self.lm_head.weight
        self.return_logits_with_loss = False
        _init_gpt_weights(self, config.n_layer)
        for p in self.parameters():
            p.requires_grad = False
        for p in self.transformer["vis_proj"].parameters():
            p.requires_grad = True
        for blk in self.transformer["h"]:
            for p in blk.xattn.parameters():
                p.requires_grad = True
            blk.cross_gate.requires_grad = True

    def forward(self, idx, z=None, targets=None, target_mask=None):
        B, T = idx.shape
        if T > self.config.block_size:
            raise AssertionError(f"Cannot forward sequence of length {T}, block size is only {self.config.block_size}")
        x = ops.embed(idx, self.transformer.wte.weight, self.transformer.wpe.weight)
        z_proj = self.transformer.vis_proj(z) if z is not None else None
        for block in self.transformer.h:
            x = block(x, z_proj)
        f = self.transformer.ln_f
        x = ops.layernorm(x, f.weight, f.bias, f.eps)
        logits = loss = None
        if targets is not None:
            loss = ops.lmhead_ce(x, self.lm_head.weight, targets, target_mask)
        if targets is None or self.return_logits_with_loss:
            logits = ops.linear(x, self.lm_head.weight)
        return logits, loss

    def configure_optimizers(self, weight_decay, learning_rate, device):
        return build_adamw(self, weight_decay, learning_rate, device)
