"""GPT-2 124M decoder on libvlk kernels — drop-in for the nn.Module surface of the reference.

Mirrors (names, constructor signatures, attribute names, state_dict keys, init distribution) of
``source/gpt2/train_gpt2.py:21-144`` (``CausalSelfAttention`` / ``MLP`` / ``Block`` / ``GPTConfig`` / ``GPT``)
and of its copy ``GPT_previous`` in ``source/gpt2_linear/model.py:7-111`` (no ``attn.bias`` buffer).
Every forward/backward op underneath is a libvlk kernel (``ops.py``); nothing falls back to ATen math.
"""
import inspect
import math
from dataclasses import dataclass

import torch
import torch.nn as nn

from . import ops


@dataclass
class GPTConfig:
    block_size: int = 1024
    vocab_size: int = 50257
    n_layer: int = 12
    n_head: int = 12
    n_embd: int = 768


class CausalSelfAttention(nn.Module):
    """c_attn -> causal attention over the packed [B,T,3C] projection -> c_proj (train_gpt2.py:21-43)."""

    # the reference registers an (unused) lower-triangular "bias" buffer; it is kept only as a state_dict key
    register_mask_buffer = True

    def __init__(self, config):
        super().__init__()
        if config.n_embd % config.n_head != 0:
            raise AssertionError("n_embd must be divisible by n_head")
        self.n_head, self.n_embd = config.n_head, config.n_embd
        self.c_attn = nn.Linear(config.n_embd, 3 * config.n_embd)
        self.c_proj = nn.Linear(config.n_embd, config.n_embd)
        self.c_proj.NANOGPT_SCALE_INIT = 1
        if self.register_mask_buffer:
            mask = torch.ones(config.block_size, config.block_size).tril_()
            self.register_buffer("bias", mask.view(1, 1, config.block_size, config.block_size))

    def attend(self, x, residual=None):
        qkv = ops.linear(x, self.c_attn.weight, self.c_attn.bias)
        y = ops.self_attention(qkv, self.n_head, True)
        return ops.linear(y, self.c_proj.weight, self.c_proj.bias, residual)

    def forward(self, x):
        return self.attend(x)


class _CausalSelfAttentionNoBuffer(CausalSelfAttention):
    register_mask_buffer = False


class MLP(nn.Module):
    """768 -> 3072 -> tanh-GELU -> 768 with the activation fused into the GEMM epilogues (train_gpt2.py:46-59)."""

    def __init__(self, config):
        super().__init__()
        self.c_fc = nn.Linear(config.n_embd, 4 * config.n_embd)
        self.gelu = nn.GELU(approximate="tanh")  # kept for state/attribute parity; the math runs in the epilogue
        self.c_proj = nn.Linear(4 * config.n_embd, config.n_embd)
        self.c_proj.NANOGPT_SCALE_INIT = 1

    def transform(self, x, residual=None):
        return ops.mlp(x, self.c_fc.weight, self.c_fc.bias, self.c_proj.weight, self.c_proj.bias, residual,
                       "gelu_tanh")

    def forward(self, x):
        return self.transform(x)


class Block(nn.Module):
    """Pre-LN residual block; both residual adds happen in the projection GEMM epilogues (train_gpt2.py:62-74)."""

    _attn_cls = CausalSelfAttention

    def __init__(self, config):
        super().__init__()
        self.ln_1 = nn.LayerNorm(config.n_embd)
        self.attn = self._attn_cls(config)
        self.ln_2 = nn.LayerNorm(config.n_embd)
        self.mlp = MLP(config)

    def forward(self, x):
        x, h = ops.residual_layernorm(x, self.ln_1.weight, self.ln_1.bias, self.ln_1.eps)
        x = self.attn.attend(h, x)
        x, h = ops.residual_layernorm(x, self.ln_2.weight, self.ln_2.bias, self.ln_2.eps)
        return self.mlp.transform(h, x)


class _BlockNoBuffer(Block):
    _attn_cls = _CausalSelfAttentionNoBuffer


def _check_config(config):
    """Shapes the sm_100a kernels are specialised for; anything else fails HERE with a clear message instead of at the
    first kernel call.  The dataclass default vocab_size=50257 is kept for signature parity with the reference, but —
    like the reference's own train scripts, which all pass vocab_size=50304 (train_gpt2.py:260, gpt2_linear/train.py:100)
    — a multiple of 8 is required: the lm_head / cross-entropy kernels move 16-byte (8 x bf16) vectors per vocab row."""
    if config.vocab_size % 8 != 0:
        raise ValueError(f"vocab_size={config.vocab_size} is not a multiple of 8: the B200 lm_head / cross-entropy "
                         f"kernels need 16-byte aligned logit rows; use the padded vocabulary the reference trains "
                         f"with, GPTConfig(vocab_size=50304)")
    if config.n_embd % config.n_head != 0 or config.n_embd // config.n_head != 64:
        raise ValueError(f"n_embd={config.n_embd} / n_head={config.n_head}: the attention kernels are specialised for "
                         f"head_dim 64 (GPT-2 124M: 768 / 12)")


def _init_gpt_weights(root, n_layer):
    """N(0, 0.02) for Linear/Embedding weights, zero biases, residual projections scaled by (2L)^-1/2
    (train_gpt2.py:100-109).  Iterates modules in registration order like nn.Module.apply does, so a given
    torch seed draws the same numbers as the reference constructor."""
    def init(m):
        if isinstance(m, nn.Linear):
            std = 0.02 * ((2 * n_layer) ** -0.5 if hasattr(m, "NANOGPT_SCALE_INIT") else 1.0)
            nn.init.normal_(m.weight, mean=0.0, std=std)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Embedding):
            nn.init.normal_(m.weight, mean=0.0, std=0.02)
    root.apply(init)


def build_adamw(module, weight_decay, learning_rate, device):
    """Parameter grouping of the reference's configure_optimizers (train_gpt2.py:127-144): trainable tensors
    with dim >= 2 are decayed, the rest are not; betas (0.9, 0.95), eps 1e-8.  On CUDA the returned optimizer
    is the fused libvlk AdamW (a torch.optim.Optimizer subclass, so param_groups / state_dict work as usual)."""
    named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
    decay = [p for _, p in named if p.dim() >= 2]
    no_decay = [p for _, p in named if p.dim() < 2]
    groups = [{"params": decay, "weight_decay": weight_decay}, {"params": no_decay, "weight_decay": 0.0}]
    print(f"num decayed parameter tensors: {len(decay)}, with {sum(p.numel() for p in decay):,} parameters")
    print(f"num non-decayed parameter tensors: {len(no_decay)}, with {sum(p.numel() for p in no_decay):,} parameters")
    on_cuda = "cuda" in str(device)
    print(f"using fused AdamW:{on_cuda}")
    if on_cuda:
        from .optim import FusedAdamW
        return FusedAdamW(groups, lr=learning_rate, betas=(0.9, 0.95), eps=1e-8)
    fused_ok = "fused" in inspect.signature(torch.optim.AdamW).parameters
    return torch.optim.AdamW(groups, lr=learning_rate, betas=(0.9, 0.95), eps=1e-8, **({"fused": False} if fused_ok else {}))


class GPT(nn.Module):
    """wte + wpe -> n_layer blocks -> ln_f -> tied lm_head; mean cross-entropy (train_gpt2.py:85-144).

    ``forward(idx, targets=None) -> (logits, loss)``.  With targets the loss comes from the chunked
    lm_head + softmax-CE path that never materialises the [B*T, vocab] logits; ``logits`` is then ``None``
    unless ``self.return_logits_with_loss`` is set (the reference loops never read it in that case).
    """

    _block_cls = Block

    def __init__(self, config):
        super().__init__()
        _check_config(config)
        self.config = config
        self.transformer = nn.ModuleDict(dict(
            wte=nn.Embedding(config.vocab_size, config.n_embd),
            wpe=nn.Embedding(config.block_size, config.n_embd),
            h=nn.ModuleList([self._block_cls(config) for _ in range(config.n_layer)]),
            ln_f=nn.LayerNorm(config.n_embd),
        ))
        self.lm_head = nn.Linear(config.n_embd, config.vocab_size, bias=False)
        self.transformer.wte.weight = self.lm_head.weight  # weight tying (train_gpt2.py:97)
        self.return_logits_with_loss = False
        _init_gpt_weights(self, config.n_layer)

    def trunk(self, x):
        """blocks + final LayerNorm on an embedded sequence [B,T,C]."""
        for block in self.transformer.h:
            x = block(x)
        f = self.transformer.ln_f
        return ops.layernorm(x, f.weight, f.bias, f.eps)

    def forward(self, idx, targets=None):
        B, T = idx.shape
        if T > self.config.block_size:
            raise AssertionError(f"Cannot forward sequence of length {T}, block size is only {self.config.block_size}")
        x = ops.embed(idx, self.transformer.wte.weight, self.transformer.wpe.weight)
        x = self.trunk(x)
        logits = loss = None
        if targets is not None:
            loss = ops.lmhead_ce(x, self.lm_head.weight, targets)
        if targets is None or self.return_logits_with_loss:
            logits = ops.linear(x, self.lm_head.weight)
        return logits, loss

    def configure_optimizers(self, weight_decay, learning_rate, device):
        return build_adamw(self, weight_decay, learning_rate, device)


class GPT_previous(GPT):
    """Same network without the per-layer ``attn.bias`` buffers (gpt2_linear/model.py:7-111)."""
    _block_cls = _BlockNoBuffer
