"""Linear-projection captioner — drop-in for source/gpt2_linear/model.py (same public names)."""
import torch.nn as nn

from . import ops
from .caption import PrefixCaptioner, pool_clip_197_to_33_avg_with_cls  # noqa: F401  (re-exported)
from .gpt2 import GPT_previous, GPTConfig, Block, MLP, CausalSelfAttention  # noqa: F401


class Linear_Bridge(nn.Module):
    """One trainable 768->768 projection of the pooled CLIP tokens (gpt2_linear/model.py:114-129).
    The unused constructor arguments are accepted for signature compatibility."""

    def __init__(self, enc_dim, d_lm, n_heads=None, n_queries=None, n_layers=None, drop=0.1):
        super().__init__()
        self.vis_proj = nn.Linear(enc_dim, d_lm)

    def forward(self, patch_tokens):
        return ops.linear(patch_tokens, self.vis_proj.weight, self.vis_proj.bias)


class GPT_Caption(PrefixCaptioner):
    def __init__(self, enc_dim: int, lm: nn.Module, m_vis_tokens: int = 8, use_cls_only: bool = False,
                 freeze_lm: bool = True):
        super().__init__()
        self._setup(lm, use_cls_only, freeze_lm)
        self.bridge = Linear_Bridge(enc_dim=enc_dim, d_lm=self.d, n_heads=lm.config.n_head, n_queries=m_vis_tokens,
                                    n_layers=2, drop=0.1)
        self._finish(freeze_lm)
