"""Q-Former-style (BLIP-2-inspired) captioner — drop-in for source/gpt2_q_former/model.py."""
import torch
import torch.nn as nn

from . import ops
from .caption import PrefixCaptioner, pool_clip_197_to_33_avg_with_cls  # noqa: F401  (re-exported)
from .gpt2 import GPT_previous, GPTConfig, Block, MLP, CausalSelfAttention  # noqa: F401


class QFormerLayer(nn.Module):
    """Learned queries: self-attention, cross-attention to the (LayerNorm'd) visual tokens, erf-GELU MLP, all
    pre-LN residual (gpt2_q_former/model.py:114-145).  ``nn.MultiheadAttention`` modules are instantiated only
    to own the parameters (same state_dict keys and default init); their math runs on libvlk:
    packed in_proj GEMM -> attention kernel -> out_proj GEMM with the residual fused."""

    def __init__(self, d, n_heads, drop=0.1):
        super().__init__()
        self.n_heads = n_heads
        self.ln1 = nn.LayerNorm(d)
        self.self_attn = nn.MultiheadAttention(d, n_heads, dropout=drop, batch_first=True)
        self.ln2_q = nn.LayerNorm(d)
        self.ln2_v = nn.LayerNorm(d)
        self.cross_attn = nn.MultiheadAttention(d, n_heads, dropout=drop, batch_first=True)
        self.ln3 = nn.LayerNorm(d)
        self.mlp = nn.Sequential(nn.Linear(d, 4 * d), nn.GELU(), nn.Linear(4 * d, d))
        self.drop = nn.Dropout(drop)

    def forward(self, q, v):
        d = q.shape[-1]
        sa, ca = self.self_attn, self.cross_attn
        # dropout is live only in train mode (nn.Dropout / nn.MultiheadAttention semantics): 3 residual branches
        # + the attention probabilities of both attentions, masks from the Philox state of ops.DropoutState
        p_res = self.drop.p if self.training else 0.0
        p_sa = sa.dropout if self.training else 0.0
        p_ca = ca.dropout if self.training else 0.0
        rng = ops.DropoutState.default(q.device) if (p_res > 0 or p_sa > 0 or p_ca > 0) else None

        def branch(out, w, b, res):          # res + dropout(out @ w^T + b)
            if p_res > 0:
                return ops.dropout_add(ops.linear(out, w, b), res, p_res, rng)
            return ops.linear(out, w, b, res)

        h = ops.layernorm(q, self.ln1.weight, self.ln1.bias, self.ln1.eps)
        qkv = ops.linear(h, sa.in_proj_weight, sa.in_proj_bias)
        a = ops.self_attention(qkv, self.n_heads, False, p_sa, rng)
        q = branch(a, sa.out_proj.weight, sa.out_proj.bias, q)

        hq = ops.layernorm(q, self.ln2_q.weight, self.ln2_q.bias, self.ln2_q.eps)
        hv = ops.layernorm(v, self.ln2_v.weight, self.ln2_v.bias, self.ln2_v.eps)
        qq = ops.linear(hq, ca.in_proj_weight[:d], ca.in_proj_bias[:d])
        kv = ops.linear(hv, ca.in_proj_weight[d:], ca.in_proj_bias[d:])
        a = ops.cross_attention(qq, kv, self.n_heads, p_ca, rng)
        q = branch(a, ca.out_proj.weight, ca.out_proj.bias, q)

        h = ops.layernorm(q, self.ln3.weight, self.ln3.bias, self.ln3.eps)
        fc, proj = self.mlp[0], self.mlp[2]
        if p_res > 0:
            m = ops.mlp(h, fc.weight, fc.bias, proj.weight, proj.bias, None, "gelu_erf")
            return ops.dropout_add(m, q, p_res, rng)
        return ops.mlp(h, fc.weight, fc.bias, proj.weight, proj.bias, q, "gelu_erf")


class BLIP2Bridge(nn.Module):
    """vis_proj + n_queries learned query tokens refined by n_layers QFormerLayers (model.py:147-168)."""

    def __init__(self, enc_dim, d_lm, n_heads, n_queries=2, n_layers=2, drop=0.1):
        super().__init__()
        self.vis_proj = nn.Linear(enc_dim, d_lm)
        self.n_queries = n_queries
        self.query_tokens = nn.Parameter(torch.randn(n_queries, d_lm))
        self.layers = nn.ModuleList([QFormerLayer(d_lm, n_heads, drop=drop) for _ in range(n_layers)])

    def forward(self, patch_tokens):
        x = ops.linear(patch_tokens, self.vis_proj.weight, self.vis_proj.bias)
        q = self.query_tokens.unsqueeze(0).expand(x.shape[0], -1, -1)
        for layer in self.layers:
            q = layer(q, x)
        return q


class GPT_Caption(PrefixCaptioner):
    def __init__(self, enc_dim: int, lm: nn.Module, m_vis_tokens: int = 8, use_cls_only: bool = False,
                 freeze_lm: bool = True):
        super().__init__()
        self._setup(lm, use_cls_only, freeze_lm)
        self.bridge = BLIP2Bridge(enc_dim=enc_dim, d_lm=self.d, n_heads=lm.config.n_head, n_queries=m_vis_tokens,
                                  n_layers=2, drop=0.1)
        self._finish(freeze_lm)
