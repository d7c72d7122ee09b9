"""Shared machinery of the two prefix-style captioners (linear bridge and Q-Former bridge).

Reference: ``GPT_Caption`` at source/gpt2_linear/model.py:134-237 and source/gpt2_q_former/model.py:172-275
(the two classes differ only in the bridge they construct), plus ``pool_clip_197_to_33_avg_with_cls``
(source/gpt2_linear/model.py:240-254).
"""
import torch
import torch.nn as nn

from . import ops
from .gpt2 import build_adamw


def pool_clip_197_to_33_avg_with_cls(tokens_197: torch.Tensor) -> torch.Tensor:
    """[B, 1+16*16, D] CLIP tokens -> [B, 33, D]: CLS kept, the patch grid averaged into 4x8 bins of 4x2
    patches, every output token L2-normalised — one libvlk kernel (vlk_pool33_l2norm)."""
    B, L, D = tokens_197.shape
    side = int(round((L - 1) ** 0.5))
    if side * side != L - 1:
        raise AssertionError(f"Expected square grid, got N={L - 1}")
    if side != 16:
        raise RuntimeError("the B200 pooling kernel is specialised for the 16x16 grid of ViT-L/14 @ 224px")
    return ops.pool33(tokens_197, normalize=True)


class PrefixCaptioner(nn.Module):
    """bridge(image tokens) -> [prefix ; text embeddings] -> frozen GPT-2 trunk -> CE over the text rows.

    Sub-classes set ``self.bridge``.  ``forward(patch_tokens, input_ids, labels=None) -> (logits, loss)``.
    With labels, the lm_head is evaluated on the text rows only, inside the chunked lm_head+CE path (the
    reference computes all rows and slices: gpt2_linear/model.py:172,205); ``logits`` is then None unless
    ``self.return_logits_with_loss`` is set.  Without labels full logits [B, M+T, V] are returned (decode).
    """

    def _setup(self, lm, use_cls_only, freeze_lm):
        self.use_cls_only = use_cls_only
        self.gpt = lm
        cfg = lm.config
        self.d = cfg.n_embd
        self.block_size = cfg.block_size
        self.return_logits_with_loss = False

    def _finish(self, freeze_lm):
        # aliases of the LM embeddings (they add the duplicate state_dict keys wte.weight / wpe.weight)
        self.wte = self.gpt.transformer.wte
        self.wpe = self.gpt.transformer.wpe
        if freeze_lm:
            for p in self.gpt.parameters():
                p.requires_grad_(False)
        for p in self.bridge.parameters():
            p.requires_grad_(True)

    def _decode_transformer(self, full_embeds):
        x = self.gpt.trunk(full_embeds)
        return ops.linear(x, self.gpt.lm_head.weight)

    def forward(self, patch_tokens, input_ids, labels=None):
        B, T_txt = input_ids.shape
        if patch_tokens.dim() == 2:
            patch_tokens = patch_tokens.unsqueeze(1)
        if patch_tokens.shape[0] != B:
            raise AssertionError("batch size image != batch size texte")
        x_img = patch_tokens[:, 0:1, :] if self.use_cls_only else patch_tokens
        prefix = self.bridge(x_img)
        M = prefix.shape[1]
        if M + T_txt > self.block_size:  # truncate the text, never the image prefix
            T_txt = self.block_size - M
            input_ids = input_ids[:, :T_txt]
            if labels is not None:
                labels = labels[:, :T_txt]
        # text positions restart at 0; the image prefix carries no position embedding
        full = ops.embed(input_ids, self.wte.weight, self.wpe.weight, prefix)
        x = self.gpt.trunk(full)
        logits = loss = None
        if labels is not None:
            loss = ops.lmhead_ce(x[:, M:M + T_txt, :], self.gpt.lm_head.weight, labels)
        if labels is None or self.return_logits_with_loss:
            logits = ops.linear(x, self.gpt.lm_head.weight)
        return logits, loss

    def configure_optimizers(self, weight_decay, learning_rate, device):
        return build_adamw(self, weight_decay, learning_rate, device)
