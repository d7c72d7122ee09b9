"""Greedy caption decoding on the B200 path.

Loop shape of ``evaluate_cider`` (source/gpt2_linear/data.py:108-131): start from a prompt ("A photo of"), re-run
the full captioner on the growing sequence, read ``logits[:, -1]``, append one token, 24 times.  The reference
samples (temperature 0.8 + nucleus 0.9); BASELINE.json's parity target is the *greedy* variant (argmax), which is
what this implements — batched, with the last-row lm_head and the argmax as libvlk kernels.  (A KV-cached decode is
SURVEY 8(f) rank 1 / next round; this version recomputes like the reference does.)
"""
import torch

from . import ops


@torch.no_grad()
def greedy_decode(model, z, prompt_ids, max_new_tokens=24, kind="prefix"):
    """model: GPT_Caption (kind='prefix': linear / Q-Former) or cross-attention GPT (kind='xattn').
    z: pooled CLIP tokens [B,33,D]; prompt_ids: int64 [B,P].  Returns int64 [B, P + max_new_tokens]."""
    x = prompt_ids
    for _ in range(max_new_tokens):
        if kind == "xattn":
            h = _xattn_hidden(model, x, z)
            w = model.lm_head.weight
        else:
            h = _prefix_hidden(model, x, z)
            w = model.gpt.lm_head.weight
        logits_last = ops.gemm(h[:, -1, :].contiguous(), w)          # [B, V]: only the last row is needed
        nxt = ops.argmax_rows(logits_last)
        x = torch.cat([x, nxt[:, None]], dim=1)
    return x


def _prefix_hidden(model, input_ids, z):
    prefix = model.bridge(z[:, 0:1, :] if model.use_cls_only else z)
    full = ops.embed(input_ids, model.wte.weight, model.wpe.weight, prefix)
    return model.gpt.trunk(full)


def _xattn_hidden(model, idx, z):
    x = ops.embed(idx, model.transformer.wte.weight, model.transformer.wpe.weight)
    zp = model.transformer.vis_proj(z)
    for block in model.transformer.h:
        x = block(x, zp)
    f = model.transformer.ln_f
    return ops.layernorm(x, f.weight, f.bias, f.eps)
