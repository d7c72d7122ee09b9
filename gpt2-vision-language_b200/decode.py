"""Greedy caption decoding on the B200 path.

Loop shape of ``evaluate_cider`` (source/gpt2_linear/data.py:108-131): start from a prompt ("A photo of"), read
``logits[:, -1]``, append one token, 24 times.  The reference re-runs the full captioner on the growing sequence and
samples (temperature 0.8 + nucleus 0.9); BASELINE.json's parity target is the *greedy* variant (argmax).

Two implementations, same ids:

* ``greedy_decode``          — KV-cached: the image prefix and the prompt are run once (prefill),
  every layer's keys / values are kept in a ``[B, L_max, 2C]`` cache, and each new token costs one row per GEMM
  plus one query row of attention against the cache (the head-pair tcgen05 kernel up to 64 cached tokens, the
  few-rows kernel beyond).  For the cross-attention captioner the ``kv_proj`` of the 33 image tokens is computed
  once per layer.  Batched; last-row lm_head + argmax are libvlk kernels.
* ``greedy_decode_recompute`` — the reference's loop shape (full forward per token), kept as the cross-check.
"""
import torch

from . import ops


# ----------------------------------------------------------------------------------------------------------
# pieces of a transformer block, forward-only, on raw ops (no autograd graph, nothing saved)
# ----------------------------------------------------------------------------------------------------------
def _ln(x2, ln):
    return ops.layernorm_fwd(x2, ln.weight, ln.bias, ln.eps, save_stats=False)[0]


def _self_attn_cached(attn, h2, x2, cache, B, L0, Ln):
    """h2: LayerNorm'd rows [B*Ln, C] of positions L0..L0+Ln-1; cache: [B, Lmax, 2C] (k | v).  Returns x2 + proj."""
    C = attn.n_embd
    qkv = ops.gemm(h2, attn.c_attn.weight, bias=attn.c_attn.bias).view(B, Ln, 3 * C)
    cache[:, L0:L0 + Ln, :] = qkv[..., C:]
    k, v = cache[:, :L0 + Ln, :C], cache[:, :L0 + Ln, C:]
    # causal with Tk - Tq = L0: query t sees cached positions <= L0 + t
    o, _ = ops.attention_fwd(qkv[..., :C], k, v, attn.n_head, True, need_lse=False)
    return ops.gemm(o.view(B * Ln, C), attn.c_proj.weight, bias=attn.c_proj.bias, residual=x2)


def _mlp(mlp, h2, x2):
    u = ops.gemm(h2, mlp.c_fc.weight, bias=mlp.c_fc.bias, act="gelu_tanh")
    return ops.gemm(u, mlp.c_proj.weight, bias=mlp.c_proj.bias, residual=x2)


def _gpt_block_cached(block, x2, cache, B, L0, Ln):
    x2 = _self_attn_cached(block.attn, _ln(x2, block.ln_1), x2, cache, B, L0, Ln)
    return _mlp(block.mlp, _ln(x2, block.ln_2), x2)


def _xattn_block_cached(block, x2, cache, zkv, B, L0, Ln):
    C = block.attn.n_embd
    xa = block.xattn
    q = ops.gemm(_ln(x2, block.ln_x), xa.q_proj.weight, bias=xa.q_proj.bias).view(B, Ln, C)
    y, _ = ops.attention_fwd(q, zkv[..., :C], zkv[..., C:], xa.n_head, False, need_lse=False)
    tg = torch.tanh(block.cross_gate.detach().float()).reshape(1)
    x2 = ops.gemm(y.view(B * Ln, C), xa.c_proj.weight, bias=xa.c_proj.bias, scale=tg, residual=x2)
    return _gpt_block_cached(block, x2, cache, B, L0, Ln)


class _Decoder:
    """Holds the per-layer caches of one decode call."""

    def __init__(self, model, kind, z, n_text_max, batch=None):
        self.kind = kind
        if kind == "gpt":                       # plain GPT-2 (source/gpt2/train_gpt2.py): no image tokens at all
            t = model.transformer
            self.blocks, self.ln_f, self.wte, self.wpe, self.lm_w = t.h, t.ln_f, t.wte.weight, t.wpe.weight, model.lm_head.weight
            self.prefix = self.zkv = None
        elif kind == "xattn":
            t = model.transformer
            self.blocks, self.ln_f, self.wte, self.wpe, self.lm_w = t.h, t.ln_f, t.wte.weight, t.wpe.weight, model.lm_head.weight
            zp = t.vis_proj(z)                                                  # [B,33,C], shared by all layers
            B = z.shape[0]
            z2 = zp.reshape(-1, zp.shape[-1])
            self.zkv = [ops.gemm(z2, b.xattn.kv_proj.weight, bias=b.xattn.kv_proj.bias).view(B, zp.shape[1], -1)
                        for b in self.blocks]                                   # cross K/V: once per layer
            self.prefix = None
        else:
            g = model.gpt
            t = g.transformer
            self.blocks, self.ln_f, self.wte, self.wpe, self.lm_w = t.h, t.ln_f, model.wte.weight, model.wpe.weight, g.lm_head.weight
            self.prefix = model.bridge(z[:, 0:1, :] if model.use_cls_only else z)
            self.zkv = None
        C = self.wte.shape[1]
        B = z.shape[0] if z is not None else batch
        self.B, self.C = B, C
        max_len = (0 if self.prefix is None else self.prefix.shape[1]) + n_text_max
        self.cache = [torch.empty(B, max_len, 2 * C, device=self.wte.device, dtype=torch.bfloat16) for _ in self.blocks]
        self.len = 0

    def _run(self, x):
        """x: [B, Ln, C] embedded rows at positions len..len+Ln-1 -> final-LayerNorm'd LAST row [B, C]."""
        B, Ln, C = x.shape
        x2 = x.reshape(B * Ln, C)
        for i, blk in enumerate(self.blocks):
            if self.kind == "xattn":
                x2 = _xattn_block_cached(blk, x2, self.cache[i], self.zkv[i], B, self.len, Ln)
            else:
                x2 = _gpt_block_cached(blk, x2, self.cache[i], B, self.len, Ln)
        self.len += Ln
        last = x2.view(B, Ln, C)[:, -1, :].contiguous()
        return _ln(last, self.ln_f)

    def prefill(self, prompt_ids):
        self.n_text = prompt_ids.shape[1]
        return self._run(ops.embed_concat(prompt_ids, self.wte, self.wpe, self.prefix))

    def step(self, token_ids):
        """token_ids [B] -> hidden of the new position.  Text positions restart at 0 after the image prefix."""
        x = ops.embed_concat(token_ids[:, None].contiguous(), self.wte, self.wpe, None, pos0=self.n_text)
        self.n_text += 1
        return self._run(x)

    def next_token(self, h_last):
        return ops.argmax_rows(ops.gemm(h_last, self.lm_w))                   # [B, V] logits of the last row only


@torch.no_grad()
def greedy_decode(model, z, prompt_ids, max_new_tokens=24, kind="prefix"):
    """model: GPT_Caption (kind='prefix': linear / Q-Former), cross-attention GPT (kind='xattn') or the plain GPT
    (kind='gpt', z=None).  z: pooled CLIP tokens [B,33,D]; prompt_ids: int64 [B,P].
    Returns int64 [B, P + max_new_tokens]."""
    ops._need_cuda(z, prompt_ids)
    was_training = model.training
    model.eval()                                   # decode is an eval-mode loop in the reference (data.py:77)
    try:
        dec = _Decoder(model, kind, z, n_text_max=prompt_ids.shape[1] + max_new_tokens, batch=prompt_ids.shape[0])
        out = [prompt_ids]
        h = dec.prefill(prompt_ids.contiguous())
        for t in range(max_new_tokens):
            nxt = dec.next_token(h)
            out.append(nxt[:, None])
            if t + 1 < max_new_tokens:
                h = dec.step(nxt)
        return torch.cat(out, dim=1)
    finally:
        model.train(was_training)


def _sample_top_p(logits, temperature, top_p, generator):
    """The sampler of ``evaluate_cider`` (source/gpt2_linear/data.py:114-125): softmax(logits / T), keep the smallest
    prefix of the sorted distribution whose mass exceeds top_p (the first token always survives), renormalise, draw."""
    probs = torch.softmax(logits.float() / temperature, dim=-1)
    sorted_probs, sorted_idx = torch.sort(probs, descending=True)
    cutoff = sorted_probs.cumsum(dim=-1) > top_p
    cutoff[..., 1:] = cutoff[..., :-1].clone()
    cutoff[..., 0] = False
    sorted_probs = sorted_probs.masked_fill(cutoff, 0.0)
    sorted_probs = sorted_probs / sorted_probs.sum(dim=-1, keepdim=True)
    return sorted_idx.gather(-1, torch.multinomial(sorted_probs, 1, generator=generator)).squeeze(-1)


def _sample_top_k(logits, k, generator):
    """The sampler of the pretraining script's text generation (source/gpt2/train_gpt2.py:444-449)."""
    topk_probs, topk_idx = torch.topk(torch.softmax(logits.float(), dim=-1), k, dim=-1)
    return topk_idx.gather(-1, torch.multinomial(topk_probs, 1, generator=generator)).squeeze(-1)


@torch.no_grad()
def sample_decode(model, z, prompt_ids, max_new_tokens=24, kind="prefix", temperature=0.8, top_p=0.9, top_k=None,
                  generator=None):
    """KV-cached decode with the reference's samplers instead of argmax: nucleus (temperature 0.8, top-p 0.9 — the
    CIDEr evaluation loop) or, with ``top_k`` set, plain top-k (k = 50 in the pretraining script).  The last-row
    logits come from the libvlk GEMM; the draw itself is host-side policy on a [B, vocab] tensor."""
    ops._need_cuda(z, prompt_ids)
    was_training = model.training
    model.eval()
    try:
        dec = _Decoder(model, kind, z, n_text_max=prompt_ids.shape[1] + max_new_tokens, batch=prompt_ids.shape[0])
        out = [prompt_ids]
        h = dec.prefill(prompt_ids.contiguous())
        for t in range(max_new_tokens):
            logits = ops.gemm(h, dec.lm_w)
            nxt = _sample_top_k(logits, top_k, generator) if top_k else _sample_top_p(logits, temperature, top_p, generator)
            out.append(nxt[:, None])
            if t + 1 < max_new_tokens:
                h = dec.step(nxt)
        return torch.cat(out, dim=1)
    finally:
        model.train(was_training)


@torch.no_grad()
def greedy_decode_recompute(model, z, prompt_ids, max_new_tokens=24, kind="prefix"):
    """The reference's loop shape: the whole captioner is re-run on the growing sequence for every token."""
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
