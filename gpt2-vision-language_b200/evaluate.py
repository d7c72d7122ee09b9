"""Forward-only evaluation helpers on the B200 kernels.

* ``get_most_likely_row(tokens, mask, logits)`` — same signature and result as the HellaSwag helper of the reference
  (source/gpt2/train_gpt2.py:190-202): per-token CE of the shifted logits (``reduction='none'``), masked mean over the
  completion region of each row, argmin over rows.
* ``most_likely_row(model, tokens, mask)`` — the same decision without ever materialising the [rows, T, vocab]
  logits: trunk forward, then the chunked last-layer product + per-row softmax-CE kernel.
* ``validation_loss(model_call, batches)`` — the loop body of ``run_validation_and_logging``
  (source/gpt2_linear/train.py:218-240): mean of the per-batch losses under ``torch.no_grad()``.
"""
import torch

from . import ops


@torch.no_grad()
def masked_row_losses(tokens, mask, logits=None, hidden=None, lm_weight=None):
    """avg over masked positions of CE(logits[:, t], tokens[:, t+1]) per row -> fp32 [rows]."""
    B, T = tokens.shape
    shift_tokens = tokens[:, 1:].contiguous()
    if logits is not None:
        lg = logits[:, :-1, :].contiguous()
        losses = ops.cross_entropy_rows(lg.view(-1, lg.shape[-1]), shift_tokens)
    else:
        losses = ops.lmhead_ce_rows(hidden[:, :-1, :].contiguous(), lm_weight, shift_tokens)
    losses = losses.view(B, T - 1)
    shift_mask = mask[:, 1:].to(losses.dtype)
    return (losses * shift_mask).sum(dim=1) / shift_mask.sum(dim=1)


@torch.no_grad()
def get_most_likely_row(tokens, mask, logits):
    return int(masked_row_losses(tokens, mask, logits=logits).argmin().item())


@torch.no_grad()
def most_likely_row(model, tokens, mask):
    """model: gpt2.GPT.  tokens [4, T] candidate endings, mask [4, T] = 1 over the completion."""
    t = model.transformer
    hidden = model.trunk(ops.embed(tokens, t.wte.weight, t.wpe.weight))
    return int(masked_row_losses(tokens, mask, hidden=hidden, lm_weight=model.lm_head.weight).argmin().item())


@torch.no_grad()
def validation_loss(model_call, batches, max_steps=20):
    """model_call(batch) -> loss tensor.  Returns the mean loss over the first max_steps batches (device scalar)."""
    acc, n = None, 0
    for i, batch in enumerate(batches):
        if i >= max_steps:
            break
        loss = model_call(batch).detach().float()
        acc = loss if acc is None else acc + loss
        n += 1
    if acc is None:
        raise RuntimeError("validation_loss: no batches")
    return acc / n
