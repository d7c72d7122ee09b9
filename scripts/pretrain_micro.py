"""Eager GPT-2 pretraining micro-steps (B=16, T=1024) exactly as step.PretrainStep runs them (flat gradient bucket,
in-place residual gradients, 1/grad_accum folded into the loss epilogue) — target for the ncu launch list.
The LAST micro-step is the one to read: `python scripts/launch_table.py <csv> --last-of 3`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpt2_vision_language_b200 import gpt2  # noqa: E402
from gpt2_vision_language_b200.step import PretrainStep  # noqa: E402

torch.manual_seed(0)
m = gpt2.GPT(gpt2.GPTConfig(vocab_size=50304)).cuda().to(torch.bfloat16)
st = PretrainStep(m, micro_batch=16, seq=1024, grad_accum=32, use_graph=False)
st.x.copy_(torch.randint(0, 50257, (16, 1024), device="cuda"))
st.y.copy_(torch.randint(0, 50257, (16, 1024), device="cuda"))
st.bucket.zero()
for i in range(3):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    st._micro()
    e.record()
    torch.cuda.synchronize()
    print("micro-step", i, "ms", s.elapsed_time(e), "loss", st.loss.item(), flush=True)
