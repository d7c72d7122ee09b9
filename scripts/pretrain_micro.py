"""One eager GPT-2 pretraining micro-step (B=16, T=1024) — target for the ncu launch list."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpt2_vision_language_b200 import gpt2
torch.manual_seed(0)
m = gpt2.GPT(gpt2.GPTConfig(vocab_size=50304)).cuda().to(torch.bfloat16)
x = torch.randint(0, 50257, (16, 1024), device="cuda")
y = torch.randint(0, 50257, (16, 1024), device="cuda")
for i in range(3):
    _, loss = m(x, y)
    loss.backward()
    torch.cuda.synchronize()
    if i == 1:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
    if i == 2:
        e.record(); torch.cuda.synchronize()
        print("micro-step ms", s.elapsed_time(e), "loss", loss.item())
