"""One fused lm_head + cross-entropy forward/backward at the given shape (for `ncu` launch lists)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpt2_vision_language_b200 import ops  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
need_dw = (sys.argv[2] == "1") if len(sys.argv) > 2 else True
C, V = 768, 50304
h = torch.randn(rows, C, device="cuda").bfloat16().requires_grad_(True)
w = (torch.randn(V, C, device="cuda") * 0.02).bfloat16().requires_grad_(need_dw)
labels = torch.randint(0, V, (rows,), device="cuda")
for _ in range(2):
    h.grad = None
    ops.lmhead_ce(h, w, labels).backward()
torch.cuda.synchronize()
