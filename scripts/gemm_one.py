"""One GEMM shape, a few launches (for ncu). Usage: gemm_one.py M N K [bias+act]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpt2_vision_language_b200 import ops
M, N, K = (int(x) for x in sys.argv[1:4])
act = sys.argv[4] if len(sys.argv) > 4 else None
a = torch.randn(M, K, device="cuda").bfloat16()
b = torch.randn(N, K, device="cuda").bfloat16()
bias = torch.randn(N, device="cuda").bfloat16() if act else None
for _ in range(5):
    out = ops.gemm(a, b, bias=bias, act=act if act != "none" else None)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    out = ops.gemm(a, b, bias=bias, act=act if act != "none" else None)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 20
print(f"M={M} N={N} K={K} act={act}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
