"""CLIP-shaped attention forward (B=64, H=16, T=257): correctness vs torch fp32 and graph-timed duration per variant."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpt2_vision_language_b200 import ops
from oracle import torch_oracle as O
B, H, T = 64, 16, 257
C = H * 64
torch.manual_seed(0)
qkv = (torch.randn(B, T, 3 * C, device="cuda") * 0.5).bfloat16()
q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
ref = O.sdpa(q[:4].float(), k[:4].float(), v[:4].float(), H, False)
for var in sys.argv[1:] or ["0"]:
    os.environ["VLK_ATTN_EXP"] = var
    o, _ = ops.attention_fwd(q, k, v, H, False, need_lse=False)
    err = ((o[:4].float() - ref).abs().max() / ref.abs().max()).item()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.attention_fwd(q, k, v, H, False, need_lse=False)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(20):
            ops.attention_fwd(q, k, v, H, False, need_lse=False)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 40 * 1e3
    print(f"variant {var}: {us:.1f} us  ({4.0 * B * H * T * T * 64 / us / 1e6:.0f} TFLOP/s)  max rel err {err:.2e}", flush=True)
