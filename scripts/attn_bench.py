"""Timing of the attention kernels at the captioning-step shapes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpt2_vision_language_b200 import ops


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


torch.manual_seed(0)
B, H, T = 64, 16, 257
qkv = torch.randn(B, T, 3 * H * 64, device="cuda").bfloat16()
C = H * 64
us = timeit(lambda: ops.attention_fwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], H, False, need_lse=False))
fl = 4.0 * B * H * T * T * 64
print(f"CLIP attn fwd B=64 H=16 T=257: {us:.1f} us  {fl/us/1e6:.0f} TFLOP/s")
for impl in ("tcgen05v1", "simt"):
    os.environ["VLK_ATTN_IMPL"] = impl
    us = timeit(lambda: ops.attention_fwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], H, False, need_lse=False), 5)
    print(f"   impl={impl}: {us:.1f} us")
os.environ.pop("VLK_ATTN_IMPL")
B, H, T = 64, 12, 64
C = H * 64
qkv = torch.randn(B, T, 3 * C, device="cuda").bfloat16()
o, lse = ops.attention_fwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], H, True)
do = torch.randn_like(o)
dqkv = torch.empty_like(qkv)
us = timeit(lambda: ops.attention_fwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], H, True))
print(f"GPT-2 attn fwd B=64 H=12 T=64 causal: {us:.1f} us")
us = timeit(lambda: ops.attention_bwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], o, do, lse, dqkv[..., :C],
                                      dqkv[..., C:2 * C], dqkv[..., 2 * C:], H, True))
print(f"GPT-2 attn bwd B=64 H=12 T=64 causal: {us:.1f} us")
B, H, T = 16, 12, 1024
C = H * 64
qkv = torch.randn(B, T, 3 * C, device="cuda").bfloat16()
o, lse = ops.attention_fwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], H, True)
do = torch.randn_like(o)
dqkv = torch.empty_like(qkv)
us = timeit(lambda: ops.attention_fwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], H, True))
fl = 4.0 * B * H * T * T * 64 / 2
print(f"pretrain attn fwd B=16 H=12 T=1024 causal: {us:.1f} us  {fl/us/1e6:.0f} TFLOP/s")
us = timeit(lambda: ops.attention_bwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], o, do, lse, dqkv[..., :C],
                                      dqkv[..., C:2 * C], dqkv[..., 2 * C:], H, True))
print(f"pretrain attn bwd B=16 H=12 T=1024 causal: {us:.1f} us  {2.5*fl/us/1e6:.0f} TFLOP/s")
