import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpt2_vision_language_b200 import ops
def timeit(fn, iters=30):
    for _ in range(3): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3
B, H = 64, 16
C = H * 64
qkv = torch.randn(B, 257, 3 * C, device="cuda").bfloat16()
q256 = qkv[:, :256, :C]
full = timeit(lambda: ops.attention_fwd(qkv[..., :C], qkv[..., C:2*C], qkv[..., 2*C:], H, False, need_lse=False))
no_tail = timeit(lambda: ops.attention_fwd(q256, qkv[..., C:2*C], qkv[..., 2*C:], H, False, need_lse=False))
q1 = qkv[:, 256:, :C]
os.environ["VLK_ATTN_IMPL"] = "simt"
tail = timeit(lambda: ops.attention_fwd(q1, qkv[..., C:2*C], qkv[..., 2*C:], H, False, need_lse=False))
print(f"Tq=257: {full:.1f} us | Tq=256 (tensor-core part only): {no_tail:.1f} us | Tq=1 row (few-rows kernel): {tail:.1f} us")
