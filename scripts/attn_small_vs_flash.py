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
for (B, H, T, causal) in [(64, 12, 64, True), (64, 12, 63, True), (64, 12, 31, True)]:
    C = H * 64
    qkv = torch.randn(B, T, 3 * C, device="cuda").bfloat16()
    for impl in ("small", "flash"):
        if impl == "flash": os.environ["VLK_ATTN_IMPL"] = "flash"
        else: os.environ.pop("VLK_ATTN_IMPL", None)
        o, lse = ops.attention_fwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], H, causal)
        do = torch.randn_like(o); dqkv = torch.empty_like(qkv)
        f = timeit(lambda: ops.attention_fwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], H, causal))
        b = timeit(lambda: ops.attention_bwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], o, do, lse, dqkv[..., :C], dqkv[..., C:2 * C], dqkv[..., 2 * C:], H, causal))
        print(f"T={T} impl={impl}: fwd {f:.1f} us  bwd {b:.1f} us")
