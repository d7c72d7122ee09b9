import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["VLK_ATTN_DEBUG"] = "1"
from gpt2_vision_language_b200 import ops, _lib
B, H, T = 64, 16, 257
C = H * 64
qkv = torch.randn(B, T, 3 * C, device="cuda").bfloat16()
for _ in range(3):
    ops.attention_fwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], H, False, need_lse=False)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_longlong * (32 * 16))()
lib.vlk_debug_dump(buf, 32 * 16)
names = ["start", "alloc+sync", "QK landed(t0)", "extra-dot", "S ready", "pass1 max", "pass2 exp/st", "sync", "O ready", "epilogue", "dealloc"]
t0 = min(buf[z * 16] for z in range(32))
for z in (0, 1, 2, 15, 31):
    st = [buf[z * 16 + i] for i in range(11)]
    print(f"cta(z={z}) start@{st[0]-t0:7d}ns: " + " ".join(f"{names[i]}=+{st[i]-st[i-1]}" for i in range(1, 11)) + f" | total {st[10]-st[0]} ns")
