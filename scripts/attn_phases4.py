import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("VLK_ATTN_DEBUG", "1")
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
names = ["start", "dots", "S ready", "pass1", "pass2", "p_ready arrive", "O ready", "epilogue", "s_free"]
t0 = min(buf[i * 16] for i in range(8) if buf[i * 16] > 0)
for g in (0, 1):
    for t in range(4):
        st = [buf[(g * 4 + t) * 16 + i] for i in range(9)]
        print(f"group {g} tile {t+2}: start@{st[0]-t0:6d} " + " ".join(f"{names[i]}=+{st[i]-st[i-1]}" for i in range(1, 9)) + f" | total {st[8]-st[0]}")
