"""Bring-up build only: per-phase %globaltimer stamps of the CLIP attention kernel's softmax groups (tiles 2..5 of CTA 0).
Build: nvcc ... -DVLK_BRINGUP of api / attention_*.cu into scripts/probe/libvlk_attn_bringup.so."""
import ctypes, os
import torch
here = os.path.dirname(os.path.abspath(__file__))
os.environ["VLK_ATTN_DEBUG"] = "1"
lib = ctypes.CDLL(os.path.join(here, "libvlk_attn_bringup.so"))
B, H, T = 64, 16, 257
C = H * 64
qkv = (torch.randn(B, T, 3 * C, device="cuda") * 0.5).bfloat16()
o = torch.empty(B, T, C, device="cuda", dtype=torch.bfloat16)
q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
vp, ll, ci, cf, cu = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_float, ctypes.c_uint
lib.vlk_attn_fwd.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, ll, ci, ll, ci, ll, ci, ll, ci, ci, cf, cf, vp, cu, vp]
for _ in range(3):
    rc = lib.vlk_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), None, B, H, T, T, q.stride(0), q.stride(1),
                          k.stride(0), k.stride(1), v.stride(0), v.stride(1), o.stride(0), o.stride(1), 0, 0.125, 0.0, None, 0,
                          torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
    torch.cuda.synchronize()
buf = (ctypes.c_longlong * (64 * 16))()
lib.vlk_debug_dump(buf, 64 * 16)
names = ["0 tile start", "1 extra key done", "2 s_ready seen", "3 pass1+max exchange done", "4 pass2 done", "5 p_ready arrived",
         "6 o_ready seen", "7 O stored", "8 s_free"]
t0 = min(x for x in buf if x > 0)
for g in range(2):
    for tile in range(4):
        st = [buf[(g * 4 + tile) * 16 + s] for s in range(15)]
        print(f"group {g} tile {tile + 2}: " + "  ".join(f"{(x - t0) / 1e3:7.2f}" for x in st) + "   (us since first stamp)")
print("phases:", " | ".join(names), "| 9-14 leftover row: start, q staged, scores, max+exp+sum, P.V, done")
