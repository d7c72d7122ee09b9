"""Bring-up build only: cycles per phase of the streaming attention forward loop (thread 0 = warp 0 with the issue duties,
thread 255 = a pure softmax thread), summed over a CTA's iterations, for the first 32 CTAs."""
import ctypes, os
import torch
here = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(here, os.environ.get("VLK_PROBE_LIB", "libvlk_bringup.so")))
import sys
B, H, T, CAUSAL = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (16, 12, 1024, 1)))
C = H * 64
qkv = (torch.randn(B, T, 3 * C, device="cuda") * 0.5).bfloat16()
o = torch.empty(B, T, C, device="cuda", dtype=torch.bfloat16)
q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
vp, ll, ci, cf, cu = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_float, ctypes.c_uint
lib.vlk_attn_fwd.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, ll, ci, ll, ci, ll, ci, ll, ci, ci, cf, cf, vp, cu, vp]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(4):
    e0.record()
    rc = lib.vlk_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), None, B, H, T, T, q.stride(0), q.stride(1),
                          k.stride(0), k.stride(1), v.stride(0), v.stride(1), o.stride(0), o.stride(1), CAUSAL, 0.125, 0.0, None, 0,
                          torch.cuda.current_stream().cuda_stream)
    e1.record()
    assert rc == 0, rc
    torch.cuda.synchronize()
print(f"B={B} H={H} T={T} causal={CAUSAL}: {e0.elapsed_time(e1) * 1e3:.1f} us (eager, all launches of the call)")
buf = (ctypes.c_longlong * (64 * 16))()
lib.vlk_debug_flash_dump(buf, 64 * 16)
names = ["wait S", "LDTM S + arrive s_free", "partial max + pair barrier", "row max", "wait PV(j-1) + rescale", "exp2+pack+STTM issue",
         "STTM wait + arrive p_ready", "loop"]
print("phases:", " | ".join(f"{i}={n}" for i, n in enumerate(names)))
for cta in range(0, 32, 3):
    for t in range(2):
        r = [buf[(cta * 2 + t) * 16 + i] for i in range(9)]
        it = max(r[8], 1)
        print(f"cta {cta:2d} thread {'0  ' if t == 0 else '255'} iters {it}: " + "  ".join(f"{x / it:7.0f}" for x in r[:8]) +
              f"   | total/iter {sum(r[:8]) / it:7.0f} clk | softmax loop lifetime {buf[(cta * 2 + t) * 16 + 9]} clk, end stamp "
              f"{(buf[(cta * 2 + t) * 16 + 10] - min(buf[(c * 2) * 16 + 10] for c in range(32))) / 1e3:.1f} us after the first CTA's end")
