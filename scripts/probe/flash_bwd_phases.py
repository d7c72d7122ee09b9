"""Bring-up build only: cycles per phase of the persistent backward kernels (softmax group 0 thread 0, group 1 thread 256),
summed over a CTA's sub-iterations, for the first 20 CTAs.  Both kernels dump into the same table and the dQ kernel runs
last, so the numbers are the dQ kernel's; stamp the dK/dV kernel alone by disabling the dQ dump in the bring-up build."""
import ctypes, os, sys
import torch
here = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(here, os.environ.get("VLK_PROBE_LIB", "libvlk_bringup.so")))
B, H, T, CAUSAL = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (16, 12, 1024, 1)))
C = H * 64
qkv = (torch.randn(B, T, 3 * C, device="cuda") * 0.5).bfloat16()
o = torch.empty(B, T, C, device="cuda", dtype=torch.bfloat16)
do = torch.randn_like(o)
dqkv = torch.empty_like(qkv)
lse = torch.empty(B, H, T, device="cuda")
delta = torch.empty(B * H * T, device="cuda")
q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
dq, dk, dv = dqkv[..., :C], dqkv[..., C:2 * C], dqkv[..., 2 * C:]
vp, ll, ci, cf, cu = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_float, ctypes.c_uint
lib.vlk_attn_fwd.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, ll, ci, ll, ci, ll, ci, ll, ci, ci, cf, cf, vp, cu, vp]
lib.vlk_attn_bwd.argtypes = [vp] * 9 + [ci] * 4 + [ll, ci] * 7 + [ci, cf, vp, cf, vp, cu, vp]
st = torch.cuda.current_stream().cuda_stream
rc = lib.vlk_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(), B, H, T, T, q.stride(0), q.stride(1),
                      k.stride(0), k.stride(1), v.stride(0), v.stride(1), o.stride(0), o.stride(1), CAUSAL, 0.125, 0.0, None, 0, st)
assert rc == 0, rc
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(4):
    e0.record()
    rc = lib.vlk_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), do.data_ptr(), lse.data_ptr(), dq.data_ptr(),
                          dk.data_ptr(), dv.data_ptr(), B, H, T, T, q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0),
                          v.stride(1), o.stride(0), o.stride(1), dq.stride(0), dq.stride(1), dk.stride(0), dk.stride(1), dv.stride(0),
                          dv.stride(1), CAUSAL, 0.125, delta.data_ptr(), 0.0, None, 0, st)
    e1.record()
    assert rc == 0, rc
    torch.cuda.synchronize()
print(f"B={B} H={H} T={T} causal={CAUSAL}: backward {e0.elapsed_time(e1) * 1e3:.1f} us (eager, all launches of the call)")
buf = (ctypes.c_longlong * (64 * 16))()
lib.vlk_debug_flash_dump(buf, 64 * 16)
which = os.environ.get("VLK_PROBE_KERNEL", "dq")   # the LAST kernel of the call that dumps wins: dq
grp_dq = ["wait S/dP", "tcgen05.ld + arrive", "exp2 / dS / pack / mask", "wait product(k-1) + tcgen05.st + arrive", "read-out (wait item, ld, stores)", "loop + item setup", "-", "-", "-", "-"]
grp = ["stats + group barrier", "wait S^T/dP^T", "tcgen05.ld", "exp2 / dS / pack", "wait products(k-1) + tcgen05.st + arrive", "read-out tcgen05.ld + arrive", "wait item",
       "sub-iteration loop", "8=stores", "9=next item setup"]
print("group phases (dQ kernel: it runs last and overwrites the dump):", " | ".join(f"{i}={n}" for i, n in enumerate(grp_dq)))
for cta in range(0, 20, 3):
    for t, name in enumerate(("group0 t0  ", "group1 t256")):   # (row 2 of a CTA is unused since the issue side became three warps)
        base = (cta * 3 + t) * 16
        r = [buf[base + i] for i in range(9)]
        it = max(r[8], 1)
        extra = [buf[base + 11 + i] for i in range(4)]
        print(f"cta {cta:2d} {name} sub-iters {it}: " + "  ".join(f"{x / it:6.0f}" for x in r[:8]) + "  +" +
              "  ".join(f"{x / it:6.0f}" for x in extra) +
              f"   | total/sub-iter {(sum(r[:8]) + sum(extra)) / it:6.0f} clk | lifetime {buf[base + 9]} clk")
