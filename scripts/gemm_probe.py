"""Bring-up probe for vlk_gemm_bf16: runs each case in its own subprocess under a timeout so a hung
kernel cannot wedge the whole gpurun call.  Usage: python scripts/gemm_probe.py [case-index]"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [
    # M, N, K, transA, transB, bn[, cluster]   (cluster 3 = cta_group::2 pair)
    (256, 256, 64, 0, 0, 256, 3), (256, 256, 256, 0, 0, 256, 3), (256, 128, 128, 0, 0, 128, 3),
    (512, 512, 768, 0, 0, 256, 3), (300, 264, 72, 0, 0, 128, 3), (256, 256, 128, 0, 1, 256, 3),
    (256, 128, 128, 1, 0, 128, 3), (768, 768, 2112, 1, 1, 128, 3), (4096, 768, 2304, 0, 1, 256, 3),
    (16448, 1024, 1024, 0, 0, 256, 3), (16448, 4096, 1024, 0, 0, 256, 3), (16448, 4096, 1024, 0, 0, 128, 3),
    (16448, 3072, 1024, 0, 0, 256, 3), (16448, 1024, 4096, 0, 0, 256, 3),
    (4096, 2304, 768, 0, 0, 256, 3), (4096, 3072, 768, 0, 0, 256, 3), (4096, 768, 3072, 0, 0, 128, 3),
    (4096, 768, 3072, 0, 0, 256, 3), (4096, 768, 768, 0, 0, 128, 3), (512, 50304, 768, 0, 0, 256, 3),
    (512, 768, 50304, 0, 1, 128, 3),
]


def run_case(i):
    import torch
    M, N, K, ta, tb, bn = CASES[i][:6]
    cl = CASES[i][6] if len(CASES[i]) > 6 else 0
    if bn:
        os.environ["VLK_GEMM_BN"] = str(bn)
    if cl:
        os.environ["VLK_GEMM_CLUSTER"] = str(cl)
    lib = ctypes.CDLL(os.path.join(ROOT, "gpt2-vision-language_b200", "libvlk.so"))
    lib.vlk_last_error_string.restype = ctypes.c_char_p
    vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.vlk_gemm_bf16.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp, vp, ci, vp, vp, ci, vp, ci, ci,
                                  cf, ci, ci, vp]
    torch.manual_seed(i)
    a = torch.randn((K, M) if ta else (M, K), device="cuda").bfloat16()
    b = torch.randn((K, N) if tb else (N, K), device="cuda").bfloat16()
    d = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    rc = lib.vlk_gemm_bf16(a.data_ptr(), b.data_ptr(), d.data_ptr(), M, N, K, a.stride(0), b.stride(0), N, ta, tb,
                           0, 0, 0, 0, 0, 0, 0, 0, 0, 1.0, 0, 1, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        print(f"case {i} {CASES[i]}: rc={rc} {lib.vlk_last_error_string()}")
        return 1
    torch.cuda.synchronize()
    A = a.float().t() if ta else a.float()
    B = b.float() if tb else b.float().t()
    ref = A @ B
    err = (d.float() - ref).abs().max().item()
    rel = err / (ref.abs().max().item() + 1e-9)
    # timing
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        lib.vlk_gemm_bf16(a.data_ptr(), b.data_ptr(), d.data_ptr(), M, N, K, a.stride(0), b.stride(0), N, ta, tb,
                          0, 0, 0, 0, 0, 0, 0, 0, 0, 1.0, 0, 1, torch.cuda.current_stream().cuda_stream)
    s.record()
    for _ in range(20):
        lib.vlk_gemm_bf16(a.data_ptr(), b.data_ptr(), d.data_ptr(), M, N, K, a.stride(0), b.stride(0), N, ta, tb,
                          0, 0, 0, 0, 0, 0, 0, 0, 0, 1.0, 0, 1, torch.cuda.current_stream().cuda_stream)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    print(f"case {i} {CASES[i]}: max_abs_err={err:.4f} rel={rel:.2e} {'OK' if rel < 2e-2 else 'MISMATCH'} "
          f"{ms*1e3:.1f} us {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    return 0 if rel < 2e-2 else 2


if __name__ == "__main__":
    if len(sys.argv) > 1:
        sys.exit(run_case(int(sys.argv[1])))
    bad = 0
    for i in range(len(CASES)):
        try:
            r = subprocess.run([sys.executable, __file__, str(i)], timeout=90)
            if r.returncode != 0:
                bad += 1
                print(f"case {i}: exit {r.returncode}", flush=True)
        except subprocess.TimeoutExpired:
            bad += 1
            print(f"case {i} {CASES[i]}: TIMEOUT (hung kernel)", flush=True)
            break  # GPU state is suspect after a hang; stop here
    print("probe done, failures:", bad)
