"""Sustained (power-capped) throughput of one GEMM shape: our kernel vs cuBLAS (torch.matmul), each looped for `secs`
seconds while nvidia-smi samples clocks and power.  Shows whether a shape is bounded by the 1000 W cap rather than by cycles."""
import ctypes, os, subprocess, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gpt2_vision_language_b200 import _lib, ops
if os.environ.get("VLK_PROBE_LIB"):
    _lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), os.environ["VLK_PROBE_LIB"])
secs = float(os.environ.get("SECS", "3"))
BF = torch.bfloat16


def sample(stop, out):
    while not stop.is_set():
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True)
        try:
            a, b = r.stdout.strip().split(",")
            out.append((float(a), float(b)))
        except Exception:
            pass
        stop.wait(0.1)


def loop(name, fn, flops):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(50):
            fn()
    g.replay()
    torch.cuda.synchronize()
    stop, out = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, out), daemon=True)
    th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    n = 0
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(20):
            g.replay()
        n += 20 * 50
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    stop.set()
    th.join()
    us = e0.elapsed_time(e1) / n * 1e3
    tail = out[len(out) // 2:]
    mhz = sorted(x[0] for x in tail)[len(tail) // 2] if tail else 0
    w = sorted(x[1] for x in tail)[len(tail) // 2] if tail else 0
    print(f"{name:50s} {us:8.1f} us {flops / us / 1e6:7.0f} TF/s   sm {mhz:5.0f} MHz  {w:5.0f} W", flush=True)


def shape(M, N, K, tag):
    a = (torch.randn(M, K, device="cuda") * 0.5).to(BF)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(BF)
    bias = torch.randn(N, device="cuda").to(BF)
    d = torch.empty(M, N, device="cuda", dtype=BF)
    fl = 2.0 * M * N * K
    loop(f"{tag} {M}x{N}x{K} vlk (bias)", lambda: ops.gemm(a, w, bias=bias, out=d), fl)
    loop(f"{tag} {M}x{N}x{K} cuBLAS addmm (bias)", lambda: torch.addmm(bias, a, w.t(), out=d), fl)


shape(16448, 4096, 1024, "CLIP fc1")
shape(16448, 1024, 4096, "CLIP fc2")
shape(16448, 3072, 1024, "CLIP qkv")
shape(16448, 1024, 1024, "CLIP out_proj")
shape(16384, 3072, 768, "GPT-2 c_fc")
shape(4096, 3072, 768, "caption c_fc")
shape(8192, 8192, 8192, "8192^3")
