"""In-process timing sweep of vlk_gemm_bf16 over tile width / cluster / rasterisation band (env overrides)."""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpt2_vision_language_b200 import ops  # noqa: E402

SHAPES = [(16448, 4096, 1024, False), (16448, 1024, 1024, False), (16448, 3072, 1024, False),
          (16448, 1024, 4096, False), (4096, 2304, 768, False), (4096, 3072, 768, False), (4096, 768, 3072, False),
          (4096, 768, 768, False), (512, 50304, 768, False), (512, 768, 50304, True), (4096, 768, 2304, True)]


def bench(a, b, tb, iters=20):
    for _ in range(3):
        ops.gemm(a, b, trans_b=tb)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        ops.gemm(a, b, trans_b=tb)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    torch.manual_seed(0)
    for (M, N, K, tb) in SHAPES:
        a = torch.randn(M, K, device="cuda").bfloat16()
        b = torch.randn((K, N) if tb else (N, K), device="cuda").bfloat16()
        ref = a.float() @ (b.float() if tb else b.float().t())
        scale = ref.abs().max().item()
        rows = []
        for bn, cl, grp in itertools.product((256, 128), (1, 2, 3), (8,)):
            os.environ.update(VLK_GEMM_BN=str(bn), VLK_GEMM_CLUSTER=str(cl), VLK_GEMM_GROUP=str(grp))
            out = ops.gemm(a, b, trans_b=tb)
            err = (out.float() - ref).abs().max().item() / scale
            ms = bench(a, b, tb)
            rows.append((2.0 * M * N * K / ms / 1e9, bn, cl, grp, ms * 1e3, err))
        for k in ("VLK_GEMM_BN", "VLK_GEMM_CLUSTER", "VLK_GEMM_GROUP"):
            os.environ.pop(k, None)
        ms = bench(a, b, tb)
        print(f"== M={M} N={N} K={K} transB={tb}: default heuristics {2.0*M*N*K/ms/1e9:.0f} TFLOP/s ({ms*1e3:.1f} us)")
        for tf, bn, cl, grp, us, err in sorted(rows, reverse=True)[:6]:
            print(f"   {tf:7.0f} TF  bn={bn} cluster={cl} group={grp:<6d} {us:7.1f} us  err={err:.1e}{'  MISMATCH' if err > 2e-2 else ''}")
        bad = [r for r in rows if r[5] > 2e-2]
        if bad:
            print("   MISMATCHES:", bad)
        sys.stdout.flush()


def epilogues():
    """Effect of the fused epilogues on the biggest CLIP GEMMs (2-CTA vs multicast pair)."""
    torch.manual_seed(1)
    for (M, N, K) in [(16448, 4096, 1024), (16448, 1024, 4096), (16448, 3072, 1024), (4096, 3072, 768), (4096, 768, 3072)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        b = (torch.randn(N, K, device="cuda") * 0.03).bfloat16()
        bias = torch.randn(N, device="cuda").bfloat16()
        res = torch.randn(M, N, device="cuda").bfloat16()
        for name, kw in (("plain", {}), ("bias", dict(bias=bias)), ("bias+quick_gelu", dict(bias=bias, act="quick_gelu")),
                         ("bias+gelu_tanh+aux", dict(bias=bias, act="gelu_tanh", aux_out=True)),
                         ("bias+residual", dict(bias=bias, residual=res))):
            line = f"   M={M} N={N} K={K} {name:20s}"
            for cl in (1, 2, 3):
                os.environ.update(VLK_GEMM_BN="256", VLK_GEMM_CLUSTER=str(cl))
                for _ in range(3):
                    ops.gemm(a, b, **kw)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(20):
                    ops.gemm(a, b, **kw)
                e.record()
                torch.cuda.synchronize()
                ms = s.elapsed_time(e) / 20
                line += f"  cl{cl}: {2.0*M*N*K/ms/1e9:5.0f} TF"
            print(line, flush=True)
    for k in ("VLK_GEMM_BN", "VLK_GEMM_CLUSTER"):
        os.environ.pop(k, None)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "epi":
        epilogues()
    else:
        main()
        epilogues()
