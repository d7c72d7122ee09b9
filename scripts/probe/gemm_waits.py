"""Bring-up build only: where the role threads of the cta_group::2 GEMM wait.  For each shape: graph-free CUDA-event time,
and per-CTA cycle counters (median over CTAs) — producer waiting for a free smem slot; MMA issuer waiting for a free
accumulator (epilogue too slow) and for operands (loads too slow) out of its total; epilogue warp 4: staged-input
prefetch, waiting for the accumulator, working.  Build: scripts/probe/build_bringup.sh."""
import ctypes, os, sys
import torch
here = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(here, os.environ.get("VLK_PROBE_LIB", "libvlk_bringup.so")))
vp, ci, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
lib.vlk_gemm_bf16.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp, vp, ci, vp, vp, ci, vp, ci, ci, cf, ci, ci, vp]
lib.vlk_gemm_bf16_lnfold.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, vp, vp, vp, vp, ci, vp]
lib.vlk_gemm_bf16_stats.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, vp, vp, ci, vp, vp]
lib.vlk_bringup_set_gemm_debug.argtypes = [vp]
dev = "cuda"
BF = torch.bfloat16
dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
ACT = {"none": 0, "gelu_tanh": 1, "gelu_erf": 2, "quick_gelu": 3}
st = lambda: torch.cuda.current_stream().cuda_stream


def run(name, M, N, K, kind="plain", act="none"):
    a = (torch.randn(M, K, device=dev) * 0.5).to(BF)
    w = (torch.randn(N, K, device=dev) * 0.05).to(BF)
    bias = torch.randn(N, device=dev).to(BF)
    d = torch.empty(M, N, device=dev, dtype=BF)
    res = torch.randn(M, N, device=dev).to(BF)
    mean, rstd = torch.zeros(M, device=dev), torch.ones(M, device=dev)
    colsum = torch.zeros(N, device=dev)
    stats = torch.zeros(M, 2, device=dev)
    aux = torch.empty(M, N, device=dev, dtype=BF)

    def call():
        if kind == "lnfold":
            return lib.vlk_gemm_bf16_lnfold(a.data_ptr(), w.data_ptr(), d.data_ptr(), M, N, K, K, K, N, bias.data_ptr(),
                                            mean.data_ptr(), rstd.data_ptr(), colsum.data_ptr(), ACT[act], st())
        if kind == "stats":
            return lib.vlk_gemm_bf16_stats(a.data_ptr(), w.data_ptr(), d.data_ptr(), M, N, K, K, K, N, bias.data_ptr(),
                                           res.data_ptr(), N, stats.data_ptr(), st())
        r = res.data_ptr() if kind == "residual" else None
        ao = aux.data_ptr() if kind == "aux" else None
        return lib.vlk_gemm_bf16(a.data_ptr(), w.data_ptr(), d.data_ptr(), M, N, K, K, K, N, 0, 0, bias.data_ptr(), r, N, None,
                                 ao, N, None, ACT[act], 0, 1.0, 0, 1, st())
    lib.vlk_bringup_set_gemm_debug(None)
    for _ in range(3):
        assert call() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    lib.vlk_bringup_set_gemm_debug(dbg.data_ptr())
    dbg.zero_()
    call()
    torch.cuda.synchronize()
    lib.vlk_bringup_set_gemm_debug(None)
    c = dbg.view(148, 8).double().cpu()
    lead = c[0::2]       # leader CTAs carry the MMA counters
    med = lambda t: t.median().item()
    tot = med(lead[:, 3])
    print(f"{name:44s} {us:7.1f} us {2.0 * M * N * K / us / 1e6:6.0f} TF/s | MMA thread: total {tot:9.0f} clk, waits acc "
          f"{100 * med(lead[:, 1]) / tot:4.1f} % operands {100 * med(lead[:, 2]) / tot:4.1f} % | producer slot wait "
          f"{100 * med(c[:, 0]) / tot:4.1f} % | epi warp: prefetch {100 * med(c[:, 4]) / tot:4.1f} % wait {100 * med(c[:, 5]) / tot:4.1f} % "
          f"work {100 * med(c[:, 6]) / tot:4.1f} %", flush=True)


M = 16448
if os.environ.get("VLK_PROBE_SWITCHES"):
    # bring-up switches: 0 = full kernel, 1 = no epilogue at all, 4 = epilogue reads the accumulator (tcgen05.ld) and
    # drops it, 8 = epilogue computes but never stores
    for dbgv in ("0", "1", "4", "8"):
        os.environ["VLK_GEMM_DEBUG"] = dbgv
        print("VLK_GEMM_DEBUG =", dbgv)
        run("CLIP fc1 plain store (bias)", M, 4096, 1024)
        run("CLIP fc1 lnfold+bias+quick_gelu", M, 4096, 1024, "lnfold", "quick_gelu")
        run("GPT-2 pretrain c_attn bias", 16384, 2304, 768)
        run("CLIP fc2 bias (K=4096)", M, 1024, 4096)
        run("plain 8192^3", 8192, 8192, 8192)
    sys.exit(0)
run("CLIP qkv lnfold+bias", M, 3072, 1024, "lnfold")
run("CLIP fc1 lnfold+bias+quick_gelu", M, 4096, 1024, "lnfold", "quick_gelu")
run("CLIP fc1 plain store (bias)", M, 4096, 1024)
run("CLIP out_proj bias+residual+stats", M, 1024, 1024, "stats")
run("CLIP out_proj bias+residual", M, 1024, 1024, "residual")
run("CLIP fc2 bias+residual+stats", M, 1024, 4096, "stats")
run("GPT-2 pretrain c_fc bias+gelu+aux", 16384, 3072, 768, "aux", "gelu_tanh")
run("GPT-2 pretrain c_attn bias", 16384, 2304, 768)
run("GPT-2 pretrain c_proj bias+residual", 16384, 768, 768, "residual")
run("GPT-2 pretrain mlp c_proj bias+residual", 16384, 768, 3072, "residual")
run("GPT-2 caption c_attn bias", 4096, 2304, 768)
run("GPT-2 caption c_proj bias+residual", 4096, 768, 768, "residual")
run("GPT-2 caption c_fc bias+gelu+aux", 4096, 3072, 768, "aux", "gelu_tanh")
run("GPT-2 caption mlp c_proj bias+residual", 4096, 768, 3072, "residual")
run("plain 8192^3", 8192, 8192, 8192)
