"""Every kernel class of the path at the shapes the bench workloads use, timed alone with CUDA events.

Memory-bound kernels rotate over enough independent buffers that each launch reads cold data (working set per
rotation > the 126 MB L2); the table reports algorithmic bytes / time against MEASURED_PEAKS.json
(burst figures: kernels timed alone).  `--once` launches each kernel exactly once after one warm-up launch — the
mode used under `ncu --set full` (profiles/r01_ncu_kernels.*).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpt2_vision_language_b200 import _lib, ops, optim  # noqa: E402
if os.environ.get("VLK_ZOO_LIB"):      # A/B of two library builds on the same box
    _lib.LIB_PATH = os.path.abspath(os.environ["VLK_ZOO_LIB"])

ap = argparse.ArgumentParser()
ap.add_argument("--once", action="store_true")
ap.add_argument("--only", default="")
args = ap.parse_args()

peaks = {"hbm_gbs": 6451.2, "bf16_tflops": 1671.5}
try:
    peaks.update(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))))
except Exception:
    pass

dev = "cuda"
BF = torch.bfloat16
rows_out = []


def bench(name, make, fn, nbytes=None, flops=None, copies=None):
    """make(i) -> state for copy i; fn(state) launches the kernel(s)."""
    if args.only and args.only not in name:
        return
    if copies is None:
        copies = 1 if nbytes is None else max(2, min(16, int(400e6 // max(nbytes, 1)) + 1))
    if args.once:
        st = make(0)
        torch.cuda.synchronize()
        print("once:", name, flush=True)
        fn(st)
        torch.cuda.synchronize()
        return
    states = [make(i) for i in range(copies)]
    fn(states[0])
    torch.cuda.synchronize()
    iters = max(copies * 3, 12)
    for i in range(copies):
        fn(states[i])
    torch.cuda.synchronize()
    # one CUDA graph of `iters` launches: no host launch overhead between kernels (small kernels would otherwise
    # be timed at the speed of the Python/ctypes call path)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(states[0])
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            fn(states[i % copies])
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(3):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / (3 * iters) * 1e3
    del g
    line = f"{name:58s} {us:9.1f} us"
    rec = {"kernel": name, "us": us}
    if nbytes is not None:
        gbs = nbytes / us / 1e3
        line += f"  {nbytes/1e6:8.1f} MB  {gbs:7.0f} GB/s  {gbs/peaks['hbm_gbs']:.2f} of HBM copy peak"
        rec.update(bytes=nbytes, gbs=gbs, frac_hbm=gbs / peaks["hbm_gbs"])
    if flops is not None:
        tf = flops / us / 1e6
        line += f"  {tf:7.0f} TFLOP/s  {tf/peaks['bf16_tflops']:.2f} of bf16 burst peak"
        rec.update(flops=flops, tflops=tf, frac_tensor=tf / peaks["bf16_tflops"])
    print(line, flush=True)
    rows_out.append(rec)
    del states
    torch.cuda.empty_cache()


torch.manual_seed(0)

# ---------------------------------------------------------------- LayerNorm (CLIP rows x 1024, GPT rows x 768)
for rows, C, tag in ((64 * 257, 1024, "CLIP B=64"), (64 * 64, 768, "GPT-2 caption B=64,T=64"), (16 * 1024, 768, "pretrain 16x1024")):
    def mk(i, rows=rows, C=C):
        return (torch.randn(rows, C, device=dev).to(BF), torch.randn(C, device=dev).to(BF), torch.randn(C, device=dev).to(BF))
    bench(f"layernorm_fwd {tag} [{rows}x{C}]", mk, lambda s: ops.layernorm_fwd(s[0], s[1], s[2], 1e-5, save_stats=True),
          nbytes=rows * C * 2 * 2)

    def mkb(i, rows=rows, C=C):
        x = torch.randn(rows, C, device=dev).to(BF)
        w = torch.randn(C, device=dev).to(BF)
        _, mean, rstd = ops.layernorm_fwd(x, w, w, 1e-5)
        return (torch.randn(rows, C, device=dev).to(BF), x, w, mean, rstd, torch.empty_like(x))
    bench(f"layernorm_bwd dx only {tag}", mkb, lambda s: ops.layernorm_bwd(s[0], s[1], s[2], s[3], s[4], dx=s[5]),
          nbytes=rows * C * 2 * 3)
    if "pretrain" in tag:
        bench(f"layernorm_bwd dx+dgamma+dbeta {tag}", mkb,
              lambda s: ops.layernorm_bwd(s[0], s[1], s[2], s[3], s[4], param_grads=True, dx=s[5]), nbytes=rows * C * 2 * 3)

for N in (768, 3072):
    bench(f"colsum (bias gradient) pretrain [16384x{N}]", lambda i, N=N: torch.randn(16384, N, device=dev).to(BF),
          lambda s: ops.colsum(s), nbytes=16384 * N * 2)
bench("row_stats CLIP B=64 [16448x1024] (statistics of a folded LayerNorm)", lambda i: torch.randn(16448, 1024, device=dev).to(BF),
      lambda s: ops.row_stats(s, 1e-5), nbytes=16448 * 1024 * 2)

# ---------------------------------------------------------------- pool 257 -> 33 + L2 normalise
bench("pool33_l2norm B=64 D=768 (bf16)", lambda i: torch.randn(64, 257, 768, device=dev).to(BF), lambda s: ops.pool33(s),
      nbytes=64 * (257 + 33) * 768 * 2)
bench("pool33_l2norm B=1024 D=768 (bf16)", lambda i: torch.randn(1024, 257, 768, device=dev).to(BF), lambda s: ops.pool33(s),
      nbytes=1024 * (257 + 33) * 768 * 2, copies=2)

# ---------------------------------------------------------------- embedding + concat
wte = torch.randn(50304, 768, device=dev).to(BF)
wpe = torch.randn(1024, 768, device=dev).to(BF)
bench("embed_concat caption B=64 P=33 T=31", lambda i: (torch.randint(0, 50257, (64, 31), device=dev), torch.randn(64, 33, 768, device=dev).to(BF)),
      lambda s: ops.embed_concat(s[0], wte, wpe, s[1]), nbytes=64 * (31 * 3 + 33 * 2) * 768 * 2)
bench("embed pretrain B=16 T=1024", lambda i: torch.randint(0, 50257, (16, 1024), device=dev),
      lambda s: ops.embed_concat(s, wte, wpe, None), nbytes=16 * 1024 * 768 * 2 * 3)

# ---------------------------------------------------------------- softmax-CE over a logits chunk (in place)
lib = ops._lib.load()
for rows, tag in ((512, "caption chunk"), (4096, "pretrain chunk")):
    V = 50304

    def mkce(i, rows=rows):
        return (torch.randn(rows, V, device=dev).to(BF), torch.randint(0, 50257, (rows,), device=dev),
                torch.empty(rows, device=dev), torch.ones(1, device=dev))

    def runce(s, rows=rows):
        ops.check(lib.vlk_softmax_ce_rows(s[0].data_ptr(), s[1].data_ptr(), 0, s[2].data_ptr(), s[3].data_ptr(), rows, V,
                                          V, 1, torch.cuda.current_stream().cuda_stream), "ce")
    bench(f"softmax_ce_rows {tag} [{rows}x{V}] (cold logits)", mkce, runce, nbytes=rows * V * 2 * 2)

# ---------------------------------------------------------------- clip-norm + AdamW
for numel, tag in ((590592, "linear bridge"), (19521792, "Q-Former bridge"), (124475904, "GPT-2 124M")):
    def mkopt(i, numel=numel):
        p = torch.nn.Parameter(torch.randn(numel, device=dev).to(BF))
        p.grad = torch.randn(numel, device=dev).to(BF)
        o = optim.FusedAdamW([p], lr=1e-3, weight_decay=0.1)
        o.clip_grad_norm(1.0)
        o.step()
        return o
    copies = 2 if numel > 1e8 else (4 if numel > 1e7 else 16)
    bench(f"grad_sumsq {tag} ({numel} params)", mkopt, lambda o: o.clip_grad_norm(1.0), nbytes=numel * 2, copies=copies)

    def runstep(o):
        o._pending_max_norm = 1.0
        o.step()
    bench(f"adamw {tag} ({numel} params)", mkopt, runstep, nbytes=numel * 14, copies=copies)

# ---------------------------------------------------------------- attention
def attn_case(name, B, H, Tq, causal, bwd, copies=2):
    C = H * 64
    fl = 4.0 * B * H * Tq * Tq * 64 * (0.5 if causal else 1.0)

    def mk(i):
        qkv = (torch.randn(B, Tq, 3 * C, device=dev) * 0.5).to(BF)
        o, lse = ops.attention_fwd(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], H, causal)
        return qkv, o, lse, torch.randn_like(o), torch.empty_like(qkv)
    bench(f"attention fwd {name}", mk,
          lambda s: ops.attention_fwd(s[0][..., :C], s[0][..., C:2 * C], s[0][..., 2 * C:], H, causal, need_lse=bwd),
          flops=fl, copies=copies)
    if bwd:
        bench(f"attention bwd {name}", mk,
              lambda s: ops.attention_bwd(s[0][..., :C], s[0][..., C:2 * C], s[0][..., 2 * C:], s[1], s[3], s[2],
                                          s[4][..., :C], s[4][..., C:2 * C], s[4][..., 2 * C:], H, causal),
              flops=2.5 * fl, copies=copies)


attn_case("CLIP B=64 H=16 T=257", 64, 16, 257, False, False)
attn_case("GPT-2 caption B=64 H=12 T=64 causal", 64, 12, 64, True, True)
attn_case("pretrain B=16 H=12 T=1024 causal", 16, 12, 1024, True, True)
attn_case("Q-Former B=64 H=12 T=32", 64, 12, 32, False, True)

# ---------------------------------------------------------------- GEMMs of the step
def gemm_case(name, M, N, K, **kw):
    def mk(i):
        a = torch.randn(M, K, device=dev).to(BF)
        b = torch.randn(N, K, device=dev).to(BF)
        bias = torch.randn(N, device=dev).to(BF)
        res = torch.randn(M, N, device=dev).to(BF)
        return a, b, bias, res, torch.empty(M, N, device=dev, dtype=BF)

    def run(s):
        k2 = dict(kw)
        if k2.pop("use_bias", False):
            k2["bias"] = s[2]
        if k2.pop("use_res", False):
            k2["residual"] = s[3]
        ops.gemm(s[0], s[1], out=s[4], **k2)
    bench(f"gemm {name} M={M} N={N} K={K}", mk, run, flops=2.0 * M * N * K, copies=2)


def gemm_ln_case(name, M, N, K, act=None):
    def mk(i):
        x = torch.randn(M, K, device=dev).to(BF)
        w = (torch.randn(N, K, device=dev) * 0.03).to(BF)
        wf, cs, bf_ = ops.fold_layernorm(w, torch.randn(N, device=dev).to(BF), torch.ones(K, device=dev).to(BF),
                                         torch.zeros(K, device=dev).to(BF))
        mean, rstd = ops.row_stats(x, 1e-5)
        return x, wf, bf_, cs, (mean, rstd)
    bench(f"gemm {name} M={M} N={N} K={K}", mk, lambda s: ops.gemm_lnfold(s[0], s[1], s[2], s[3], 1e-5, act=act, stats=s[4]),
          flops=2.0 * M * N * K, copies=2)


gemm_ln_case("CLIP ln1+qkv (folded LayerNorm, bias)", 16448, 3072, 1024)
gemm_ln_case("CLIP ln2+fc1 (folded LayerNorm, bias+quick_gelu)", 16448, 4096, 1024, act="quick_gelu")
gemm_case("CLIP qkv (bias)", 16448, 3072, 1024, use_bias=True)
gemm_case("CLIP out_proj (bias+residual)", 16448, 1024, 1024, use_bias=True, use_res=True)
gemm_case("CLIP fc1 (bias+quick_gelu)", 16448, 4096, 1024, use_bias=True, act="quick_gelu")
gemm_case("CLIP fc2 (bias+residual)", 16448, 1024, 4096, use_bias=True, use_res=True)
gemm_case("GPT-2 c_fc pretrain (bias+gelu)", 16384, 3072, 768, use_bias=True, act="gelu_tanh")
gemm_case("GPT-2 lm_head chunk", 4096, 50304, 768)
gemm_case("plain 8192^3", 8192, 8192, 8192)
# the GPT-2 GEMMs of the caption step (M = 64 x 64 rows)
gemm_case("GPT-2 c_attn caption (bias)", 4096, 2304, 768, use_bias=True)
gemm_case("GPT-2 attn c_proj caption (bias+residual)", 4096, 768, 768, use_bias=True, use_res=True)
gemm_case("GPT-2 c_fc caption (bias+gelu)", 4096, 3072, 768, use_bias=True, act="gelu_tanh")
gemm_case("GPT-2 mlp c_proj caption (bias+residual)", 4096, 768, 3072, use_bias=True, use_res=True)
def gemm_tile_case(name, M, N, K, bn, pair, **kw):
    def mk(i):
        return (torch.randn(M, K, device=dev).to(BF), torch.randn(N, K, device=dev).to(BF), torch.randn(N, device=dev).to(BF),
                torch.randn(M, N, device=dev).to(BF))
    bench(f"gemm [tile_n={bn} pair={pair}] {name} M={M} N={N} K={K}", mk,
          lambda s: ops.gemm_tile(s[0], s[1], tile_n=bn, cta_pair=pair, bias=s[2],
                                  residual=s[3] if kw.get("use_res") else None, act=kw.get("act")),
          flops=2.0 * M * N * K, copies=2)


for bn, pair in ((256, 0), (128, 0), (128, 1), (64, 0)):
    gemm_tile_case("c_attn", 4096, 2304, 768, bn, pair)
    gemm_tile_case("attn c_proj", 4096, 768, 768, bn, pair, use_res=True)
    gemm_tile_case("c_fc", 4096, 3072, 768, bn, pair, act="gelu_tanh")
    gemm_tile_case("mlp c_proj", 4096, 768, 3072, bn, pair, use_res=True)

if not args.once:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"peaks": peaks, "kernels": rows_out}, open("gpurun_out/kernel_zoo.json", "w"), indent=1)
