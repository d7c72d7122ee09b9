"""A/B of the fused lm_head + cross-entropy (vlk_lmhead_ce_fwd/_bwd: no logits in HBM, logit tiles recomputed in the
backward) against the round-1 sequence (lm_head GEMM -> bf16 logits in HBM -> vlk_softmax_ce_rows in place -> d h / d W
GEMMs), graph-timed in one process on one clock.  Shapes: the caption step's 1,984 text rows (frozen head) and a
pretraining micro-batch of 16,384 rows (trainable head)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpt2_vision_language_b200 import _lib, ops  # noqa: E402

dev, BF = "cuda", torch.bfloat16
C, V = 768, 50304
out = []


def timeit(name, fn, iters=10):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (2 * iters) * 1e3
    print(f"{name:60s} {us:9.1f} us", flush=True)
    out.append({"name": name, "us": us})
    return us


for rows, need_dw in ((1984, False), (16384, True)):
    h = torch.randn(rows, C, device=dev).to(BF)
    w = (torch.randn(V, C, device=dev) * 0.02).to(BF)
    labels = torch.randint(0, V, (rows,), device=dev)
    dw = torch.zeros(V, C, device=dev, dtype=BF)
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream

    geo = [0, 0]

    def fused():
        stats, loss_row, lse = ops._lmhead_ce_forward(h, w, labels, None)
        dh = torch.empty_like(h)
        nb = int(lib.vlk_lmhead_ce_workspace_bytes(rows, C, V, 1, geo[0], geo[1]))
        ws = torch.empty(nb, device=dev, dtype=torch.uint8)
        _lib.check(lib.vlk_lmhead_ce_bwd(h.data_ptr(), w.data_ptr(), labels.data_ptr(), 0, lse.data_ptr(), stats[1:].data_ptr(),
                                         0, dh.data_ptr(), dw.data_ptr() if need_dw else 0, 1, rows, C, V, C, C, C, C,
                                         geo[0], geo[1], ws.data_ptr(), nb, torch.cuda.current_stream().cuda_stream), "bwd")

    def fused_fwd_only():
        ops._lmhead_ce_forward(h, w, labels, None)

    def round1():
        stats = torch.zeros(2, device=dev)
        loss_row = torch.empty(rows, device=dev)
        s_ = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.vlk_ce_count(labels.data_ptr(), 0, stats.data_ptr(), rows, s_), "count")
        chunk = 2048 if not need_dw else 16384
        logits = torch.empty((min(chunk, rows), V), device=dev, dtype=BF)
        dh = torch.empty_like(h)
        for r0 in range(0, rows, chunk):
            r1 = min(rows, r0 + chunk)
            lg = logits[: r1 - r0]
            ops.gemm(h[r0:r1], w, out=lg)
            _lib.check(lib.vlk_softmax_ce_rows(lg.data_ptr(), labels[r0:r1].data_ptr(), 0, loss_row[r0:r1].data_ptr(),
                                               stats[1:].data_ptr(), r1 - r0, V, V, 1, s_), "ce")
            ops.gemm(lg, w, trans_b=True, out=dh[r0:r1], split_k=ops.auto_split_k(r1 - r0, C, V))
            if need_dw:
                ops.gemm(lg, h[r0:r1], trans_a=True, trans_b=True, out=dw, residual=dw)
        _lib.check(lib.vlk_ce_finalize(loss_row.data_ptr(), 0, stats.data_ptr(), rows, s_), "fin")

    tag = f"rows={rows} dW={need_dw}"
    a = timeit(f"fused fwd+bwd (auto geometry) {tag}", fused)
    sweep = ((0, 25152), (0, 50304)) if not need_dw else ((4096, 12576), (4096, 50304), (8192, 50304), (16384, 12576), (16384, 50304))
    for rb_, vc_ in sweep:
        geo[0], geo[1] = rb_, vc_
        timeit(f"fused fwd+bwd row_block={rb_ or 'auto'} chunk_cols={vc_} {tag}", fused)
    geo[0] = geo[1] = 0
    f = timeit(f"fused fwd only           {tag}", fused_fwd_only)
    b = timeit(f"round-1 GEMM->HBM->CE->GEMM {tag}", round1)
    units = 3 if need_dw else 2
    print(f"   algorithmic {units} products = {units * 2.0 * rows * V * C / 1e12:.3f} TFLOP; fused {units * 2.0 * rows * V * C / a / 1e6:.0f} "
          f"TFLOP/s (algorithmic), round-1 {units * 2.0 * rows * V * C / b / 1e6:.0f} TFLOP/s")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/lmhead_ce_ab.json", "w"), indent=1)
