#!/usr/bin/env python
"""Benchmark of the captioning training step (BASELINE.json metric, configs[1] by default).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

One "step" = frozen CLIP ViT-L/14 forward on B=64 synthetic 224x224 images -> 257->33 pool -> bridge -> frozen
GPT-2 124M forward+backward with chunked lm_head+CE -> gradient all-reduce -> clip-norm + AdamW (bridge only).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Algorithmic FLOPs per step at B=64: 2MNK per GEMM, 4*Tq*Tk*d per attention head,
# dgrad-only backward through frozen weights.
ALGO_TFLOP_PER_STEP_B64 = {"linear": 12.28, "qformer": 12.49, "xattn": 11.74}
PRETRAIN_GFLOP_PER_TOKEN = 0.8075      # T=1024, fwd + full bwd
PRETRAIN_TOKENS_PER_STEP = 524288      # 16 x 1024 x 32 micro-batches, split over the ranks (train_gpt2.py:244-248)
TEXT_LEN = 31


def sample_clocks(stop, out, index):
    """nvidia-smi clocks + throttle reasons every 200 ms while the timed region runs (B200_PROFILING.md)."""
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(index)],
                               capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split("\n")[0].split(",")]
            out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    sm = sorted(float(s[0]) for s in samples if s[0].replace(".", "").isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(len(s) > 3 + i and s[3 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(samples[0][1]) if samples else None,
            "power_w_max": max((float(s[2]) for s in samples if s[2].replace(".", "").isdigit()), default=None),
            "samples": len(samples), "reasons": reasons}


def synthetic_host_batch(B, seed, torch, pin):
    """Random 224x224 images and a 5-caption pool per image, one caption drawn per step (data.py:53 semantics)."""
    from gpt2_vision_language_b200.data import synthetic_caption_batch
    g = torch.Generator().manual_seed(seed)
    pixels = torch.randn(B, 3, 224, 224, generator=g)
    x, y, m, _ = synthetic_caption_batch(B, seed=seed + 1)
    if pin:
        pixels, x, y, m = (t.pin_memory() for t in (pixels, x, y, m))
    return pixels, x, y, m


# ======================================================================================================
# reference arm / CPU baseline: the oracle port of the reference algorithm, fp32, all host threads
# ======================================================================================================
def build_cpu_problem(torch, workload, seed=1337, from_gpu_state=None):
    """The oracle port of one caption train step (CLIP forward -> pool -> bridge -> GPT-2 -> CE -> backward ->
    clip_grad_norm_ + AdamW) on fp32 CPU tensors.  Returns (step_fn, state); state['first'] keeps the loss and the
    trainable gradients of the FIRST call (the parity leg compares them with the GPU's)."""
    from oracle import torch_oracle as O
    from gpt2_vision_language_b200.clip import ClipVisionTower
    if from_gpu_state is not None:
        sd, clip_sd = from_gpu_state
    else:
        model = build_host_model(torch, workload, seed)
        sd = {k: v.detach().float() for k, v in model.state_dict().items()}
        clip_sd = ClipVisionTower.random_state_dict(seed)
    train = [k for k in sd if trainable_key(workload, k)]
    for k in train:
        sd[k] = sd[k].clone().requires_grad_(True)
    state = dict(m=[torch.zeros_like(sd[k]) for k in train], v=[torch.zeros_like(sd[k]) for k in train], step=0,
                 first=None, names=train)

    def step(pixels, x, y, mask):
        with torch.no_grad():
            z = O.pool33(O.clip_features(clip_sd, pixels))
        for k in train:
            sd[k].grad = None
        if workload == "linear":
            _, loss = O.caption_linear_forward(sd, z, x, y.masked_fill(~mask, -100), 12, 12)
        elif workload == "qformer":
            _, loss = O.caption_qformer_forward(sd, z, x, y.masked_fill(~mask, -100), 12, 12)
        else:
            _, loss = O.xattn_forward(sd, x, z, y, mask, 12, 12)
        loss.backward()
        if state["first"] is None:
            state["first"] = (loss.item(), {k: sd[k].grad.detach().clone() for k in train})
        state["step"] += 1
        with torch.no_grad():
            O.clip_and_adamw([sd[k] for k in train], [sd[k].grad for k in train], state["m"], state["v"],
                             state["step"], 1e-3, [0.1 if sd[k].dim() >= 2 else 0.0 for k in train], 1.0)
        return loss.item()
    return step, state


def trainable_key(workload, k):
    if workload == "xattn":
        return ("xattn." in k) or ("vis_proj." in k) or k.endswith("cross_gate")
    return k.startswith("bridge.")


def build_host_model(torch, workload, seed=1337):
    """The captioner of one workload with the reference's init (random-init GPT-2 124M, vocab 50304), on the host."""
    from gpt2_vision_language_b200 import gpt2, gpt2_cross_att, gpt2_linear, gpt2_q_former
    torch.manual_seed(seed)
    if workload == "xattn":
        model = gpt2_cross_att.GPT(gpt2_cross_att.GPTConfig(vocab_size=50304))
        with torch.no_grad():
            for blk in model.transformer.h:     # non-zero gates: otherwise every x-attn gradient is exactly zero
                blk.cross_gate.copy_(torch.randn(()) * 0.5)
        return model
    lm = gpt2.GPT_previous(gpt2.GPTConfig(vocab_size=50304))
    mod = gpt2_linear if workload == "linear" else gpt2_q_former
    return mod.GPT_Caption(enc_dim=768, lm=lm, m_vis_tokens=32)


def time_cpu(torch, step_fn, batch, steps, warmup):
    for _ in range(warmup):
        step_fn(*batch)
    t0 = time.perf_counter()
    for _ in range(steps):
        step_fn(*batch)
    return (time.perf_counter() - t0) / steps


REF_BATCH = 16     # the reference arm's batch: a CONSTANT slice of the B=64 workload, the same on every box


def run_reference(args, out=sys.stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step_fn, _ = build_cpu_problem(torch, args.workload)
    B = args.ref_batch
    pixels, x, y, m = synthetic_host_batch(B, seed=0, torch=torch, pin=False)
    dt = time_cpu(torch, step_fn, (pixels, x, y, m), args.steps, args.warmup)
    v = B / dt
    line = {"impl": "reference", "metric": "caption_train_samples_per_s", "value": v, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"caption-{args.workload} step (CLIP ViT-L/14 fwd + pool + bridge + GPT-2 124M fwd/bwd + "
                                   f"clip+AdamW), reference algorithm on host CPU", "per_step_batch": B,
                       "note": f"fixed B={B} sample of the B=64 per-GPU workload; samples/s is per-sample comparable"},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} steps of B={B} (of the B=64 workload), fp32 torch, {cores} threads"},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=out, flush=True)
    return 0


# ======================================================================================================
# B200 arm
# ======================================================================================================
def build_gpu_problem(torch, args, workload, dev):
    from gpt2_vision_language_b200.clip import ClipVisionTower
    from gpt2_vision_language_b200.dp import broadcast_parameters
    from gpt2_vision_language_b200.step import CaptionTrainStep
    model = build_host_model(torch, workload).to(dev).to(torch.bfloat16)
    clip_sd = ClipVisionTower.random_state_dict(1337, device=dev)
    clip = ClipVisionTower.from_state_dict(clip_sd, device=dev)
    broadcast_parameters(model)
    # VLK_NO_OVERLAP=1 switches the split-backward / early all-reduce path off (A/B measurement only)
    step = CaptionTrainStep(model, clip, workload, args.batch, TEXT_LEN, use_graph=not args.no_graph,
                            overlap_comm=False if os.environ.get("VLK_NO_OVERLAP") else None)
    return model, clip, clip_sd, step


def dp_bucket_check(torch, dist, bucket, produce_local_grads):
    """Data-parallel correctness ON THE GPUs (DDP at train_gpt2.py:270,467-471): every rank computes
    its local gradients, the flat buckets are all-gathered, and the bucket after the NCCL all-reduce must equal the
    mean of the ranks' local buckets (up to the bf16 rounding of the collective's partial sums)."""
    world = dist.get_world_size()
    produce_local_grads()
    lo, hi = bucket.params_off, bucket.params_end
    local = bucket.flat[lo:hi].clone()
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    bucket.all_reduce()
    red = bucket.flat[lo:hi].float()
    mean = gathered[0].float()
    for g in gathered[1:]:
        mean += g.float()
    mean /= world
    differ = max((gathered[0].float() - g.float()).abs().max().item() for g in gathered[1:])
    scale = mean.abs().max().clamp_min(1e-30)
    out = {"ranks": world, "elements": int(hi - lo), "max_abs_err_over_max_abs": ((red - mean).abs().max() / scale).item(),
           "cosine": torch.nn.functional.cosine_similarity(red.double(), mean.double(), dim=0).item(),
           "local_buckets_differ_by": differ / scale.item()}
    out["ok"] = bool(out["max_abs_err_over_max_abs"] < 2e-2 and out["cosine"] > 0.9999 and differ > 0)
    t = torch.tensor([0.0 if out["ok"] else 1.0], device=red.device)
    dist.all_reduce(t)
    out["ok_all_ranks"] = bool(t.item() == 0)
    return out


def timed(torch, dist, world, steps, body):
    """K calls of body() bracketed by barrier + synchronize on both sides; CUDA events on the current stream."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(steps):
        body(i)
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / steps


def max_over_ranks(torch, dist, world, dev, values):
    t = torch.tensor(values, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def parity_leg(torch, workload, model, clip, clip_sd, step, batch, time_steps=2):
    """Rank 0, one GPU: the fp32 oracle (reference algorithm, CPU) and the B200 path on the SAME weights (the GPU
    model's bf16 values) and the same B=2 batch — first-step loss and trainable-gradient cosine (north_star: loss within
    2e-3 relative, gradient cosine >= 0.999) — and, from the same oracle calls, the bounded CPU baseline."""
    from gpt2_vision_language_b200.caption import pool_clip_197_to_33_avg_with_cls
    pixels, x, y, m = batch
    pixels = pixels.to(torch.bfloat16).float()            # the device path rounds pixels to bf16 in im2col
    dev = step.dev
    was_training = model.training
    model.eval()                                           # Q-Former dropout off on both sides (SURVEY 8c pitfall 2)
    step.bucket.zero()
    z = pool_clip_197_to_33_avg_with_cls(clip(pixels.to(dev)))
    if workload == "xattn":
        _, loss = model(x.to(dev), z=z, targets=y.to(dev), target_mask=m.to(dev))
    else:
        _, loss = model(z, x.to(dev), labels=y.masked_fill(~m, -100).to(dev))
    loss.backward()
    torch.cuda.synchronize()
    loss_gpu = loss.item()
    g_gpu = {n: p.grad.detach().float().cpu().clone() for n, p in model.named_parameters() if p.requires_grad}
    model.train(was_training)
    step.bucket.zero()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    csd = {k: v.detach().float().cpu() for k, v in clip_sd.items()}
    step_fn, state = build_cpu_problem(torch, workload, from_gpu_state=(sd, csd))
    Bc = pixels.shape[0]
    if time_steps > 0:
        dt = time_cpu(torch, step_fn, (pixels, x, y, m), time_steps, 1)
    else:
        step_fn(pixels, x, y, m)
        dt = float("nan")
    loss_cpu, g_cpu = state["first"]
    names = [n for n in state["names"] if n in g_gpu and g_cpu[n].numel() > 1]
    flat_gpu = torch.cat([g_gpu[n].flatten() for n in names]).double()
    flat_cpu = torch.cat([g_cpu[n].flatten() for n in names]).double()
    cos_all = torch.nn.functional.cosine_similarity(flat_gpu, flat_cpu, dim=0).item()
    per = {n: torch.nn.functional.cosine_similarity(g_gpu[n].flatten().double(), g_cpu[n].flatten().double(), dim=0).item()
           for n in names}
    worst = min(per, key=per.get)
    cos_min = per[worst]
    share = (g_cpu[worst].double().norm() / flat_cpu.norm()).item()
    rel = abs(loss_gpu - loss_cpu) / abs(loss_cpu)
    parity = {"batch": Bc, "loss_gpu": loss_gpu, "loss_oracle": loss_cpu, "rel": rel, "grad_cos": cos_all,
              "grad_cos_min_tensor": cos_min, "grad_cos_min_tensor_name": worst,
              "grad_cos_min_tensor_share_of_grad_norm": share, "trainable_tensors": len(names),
              "ok": bool(rel < 2e-3 and cos_all > 0.999),
              "what": "first-step loss and trainable gradients, pixels -> CLIP ViT-L/14 -> pool -> bridge -> GPT-2 124M -> CE, "
                      "B200 path vs fp32 CPU oracle on the same bf16-rounded weights"}
    cpu = {"value": Bc / dt, "unit": "samples/s", "cores": cores, "kind": "port",
           "sample": f"{time_steps} steps of B={Bc} (of the B=64 workload) after 1 warm-up, oracle port of the reference "
                     f"algorithm (caption-{workload} incl. CLIP forward), fp32 torch, {cores} threads"}
    return parity, cpu


def run_caption(args, torch, dist, dev, world, rank, local, peaks, workload, steps, warmup, headline):
    """One captioning workload: device-resident throughput, end-to-end throughput, and (headline only) the GEMM
    roofline pass, clocks and the parity / CPU-baseline leg.  Returns the JSON-ready dict on rank 0, None elsewhere."""
    from gpt2_vision_language_b200 import _lib
    from gpt2_vision_language_b200.step import HostBatchFeeder
    stage(f"{workload}: building")
    model, clip, clip_sd, step = build_gpu_problem(torch, args, workload, dev)
    B = args.batch
    pixels_h, x_h, y_h, m_h = synthetic_host_batch(B, seed=rank, torch=torch, pin=True)
    step.load_batch(pixels_h, x_h, y_h, m_h)

    dp = None
    if world > 1:
        dp = dp_bucket_check(torch, dist, step.bucket, step._fwd_bwd)
    # launches per step, counted on one eager pass (graph replays do not go through the launcher)
    c0 = _lib.launch_count()
    step._body()
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - c0
    stage(f"{workload}: eager step done")

    warm = max(warmup, 3)
    for _ in range(warm + (3 if not args.no_graph else 0)):     # +3: two eager warm steps and the capture itself
        step.run()
    torch.cuda.synchronize()
    stage(f"{workload}: warm-up + capture done")

    clocks, stop = [], threading.Event()
    th = threading.Thread(target=sample_clocks, args=(stop, clocks, local), daemon=True)
    if rank == 0:
        th.start()
    # ---- device-resident throughput ("value") ---------------------------------------------------------
    ms = timed(torch, dist, world, steps, lambda i: step.run())
    # ---- end-to-end: pinned host batch -> H2D, step, loss -> D2H, every step ----------------------------
    loss_h = torch.zeros(steps, dtype=torch.float32).pin_memory()
    feeder = HostBatchFeeder(step)

    def e2e_body(i):
        if i == 0:
            feeder.submit(pixels_h, x_h, y_h, m_h)             # every step's batch is uploaded inside the timed region
        feeder.load()
        if i + 1 < steps:
            feeder.submit(pixels_h, x_h, y_h, m_h)             # next batch crosses PCIe while this step computes
        loss = step.run()
        loss_h[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
    ms_e2e = timed(torch, dist, world, steps, e2e_body)
    stop.set()
    ms, ms_e2e = max_over_ranks(torch, dist, world, dev, [ms, ms_e2e])
    if rank == 0:
        th.join(timeout=2)
    roof = parity = cpu = None
    if rank == 0 and headline:
        roof = gemm_roofline(torch, step, peaks)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Bc = 2
        parity, cpu = parity_leg(torch, workload, model, clip, clip_sd, step,
                                 (pixels_h[:Bc].clone(), x_h[:Bc].clone(), y_h[:Bc].clone(), m_h[:Bc].clone()),
                                 time_steps=2 if headline else 0)
        if not headline:
            cpu = None
    res = None
    if rank == 0:
        gb = B * world
        tflop = ALGO_TFLOP_PER_STEP_B64[workload] * B / 64.0
        peak = peaks.get("bf16_tflops_sustained") or 1400.0
        h2d = sum(t_.numel() * t_.element_size() for t_ in (pixels_h, x_h, y_h, m_h))
        res = {
            "metric": "caption_train_samples_per_s", "value": gb / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"caption-{workload} train step: frozen CLIP ViT-L/14 fwd + 257->33 pool + "
                                   f"{workload} bridge + frozen GPT-2 124M fwd/bwd + fused lm_head+CE + clip-norm+AdamW",
                       "global_batch": gb, "per_gpu_batch": B, "text_len": TEXT_LEN, "image": "3x224x224",
                       "parallelism": f"dp{world}", "cuda_graph": not args.no_graph,
                       "nccl_in_graph": bool(world > 1 and step.nccl_in_graph and not args.no_graph),
                       "overlap_comm": bool(getattr(step, "overlap", False)),
                       "l2": "no explicit flush: one step streams ~0.9 GB of weights plus >2 GB of activations, "
                             "far above the 126 MB L2"},
            "tokens_per_s": gb * TEXT_LEN / (ms * 1e-3),
            "algorithmic_tflops": tflop / (ms * 1e-3), "frac_of_bf16_sustained_peak": tflop / (ms * 1e-3) / peak,
            "frac_of_bf16_nominal_2250": tflop / (ms * 1e-3) / 2250.0,
            "clocks": summarize_clocks(clocks),
            "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(launches_per_step * steps), "launches_per_step": int(launches_per_step),
            "roofline": roof, "cpu_baseline": cpu, "parity": parity, "dp_check": dp, "final_loss": float(loss_h[-1]),
        }
    del feeder, step, model, clip, clip_sd
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def run_pretrain(args, torch, dist, dev, world, rank, local, peaks, steps, warmup, headline):
    """GPT-2 124M pretraining step (BASELINE.json configs[4]): tokens/s, strong scaling over the ranks."""
    from gpt2_vision_language_b200 import _lib, gpt2
    from gpt2_vision_language_b200.dp import broadcast_parameters
    from gpt2_vision_language_b200.step import PretrainStep
    torch.manual_seed(1337)
    model = gpt2.GPT(gpt2.GPTConfig(vocab_size=50304)).to(dev).to(torch.bfloat16)
    broadcast_parameters(model)
    mb, T = args.micro_batch, 1024
    accum = max(1, PRETRAIN_TOKENS_PER_STEP // (mb * T * world))
    step = PretrainStep(model, mb, T, accum, use_graph=not args.no_graph,
                        overlap_comm=False if os.environ.get("VLK_NO_OVERLAP") else None, zero1=args.zero1)
    g = torch.Generator().manual_seed(rank)
    x_h = torch.randint(0, 50257, (accum, mb, T), generator=g).pin_memory()
    y_h = torch.randint(0, 50257, (accum, mb, T), generator=g).pin_memory()
    step.load_tokens(x_h, y_h)

    dp = None
    if world > 1 and not args.zero1:
        def local_grads():
            step.bucket.zero()
            step._set_slot(0)
            step._micro()
        dp = dp_bucket_check(torch, dist, step.bucket, local_grads)
    c0 = _lib.launch_count()
    step.run()                      # eager warm-up step, also counts launches
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - c0
    warm = max(warmup, 3)
    for _ in range(warm):
        step.run()
    clocks, stop = [], threading.Event()
    th = threading.Thread(target=sample_clocks, args=(stop, clocks, local), daemon=True)
    if rank == 0 and headline:
        th.start()
    ms = timed(torch, dist, world, steps, lambda i: step.run())
    loss_h = torch.zeros(steps, dtype=torch.float32).pin_memory()

    def e2e_body(i):
        step.load_tokens(x_h, y_h)
        loss = step.run()
        loss_h[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
    ms_e2e = timed(torch, dist, world, steps, e2e_body)
    stop.set()
    # where one step's time goes (CUDA events between the phases of one more step)
    step.phase_events = []
    step.run()
    torch.cuda.synchronize()
    ev = step.phase_events
    step.phase_events = None
    phases = {f"{ev[i][0]}->{ev[i + 1][0]}": ev[i][1].elapsed_time(ev[i + 1][1]) for i in range(len(ev) - 1)}
    ms, ms_e2e = max_over_ranks(torch, dist, world, dev, [ms, ms_e2e])
    res = None
    if rank == 0:
        if headline:
            th.join(timeout=2)
        # roofline of the GEMMs: an instrumented eager micro-step (run twice: the first pass warms the allocator)
        with GemmRecorder(torch) as rec:
            for i in range(2):
                rec.records.clear()
                step.bucket.zero()
                step._set_slot(0)
                if i == 1:
                    torch.cuda._sleep(int(4e8))   # the host enqueues the pass while the GPU is parked (see gemm_roofline)
                step._micro()
            torch.cuda.synchronize()
        records = [(a, b, 2.0 * m * n * k) for a, b, (m, n, k), _ in rec.records]
        tot_ms = sum(a.elapsed_time(b) for a, b, _ in records)
        tot_fl = sum(f for _, _, f in records)
        peak = peaks.get("bf16_tflops_sustained") or 1400.0
        tokens = mb * T * accum * world
        tflop = PRETRAIN_GFLOP_PER_TOKEN * tokens / 1e3
        res = {
            "metric": "gpt2_pretrain_tokens_per_s", "value": tokens / (ms * 1e-3), "unit": "tokens/s", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "GPT-2 124M pretraining step, T=1024, AdamW, clip 1.0, grad accumulation",
                       "tokens_per_step": tokens, "micro_batch": mb, "seq_len": T, "grad_accum_per_rank": accum, "zero1": bool(args.zero1),
                       "parallelism": f"dp{world}", "cuda_graph": not args.no_graph,
                       "nccl_in_graph": bool(world > 1 and step.nccl_in_graph and not args.no_graph),
                       "overlap_comm": bool(getattr(step, "overlap", False)),
                       "l2": "no explicit flush: 250 MB of weights + GBs of activations per micro-step >> 126 MB L2"},
            "algorithmic_tflops": tflop / (ms * 1e-3), "frac_of_bf16_sustained_peak": tflop / (ms * 1e-3) / peak / world,
            "clocks": summarize_clocks(clocks) if headline else None,
            "e2e": {"value": tokens / (ms_e2e * 1e-3), "unit": "tokens/s",
                    "h2d_bytes_per_step": int(x_h.numel() * 16), "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
            "gpu_launches": int(launches_per_step * steps), "launches_per_step": int(launches_per_step),
            "phases_ms": phases,
            "roofline": {"bound": "tensor", "kernel": "vlk_gemm_bf16 (all launches of one micro-step)",
                         "launches_per_step": len(records), "achieved": tot_fl / (tot_ms * 1e-3) / 1e12, "peak": peak,
                         "unit": "TFLOP/s", "frac": tot_fl / (tot_ms * 1e-3) / 1e12 / peak,
                         "gemm_ms_per_micro_step": tot_ms, "traffic": None},
            "cpu_baseline": None, "dp_check": dp, "final_loss": float(loss_h[-1]),
        }
    del step, model
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


class GemmRecorder:
    """CUDA events around every tensor-core product the step launches through ops (ops.gemm, ops.gemm_lnfold,
    ops.wgrad); nested calls (wgrad -> gemm) are recorded once, at the outermost level."""

    def __init__(self, torch):
        from gpt2_vision_language_b200 import ops
        self.torch, self.ops, self.records, self.depth = torch, ops, [], 0
        self.orig = (ops.gemm, ops.gemm_lnfold, ops.wgrad, ops.gemm_stats)

    def _wrap(self, fn, shape_of, kind):
        def timed(*a, **kw):
            if self.depth:
                return fn(*a, **kw)
            e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            self.depth += 1
            e0.record()
            try:
                out = fn(*a, **kw)
            finally:
                self.depth -= 1
            e1.record()
            self.records.append((e0, e1, shape_of(*a, **kw), kind))
            return out
        return timed

    def __enter__(self):
        def gemm_shape(a, b, **kw):
            ta, tb = kw.get("trans_a", False), kw.get("trans_b", False)
            M, K = (a.shape[1], a.shape[0]) if ta else (a.shape[0], a.shape[1])
            return (M, b.shape[1] if tb else b.shape[0], K)
        o = self.ops
        o.gemm = self._wrap(self.orig[0], gemm_shape, "gemm")
        o.gemm_lnfold = self._wrap(self.orig[1], lambda x, wf, *a, **kw: (x.shape[0], wf.shape[0], x.shape[1]), "lnfold")
        o.wgrad = self._wrap(self.orig[2], lambda dy, x, *a, **kw: (dy.shape[1], x.shape[1], dy.shape[0]), "wgrad")
        o.gemm_stats = self._wrap(self.orig[3], lambda a_, w, *a, **kw: (a_.shape[0], w.shape[0], a_.shape[1]), "gemm")
        return self

    def __exit__(self, *exc):
        self.ops.gemm, self.ops.gemm_lnfold, self.ops.wgrad, self.ops.gemm_stats = self.orig


def gemm_roofline(torch, step, peaks):
    """Instrumented eager pass: CUDA events around every vlk_gemm_bf16 launch of one full step.  The roofline object
    is for the DOMINANT kernel = the GEMM launch shape with the largest total time in the step (CLIP fc1,
    16448 x 4096 x 1024 with bias + quick-GELU at B=64); the aggregate over all GEMM launches is reported beside it."""
    with GemmRecorder(torch) as rec:
        for i in range(2):           # the first pass warms the allocator: an allocation inside a timed window is host time
            rec.records.clear()
            if i == 1:
                # Park the GPU behind a spin kernel while the host enqueues the whole pass: the events then bracket GPU
                # time only.  (Without it the host — Python, ctypes, two cuTensorMapEncode per product — is slower than
                # the GPU on this path and its latency lands inside the event windows: the 117 us fc1 product read 134 us.)
                torch.cuda._sleep(int(4e8))
            step._fwd_bwd()          # rank-local: no collective here (only rank 0 runs this pass)
            step._update()
        torch.cuda.synchronize()
    records = rec.records
    by_shape = {}
    for e0, e1, shp, kind in records:
        t = e0.elapsed_time(e1)
        n, tot = by_shape.get((shp, kind), (0, 0.0))
        by_shape[(shp, kind)] = (n + 1, tot + t)
    tot_ms = sum(t for _, t in by_shape.values())
    tot_fl = sum(2.0 * m * n * k * cnt for ((m, n, k), _), (cnt, _) in by_shape.items())
    ((dm, dn, dk), dkind), (dcnt, dms) = max(by_shape.items(), key=lambda kv: kv[1][1])
    flops = 2.0 * dm * dn * dk
    achieved = flops / (dms / dcnt * 1e-3) / 1e12
    peak = peaks.get("bf16_tflops_sustained") or 1400.0
    traffic = None
    epi = ""
    try:   # DRAM bytes of one launch of that shape from the committed `ncu --set full` capture
        name = "r02_ncu_gemm_shapes.json" if os.path.exists(os.path.join(ROOT, "profiles", "r02_ncu_gemm_shapes.json")) \
            else "r01_ncu_gemm_shapes.json"
        cap = json.load(open(os.path.join(ROOT, "profiles", name)))
        for c in cap["shapes"]:
            if [dm, dn, dk] == c["shape"] and c["kind"] == dkind:
                traffic = c["dram_bytes_read"] + c["dram_bytes_write"]
                epi = ", " + c["epilogue"]
    except Exception:
        pass
    return {"bound": "tensor",
            "kernel": f"gemm_bf16_2cta_kernel ({dkind}{epi}), M={dm} N={dn} K={dk} ({dcnt} launches per step)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else "fallback 1.4 PFLOP/s sustained",
            "algorithmic_flop_per_launch": flops, "avg_launch_us": dms / dcnt * 1e3,
            "share_of_step_gemm_time": dms / tot_ms, "frac_of_nominal_2250": achieved / 2250.0, "traffic": traffic,
            "all_gemms": {"launches_per_step": len(records), "gemm_ms_per_step": tot_ms,
                          "achieved": tot_fl / (tot_ms * 1e-3) / 1e12, "frac": tot_fl / (tot_ms * 1e-3) / 1e12 / peak,
                          "algorithmic_tflop": tot_fl / 1e12},
            # every launch shape of the step, largest share of the GEMM time first (eager pass, CUDA events per launch)
            "by_shape": [{"M": m, "N": n, "K": k, "kind": kind, "launches": cnt, "avg_us": ms_ / cnt * 1e3,
                          "tflops": 2.0 * m * n * k / (ms_ / cnt * 1e-3) / 1e12, "share": ms_ / tot_ms}
                         for ((m, n, k), kind), (cnt, ms_) in sorted(by_shape.items(), key=lambda kv: -kv[1][1])[:16]]}


def stage(msg):
    if os.environ.get("VLK_BENCH_DEBUG"):
        sys.stderr.write(f"[bench rank {os.environ.get('RANK', '0')}] {msg}\n")
        sys.stderr.flush()


SUMMARY_KEYS = ("metric", "value", "unit", "ms_per_step", "scaling", "e2e", "algorithmic_tflops",
                "frac_of_bf16_sustained_peak", "launches_per_step", "config", "parity", "dp_check", "phases_ms", "final_loss")


def run_b200(args, out=sys.stdout):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (B200 arm) needs a GPU: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    stage("process group up")
    common = (args, torch, dist, dev, world, rank, local, peaks)
    if args.workload == "pretrain":
        line = run_pretrain(*common, steps=args.steps, warmup=args.warmup, headline=True)
    else:
        line = run_caption(*common, workload=args.workload, steps=args.steps, warmup=args.warmup, headline=True)
    # ---- the other workloads of BASELINE.json's metric, in the same process, on the same clock ------------------
    extra = {}
    if args.workload == "linear" and not args.no_extra_workloads:
        k = max(3, min(args.steps, 8))
        for w in ("qformer", "xattn"):
            r = run_caption(*common, workload=w, steps=k, warmup=3, headline=False)
            if r is not None:
                extra[w] = {key: r[key] for key in SUMMARY_KEYS if r.get(key) is not None}
        r = run_pretrain(*common, steps=max(3, min(args.steps, 5)), warmup=3, headline=False)
        if r is not None:
            extra["pretrain"] = {key: r[key] for key in SUMMARY_KEYS if r.get(key) is not None}
            extra["pretrain"]["roofline"] = r["roofline"]
    if rank == 0:
        if extra:
            line["workloads"] = extra
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="linear", choices=["linear", "qformer", "xattn", "pretrain"])
    ap.add_argument("--micro-batch", type=int, default=16, help="pretrain micro-batch (train_gpt2.py:245: B = 16)")
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (the reference's B is per rank)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--zero1", action="store_true", help="pretrain: ZeRO-1 update (reduce-scatter / sharded AdamW / all-gather)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the oracle parity + CPU-baseline leg")
    ap.add_argument("--no-extra-workloads", action="store_true",
                    help="default run: only the headline caption-linear line, without the qformer / xattn / pretrain entries")
    ap.add_argument("--ref-batch", type=int, default=REF_BATCH, help="--impl reference: batch per CPU step (constant)")
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON record).  Everything else that libraries print on the way goes to
    # stderr: the reference's configure_optimizers prints its parameter groups (train_gpt2.py:137-142) and NCCL
    # announces its version with a C-level printf, so file descriptor 1 itself is pointed at stderr for the run.
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved_fd, "w")
    py_stdout, sys.stdout = sys.stdout, sys.stderr
    try:
        rc = run_reference(args, real_stdout) if args.impl == "reference" else run_b200(args, real_stdout)
    finally:
        sys.stdout = py_stdout
        real_stdout.flush()
        os.dup2(saved_fd, 1)
    return rc


if __name__ == "__main__":
    sys.exit(main())
