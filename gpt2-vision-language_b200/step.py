"""The captioning training step as ONE replayable unit of GPU work.

Reference step body: source/gpt2_linear/train.py:291-326 and source/gpt2_cross-att/train.py:272-303
(batch -> device, labels = y.masked_fill(~m, -100), pool, autocast forward, backward, all-reduce,
clip_grad_norm_(1.0), set lr, optimizer.step()) with the CLIP ViT-L/14 forward in front of it (north_star).

``CaptionTrainStep`` owns static input buffers and (optionally) captures the whole step — CLIP forward, pooling,
bridge, GPT-2 forward/backward, loss, gradient all-reduce, clip-norm + AdamW — into a CUDA graph, so a step is a
single graph launch with no host work in between (shapes are static: B x 224 x 224 images, 31-token captions).
"""
import os

import torch
import torch.distributed as dist

from . import ops
from .caption import pool_clip_197_to_33_avg_with_cls
from .dp import FlatGradBucket, FlatParamBucket, average_scalar_


def _nccl_in_graph_default():
    """Where the gradient all-reduce lives when the step is replayed from CUDA graphs.  Default: BETWEEN the graphs
    (forward+backward | all-reduce | clip+AdamW), launched eagerly on the same stream.  VLK_NCCL_IN_GRAPH=1 (or
    nccl_in_graph=True) captures the collectives inside ONE graph per step instead; both layouts are exercised on two
    GPUs by tests/test_dp_gpu2.py.  Measured at 2 GPUs, same box (profiles/r02/bench_n2_*.json): caption-linear 15.14 ms
    between graphs vs 15.25 ms in-graph, Q-Former 16.57 vs 16.71, x-attn 16.94 vs 17.06, pretraining 326.4 vs 326.1 —
    capturing NCCL buys nothing here (the collective is one launch either way), so the simpler layout stays."""
    return bool(os.environ.get("VLK_NCCL_IN_GRAPH"))


def _capture(graph, fn, pool=None, with_collectives=False):
    """Capture fn() into graph.  With collectives inside, every rank first drains its stream and meets at a barrier
    (no NCCL work may be in flight when capture starts), and the capture only polices the capturing thread — the
    process group's watchdog thread keeps polling its own events."""
    kw = {}
    if pool is not None:
        kw["pool"] = pool
    if with_collectives:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        kw["capture_error_mode"] = "thread_local"
    with torch.cuda.graph(graph, **kw):
        fn()
    return graph


def _gpt_split_forward(model, idx, targets, split):
    """``GPT.forward`` (source/gpt2/train_gpt2.py:111-125) with the autograd graph CUT after block ``split-1``:
    returns (loss, [xm], [xd]) where xd is a detached copy of the activation xm entering block ``split``.
    ``loss.backward()`` then stops at xd (gradients of blocks split.., ln_f and the lm_head are complete), and
    ``torch.autograd.backward([xm], [xd.grad])`` finishes the lower half — in between, the upper gradients can
    already travel."""
    t = model.transformer
    x = ops.embed(idx, t.wte.weight, t.wpe.weight)
    for blk in t.h[:split]:
        x = blk(x)
    xd = x.detach().requires_grad_(True)
    y = xd
    for blk in t.h[split:]:
        y = blk(y)
    y = ops.layernorm(y, t.ln_f.weight, t.ln_f.bias, t.ln_f.eps)
    return ops.lmhead_ce(y, model.lm_head.weight, targets), [x], [xd]


def _xattn_split_forward(model, idx, z, targets, mask, split):
    """Same cut for the cross-attention captioner (source/gpt2_cross-att/model.py:152-186); the projected image
    tokens feed every layer, so they are cut as well."""
    t = model.transformer
    x = ops.embed(idx, t.wte.weight, t.wpe.weight)
    zp = t.vis_proj(z)
    for blk in t.h[:split]:
        x = blk(x, zp)
    xd, zd = x.detach().requires_grad_(True), zp.detach().requires_grad_(True)
    y = xd
    for blk in t.h[split:]:
        y = blk(y, zd)
    y = ops.layernorm(y, t.ln_f.weight, t.ln_f.bias, t.ln_f.eps)
    return ops.lmhead_ce(y, model.lm_head.weight, targets, mask), [x, zp], [xd, zd]


def _upper_range(bucket, blocks, split):
    """[lo, hi) of the flat gradient bucket holding blocks[split:] and everything after them."""
    for blk in blocks[split:]:
        for p in blk.parameters():
            if p.requires_grad:
                return bucket.offset_of(p), bucket.params_end
    return bucket.params_end, bucket.params_end


class _CommOverlap:
    """Second stream for the early half of the gradient exchange."""

    def __init__(self):
        self.stream = torch.cuda.Stream()

    def launch(self, fn):
        """Run fn (a collective) on the comm stream, ordered after everything already on the current stream."""
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            fn()

    def join(self):
        torch.cuda.current_stream().wait_stream(self.stream)


class CaptionTrainStep:
    def __init__(self, model, clip_tower, kind, batch, text_len=31, lr=1e-3, weight_decay=0.1, max_norm=1.0,
                 use_graph=True, pixels_dtype=torch.float32, group=None, overlap_comm=None, nccl_in_graph=None):
        """kind: 'linear' | 'qformer' (GPT_Caption(patch_tokens, input_ids, labels)) or 'xattn'
        (GPT(idx, z, targets, target_mask)).
        overlap_comm (xattn only; default OFF): backward is cut in the middle of the stack and the all-reduce of the
        upper layers' gradients runs on a second stream UNDER the lower half of backward.  Measured on one box
        (profiles/r01_dp2_*, profiles/r02/overlap_ab_n4.md): 58 MB exchanged, 6,866 vs 6,919 samples/s at 2 GPUs and
        16,162 vs 16,212 at 4 — the extra graph boundary costs what the overlap hides, so the single exchange after
        backward is the default and the split path stays available (and tested) for larger rings.
        (Linear / Q-Former bridge gradients all materialise at the very end of backward: nothing to overlap.)"""
        assert kind in ("linear", "qformer", "xattn")
        self.model, self.clip, self.kind, self.group = model, clip_tower, kind, group
        self.max_norm, self.use_graph = max_norm, use_graph
        dev = next(model.parameters()).device
        self.dev = dev
        self.pixels = torch.zeros(batch, 3, 224, 224, device=dev, dtype=pixels_dtype)
        self.x = torch.zeros(batch, text_len, device=dev, dtype=torch.int64)
        self.y = torch.zeros(batch, text_len, device=dev, dtype=torch.int64)
        self.mask = torch.ones(batch, text_len, device=dev, dtype=torch.bool)
        self.loss = torch.zeros((), device=dev, dtype=torch.float32)
        self.norm = torch.zeros((), device=dev, dtype=torch.float32)
        # the averaged loss rides in the front slots of the gradient bucket: ONE collective per step
        self.bucket = FlatGradBucket(model.parameters(), scalar_slot=True)
        self.opt = model.configure_optimizers(weight_decay, lr, "cuda")
        self.graph = None
        self._warm = 0
        self.nccl_in_graph = _nccl_in_graph_default() if nccl_in_graph is None else bool(nccl_in_graph)
        self.overlap = bool(overlap_comm) and kind == "xattn"   # None = default = off
        if self.overlap:
            self.split = len(model.transformer.h) // 2
            self.upper = _upper_range(self.bucket, model.transformer.h, self.split)
            self.comm = _CommOverlap()
            self._cut = None

    # -------------------------------------------------------------------------------------------------
    def _multi(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _fwd_bwd(self):
        self._fwd_phase1()
        if self.overlap:
            self._phase2()

    def _fwd_phase1(self):
        """CLIP forward, pooling, captioner forward and backward (down to the cut when overlapping)."""
        feats = self.clip(self.pixels)                                    # [B,257,768]
        z = pool_clip_197_to_33_avg_with_cls(feats)                       # [B,33,768], unit-norm rows
        self.bucket.zero()
        if self.overlap:
            loss, outs, cuts = _xattn_split_forward(self.model, self.x, z, self.y, self.mask, self.split)
            self._cut = (outs, cuts)
        elif self.kind == "xattn":
            _, loss = self.model(self.x, z=z, targets=self.y, target_mask=self.mask)
        else:
            labels = self.y.masked_fill(~self.mask, -100)
            _, loss = self.model(z, self.x, labels=labels)
        with ops.residual_grad_inplace():     # this class owns every gradient tensor of the step
            loss.backward()
        self.loss.copy_(loss.detach())

    def _phase2(self):
        """Lower half of backward, from the cut to the first layer."""
        outs, cuts = self._cut
        with ops.residual_grad_inplace():
            torch.autograd.backward(outs, [c.grad for c in cuts])
        self._cut = None

    def _exchange(self, lo=0, hi=None):
        """The one exchange step of the path: average the flat gradient bucket over ranks — and with it the step's
        loss (train_gpt2.py:470-471), which travels as exact base-16 digits in the bucket's front slots."""
        if self._multi():
            rides = self.bucket.scalar_rides_along(self.group)
            if rides:
                self.bucket.pack_scalar(self.loss)
            self.bucket.all_reduce(self.group, lo, hi)
            if rides:
                self.bucket.unpack_scalar(self.loss)
            else:
                average_scalar_(self.loss, self.group)

    def _update(self):
        self.norm.copy_(self.opt.clip_grad_norm(self.max_norm))
        self.opt.step()
        if self.kind == "qformer":
            # fresh dropout masks next step, also under CUDA-graph replay (the Philox step counter lives on device)
            ops.DropoutState.default(self.dev).advance()

    def _exchange_upper_async(self):
        if self._multi():
            lo, hi = self.upper
            self.comm.launch(lambda: self.bucket.all_reduce(self.group, lo, hi))

    def _exchange_lower(self):
        if self._multi():
            self._exchange(0, self.upper[0])
            self.comm.join()

    def _body(self):
        if self.overlap:
            self._fwd_phase1()
            self._exchange_upper_async()
            self._phase2()
            self._exchange_lower()
        else:
            self._fwd_bwd()
            self._exchange()
        self._update()

    def set_lr(self, lr):
        """optimizer.param_groups[i]['lr'] = lr (train_gpt2.py:474-475).  The learning rate lives in a device scalar
        that the captured update reads, so the new value also reaches graph replays."""
        for g in self.opt.param_groups:
            g["lr"] = lr
        self.opt.sync_lr()

    def load_batch(self, pixels, x, y, mask, non_blocking=True):
        """Host (pinned) or device tensors -> the static input buffers."""
        self.pixels.copy_(pixels, non_blocking=non_blocking)
        self.x.copy_(x, non_blocking=non_blocking)
        self.y.copy_(y, non_blocking=non_blocking)
        self.mask.copy_(mask, non_blocking=non_blocking)

    def run(self):
        """One optimizer step on the current contents of the static buffers. Returns the (device) loss tensor.
        The whole step — at any world size, including the gradient all-reduce — is one CUDA graph."""
        if not self.use_graph:
            self._body()
            return self.loss
        if self.graph is None:
            # two eager warm-up steps on a side stream (allocator / lazy state / NCCL communicator), then capture
            if self._warm < 2:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._body()
                torch.cuda.current_stream().wait_stream(s)
                self._warm += 1
                return self.loss
            if not self._multi() or self.nccl_in_graph:
                # ONE graph for the whole step at any world size: the NCCL collectives (and, with overlap, the fork
                # onto the comm stream) are captured with the kernels
                self.graph = _capture(torch.cuda.CUDAGraph(), self._body, with_collectives=self._multi())
            elif self.overlap:
                # three graphs: forward + upper backward | lower backward | update; the two halves of the gradient
                # exchange are launched between them, the first one on the comm stream
                self.graph = (torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph())
                with torch.cuda.graph(self.graph[0]):
                    self._fwd_phase1()
                with torch.cuda.graph(self.graph[1], pool=self.graph[0].pool()):
                    self._phase2()
                with torch.cuda.graph(self.graph[2], pool=self.graph[0].pool()):
                    self._update()
            else:
                self.graph = (torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph())
                with torch.cuda.graph(self.graph[0]):
                    self._fwd_bwd()
                with torch.cuda.graph(self.graph[1]):
                    self._update()
        if isinstance(self.graph, tuple) and len(self.graph) == 3:
            self.graph[0].replay()
            self._exchange_upper_async()
            self.graph[1].replay()
            self._exchange_lower()
            self.graph[2].replay()
        elif isinstance(self.graph, tuple):
            self.graph[0].replay()
            self._exchange()
            self.graph[1].replay()
        else:
            self.graph.replay()
        return self.loss


class HostBatchFeeder:
    """Double-buffered host -> device input pipeline for CaptionTrainStep.

    The reference moves each batch with blocking ``.to(device)`` calls at the top of the step
    (source/gpt2_linear/train.py:300-303).  Here the NEXT batch's pinned-host -> HBM copy (38.6 MB of fp32 pixels at
    B=64) runs on a copy stream while the current step computes; at the step boundary a device-to-device copy
    (a few microseconds) drops it into the step's static input buffers.  Every byte still crosses PCIe every step."""

    def __init__(self, step):
        self.step = step
        self.copy_stream = torch.cuda.Stream()
        self.stage = [tuple(torch.empty_like(t) for t in (step.pixels, step.x, step.y, step.mask)) for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]      # H2D into stage[i] finished
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]   # D2D out of stage[i] finished
        self._n_submit = self._n_load = 0
        for e in self.consumed:
            e.record()

    def submit(self, pixels, x, y, mask):
        """Start the asynchronous upload of one pinned host batch."""
        i = self._n_submit & 1
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[i])
            for dst, src in zip(self.stage[i], (pixels, x, y, mask)):
                dst.copy_(src, non_blocking=True)
            self.ready[i].record(self.copy_stream)
        self._n_submit += 1

    def load(self):
        """Make the oldest submitted batch the step's current input (on the compute stream)."""
        i = self._n_load & 1
        cur = torch.cuda.current_stream()
        cur.wait_event(self.ready[i])
        self.step.load_batch(*self.stage[i])
        self.consumed[i].record(cur)
        self._n_load += 1


class PretrainStep:
    """GPT-2 124M pretraining step (source/gpt2/train_gpt2.py:456-478): ``grad_accum`` micro-batches of
    [micro_batch, seq] tokens -> loss/grad_accum -> backward (gradients accumulate in the flat bucket) -> all-reduce
    -> clip_grad_norm_(1.0) -> AdamW.  Two CUDA graphs: one micro-step (replayed grad_accum times on static token
    slots) and one update (all-reduce + clip + AdamW)."""

    def __init__(self, model, micro_batch=16, seq=1024, grad_accum=32, lr=6e-4, weight_decay=0.1, max_norm=1.0,
                 use_graph=True, group=None, overlap_comm=None, zero1=False, nccl_in_graph=None):
        """zero1: ZeRO-1 update (optim.Zero1AdamW) — the gradient bucket is reduce-scattered, every rank runs clip +
        AdamW on its 1/N slice of a flat parameter buffer (moments for that slice only), the slices are all-gathered.
        The update then runs eagerly (three collectives + two kernels per step); overlap_comm is ignored.
        overlap_comm (default: on when data parallel): the LAST micro-step's backward is cut in the middle of the
        stack; the all-reduce of layers n/2.. + ln_f (half of the 249 MB) runs on a second stream under the lower
        half of that backward, the rest (layers 0..n/2-1, wte — whose gradient also collects the lm_head's —, wpe)
        follows.  DDP overlaps the same way with 25 MiB buckets (train_gpt2.py:467-468 enables the sync on the last
        micro-step only)."""
        self.model, self.group, self.max_norm, self.use_graph = model, group, max_norm, use_graph
        self.grad_accum = grad_accum
        dev = next(model.parameters()).device
        self.dev = dev
        self.tokens_x = torch.zeros(grad_accum, micro_batch, seq, device=dev, dtype=torch.int64)
        self.tokens_y = torch.zeros(grad_accum, micro_batch, seq, device=dev, dtype=torch.int64)
        self.x = torch.zeros(micro_batch, seq, device=dev, dtype=torch.int64)
        self.y = torch.zeros(micro_batch, seq, device=dev, dtype=torch.int64)
        self.loss = torch.zeros((), device=dev, dtype=torch.float32)
        self.norm = torch.zeros((), device=dev, dtype=torch.float32)
        self.inv_accum = torch.full((), 1.0 / grad_accum, device=dev, dtype=torch.float32)
        self.zero1 = bool(zero1)
        if self.zero1:
            from .optim import Zero1AdamW
            world = dist.get_world_size(group) if self._multi_static(group) else 1
            rank = dist.get_rank(group) if world > 1 else 0
            self.bucket = FlatGradBucket(model.parameters(), scalar_slot=True, pad_multiple=8 * world)
            self.pbucket = FlatParamBucket(self.bucket)
            wds = [weight_decay if p.dim() >= 2 else 0.0 for p in self.bucket.params]   # train_gpt2.py:131-136
            self.opt = Zero1AdamW(self.bucket, self.pbucket, wds, lr=lr, rank=rank, world=world, group=group)
        else:
            self.bucket = FlatGradBucket(model.parameters(), scalar_slot=True)
            self.opt = model.configure_optimizers(weight_decay, lr, "cuda")
        self.g_micro = self.g_update = self.g_last = self.g_tail = None
        self._warm = 0
        self.nccl_in_graph = _nccl_in_graph_default() if nccl_in_graph is None else bool(nccl_in_graph)
        self.phase_events = None      # set to [] to record (label, cuda event) pairs of one run() (bench.py)
        if overlap_comm is None:
            overlap_comm = self._multi()
        self.overlap = bool(overlap_comm) and not self.zero1
        if self.overlap:
            self.split = len(model.transformer.h) // 2
            self.upper = _upper_range(self.bucket, model.transformer.h, self.split)
            self.comm = _CommOverlap()
            self._cut = None

    @staticmethod
    def _multi_static(group):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def _multi(self):
        return self._multi_static(self.group)

    def _micro(self):
        _, loss = self.model(self.x, self.y)
        with ops.residual_grad_inplace():     # this class owns every gradient tensor of the step
            loss.backward(self.inv_accum)     # loss / grad_accum (train_gpt2.py:464): the factor rides in the CE epilogue
        self.loss.add_(loss.detach(), alpha=1.0 / self.grad_accum)

    def _last_phase1(self):
        loss, outs, cuts = _gpt_split_forward(self.model, self.x, self.y, self.split)
        self._cut = (outs, cuts)
        with ops.residual_grad_inplace():
            loss.backward(self.inv_accum)
        self.loss.add_(loss.detach(), alpha=1.0 / self.grad_accum)

    def _last_phase2(self):
        outs, cuts = self._cut
        with ops.residual_grad_inplace():
            torch.autograd.backward(outs, [c.grad for c in cuts])
        self._cut = None

    def _exchange_upper_async(self):
        if self._multi():
            lo, hi = self.upper
            self.comm.launch(lambda: self.bucket.all_reduce(self.group, lo, hi))

    def _exchange_lower(self):
        if self._multi():
            self._exchange(0, self.upper[0])
            self.comm.join()

    def _exchange(self, lo=0, hi=None):
        """Average the gradient bucket over ranks; the step's mean loss (train_gpt2.py:470-471) rides in its front
        slots.  ZeRO-1 reduce-scatters the bucket inside its update instead: only the loss is exchanged here."""
        if not self._multi():
            return
        if self.zero1:
            average_scalar_(self.loss, self.group)
            return
        rides = self.bucket.scalar_rides_along(self.group)
        if rides:
            self.bucket.pack_scalar(self.loss)
        self.bucket.all_reduce(self.group, lo, hi)
        if rides:
            self.bucket.unpack_scalar(self.loss)
        else:
            average_scalar_(self.loss, self.group)

    def _tail(self):
        """Last micro-step + gradient exchange + update (with overlap: the upper half of the exchange runs on the
        comm stream under the lower half of the last backward)."""
        if self.overlap:
            self._last_phase1()
            self._exchange_upper_async()
            self._last_phase2()
            self._exchange_lower()
        else:
            self._micro()
            self._exchange()
        self._mark("exchange")
        self._update()

    def _mark(self, label):
        if self.phase_events is not None and not torch.cuda.is_current_stream_capturing():
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.phase_events.append((label, e))

    def _update(self):
        if self.zero1:
            self.norm.copy_(self.opt.step(self.max_norm))
            return
        self.norm.copy_(self.opt.clip_grad_norm(self.max_norm))
        self.opt.step()

    def set_lr(self, lr):
        if self.zero1:
            self.opt.set_lr(lr)
            return
        for g in self.opt.param_groups:
            g["lr"] = lr
        self.opt.sync_lr()

    def load_tokens(self, x, y, non_blocking=True):
        """[grad_accum, micro_batch, seq] int64 (host pinned or device)."""
        self.tokens_x.copy_(x, non_blocking=non_blocking)
        self.tokens_y.copy_(y, non_blocking=non_blocking)

    def _set_slot(self, i):
        self.x.copy_(self.tokens_x[i])
        self.y.copy_(self.tokens_y[i])

    def run(self):
        """One optimizer step over the grad_accum micro-batches currently in the token buffers.
        Graph layout: ``g_micro`` replayed grad_accum-1 times, then ``g_tail`` = last micro-step + all-reduce +
        clip + AdamW as ONE graph (collectives captured inside).  With VLK_NCCL_OUTSIDE_GRAPH=1 and several ranks
        the round-1 layout is used instead (collectives launched between graphs)."""
        self._mark("start")
        self.bucket.zero()
        self.loss.zero_()
        if not self.use_graph or self._warm < 1:
            # eager pass (also the warm-up that creates optimizer state and allocator pools), on a side stream
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for i in range(self.grad_accum - 1):
                    self._set_slot(i)
                    self._micro()
                self._set_slot(self.grad_accum - 1)
                self._tail()
            torch.cuda.current_stream().wait_stream(s)
            self._warm += 1
            return self.loss
        single_tail = not self._multi() or self.nccl_in_graph
        if self.g_micro is None and self.g_tail is None and self.g_update is None:
            self._set_slot(0)
            pool = None
            if self.grad_accum > 1 or not single_tail:
                self.g_micro = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.g_micro):      # capture only records; nothing below has run yet
                    self._micro()
                pool = self.g_micro.pool()
            if single_tail:
                self.g_tail = _capture(torch.cuda.CUDAGraph(), self._tail, pool=pool, with_collectives=self._multi())
            else:
                if self.overlap:
                    # the last micro-step as two graphs (down to the cut | the rest); they share the micro-step's
                    # memory pool — the three never run concurrently
                    self.g_last = (torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph())
                    with torch.cuda.graph(self.g_last[0], pool=pool):
                        self._last_phase1()
                    with torch.cuda.graph(self.g_last[1], pool=self.g_last[0].pool()):
                        self._last_phase2()
                self.g_update = torch.cuda.CUDAGraph()
                if not self.zero1:                       # the ZeRO-1 update holds collectives: it runs eagerly here
                    with torch.cuda.graph(self.g_update):
                        self._update()
        if single_tail:
            for i in range(self.grad_accum - 1):
                self._set_slot(i)
                self.g_micro.replay()
            self._mark("micro_steps")
            self._set_slot(self.grad_accum - 1)
            self.g_tail.replay()
            self._mark("tail")
            return self.loss
        if self.overlap:
            for i in range(self.grad_accum - 1):
                self._set_slot(i)
                self.g_micro.replay()
            self._set_slot(self.grad_accum - 1)
            self.g_last[0].replay()
            self._exchange_upper_async()           # NCCL outside graph capture, on the comm stream
            self.g_last[1].replay()
            self._exchange_lower()
        else:
            for i in range(self.grad_accum):
                self._set_slot(i)
                self.g_micro.replay()
            self._exchange()
        self._mark("exchange")
        if self.zero1:
            self._update()
        else:
            self.g_update.replay()
        self._mark("update")
        return self.loss
