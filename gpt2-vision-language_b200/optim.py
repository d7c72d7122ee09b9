"""Fused global-norm clip + AdamW on libvlk (vlk_grad_sumsq / vlk_adamw_step).

Replaces ``torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)`` + ``torch.optim.AdamW(..., fused=True)``
(reference: source/gpt2/train_gpt2.py:143,472-476).  ``FusedAdamW`` is a ``torch.optim.Optimizer`` so the
reference loop's ``param_groups[i]['lr'] = lr`` / ``zero_grad()`` / ``state_dict()`` keep working.  Semantics
follow torch's AdamW: decoupled weight decay, bias correction, moments stored in the parameter dtype
(bf16 for the reference's bf16 models), fp32 math in registers.

Two deliberate deviations from torch.optim.AdamW, both enforced rather than silent:
  * ONE device step counter is shared by all parameters (torch keeps one per parameter).  They are identical as long as
    every parameter steps every time — which is how the reference trains — so a parameter that joins later (its
    ``.grad`` was None during earlier steps) raises instead of being bias-corrected with the wrong step;
  * the learning rate is read by the kernel from a device scalar, so that a captured CUDA graph follows the schedule:
    ``param_groups[i]['lr'] = x`` takes effect at the next eager ``step()``, or immediately after ``sync_lr()``
    when the step is replayed from a graph (step.CaptionTrainStep.set_lr does both).
"""
import ctypes

import torch

from . import _lib
from ._lib import TensorDesc, check


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._table_key = None
        self._tables = None
        self._norm_sq = None
        self._pending_max_norm = 0.0
        self._step_t = None  # ONE device step counter shared by every parameter (they always step together)
        self._n_steps = 0

    # ---------------------------------------------------------------------------------------------
    def _active(self):
        """[(group, [params with grads])] — torch semantics: parameters whose .grad is None are skipped."""
        out = []
        for g in self.param_groups:
            ps = [p for p in g["params"] if p.grad is not None]
            if ps:
                out.append((g, ps))
        return out

    def _ensure_state(self, p):
        st = self.state[p]
        if not st:
            if self._step_t is None:
                self._step_t = torch.zeros(1, dtype=torch.float32, device=p.device)
            st["step"] = self._step_t
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _build_tables(self, active):
        """One device table per (group, dtype); rebuilt only when a pointer changed (e.g. after
        zero_grad(set_to_none=True)).  Under CUDA graphs all pointers are static, so this is a no-op."""
        # weight_decay is baked into the device table: it is part of the key (a scheduler or load_state_dict may change it)
        key = tuple((id(g), float(g["weight_decay"]), p.data_ptr(), p.grad.data_ptr()) for g, ps in active for p in ps)
        if key == self._table_key:
            return self._tables
        if self._n_steps > 0 and any(not self.state[p] for _, ps in active for p in ps):
            raise RuntimeError("FusedAdamW: a parameter received its first gradient after the optimizer had already "
                               "stepped; all parameters share one step counter (see the module docstring)")
        tables = []
        for g, ps in active:
            for fp32 in (False, True):
                sel = [p for p in ps if (p.dtype == torch.float32) == fp32]
                if not sel:
                    continue
                for p in sel:
                    if p.dtype not in (torch.bfloat16, torch.float32) or not p.is_cuda:
                        raise RuntimeError("FusedAdamW supports bf16 / fp32 CUDA parameters only")
                    if p.grad.dtype != p.dtype or not p.is_contiguous() or not p.grad.is_contiguous():
                        raise RuntimeError("FusedAdamW needs contiguous grads of the parameter dtype")
                arr = (TensorDesc * len(sel))()
                for i, p in enumerate(sel):
                    st = self._ensure_state(p)
                    arr[i] = TensorDesc(p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(),
                                        st["exp_avg_sq"].data_ptr(), p.numel(), float(g["weight_decay"]), 0)
                host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).pin_memory()
                dev = host.to(sel[0].device, non_blocking=True)
                max_numel = max(p.numel() for p in sel)
                nws = int(_lib.load().vlk_grad_sumsq_workspace_floats(len(sel), max_numel))
                tables.append(dict(group=g, params=sel, fp32=fp32, host=host, dev=dev, n=len(sel), max_numel=max_numel,
                                   partials=torch.zeros(nws, dtype=torch.float32, device=sel[0].device),
                                   lr=torch.full((1,), float(g["lr"]), dtype=torch.float32, device=sel[0].device)))
        self._table_key, self._tables = key, tables
        return tables

    # ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def clip_grad_norm(self, max_norm):
        """Fused equivalent of clip_grad_norm_: computes the global L2 norm of all gradients now and arms the
        clip factor min(1, max_norm/(norm+1e-6)) for the next step() (applied in registers; .grad itself is not
        rewritten).  Returns the pre-clip norm as a 0-d device tensor, like the reference logs it."""
        lib = _lib.load()
        tables = self._build_tables(self._active())
        if not tables:
            return torch.zeros(())
        dev = tables[0]["dev"].device
        if self._norm_sq is None:
            self._norm_sq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._norm_sq.zero_()
        stream = torch.cuda.current_stream().cuda_stream
        for t in tables:
            check(lib.vlk_grad_sumsq(t["dev"].data_ptr(), t["n"], t["max_numel"], int(t["fp32"]),
                                     self._norm_sq.data_ptr(), t["partials"].data_ptr(), stream), "vlk_grad_sumsq")
        self._pending_max_norm = float(max_norm)
        return self._norm_sq.sqrt().reshape(())

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        tables = self._build_tables(self._active())
        stream = torch.cuda.current_stream().cuda_stream
        if tables:
            self._step_t += 1
            self._n_steps += 1
        capturing = torch.cuda.is_current_stream_capturing()
        for t in tables:
            g = t["group"]
            if not capturing:                       # a captured fill would freeze the learning rate into the graph
                t["lr"].fill_(float(g["lr"]))
            step_t = self._step_t
            b1, b2 = g["betas"]
            check(lib.vlk_adamw_step(t["dev"].data_ptr(), t["n"], t["max_numel"], int(t["fp32"]),
                                     self._norm_sq.data_ptr() if self._pending_max_norm > 0 else 0,
                                     self._pending_max_norm, t["lr"].data_ptr(), float(b1), float(b2), float(g["eps"]),
                                     step_t.data_ptr(), stream), "vlk_adamw_step")
        self._pending_max_norm = 0.0
        return loss

    @torch.no_grad()
    def sync_lr(self):
        """Push param_groups[*]['lr'] into the device scalars the (possibly graph-captured) update kernels read."""
        for t in self._tables or ():
            t["lr"].fill_(float(t["group"]["lr"]))

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # re-share the step counter (a loaded state has one copy per parameter) and drop cached tables.  torch leaves
        # 'step' wherever the checkpoint was mapped to (the CPU for data.load_checkpoint's default map_location): the
        # kernel dereferences it, so it is rebuilt on the PARAMETER's device.
        self._step_t = None
        for p, st in self.state.items():
            if "step" in st:
                if self._step_t is None:
                    step = st["step"]
                    step = step.detach() if torch.is_tensor(step) else torch.tensor(float(step))
                    self._step_t = step.to(device=p.device, dtype=torch.float32).reshape(1).clone()
                st["step"] = self._step_t
            for k in ("exp_avg", "exp_avg_sq"):
                if k in st and (st[k].device != p.device or st[k].dtype != p.dtype):
                    st[k] = st[k].to(device=p.device, dtype=p.dtype)
        self._n_steps = int(self._step_t.item()) if self._step_t is not None else 0
        self._table_key = None


class Zero1AdamW:
    """ZeRO-1 flavour of the fused clip + AdamW for the data-parallel pretraining step: instead of
    all-reducing the 249 MB gradient bucket and running the full 1.7 GB/step optimizer pass on every rank, the flat
    gradient is REDUCE-SCATTERED (each rank receives the average of its 1/N slice), the rank updates only that slice of
    the flat parameter buffer (moments exist for the slice only: optimizer state and traffic / N) and the updated
    slices are ALL-GATHERED.  Same bytes on the wire as the all-reduce; same arithmetic as FusedAdamW per element
    (decoupled weight decay per tensor, bias correction, clip by the global norm = all-reduced sum of the slices'
    squared norms).  The phases are separate methods so they can be driven without a process group (tests)."""

    def __init__(self, grad_bucket, param_bucket, weight_decays, lr=6e-4, betas=(0.9, 0.95), eps=1e-8, rank=0, world=1,
                 group=None):
        from .dp import shard_segments
        g = grad_bucket
        total = g.flat.numel()
        if total % (8 * world):
            raise ValueError("flat bucket must be padded to a multiple of 8 * world (FlatGradBucket(pad_multiple=...))")
        self.g, self.p, self.rank, self.world, self.group = g, param_bucket, rank, world, group
        self.lr, self.betas, self.eps = lr, betas, eps
        self.S = total // world
        self.lo, self.hi = rank * self.S, (rank + 1) * self.S
        dev, dt = g.flat.device, g.flat.dtype
        self.gshard = torch.zeros(self.S, device=dev, dtype=dt)
        self.exp_avg = torch.zeros(self.S, device=dev, dtype=dt)
        self.exp_avg_sq = torch.zeros(self.S, device=dev, dtype=dt)
        offs = [g.offset_of(q) for q in g.params]
        # the bucket's scalar slots (offsets < params_off) and tail padding belong to no tensor: never updated
        self.segments = shard_segments(offs, g.sizes, weight_decays, self.lo, self.hi)
        esz = g.flat.element_size()
        arr = (TensorDesc * max(1, len(self.segments)))()
        for i, (a, n, wd) in enumerate(self.segments):
            k = a - self.lo
            arr[i] = TensorDesc(self.p.flat.data_ptr() + a * esz, self.gshard.data_ptr() + k * esz,
                                self.exp_avg.data_ptr() + k * esz, self.exp_avg_sq.data_ptr() + k * esz, n, float(wd), 0)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.table = host.to(dev)
        self.n = len(self.segments)
        self.max_numel = max((n for _, n, _ in self.segments), default=1)
        self.fp32 = dt == torch.float32
        self.norm_sq = torch.zeros(1, device=dev, dtype=torch.float32)
        self.partials = torch.zeros(max(1, int(_lib.load().vlk_grad_sumsq_workspace_floats(max(self.n, 1), self.max_numel))),
                                    device=dev, dtype=torch.float32)
        self.step_t = torch.zeros(1, device=dev, dtype=torch.float32)
        self.lr_t = torch.zeros(1, device=dev, dtype=torch.float32)

    # ---- phases ------------------------------------------------------------------------------------------
    def reduce_scatter(self):
        """gshard <- average over ranks of flat_grad[lo:hi]."""
        import torch.distributed as dist
        if self.world > 1 and dist.is_initialized():
            if dist.get_backend(self.group) == "nccl":
                dist.reduce_scatter_tensor(self.gshard, self.g.flat, op=dist.ReduceOp.AVG, group=self.group)
            else:   # gloo: no reduce-scatter / AVG / bf16 arithmetic
                tmp = self.g.flat.float()
                dist.all_reduce(tmp, group=self.group)
                self.gshard.copy_(tmp[self.lo:self.hi] / self.world)
        else:
            self.gshard.copy_(self.g.flat[self.lo:self.hi])

    @torch.no_grad()
    def local_sumsq(self):
        self.norm_sq.zero_()
        if self.n:
            check(_lib.load().vlk_grad_sumsq(self.table.data_ptr(), self.n, self.max_numel, int(self.fp32),
                                             self.norm_sq.data_ptr(), self.partials.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream), "vlk_grad_sumsq")
        return self.norm_sq

    def set_lr(self, lr):
        self.lr = lr
        self.lr_t.fill_(float(lr))

    @torch.no_grad()
    def apply(self, norm_sq, max_norm):
        """AdamW on this rank's slice with the clip factor of the GLOBAL norm (norm_sq: 1-element fp32 tensor)."""
        self.step_t += 1
        if not torch.cuda.is_current_stream_capturing():
            self.lr_t.fill_(float(self.lr))
        if self.n:
            b1, b2 = self.betas
            check(_lib.load().vlk_adamw_step(self.table.data_ptr(), self.n, self.max_numel, int(self.fp32),
                                             norm_sq.data_ptr() if max_norm > 0 else 0, float(max_norm),
                                             self.lr_t.data_ptr(), float(b1), float(b2), float(self.eps),
                                             self.step_t.data_ptr(), torch.cuda.current_stream().cuda_stream),
                  "vlk_adamw_step")

    def all_gather(self):
        import torch.distributed as dist
        if self.world > 1 and dist.is_initialized():
            shard = self.p.flat[self.lo:self.hi]
            if dist.get_backend(self.group) == "nccl":
                dist.all_gather_into_tensor(self.p.flat, shard, group=self.group)
            else:
                parts = [torch.empty_like(shard) for _ in range(self.world)]
                dist.all_gather(parts, shard.clone(), group=self.group)
                self.p.flat.copy_(torch.cat(parts))

    def step(self, max_norm=1.0):
        """One optimizer step over the gradients currently in the flat bucket; returns the pre-clip global norm."""
        import torch.distributed as dist
        self.reduce_scatter()
        nsq = self.local_sumsq()
        if self.world > 1 and dist.is_initialized():
            dist.all_reduce(nsq, group=self.group)
        self.apply(nsq, max_norm)
        self.all_gather()
        return nsq.sqrt().reshape(())
