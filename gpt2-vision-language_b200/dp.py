"""Data-parallel gradient exchange: one flat bf16 bucket, one NCCL all-reduce (AVG) per step.

Replaces ``DDP(model, device_ids=[local_rank])`` (source/gpt2/train_gpt2.py:270, gpt2_linear/train.py:122,
gpt2_cross-att/train.py:99).  Differences from stock DDP, all deliberate:
  * only trainable parameters are reduced (0.6 M / 19.5 M / 28.9 M elements for the caption bridges);
  * gradients live in ONE contiguous buffer (``p.grad`` are views), so the exchange is a single collective sized
    for launch latency, not 25 MiB buckets;
  * the per-forward broadcast of the unused ``attn.bias`` buffers is dropped.
The collective is ``torch.distributed.all_reduce`` (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU
tests); averaging of per-rank mean losses / gradients replicates DDP semantics exactly (average of per-rank means,
not re-weighted by token count — gpt2_linear/model.py:206-210).
"""
import torch
import torch.distributed as dist


SCALAR_SLOTS = 8          # one fp32 scalar travels as 8 base-16 digits (see FlatGradBucket.pack_scalar)


class FlatGradBucket:
    """Layout of ``flat``: [ scalar slots (optional) | gradient of params[0] | params[1] | ... | padding ].
    The scalar slots sit in FRONT so that both the whole-bucket exchange and the ``[0, k)`` head range of the
    split-backward path carry them."""

    def __init__(self, params, scalar_slot=False, pad_multiple=1):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dtype, device = self.params[0].dtype, self.params[0].device
        if any(p.dtype != dtype for p in self.params):
            raise ValueError("FlatGradBucket needs a single gradient dtype")
        # 16-byte align every view so the fused optimizer can use vector loads
        align = 16 // torch.empty((), dtype=dtype).element_size()
        total = (SCALAR_SLOTS + align - 1) // align * align if scalar_slot else 0
        self.params_off = total
        offs = []
        for p in self.params:
            offs.append(total)
            total += (p.numel() + align - 1) // align * align
        self.params_end = total
        total = (total + pad_multiple - 1) // pad_multiple * pad_multiple   # ZeRO-1: equal, 16-byte aligned shards
        self.sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(total, dtype=dtype, device=device)
        self.offsets = {id(p): o for p, o in zip(self.params, offs)}
        for p, o in zip(self.params, offs):
            p.grad = self.flat[o:o + p.numel()].view_as(p)
        self.scalar = self.flat[:SCALAR_SLOTS] if scalar_slot else None

    def offset_of(self, param):
        """Element offset of a parameter's gradient inside the flat buffer (parameters are laid out in
        ``model.parameters()`` order, so the gradients of the last layers form a contiguous tail)."""
        return self.offsets[id(param)]

    def zero(self):
        """Replaces optimizer.zero_grad(): one memset; the .grad views stay attached (static addresses)."""
        self.flat.zero_()

    # ---- a scalar (the step's loss) riding in the same collective ---------------------------------------------
    def pack_scalar(self, value):
        """Encode a non-negative fp32 scalar (< 256) into the scalar slots as 8 base-16 digits of its Q8.24
        fixed-point value.  Digits are small integers: their sum over <= 16 ranks (< 256) and the division by a
        power-of-two world size are EXACT in bf16 / fp32, whatever order the collective adds them in — so the
        averaged loss (train_gpt2.py:470-471) needs no collective of its own."""
        if self.scalar is None:
            raise RuntimeError("bucket was built without a scalar slot")
        if value.is_cuda:
            from . import _lib
            _lib.check(_lib.load().vlk_scalar_pack_digits(value.data_ptr(), self.scalar.data_ptr(),
                                                          int(self.flat.dtype == torch.float32),
                                                          torch.cuda.current_stream().cuda_stream), "vlk_scalar_pack_digits")
        else:   # host restatement for the gloo tests
            q = int(round(min(max(float(value), 0.0), 255.99999) * (1 << 24)))
            self.scalar.copy_(torch.tensor([(q >> (4 * i)) & 15 for i in range(SCALAR_SLOTS)], dtype=self.flat.dtype))

    def unpack_scalar(self, out):
        """Inverse of pack_scalar on the (averaged) digits -> ``out`` (0-d / 1-element fp32 tensor)."""
        if out.is_cuda:
            from . import _lib
            _lib.check(_lib.load().vlk_scalar_unpack_digits(self.scalar.data_ptr(), out.data_ptr(),
                                                            int(self.flat.dtype == torch.float32),
                                                            torch.cuda.current_stream().cuda_stream), "vlk_scalar_unpack_digits")
        else:
            d = self.scalar.double()
            out.fill_(float(sum(d[i].item() * 16.0 ** i for i in range(SCALAR_SLOTS)) / (1 << 24)))

    @staticmethod
    def scalar_rides_along(group=None):
        """The digit encoding is exact only when the average divides by a power of two."""
        n = dist.get_world_size(group)
        return n & (n - 1) == 0 and n <= 16

    def all_reduce(self, group=None, lo=0, hi=None):
        """Average elements [lo, hi) over ranks, in place, on the current stream (default: the whole bucket)."""
        buf = self.flat if (lo == 0 and hi is None) else self.flat[lo:hi]
        if buf.numel() and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=group)
            else:  # gloo has no AVG and no bf16 arithmetic: sum in fp32, divide
                tmp = buf.float()
                dist.all_reduce(tmp, op=dist.ReduceOp.SUM, group=group)
                buf.copy_(tmp / dist.get_world_size(group))
        return self.flat


def average_scalar_(t, group=None):
    """In-place mean of a small fp32 tensor over ranks (fallback when it cannot ride in the gradient bucket)."""
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(dist.get_world_size(group))
    return t


class FlatParamBucket:
    """The trainable parameters re-pointed into ONE contiguous buffer with exactly the layout of a FlatGradBucket
    (``p.data`` become views; values are preserved).  Gradient element i and parameter element i then belong to the
    same weight, which is what lets a rank update an arbitrary [lo, hi) slice (ZeRO-1) and all-gather the result."""

    def __init__(self, grad_bucket):
        g = grad_bucket
        self.flat = torch.zeros_like(g.flat)
        with torch.no_grad():
            for p in g.params:
                o = g.offset_of(p)
                view = self.flat[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view


def shard_segments(offsets, numels, weight_decays, lo, hi):
    """Intersections of the flat range [lo, hi) with the tensors laid out at ``offsets``:
    -> [(start, length, weight_decay)] in flat coordinates.  Pure arithmetic (tested on CPU)."""
    out = []
    for o, n, wd in zip(offsets, numels, weight_decays):
        a, b = max(o, lo), min(o + n, hi)
        if a < b:
            out.append((a, b - a, wd))
    return out


def broadcast_parameters(module, src=0, group=None):
    """One-time parameter broadcast from rank 0 (what DDP's constructor does); buffers are NOT re-broadcast
    every forward."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)
