"""Fused global-norm clip + AdamW on libvlk (vlk_grad_sumsq / vlk_adamw_step).

Replaces ``torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)`` + ``torch.optim.AdamW(..., fused=True)``
(reference: source/gpt2/train_gpt2.py:143,472-476).  ``FusedAdamW`` is a ``torch.optim.Optimizer`` so the
reference loop's ``param_groups[i]['lr'] = lr`` / ``zero_grad()`` / ``state_dict()`` keep working.  Semantics
follow torch's AdamW: decoupled weight decay, bias correction, moments stored in the parameter dtype
(bf16 for the reference's bf16 models), fp32 math in registers.
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
        key = tuple((id(g), p.data_ptr(), p.grad.data_ptr()) for g, ps in active for p in ps)
        if key == self._table_key:
            return self._tables
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
                tables.append(dict(group=g, params=sel, fp32=fp32, host=host, dev=dev, n=len(sel),
                                   max_numel=max(p.numel() for p in sel),
                                   lr=torch.zeros(1, dtype=torch.float32, device=sel[0].device)))
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
                                     self._norm_sq.data_ptr(), stream), "vlk_grad_sumsq")
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
        for t in tables:
            g = t["group"]
            t["lr"].fill_(float(g["lr"]))
            step_t = self._step_t
            b1, b2 = g["betas"]
            check(lib.vlk_adamw_step(t["dev"].data_ptr(), t["n"], t["max_numel"], int(t["fp32"]),
                                     self._norm_sq.data_ptr() if self._pending_max_norm > 0 else 0,
                                     self._pending_max_norm, t["lr"].data_ptr(), float(b1), float(b2), float(g["eps"]),
                                     step_t.data_ptr(), stream), "vlk_adamw_step")
        self._pending_max_norm = 0.0
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # re-share the step counter (a loaded state has one copy per parameter) and drop cached tables
        self._step_t = None
        for st in self.state.values():
            if "step" in st:
                if self._step_t is None:
                    self._step_t = st["step"].detach().float().reshape(1).clone()
                st["step"] = self._step_t
        self._table_key = None
