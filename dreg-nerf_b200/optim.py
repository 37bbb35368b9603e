"""Fused ``clip_grad_norm_`` + ``AdamW`` step through the C ABI (``drb_adamw_*``).

Replaces the three host-driven passes of the reference's training loop - ``clip_grad_norm_``
(train_nerf_regtr.py:232-235), ``optimizer.step()`` of ``torch.optim.AdamW`` (:96-102, 237) - by one
gradient-norm kernel and one update kernel over all parameters, with the norm kept on the device.  It is
a ``torch.optim.Optimizer`` so that ``StepLR`` (train_nerf_regtr.py:100-102) and ``zero_grad`` work
unchanged; arithmetic follows torch's AdamW (decoupled weight decay, bias correction).
"""
import ctypes as C

import torch

from . import _lib


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise _lib.DrbError("FusedAdamW supports one parameter group (as train_nerf_regtr.py:96-99 uses)")
        self._params = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        for p in self._params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.DrbError("FusedAdamW needs contiguous fp32 CUDA parameters")
        self._handle = None
        self._active = []          # parameters that receive gradients (torch's AdamW skips the others too)

    def _ensure(self, active):
        if self._handle is not None and [id(p) for p in active] == [id(p) for p in self._active]:
            return
        lib = _lib.load()
        if self._handle is not None:       # the set of parameters with gradients changed: the moments restart
            lib.drb_adamw_destroy(self._handle)
            self._handle = None
        n = len(active)
        ptrs = (C.c_void_p * n)(*[p.data_ptr() for p in active])
        numels = (C.c_longlong * n)(*[p.numel() for p in active])
        h = C.c_void_p()
        with torch.cuda.device(active[0].device):
            _lib.check(lib.drb_adamw_create(n, ptrs, numels, C.byref(h)), "drb_adamw_create")
        self._handle = h
        self._active = list(active)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        active = [p for p in self._params if p.grad is not None]
        if not active:
            return loss
        self._ensure(active)
        lib = _lib.load()
        group = self.param_groups[0]
        n = len(active)
        grads = []
        for p in active:
            g = p.grad
            if not g.is_contiguous() or g.dtype != torch.float32:
                g = g.contiguous().float()
                p.grad = g
            grads.append(g)
        gptrs = (C.c_void_p * n)(*[g.data_ptr() for g in grads])
        b1, b2 = group["betas"]
        with torch.cuda.device(active[0].device):
            _lib.check(lib.drb_adamw_step(self._handle, gptrs, float(group["lr"]), float(b1), float(b2),
                                          float(group["eps"]), float(group["weight_decay"]),
                                          float(group["max_grad_norm"]), _lib.stream_ptr()), "drb_adamw_step")
        # the kernels wrote through raw pointers: tell autograd / NeRFRegTr's weight cache
        try:
            torch.autograd.graph.increment_version(active)
        except TypeError:
            for p in active:
                torch.autograd.graph.increment_version(p)
        return loss

    def grad_norm(self):
        """Total gradient norm the last step saw (synchronises)."""
        if self._handle is None:
            return 0.0
        v = C.c_double(0.0)
        _lib.check(_lib.load().drb_adamw_grad_norm(self._handle, C.byref(v), _lib.stream_ptr()))
        return v.value

    def moments(self, param):
        """(exp_avg, exp_avg_sq, step) of a parameter, copied out of the optimiser state (debug / checkpoints)."""
        i = [id(p) for p in self._active].index(id(param))
        p = self._active[i]
        m, v = torch.empty_like(p), torch.empty_like(p)
        st = C.c_longlong(0)
        with torch.cuda.device(p.device):
            _lib.check(_lib.load().drb_adamw_copy_state(self._handle, i, _lib.ptr(m), _lib.ptr(v), C.byref(st),
                                                        _lib.stream_ptr()), "drb_adamw_copy_state")
        return m, v, st.value

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().drb_adamw_destroy(self._handle)
        except Exception:
            pass
