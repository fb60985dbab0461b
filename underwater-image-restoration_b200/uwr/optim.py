"""Fused optimizer step: clip_grad_norm_(params, max_norm) + Adam/AdamW in two kernels over a
device pointer table (replaces ModelTrainer.py:87-88 + torch.optim.Adam/AdamW, 197-204)."""
import ctypes as C

import torch

from ._lib import check, fn


class FusedClipAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False,
                 max_norm=1.0, grad_prescale=1.0):
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.betas, self.eps = lr, betas, eps
        self.weight_decay, self.decoupled = weight_decay, decoupled
        self.max_norm = max_norm
        self.grad_prescale = grad_prescale
        self.step_count = 0
        self._built_for = None
        self.exp_avg = None

    # -- state ------------------------------------------------------------------------------
    def _build(self, active):
        dev = active[0].device
        if self.exp_avg is None:
            self.exp_avg = {id(p): torch.zeros_like(p) for p in self.params}
            self.exp_avg_sq = {id(p): torch.zeros_like(p) for p in self.params}
        sizes = [p.numel() for p in active]
        offs = [0]
        for s in sizes:
            offs.append(offs[-1] + s)
        i64 = lambda xs: torch.tensor(xs, dtype=torch.int64, device=dev)
        self._p = i64([p.data_ptr() for p in active])
        self._g = i64([p.grad.data_ptr() for p in active])
        self._m = i64([self.exp_avg[id(p)].data_ptr() for p in active])
        self._v = i64([self.exp_avg_sq[id(p)].data_ptr() for p in active])
        self._off = i64(offs)
        self._n, self._total = len(active), offs[-1]
        self._norm = torch.zeros(2, device=dev, dtype=torch.float32)
        self._ws = torch.empty(8192, device=dev, dtype=torch.float32)
        self._step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self._step_dev.fill_(self.step_count)
        self._built_for = tuple((p.data_ptr(), p.grad.data_ptr()) for p in active)

    def zero_grad(self, set_to_none=False):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self):
        """Returns the device tensor [grad_norm, clip_coef] (no host sync)."""
        active = [p for p in self.params if p.grad is not None]
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in active)
        if key != self._built_for:
            self._build(active)
        stream = torch.cuda.current_stream().cuda_stream
        check(fn["uwr_increment_i32"](self._step_dev.data_ptr(), stream), "uwr_increment_i32")
        self.step_count += 1
        check(fn["uwr_grad_norm"](self._g.data_ptr(), self._off.data_ptr(), self._n, self._total,
                                  float(self.max_norm), float(self.grad_prescale), self._norm.data_ptr(),
                                  self._ws.data_ptr(), stream), "uwr_grad_norm")
        check(fn["uwr_adam_step"](self._p.data_ptr(), self._g.data_ptr(), self._m.data_ptr(), self._v.data_ptr(),
                                  self._off.data_ptr(), self._n, self._total, self._norm[1:].data_ptr(),
                                  float(self.grad_prescale), float(self.lr), float(self.betas[0]),
                                  float(self.betas[1]), float(self.eps), float(self.weight_decay),
                                  int(self.decoupled), 0, self._step_dev.data_ptr(), stream), "uwr_adam_step")
        from . import ops
        ops.bump_weight_epoch()  # parameters changed behind torch's version counter: rounded copies are stale
        ops.refresh_rounded_copies()  # ... and are re-rounded here, all at once (two multi-tensor launches)
        return self._norm
