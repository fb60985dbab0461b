"""Fused optimizer step: clip_grad_norm_(params, max_norm) + Adam/AdamW in two kernels over a device pointer
table (replaces ModelTrainer.py:87-88 + torch.optim.Adam/AdamW, 197-204).

A `torch.optim.Optimizer` subclass, so the rest of the reference's loop keeps working unchanged:
  * `param_groups[0]["lr"]` is live — torch LR schedulers (the reference's MultiStepLR([1,100,250], 0.25),
    ModelTrainer.py:55,129) and its logging read / write it.  The kernel reads the rate from a DEVICE scalar that
    `sync_hyper()` refreshes whenever the host value changed, so a step captured in a CUDA graph follows the
    schedule (GraphedTrainStep.replay calls sync_hyper first);
  * `state_dict()` / `load_state_dict()` use torch.optim.Adam's layout ({"state": {i: {"step", "exp_avg",
    "exp_avg_sq"}}, "param_groups": [...]}, the same group keys as this torch's Adam) — the `optimizer_state_dict` of the reference checkpoint
    (ModelTrainer.py:172-190) round-trips, also from / to a plain torch.optim.Adam;
  * parameters that never receive a gradient are left alone like torch does for `grad is None` (no moments, no
    weight decay): `exclude()` removes them from the kernel's tables (TrainStep detects them after the first
    backward — its gradient buckets hold zeros where torch has None).
"""
import torch

from ._lib import check, fn


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False,
                 max_norm=1.0, grad_prescale=1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=True, differentiable=False, fused=True, decoupled_weight_decay=decoupled)
        super().__init__([p for p in params if p.requires_grad], defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FusedClipAdam takes one parameter group (the reference builds one, ModelTrainer.py:197-204)")
        self.max_norm = max_norm
        self.grad_prescale = grad_prescale
        self._excluded = set()
        self._built_for = None
        self._step_dev = None      # device int32 step counter (advanced by a kernel: graph-replay safe)
        self._lr_dev = None        # device float32 learning rate
        self._lr_pushed = None

    # ------------------------------------------------------------------------------------ reference-style accessors
    @property
    def params(self):
        return self.param_groups[0]["params"]

    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, value):
        self.param_groups[0]["lr"] = value

    def set_lr(self, value):
        """Change the learning rate (host value + device scalar); valid between replays of a captured step."""
        self.lr = float(value)
        self.sync_hyper()

    @property
    def step_count(self):
        """Number of optimizer steps taken, read from the device counter (graph replays advance it too)."""
        return int(self._step_dev.item()) if self._step_dev is not None else 0

    def exclude(self, params):
        """Leave these parameters out of the update (no moments, no decay): torch's behaviour for grad None."""
        self._excluded |= {id(p) for p in params}
        self._built_for = None

    # -------------------------------------------------------------------------------------------------- device state
    def _device_scalars(self, dev):
        if self._step_dev is None:
            self._step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
            self._lr_dev = torch.zeros(1, device=dev, dtype=torch.float32)
            self._norm = torch.zeros(2, device=dev, dtype=torch.float32)
            self._ws = torch.empty(8192, device=dev, dtype=torch.float32)

    def sync_hyper(self):
        """Push the host learning rate to the device scalar if it changed (tiny fill kernel, outside any graph)."""
        if self._lr_dev is not None and self._lr_pushed != self.lr:
            self._lr_dev.fill_(float(self.lr))
            self._lr_pushed = self.lr

    def _moments(self, p):
        st = self.state[p]
        if "exp_avg" not in st:
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st["exp_avg"], st["exp_avg_sq"]

    def _build(self, active):
        dev = active[0].device
        self._device_scalars(dev)
        sizes = [p.numel() for p in active]
        offs = [0]
        for s in sizes:
            offs.append(offs[-1] + s)
        i64 = lambda xs: torch.tensor(xs, dtype=torch.int64, device=dev)
        mom = [self._moments(p) for p in active]
        self._p = i64([p.data_ptr() for p in active])
        self._g = i64([p.grad.data_ptr() for p in active])
        self._m = i64([m.data_ptr() for m, _ in mom])
        self._v = i64([v.data_ptr() for _, v in mom])
        self._off = i64(offs)
        self._n, self._total = len(active), offs[-1]
        self._built_for = tuple((p.data_ptr(), p.grad.data_ptr()) for p in active)

    @torch.no_grad()
    def step(self, closure=None):
        """Returns the device tensor [grad_norm, clip_coef] (no host sync)."""
        if closure is not None:
            raise NotImplementedError("FusedClipAdam.step does not take a closure")
        g = self.param_groups[0]
        active = [p for p in g["params"] if p.grad is not None and id(p) not in self._excluded]
        if not active:
            return None
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in active)
        if key != self._built_for:
            self._build(active)
        if not torch.cuda.is_current_stream_capturing():
            self.sync_hyper()
        stream = torch.cuda.current_stream().cuda_stream
        check(fn["uwr_increment_i32"](self._step_dev.data_ptr(), stream), "uwr_increment_i32")
        check(fn["uwr_grad_norm"](self._g.data_ptr(), self._off.data_ptr(), self._n, self._total,
                                  float(self.max_norm), float(self.grad_prescale), self._norm.data_ptr(),
                                  self._ws.data_ptr(), stream), "uwr_grad_norm")
        check(fn["uwr_adam_step"](self._p.data_ptr(), self._g.data_ptr(), self._m.data_ptr(), self._v.data_ptr(),
                                  self._off.data_ptr(), self._n, self._total, self._norm[1:].data_ptr(),
                                  float(self.grad_prescale), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
                                  float(g["eps"]), float(g["weight_decay"]), int(g["decoupled_weight_decay"]), 0,
                                  self._step_dev.data_ptr(), self._lr_dev.data_ptr(), stream), "uwr_adam_step")
        from . import ops
        ops.bump_weight_epoch()  # parameters changed behind torch's version counter: rounded copies are stale
        ops.refresh_rounded_copies()  # ... and are re-rounded here, all at once (two multi-tensor launches)
        return self._norm

    # ---------------------------------------------------------------------------------------- checkpoint (torch layout)
    def state_dict(self):
        """torch.optim.Adam layout; the per-parameter `step` entries are materialised from the device counter."""
        n = float(self.step_count)
        for p in self.param_groups[0]["params"]:
            st = self.state.get(p)
            if st is not None and "exp_avg" in st:
                st["step"] = torch.tensor(n, dtype=torch.float32)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        mine = self.param_groups[0]
        groups = [dict(g, decoupled_weight_decay=g.get("decoupled_weight_decay", mine["decoupled_weight_decay"]),
                       capturable=True, fused=True, foreach=None) for g in state_dict["param_groups"]]
        super().load_state_dict({"state": state_dict["state"], "param_groups": groups})
        for st in self.state.values():      # torch's loader may alias the caller's tensors: own the moments
            for k in ("exp_avg", "exp_avg_sq"):
                if k in st:
                    st[k] = st[k].clone()
        steps = [float(st["step"]) for st in self.state.values() if "step" in st]
        dev = next((p.device for p in self.param_groups[0]["params"] if p.is_cuda), None)
        if dev is not None:
            self._device_scalars(dev)
            self._step_dev.fill_(int(max(steps)) if steps else 0)
        self._built_for = None      # moment tensors were replaced: rebuild the pointer tables
        self._lr_pushed = None
