"""Training step of the reference loop (src/ModelTrainer.py:78-88) on the uwr kernels:

    zero_grad -> model(raw) -> LossFunction.getloss -> backward -> [grad all-reduce] ->
    clip_grad_norm_(1.0) + Adam   (fused, no host sync; the reference's per-step .item()/print
    calls at ModelTrainer.py:90-126 are left to the caller)

Gradients live in a few flat buckets (views handed to autograd as .grad) so that
  * the optimizer's pointer tables are built once,
  * data parallelism is one NCCL all-reduce per bucket, launched from post-accumulate hooks as
    soon as a bucket is complete, i.e. overlapped with the rest of backward.
The loss divisor uses the GLOBAL batch so N-rank training reproduces the single-GPU global-batch
gradients (losses.py:57 divides by the local B).
"""
import torch
import torch.distributed as dist

from . import ops
from .losses import LossFunction
from .optim import FusedClipAdam


class GradBuckets:
    def __init__(self, params, bucket_bytes=32 << 20, group=None, world_size=1, adjacent=(), comm=None):
        """adjacent: pairs (p0, p1) whose slots must be laid out back to back, p0 first (their gradients come
        out of one kernel, e.g. the packed to_q | to_kv weight gradient: ops.fused_grad_slot).
        comm: a uwr.nccl.Communicator — the buckets are then reduced with raw, stream-ordered ncclAllReduce calls
        (graph-capturable) on a side stream instead of torch.distributed work objects."""
        self.params = [p for p in params if p.requires_grad]
        self.group, self.world = group, world_size
        self.comm = comm
        self._side = None
        order = list(reversed(self.params))  # roughly the order gradients are produced in backward
        follower = {id(a): b for a, b in adjacent}
        led = {id(b) for _, b in adjacent}
        fixed = []
        for p in order:
            if id(p) in led:
                continue              # placed right after its leader
            fixed.append(p)
            if id(p) in follower:
                fixed.append(follower[id(p)])
        order = fixed
        self.buckets, cur, cur_n = [], [], 0
        for p in order:
            cur.append(p)
            cur_n += p.numel()
            if cur_n * 4 >= bucket_bytes:
                self.buckets.append(cur)
                cur, cur_n = [], 0
        if cur:
            self.buckets.append(cur)
        self.flat, self._bucket_of, self._pending, self._handles = [], {}, [], []
        for bi, ps in enumerate(self.buckets):
            n = sum(-(-p.numel() // 4) * 4 for p in ps)  # 16-byte aligned slots
            flat = torch.zeros(n, device=ps[0].device, dtype=torch.float32)
            off = 0
            for p in ps:
                p.grad = flat[off:off + p.numel()].view_as(p)
                p._uwr_direct = True   # kernels may write this gradient in place (ops.grad_slot)
                off += -(-p.numel() // 4) * 4
                self._bucket_of[id(p)] = bi
            self.flat.append(flat)
        self._count = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self.overlap = True   # False: hooks only count, every bucket is reduced in finish() (graph-replayed backward)
        if self.world > 1:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._hook)

    def _reduce(self, bi):
        self._launched[bi] = True
        if self.comm is not None:
            # fork: the side stream picks up where the (possibly capturing) compute stream is now, reduces the bucket
            # there, and finish() joins it back -> the reduce overlaps whatever backward work is issued afterwards
            cur = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream()
            self._side.wait_stream(cur)
            self.comm.all_reduce_sum_(self.flat[bi], self._side)
            self._forked = True
            return
        self._handles.append(dist.all_reduce(self.flat[bi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _hook(self, p):
        bi = self._bucket_of[id(p)]
        self._count[bi] += 1
        if self.overlap and self._count[bi] == len(self.buckets[bi]):
            self._reduce(bi)

    def zero(self):
        for f in self.flat:
            f.zero_()
        self._count = [0] * len(self.buckets)
        self._launched = [False] * len(self.buckets)

    def finish(self):
        """Wait for the in-flight all-reduces; buckets not reduced yet — parameters that never received a
        gradient (dead parameters), or all of them when overlap is off — are reduced
        here so every rank stays in lock-step."""
        if self.world > 1:
            for bi in range(len(self.buckets)):
                if not self._launched[bi]:
                    self._reduce(bi)
            for h in self._handles:
                h.wait()
            self._handles = []
            if getattr(self, "_forked", False):
                torch.cuda.current_stream().wait_stream(self._side)   # join
                self._forked = False


class TrainStep:
    def __init__(self, model, loss_name="L1", lr=1e-3, optim="adam", world_size=1, group=None,
                 local_batch=None, bucket_bytes=32 << 20, vgg_weights=None, comm=None):
        self.model = model
        self.world = world_size
        self.lossf = LossFunction(loss_name, "cuda", batch_divisor=(local_batch * world_size) if local_batch else None,
                                  vgg_weights=vgg_weights)
        adjacent = model.adjacent_grad_pairs() if hasattr(model, "adjacent_grad_pairs") else ()
        self.buckets = GradBuckets(model.parameters(), bucket_bytes, group, world_size, adjacent=adjacent, comm=comm)
        self.opt = FusedClipAdam(model.parameters(), lr=lr, weight_decay=0.01 if optim == "adamw" else 0.0,
                                 decoupled=(optim == "adamw"), max_norm=1.0, grad_prescale=1.0 / world_size)

        self._dead_checked = False

    def _exclude_dead_parameters(self):
        """Parameters no gradient ever reaches (SpectralTransformer has 230 233 of them, NewBigFRFN 4.2 M: SURVEY.md
        §3.3/3.4) have `grad is None` under torch and are skipped by torch.optim (no moments, no AdamW decay); here their
        bucket slots simply stay zero.  Detected once, after the first backward (one host sync), and removed from the
        optimizer's tables so that the update rule matches."""
        dead = [p for p in self.buckets.params if not bool(p.grad.any())]
        if dead:
            self.opt.exclude(dead)
        self._dead_checked = True
        self.dead_parameters = dead

    def forward_backward(self, raw, ref):
        self.buckets.zero()
        out = self.model(raw)
        loss = self.lossf.getloss(out, ref)
        if isinstance(loss, tuple):
            loss = loss[0]
        # kernels may write parameter gradients straight into the (just zeroed) bucket slots only inside THIS backward:
        # one gradient per parameter per step, overwrite semantics (uwr.ops.grad_slot)
        with ops.direct_grad_writes():
            loss.backward()
        return loss.detach()

    def __call__(self, raw, ref):
        loss = self.forward_backward(raw, ref)
        self.buckets.finish()
        if not self._dead_checked and not torch.cuda.is_current_stream_capturing():
            self._exclude_dead_parameters()
        norm = self.opt.step()
        return loss, norm
