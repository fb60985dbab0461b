"""CPU, world_size 2 over gloo: the gradient buckets of uwr.train (views handed to autograd,
all-reduce launched from post-accumulate hooks, dead-parameter buckets flushed in finish())
reproduce the single-process global-batch gradients."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, PKG


def _worker(rank, world, port, out, overlap=True):
    for p in (ROOT, PKG):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from uwr.train import GradBuckets
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 8))
    dead = torch.nn.Parameter(torch.ones(5))          # never receives a gradient
    params = list(net.parameters()) + [dead]
    buckets = GradBuckets(params, bucket_bytes=1024, world_size=world)
    buckets.overlap = overlap   # False: hooks only count, finish() reduces every bucket (graph-replayed backward)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 16, generator=g)
    y = torch.randn(8, 8, generator=g)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    for _ in range(2):                                  # second pass checks zero()/re-arming
        buckets.zero()
        loss = ((net(xs) - ys) ** 2).sum() / 8          # divisor = GLOBAL batch
        loss.backward()
        buckets.finish()
    if rank == 0:
        torch.save([p.grad.clone() for p in params], out)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("overlap", [True, False])
def test_grad_buckets_match_global_batch(tmp_path, overlap):
    out = str(tmp_path / "grads.pt")
    port = 29500 + (os.getpid() + (0 if overlap else 17)) % 1000
    mp.spawn(_worker, args=(2, port, out, overlap), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.GELU(), torch.nn.Linear(32, 8))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 16, generator=g)
    y = torch.randn(8, 8, generator=g)
    (((net(x) - y) ** 2).sum() / 8).backward()
    for a, p in zip(got, net.parameters()):
        assert torch.allclose(a, p.grad, rtol=1e-5, atol=1e-7)   # SUM over ranks of local/global-divisor grads
    assert torch.count_nonzero(got[-1]) == 0


def test_grad_buckets_adjacent_pairs_are_contiguous():
    """GradBuckets(adjacent=[(a, b)]): b's gradient slot starts right where a's ends (what ops.fused_grad_slot
    needs for the packed to_q | to_kv gradient), every parameter still owns exactly one slot."""
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from uwr.train import GradBuckets
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(*s)) for s in ((8, 4), (8,), (16, 4), (16,), (5,), (4, 4))]
    buckets = GradBuckets(ps, bucket_bytes=1 << 20, adjacent=[(ps[0], ps[2]), (ps[1], ps[3])])
    for a, b in ((ps[0], ps[2]), (ps[1], ps[3])):
        assert a.grad.data_ptr() + a.numel() * 4 == b.grad.data_ptr()
    seen = set()
    for p in ps:
        assert p.grad.shape == p.shape and p.grad.data_ptr() % 16 == 0
        rng = (p.grad.data_ptr(), p.grad.data_ptr() + p.numel() * 4)
        assert all(rng[1] <= s or rng[0] >= e for s, e in seen)   # no overlap
        seen.add(rng)
    buckets.zero()
    assert all(torch.count_nonzero(p.grad) == 0 for p in ps)
