"""2-rank check on real GPUs: N-rank data-parallel step == 1-rank step at the same global batch
(same weights, eval-mode DropPath off): gradients after all-reduce and parameters after one
clip+Adam step agree."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "underwater-image-restoration_b200"))
import torch
import torch.distributed as dist
import uwr
from uwr.train import TrainStep

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from uwr.nccl import Communicator
COMM = Communicator(rank, world)          # raw NCCL communicator: graph-capturable all-reduces (csrc/dp.cu)
S, Bl = 128, 2
g = torch.Generator().manual_seed(5)
raw = torch.rand(Bl * world, 3, S, S, generator=g) * 2 - 1
ref = torch.rand(Bl * world, 3, S, S, generator=g) * 2 - 1


def run(ws, batch_slice):
    torch.manual_seed(1234)
    m = uwr.AST(img_size=S).cuda().eval()   # eval: DropPath off so that both runs are deterministic
    step = TrainStep(m, "L2", lr=1e-3, world_size=ws, local_batch=batch_slice.stop - batch_slice.start)
    loss, norm = step(raw[batch_slice].cuda(), ref[batch_slice].cuda())
    torch.cuda.synchronize()
    return m, loss, norm


def run_steps(ws, batch_slice, graphed, nsteps=2, comm=None):
    """nsteps optimizer steps, eagerly (overlapped all-reduce) or through GraphedTrainStep: with `comm` the whole
    step incl. the raw NCCL all-reduces is ONE CUDA graph, without it forward+backward and clip+Adam are two graphs
    around eager torch.distributed all-reduces."""
    from uwr.graph import GraphedTrainStep
    torch.manual_seed(1234)
    m = uwr.AST(img_size=S).cuda().eval()
    step = TrainStep(m, "L2", lr=1e-3, world_size=ws, local_batch=batch_slice.stop - batch_slice.start, comm=comm)
    r, t = raw[batch_slice].cuda(), ref[batch_slice].cuda()
    if graphed:
        gs = GraphedTrainStep(step, r, t, warmup=nsteps - 1)
        gs(r, t)
    else:
        for _ in range(nsteps):
            step(r, t)
    torch.cuda.synchronize()
    return m


m_dp, loss_dp, norm_dp = run(world, slice(rank * Bl, (rank + 1) * Bl))
m_e = run_steps(world, slice(rank * Bl, (rank + 1) * Bl), graphed=False)
m_g = run_steps(world, slice(rank * Bl, (rank + 1) * Bl), graphed=True)
m_c = run_steps(world, slice(rank * Bl, (rank + 1) * Bl), graphed=False, comm=COMM)
m_cg = run_steps(world, slice(rank * Bl, (rank + 1) * Bl), graphed=True, comm=COMM)
if rank == 0:
    den = sum((b.double() ** 2).sum() for b in m_e.parameters()).sqrt().item()
    for tag, mm in (("2 graphs + eager torch.distributed all-reduce", m_g), ("raw NCCL, eager", m_c),
                    ("raw NCCL captured in ONE graph", m_cg)):
        num = sum(((a - b).double() ** 2).sum() for a, b in zip(mm.parameters(), m_e.parameters())).sqrt().item()
        print(f"DP [{tag}] vs eager overlapped, 2 steps: param rel diff {num / den:.2e}")
        assert num / den < 1e-6
if rank == 0:
    m_1, loss_1, norm_1 = run(1, slice(0, Bl * world))
    num = sum(((a - b).double() ** 2).sum() for a, b in zip(m_dp.parameters(), m_1.parameters())).sqrt().item()
    den = sum((b.double() ** 2).sum() for b in m_1.parameters()).sqrt().item()
    print(f"DP parity world={world}: param rel diff after 1 step {num / den:.2e}; "
          f"grad norm dp {norm_dp[0].item():.6e} vs single {norm_1[0].item():.6e}")
    assert num / den < 1e-6 and abs(norm_dp[0].item() - norm_1[0].item()) < 2e-4 * norm_1[0].item()
    print("DP PARITY OK")
dist.barrier()
dist.destroy_process_group()
