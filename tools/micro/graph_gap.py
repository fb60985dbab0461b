"""How long does a dependent kernel node cost inside a replayed CUDA graph?  N tiny uwr launches (scale_round of 4 KB) in one
graph: time per node = launch gap + the minimal kernel time.  usage: python tools/micro/graph_gap.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "underwater-image-restoration_b200"))
import torch
from uwr import ops

x = torch.randn(32, 32, device="cuda")
y = torch.empty_like(x)
big = torch.randn(1 << 20, 64, device="cuda")
bigo = torch.empty_like(big)
for n, (src, dst, cols) in (("tiny (4 KB)", (x, y, 32)), ("268 MB", (big, bigo, 64))):
    N = 500 if n.startswith("tiny") else 50
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            ops.scale_round(src, cols, out=dst)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(N):
                ops.scale_round(src, cols, out=dst)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{n}: {e0.elapsed_time(e1) / 5 / N * 1e3:.2f} us per node in a graph of {N}")
    e0.record()
    for _ in range(N):
        ops.scale_round(src, cols, out=dst)
    e1.record(); torch.cuda.synchronize()
    print(f"{n}: {e0.elapsed_time(e1) / N * 1e3:.2f} us per launch, eager")
