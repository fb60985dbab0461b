import gc, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "underwater-image-restoration_b200"))
import torch
import uwr
from uwr import fflmix, ssim as dev
from uwr.losses import LossFunction, PixelLossFn
from uwr.ffl import FocalFrequencyFn
from uwr.train import TrainStep

# memory held after eager steps (B=64 did not fit next to the graph pool)
model = uwr.AST().cuda().train()
step = TrainStep(model, "L1", lr=1e-3, local_batch=16)
import gc
raw = torch.rand(16, 3, 256, 256, device="cuda") * 2 - 1
ref = torch.rand(16, 3, 256, 256, device="cuda") * 2 - 1
base = torch.cuda.memory_allocated()
for _ in range(2):
    step(raw, ref)
torch.cuda.synchronize(); gc.collect(); torch.cuda.empty_cache()
print(f"allocated before {base / 2**30:.2f} GiB, after 2 eager steps {torch.cuda.memory_allocated() / 2**30:.2f} GiB, "
      f"peak {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
big = [(o.numel() * o.element_size() / 2**20, tuple(o.shape), o.dtype) for o in gc.get_objects()
       if torch.is_tensor(o) and o.is_cuda and o.numel() * o.element_size() > 64 * 2**20]
big.sort(reverse=True)
print("live CUDA tensors > 64 MiB:", len(big), big[:12])
