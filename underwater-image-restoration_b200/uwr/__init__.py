"""uwr — B200-native hot path (forward/backward/losses/optimizer step) of the restoration
transformers of KarthikSundar2002/Underwater-Image-Restoration, behind the reference's own
registry / nn.Module / LossFunction surface.

    from uwr import init_model, get_names, LossFunction
    model = init_model("AST").cuda()

The kernels are also registered as `torch.library` ops (`torch.ops.uwr.linear`, `.layernorm`,
`.window_attn_sparse`, `.fused_window_block`, ... — uwr/torchlib.py) with fake impls and autograd.
"""
from . import _lib  # noqa: F401  (fails loudly when libuwr_b200.so is missing)
from .registry import get_names, init_model  # noqa: F401
from .losses import LossFunction  # noqa: F401
from .ast import AST  # noqa: F401
from .optim import FusedClipAdam  # noqa: F401
from .metrics import torchPSNR, torchSSIM  # noqa: F401
from . import torchlib  # noqa: F401  (registers torch.ops.uwr.*)
