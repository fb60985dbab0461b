"""CPU restatement of the reference's Haar "wavelet" modules (src/model/wave_modules.py:9-181) on token tensors.

TEST INFRASTRUCTURE (see oracle/__init__.py).  DWT_2D / IDWT_2D broadcast one 2x2 Haar tap pattern over every
(out, in) channel pair, and their custom autograd backward passes are NOT the adjoints of their forwards; both are
restated here in closed form (the same formulas the CUDA kernels of csrc/wavelet.cu implement) and pinned against the
reference modules themselves in tests/test_oracle_golden.py.  Tokens: (B, H*W, C), NHWC.
"""
import torch

_S = (  # sign pattern of sub-band n at (dy, dx); magnitude 0.5.  Order ll, lh, hl, hh (wave_modules.py:21)
    ((1, 1), (1, 1)),
    ((1, 1), (-1, -1)),     # lh: w[i][j] = lo[j] * hi[i]   (wave_modules.py:128)
    ((1, -1), (1, -1)),     # hl: w[i][j] = hi[j] * lo[i]   (wave_modules.py:127)
    ((1, -1), (-1, 1)),
)


def _w(n, dtype):
    return 0.5 * torch.tensor(_S[n], dtype=dtype)


def dwt_fwd(x, B, h, w):
    """x: (B, 2h*2w, C) -> (B, h*w, C)  (DWT_function.forward, wave_modules.py:10-24)"""
    C = x.shape[-1]
    S = x.view(B, h, 2, w, 2, C).sum(-1)                      # (B, h, dy, w, dx)
    subs = [sum(_w(n, x.dtype)[dy, dx] * S[:, :, dy, :, dx] for dy in range(2) for dx in range(2)) for n in range(4)]
    out = torch.stack(subs, -1)                                # (B, h, w, 4)
    return out.repeat_interleave(C // 4, dim=-1).reshape(B, h * w, C)


def dwt_bwd(dout, B, h, w):
    """reference 'gradient' of dwt_fwd (DWT_function.backward, wave_modules.py:27-53): dout (B, h*w, C) -> (B, 2h*2w, C)"""
    C = dout.shape[-1]
    q = C // 4
    j = torch.arange(C)
    perm = (j % 4) * q + j // 4                                # channel read at reordered position j
    R = dout[..., perm].view(B, h, w, 4, q).sum(-1)            # (B, h, w, n)
    out = dout.new_zeros(B, h, 2, w, 2)
    for n in range(4):
        for dy in range(2):
            for dx in range(2):
                out[:, :, dy, :, dx] += _w(n, dout.dtype)[dy, dx] * R[..., n]
    return out.reshape(B, 4 * h * w, 1).expand(B, 4 * h * w, C).contiguous()


def idwt_fwd(x, B, h, w):
    """x: (B, h*w, C) -> (B, 2h*2w, C)  (IDWT_function.forward, wave_modules.py:57-76)"""
    C = x.shape[-1]
    T = x.view(B, h, w, C // 4, 4).sum(-1)                     # (B, h, w, g)
    out = x.new_zeros(B, h, 2, w, 2, C // 4, 4)
    for o in range(4):
        for dy in range(2):
            for dx in range(2):
                out[:, :, dy, :, dx, :, o] = _w(o, x.dtype)[dy, dx] * T
    return out.reshape(B, 4 * h * w, C)


def idwt_bwd(dout, B, h, w):
    """reference 'gradient' of idwt_fwd (IDWT_function.backward, wave_modules.py:78-116): dout (B, 2h*2w, C) -> (B, h*w, C)"""
    C = dout.shape[-1]
    flat = dout.view(B, 2 * h, 2 * w, C).permute(0, 3, 1, 2).contiguous().view(B, 16 * C, h // 2, w // 2)
    r = flat.view(B, 16 * C, h // 4, 2, w // 4, 2).sum(1)      # (B, h/4, dy, w/4, dx)
    V = [sum(_w(n, dout.dtype)[dy, dx] * r[:, :, dy, :, dx] for dy in range(2) for dx in range(2)).reshape(B, -1)
         for n in range(4)]                                     # each (B, h*w/16)
    m = h * w // 16
    q = C // 4
    idx = (torch.arange(q).view(q, 1) * (h * w) + torch.arange(h * w).view(1, h * w)) % m     # (c', yx)
    out = torch.stack([V[n][:, idx] for n in range(4)], 1)     # (B, n, c', yx)
    return out.reshape(B, C, h * w).transpose(1, 2).contiguous()
