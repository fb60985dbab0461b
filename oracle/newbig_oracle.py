"""Functional CPU restatement of the reference's NewBigFRFNModel (src/model/model.py:465-640 with the
blocks of src/model/block.py) in its only reachable mode, use_dwt="Fourier" (SURVEY.md §0).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Includes the documented patch P1:
the token->NCHW transpose that MyBigModel has at model.py:435-437 is applied before output_proj
(the model as shipped raises at model.py:637).  DropPath: encoder blocks draw two masks
(drop_path2 for the frequency branch, drop_path for the spatial branch, model.py:90); decoder
blocks are constructed with drop_path=0.
"""
import math

import torch
import torch.nn.functional as F

from .ast_oracle import frfn, from_windows, to_windows

WIN = 8


def _ln(sd, pre, x):
    return F.layer_norm(x, (x.shape[-1],), sd[pre + "weight"], sd[pre + "bias"], 1e-5)


def sparse_window_attention(sd, pre, xq, heads, kv_in=None):
    """WindowAttention_Sparse.forward (block.py:325-370) with LinearProjection (block.py:166-200):
    self-attention uses to_q / to_kv_from_q, cross-attention to_q / to_kv (2C -> 2C)."""
    B_, N, C = xq.shape
    hd = C // heads
    q = F.linear(xq, sd[pre + "to_qkv.to_q.weight"], sd[pre + "to_qkv.to_q.bias"])
    if kv_in is None:
        kv = F.linear(xq, sd[pre + "to_qkv.to_kv_from_q.weight"], sd[pre + "to_qkv.to_kv_from_q.bias"])
    else:
        kv = F.linear(kv_in, sd[pre + "to_qkv.to_kv.weight"], sd[pre + "to_qkv.to_kv.bias"])
    q = q.view(B_, N, heads, hd).transpose(1, 2) * hd ** -0.5
    k = kv[..., :C].reshape(B_, N, heads, hd).transpose(1, 2)
    v = kv[..., C:].reshape(B_, N, heads, hd).transpose(1, 2)
    s = q @ k.transpose(-2, -1)
    idx = sd[pre + "relative_position_index"].view(-1)
    s = s + sd[pre + "relative_position_bias_table"][idx].view(N, N, heads).permute(2, 0, 1).unsqueeze(0)
    e = torch.exp(sd[pre + "w"])
    p = torch.softmax(s, -1) * (e[0] / e.sum()) + torch.relu(s) ** 2 * (e[1] / e.sum())
    o = (p @ v).transpose(1, 2).reshape(B_, N, C)
    return F.linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def fdfp(sd, pre, x):
    """FDFP.forward, Fourier mode (block.py:532-556): x (B,H,W,C)."""
    f = torch.fft.fftn(x, dim=(1, 2)).real                       # per channel over (H, W)
    f = F.linear(f, sd[pre + "conv1.weight"].flatten(1), sd[pre + "conv1.bias"])   # 1x1 conv C -> 2C
    f = F.linear(F.gelu(f), sd[pre + "conv2.weight"].flatten(1), sd[pre + "conv2.bias"])
    return torch.fft.ifftn(f, dim=(1, 2)).real


def mdassa(sd, pre, x, heads, H, W):
    """MDASSA.forward with shift_size = 0 (block.py:408-515). x: (B, L, D) -> (B, L, D)."""
    B, L, D = x.shape
    x = _ln(sd, pre + "norm1.", x)
    xs = x.view(B, H, W, D)
    aw = sparse_window_attention(sd, pre + "attn.", to_windows(xs, B, H, W, D), heads)
    xa = x + from_windows(aw, B, H, W, D).reshape(B, L, D)        # shortcut + attention
    fq = fdfp(sd, pre + "fdfp.", xs)                              # (B,H,W,D)
    kv = F.linear(xa, sd[pre + "conv1x1.weight"].flatten(1), sd[pre + "conv1x1.bias"]).view(B, H, W, 2 * D)
    fw = sparse_window_attention(sd, pre + "freq_attn.", to_windows(fq, B, H, W, D), heads,
                                 kv_in=to_windows(kv, B, H, W, 2 * D))
    return (fq + from_windows(fw, B, H, W, D)).reshape(B, L, D)


def encoder_block(sd, pre, x, dp_freq=None, dp_spatial=None):
    """EncoderBlock.forward, Fourier mode (model.py:57-93): norm2(x) is computed and discarded there."""
    B, L, C = x.shape
    H = W = int(math.sqrt(L))
    a = frfn(sd, pre + "mlp.", _ln(sd, pre + "norm1.", x), H, W)
    f = torch.fft.fftn(a, dim=(-2, -1)).real                      # 2-D DFT over the (token, channel) axes
    f = frfn(sd, pre + "freq_mlp.", f, H, W)
    f = torch.fft.ifftn(f, dim=(-2, -1)).real
    if dp_freq is not None:
        f = f * dp_freq.view(B, 1, 1)
    if dp_spatial is not None:
        a = a * dp_spatial.view(B, 1, 1)
    return x + f + a


def decoder_block(sd, pre, x, skip=None, heads=4):
    """DecoderBlock.forward (model.py:141-160); drop_path is Identity for every decoder (ctor passes 0)."""
    if skip is not None:
        x = torch.cat([x, skip], 2)
    B, L, D = x.shape
    H = W = int(math.sqrt(L))
    y = mdassa(sd, pre + "mdassa.", _ln(sd, pre + "norm1.", x), heads, H, W) + x
    z = y + frfn(sd, pre + "mlp.", _ln(sd, pre + "norm2.", y), H, W)
    return F.linear(z, sd[pre + "mlp_proj.weight"], sd[pre + "mlp_proj.bias"])


def _img(t):
    B, L, C = t.shape
    H = int(math.sqrt(L))
    return t.transpose(1, 2).reshape(B, C, H, H)


def _tok(img):
    return img.flatten(2).transpose(1, 2)


def downsample(sd, pre, x):
    """Downsample (block.py:107-122): Conv3x3 C -> C/2 (no bias) + PixelUnshuffle(2)."""
    return _tok(F.pixel_unshuffle(F.conv2d(_img(x), sd[pre + "body.0.weight"], None, padding=1), 2))


def upsample(sd, pre, x):
    """Upsample (block.py:138-153): Conv3x3 C -> 2C (no bias) + PixelShuffle(2)."""
    return _tok(F.pixel_shuffle(F.conv2d(_img(x), sd[pre + "body.0.weight"], None, padding=1), 2))


def newbig_frfn_forward(sd, x, drop_scales=None):
    """MyBigFRFNModel.forward (model.py:594-640) + patch P1. drop_scales: {encoder prefix: (freq, spatial)}."""
    ds = drop_scales or {}
    y = x
    for i in range(3):                                             # InputProjection (block.py:42-63)
        y = F.conv2d(y, sd[f"input_proj.proj.{i}.weight"], sd[f"input_proj.proj.{i}.bias"], padding=1)
    y = _tok(F.leaky_relu(y, 0.01))
    skips = []
    for l in range(4):
        for name in (f"encoder_{l}.", f"encoder_{l}_1."):
            y = encoder_block(sd, name, y, *ds.get(name, (None, None)))
        skips.append(y)
        y = downsample(sd, f"downsample_{l}.", y)
    y = decoder_block(sd, "bottleneck.", y)
    for l in (3, 2, 1, 0):
        y = upsample(sd, f"upsample_{l}.", y)
        y = decoder_block(sd, f"decoder_{l}.", y, skips[l])
        y = decoder_block(sd, f"decoder_{l}_1.", y)
    o = _img(y)                                                    # patch P1 (model.py:435-437)
    for i in range(3):                                             # OutputProjection (block.py:65-91)
        o = F.conv2d(o, sd[f"output_proj.proj.{i}.weight"], sd[f"output_proj.proj.{i}.bias"], padding=1)
    return o + x
