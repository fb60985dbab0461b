"""Functional CPU restatement of the reference AST (src/Models/AST.py) on a plain state_dict.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the reference lines it
follows.  Tensors are fp32 (or fp64 when the caller converts the state_dict) on the CPU; autograd
works through everything, so gradients of the oracle are the reference gradients.
"""
import math

import torch
import torch.nn.functional as F

WIN = 8


def shift_mask(H, W, shift, dtype=torch.float32, device=None):
    """(nW, 64, 64) additive mask of {0, -100}: AST.py:568-588 (region ids on the rolled grid)."""
    ys = torch.arange(H, device=device)
    xs = torch.arange(W, device=device)
    ry = (ys >= H - WIN).long() + (ys >= H - shift).long()
    rx = (xs >= W - WIN).long() + (xs >= W - shift).long()
    reg = (ry[:, None] * 3 + rx[None, :]).to(dtype)                     # (H, W)
    reg = reg.view(H // WIN, WIN, W // WIN, WIN).permute(0, 2, 1, 3).reshape(-1, WIN * WIN)
    diff = reg[:, None, :] - reg[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def to_windows(x, B, H, W, C):
    """window_partition, AST.py:377-389."""
    x = x.view(B, H // WIN, WIN, W // WIN, WIN, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, WIN * WIN, C)


def from_windows(w, B, H, W, C):
    """window_reverse, AST.py:392-402."""
    x = w.view(B, H // WIN, W // WIN, WIN, WIN, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)


def window_attention(sd, pre, xw, heads, mask, sparse=True, kv=None):
    """WindowAttention_sparse.forward, AST.py:187-219 (and WindowAttention 109-137 when sparse=False)."""
    B_, N, C = xw.shape
    hd = C // heads
    q = F.linear(xw, sd[pre + "qkv.to_q.weight"], sd.get(pre + "qkv.to_q.bias"))
    kvt = F.linear(xw if kv is None else kv, sd[pre + "qkv.to_kv.weight"], sd.get(pre + "qkv.to_kv.bias"))
    q = q.view(B_, N, heads, hd).transpose(1, 2)                        # AST.py:59
    k = kvt[..., :C].reshape(B_, N, heads, hd).transpose(1, 2)          # AST.py:60-62
    v = kvt[..., C:].reshape(B_, N, heads, hd).transpose(1, 2)
    s = (q * hd ** -0.5) @ k.transpose(-2, -1)                          # AST.py:190-191
    idx = sd[pre + "relative_position_index"].view(-1)
    bias = sd[pre + "relative_position_bias_table"][idx].view(N, N, heads).permute(2, 0, 1)
    s = s + bias.unsqueeze(0)                                           # AST.py:193-199
    if mask is not None:                                                # AST.py:201-205
        nW = mask.shape[0]
        s = (s.view(B_ // nW, nW, heads, N, N) + mask[None, :, None]).view(-1, heads, N, N)
    p = torch.softmax(s, dim=-1)
    if sparse:                                                          # AST.py:206-213
        w = sd[pre + "w"]
        e = torch.exp(w)
        p = p * (e[0] / e.sum()) + torch.relu(s) ** 2 * (e[1] / e.sum())
    o = (p @ v).transpose(1, 2).reshape(B_, N, C)                       # AST.py:216
    return F.linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"])  # AST.py:217


def leff(sd, pre, x, H, W):
    """LeFF.forward, AST.py:307-326."""
    B, L, C = x.shape
    h = F.gelu(F.linear(x, sd[pre + "linear1.0.weight"], sd[pre + "linear1.0.bias"]))
    Ch = h.shape[-1]
    h = h.transpose(1, 2).reshape(B, Ch, H, W)
    h = F.gelu(F.conv2d(h, sd[pre + "dwconv.0.weight"], sd[pre + "dwconv.0.bias"], padding=1, groups=Ch))
    h = h.flatten(2).transpose(1, 2)
    return F.linear(h, sd[pre + "linear2.0.weight"], sd[pre + "linear2.0.bias"])


def frfn(sd, pre, x, H, W):
    """FRFN.forward, AST.py:345-372 (== block.py:263-282)."""
    B, L, C = x.shape
    cc = C // 4
    img = x.transpose(1, 2).reshape(B, C, H, W)
    x1 = F.conv2d(img[:, :cc], sd[pre + "partial_conv3.weight"], None, padding=1)
    img = torch.cat([x1, img[:, cc:]], 1)
    t = img.flatten(2).transpose(1, 2)
    u = F.gelu(F.linear(t, sd[pre + "linear1.0.weight"], sd[pre + "linear1.0.bias"]))
    u1, u2 = u.chunk(2, dim=-1)
    Ch = u1.shape[-1]
    u1 = u1.transpose(1, 2).reshape(B, Ch, H, W)
    u1 = F.gelu(F.conv2d(u1, sd[pre + "dwconv.0.weight"], sd[pre + "dwconv.0.bias"], padding=1, groups=Ch))
    u1 = u1.flatten(2).transpose(1, 2)
    return F.linear(u1 * u2, sd[pre + "linear2.0.weight"], sd[pre + "linear2.0.bias"])


def transformer_block(sd, pre, x, heads, shift, att, token_mlp, dp_attn=None, dp_mlp=None):
    """TransformerBlock.forward, AST.py:552-624. dp_* are per-sample DropPath scales (mask/keep) or None."""
    B, L, C = x.shape
    H = W = int(math.sqrt(L))
    if att:
        y = F.layer_norm(x, (C,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-5).view(B, H, W, C)
        mask = None
        if shift > 0:
            y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
            mask = shift_mask(H, W, shift, x.dtype, x.device)
        yw = window_attention(sd, pre + "attn.", to_windows(y, B, H, W, C), heads, mask,
                              sparse=(pre + "attn.w") in sd)
        y = from_windows(yw, B, H, W, C)
        if shift > 0:
            y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
        y = y.reshape(B, L, C)
        if dp_attn is not None:
            y = y * dp_attn.view(B, 1, 1)
        x = x + y
    y = F.layer_norm(x, (C,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-5)
    if token_mlp == "leff":
        y = leff(sd, pre + "mlp.", y, H, W)
    elif token_mlp == "frfn":
        y = frfn(sd, pre + "mlp.", y, H, W)
    else:  # Mlp.forward, AST.py:285-291
        y = F.linear(F.gelu(F.linear(y, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])),
                     sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    if dp_mlp is not None:
        y = y * dp_mlp.view(B, 1, 1)
    return x + y


def downsample(sd, pre, x):
    """Downsample.forward, AST.py:417-424."""
    B, L, C = x.shape
    H = W = int(math.sqrt(L))
    img = x.transpose(1, 2).reshape(B, C, H, W)
    return F.conv2d(img, sd[pre + "conv.0.weight"], sd[pre + "conv.0.bias"], stride=2, padding=1).flatten(2).transpose(1, 2)


def upsample(sd, pre, x):
    """Upsample.forward, AST.py:437-443."""
    B, L, C = x.shape
    H = W = int(math.sqrt(L))
    img = x.transpose(1, 2).reshape(B, C, H, W)
    return F.conv_transpose2d(img, sd[pre + "deconv.0.weight"], sd[pre + "deconv.0.bias"], stride=2).flatten(2).transpose(1, 2)


def ast_forward(sd, x, *, img_size=256, num_heads=(1, 2, 4, 8, 16, 16, 8, 4, 2), depths=(2,) * 9, token_mlp="leff",
                shift_flag=True, drop_scales=None):
    """AST.forward, AST.py:885-921.  drop_scales: {block_prefix: (attn_scale|None, mlp_scale|None)}."""
    drop_scales = drop_scales or {}
    B, _, Himg, Wimg = x.shape
    y = F.leaky_relu(F.conv2d(x, sd["input_proj.proj.0.weight"], sd["input_proj.proj.0.bias"], padding=1), 0.01)
    y = y.flatten(2).transpose(1, 2)                                    # AST.py:461-466

    def stage(name, t, idx, level, att):
        res = img_size // (2 ** level)          # constructor-time resolution decides the shift (AST.py:515-517,647)
        for i in range(depths[idx]):
            shift = (WIN // 2 if (i % 2 == 1 and shift_flag) else 0)
            if res <= WIN:
                shift = 0
            pre = f"{name}.blocks.{i}."
            da, dm = drop_scales.get(pre, (None, None))
            t = transformer_block(sd, pre, t, num_heads[idx], shift, att, token_mlp, da, dm)
        return t

    conv0 = stage("encoderlayer_0", y, 0, 0, False)
    conv1 = stage("encoderlayer_1", downsample(sd, "dowsample_0.", conv0), 1, 1, False)
    conv2 = stage("encoderlayer_2", downsample(sd, "dowsample_1.", conv1), 2, 2, False)
    conv3 = stage("encoderlayer_3", downsample(sd, "dowsample_2.", conv2), 3, 3, False)
    conv4 = stage("conv", downsample(sd, "dowsample_3.", conv3), 4, 4, True)
    d0 = stage("decoderlayer_0", torch.cat([upsample(sd, "upsample_0.", conv4), conv3], -1), 5, 3, True)
    d1 = stage("decoderlayer_1", torch.cat([upsample(sd, "upsample_1.", d0), conv2], -1), 6, 2, True)
    d2 = stage("decoderlayer_2", torch.cat([upsample(sd, "upsample_2.", d1), conv1], -1), 7, 1, True)
    d3 = stage("decoderlayer_3", torch.cat([upsample(sd, "upsample_3.", d2), conv0], -1), 8, 0, True)
    C = d3.shape[-1]
    img = d3.transpose(1, 2).reshape(B, C, Himg, Wimg)                  # AST.py:485-490
    return x + F.conv2d(img, sd["output_proj.proj.0.weight"], sd["output_proj.proj.0.bias"], padding=1)
