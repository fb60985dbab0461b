"""Functional CPU restatement of the reference SpectralTransformer (src/Models/SpectralTransformer.py:76-269).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Only the live data path is restated: in MDTA.forward the
FFT branch (lines 103-108), `attnf` (111-112) and q1X1_2 are dead in value and gradient (the second
product reuses `attn`, line 113), and ups_4 / ups1 / ups2 / output1 are never called — SURVEY.md §3.3.
"""
import torch
import torch.nn.functional as F


def _ln_nchw(sd, pre, x):
    """TransformerBlock's LayerNorm over channels of an NCHW tensor (lines 142-147)."""
    y = F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), sd[pre + "weight"], sd[pre + "bias"], 1e-5)
    return y.permute(0, 3, 1, 2)


def mdta(sd, pre, x, heads):
    """MDTA.forward live path (lines 92-101, 109, 113-114)."""
    b, c, h, w = x.shape
    qkv = F.conv2d(F.conv2d(x, sd[pre + "qkv.weight"]), sd[pre + "qkv_conv.weight"], padding=1, groups=3 * c)
    q, k, v = (t.reshape(b, heads, -1, h * w) for t in qkv.chunk(3, dim=1))
    q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
    attn = torch.softmax(q @ k.transpose(-2, -1) * sd[pre + "temperature"], dim=-1)
    out = F.conv2d((attn @ v).reshape(b, -1, h, w), sd[pre + "project_out.weight"])
    kvf = F.conv2d(F.conv2d(out, sd[pre + "kv.weight"]), sd[pre + "kv_conv.weight"], padding=1, groups=2 * c)
    vf = kvf.chunk(2, dim=1)[1].reshape(b, heads, -1, h * w)
    return F.conv2d((attn @ vf).reshape(b, -1, h, w), sd[pre + "project_outf.weight"])


def gdfn(sd, pre, x):
    """GDFN.forward (lines 127-130)."""
    t = F.conv2d(x, sd[pre + "project_in.weight"])
    t = F.conv2d(t, sd[pre + "conv.weight"], padding=1, groups=t.shape[1])
    x1, x2 = t.chunk(2, dim=1)
    return F.conv2d(F.gelu(x1) * x2, sd[pre + "project_out.weight"])


def block(sd, pre, x, heads):
    x = x + mdta(sd, pre + "attn.", _ln_nchw(sd, pre + "norm1.", x), heads)
    return x + gdfn(sd, pre + "ffn.", _ln_nchw(sd, pre + "norm2.", x))


def stage(sd, pre, x, n, heads):
    for i in range(n):
        x = block(sd, f"{pre}{i}.", x, heads)
    return x


def fft_upsample(sd, pre, x):
    """UpSample.forward (lines 174-188): amplitude / phase mixing, spectrum tiled 2x2, inverse FFT."""
    f = torch.fft.fft2(x)
    mag, pha = torch.abs(f), torch.angle(f)

    def mlp(t, name):
        t = F.conv2d(t, sd[pre + name + ".0.weight"], sd[pre + name + ".0.bias"])
        return F.conv2d(F.leaky_relu(t, 0.1), sd[pre + name + ".2.weight"], sd[pre + name + ".2.bias"])
    Mag, Pha = torch.tile(mlp(mag, "amp_fuse"), (2, 2)), torch.tile(mlp(pha, "pha_fuse"), (2, 2))
    out = torch.abs(torch.fft.ifft2(torch.complex(Mag * torch.cos(Pha), Mag * torch.sin(Pha))))
    return F.conv2d(out, sd[pre + "post.weight"], sd[pre + "post.bias"])


def ups(sd, pre, x):
    """UpS.forward (lines 208-210)."""
    s = F.pixel_shuffle(F.conv2d(x, sd[pre + "Sups.body.0.weight"], padding=1), 2)
    return F.conv2d(torch.cat([fft_upsample(sd, pre + "Fups.", x), s], 1), sd[pre + "reduce.weight"])


def down(sd, pre, x):
    return F.pixel_unshuffle(F.conv2d(x, sd[pre + "body.0.weight"], padding=1), 2)


def spectral_forward(sd, x, num_blocks=(2, 3, 3, 4), num_heads=(1, 2, 4, 8), num_refinement=4):
    """SpectralTransformer.forward (lines 254-269)."""
    f0 = F.conv2d(x, sd["embed_conv_rgb.weight"], padding=1)
    e1 = stage(sd, "encoders.0.", f0, num_blocks[0], num_heads[0])
    e2 = stage(sd, "encoders.1.", down(sd, "down1.", e1), num_blocks[1], num_heads[1])
    e3 = stage(sd, "encoders.2.", down(sd, "down2.", e2), num_blocks[2], num_heads[2])
    e4 = stage(sd, "encoders.3.", down(sd, "down3.", e3), num_blocks[3], num_heads[3])
    d3 = stage(sd, "decoders.0.", F.conv2d(torch.cat([ups(sd, "ups_1.", e4), e3], 1), sd["reduces1.weight"]),
               num_blocks[2], num_heads[2])
    d2 = stage(sd, "decoders.1.", F.conv2d(torch.cat([ups(sd, "ups_2.", d3), e2], 1), sd["reduces2.weight"]),
               num_blocks[1], num_heads[1])
    fd = stage(sd, "decoders.2.", torch.cat([ups(sd, "ups_3.", d2), e1], 1), num_blocks[0], num_heads[0])
    fr = stage(sd, "refinement.", fd, num_refinement, num_heads[0])
    return F.conv2d(F.conv2d(fr, sd["outputl.weight"], padding=1), sd["output.weight"], padding=1)
