"""SpectralTransformer (src/Models/SpectralTransformer.py:213-269) behind the reference's module tree
(443 state_dict entries, RNG-identical default init), computed on NHWC token matrices.

Native (uwr kernels): LayerNorm over channels (no NCHW<->NHWC round trips), every 1x1 convolution
(tensor-core GEMM on tokens), the depthwise 3x3 convs of MDTA / GDFN (plain mode of the dwconv kernel),
MDTA's channel attention as two GEMMs per image (Gram matrix of [q|k] over the tokens -> tiny softmax
-> block-diagonal apply), the dense 3x3 convs with Cin % 4 == 0 (im2col + GEMM).
The FFT amplitude/phase up-sampler runs its fft2 / ifft2 (and their adjoints) on the shared-memory FFT
passes of csrc/fft.cu; the (2, 2)-tiled inverse transform is an H x W one scattered to the even pixels.
The up-sampler's abs/angle, mag*exp(i pha), |ifft2| and LeakyReLU chains are fused elementwise kernels (csrc/spectral_ew.cu),
PixelShuffle / PixelUnshuffle are token-layout scatters (csrc/conv_io.cu), the 3-channel first / last 3x3 convs are thin direct
convolutions (csrc/conv_small.cu).  ATen is left with torch.cat of skip connections and the residual adds.
Only the live data path is executed: MDTA's FFT branch, `attnf`, q1X1_*, ups_4, ups1, ups2, output1
are dead in value and gradient in the reference; their parameters are kept.
"""
import torch
import torch.nn as nn

from . import fn


def _w2(conv):
    return conv.weight.flatten(1)


class MDTA(nn.Module):
    def __init__(self, channels, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.temperature = nn.Parameter(torch.ones(1, num_heads, 1, 1))
        self.qkv = nn.Conv2d(channels, channels * 3, kernel_size=1, bias=False)
        self.qkv_conv = nn.Conv2d(channels * 3, channels * 3, kernel_size=3, padding=1, groups=channels * 3, bias=False)
        self.project_out = nn.Conv2d(channels, channels, kernel_size=1, bias=False)
        self.kv = nn.Conv2d(channels, channels * 2, kernel_size=1, bias=False)
        self.q1X1_1 = nn.Conv2d(channels, channels, kernel_size=1, bias=False)   # dead
        self.q1X1_2 = nn.Conv2d(channels, channels, kernel_size=1, bias=False)   # dead
        self.kv_conv = nn.Conv2d(channels * 2, channels * 2, kernel_size=3, padding=1, groups=channels * 2, bias=False)
        self.project_outf = nn.Conv2d(channels, channels, kernel_size=1, bias=False)

    def forward(self, y, B, H, W):   # y: LayerNorm output tokens (B*L, C)
        L, C = H * W, y.shape[1]
        qkv = fn.LinearDWConvFn.apply(y, _w2(self.qkv), self.qkv_conv.weight, B, H, W, True)
        # channel attention for the whole batch: Gram GEMMs on strided views, batched normalise / softmax
        out, attn = fn.MDTAAttnFn.apply(qkv, self.temperature, B, L, C, self.num_heads)        # attn @ v
        out = fn.linear(out, _w2(self.project_out), rounded=True)       # uwr_mdta_apply rounds its output to TF32
        kvf = fn.LinearDWConvFn.apply(out, _w2(self.kv), self.kv_conv.weight, B, H, W, False)
        outf = fn.ChannelApplyFn.apply(kvf, attn, C, B, L)                                       # attn @ vf
        return fn.linear(outf, _w2(self.project_outf), rounded=True)


class GDFN(nn.Module):
    def __init__(self, channels, expansion_factor):
        super().__init__()
        hidden = int(channels * expansion_factor)
        self.hidden = hidden
        self.project_in = nn.Conv2d(channels, hidden * 2, kernel_size=1, bias=False)
        self.conv = nn.Conv2d(hidden * 2, hidden * 2, kernel_size=3, padding=1, groups=hidden * 2, bias=False)
        self.project_out = nn.Conv2d(hidden, channels, kernel_size=1, bias=False)

    def forward(self, y, B, H, W):
        # hidden = int(2.66 C) is odd-sized (42, 85, 170, 340): pad each half to a multiple of 4 channels with
        # zero weights so every row of the token matrices stays 16-byte aligned (zeros stay zeros:
        # conv(0) = 0 and gelu(0) * 0 = 0)
        h = self.hidden
        hp = -(-h // 4) * 4
        win, wdw, wout = _w2(self.project_in), self.conv.weight, _w2(self.project_out)
        if hp != h:
            z1 = win.new_zeros(hp - h, win.shape[1])
            win = torch.cat([win[:h], z1, win[h:], z1], 0)
            z2 = wdw.new_zeros(hp - h, 1, 3, 3)
            wdw = torch.cat([wdw[:h], z2, wdw[h:], z2], 0)
            wout = torch.cat([wout, wout.new_zeros(wout.shape[0], hp - h)], 1)
        t = fn.LinearDWConvFn.apply(y, win, wdw, B, H, W, True)
        return fn.linear(fn.GeluMulFn.apply(t, hp), wout, rounded=True)   # the gate kernel rounds its output


class TransformerBlock(nn.Module):
    def __init__(self, channels, num_heads, expansion_factor):
        super().__init__()
        self.norm1 = nn.LayerNorm(channels)
        self.attn = MDTA(channels, num_heads)
        self.norm2 = nn.LayerNorm(channels)
        self.ffn = GDFN(channels, expansion_factor)

    def forward(self, x, B, H, W):   # tokens (B*L, C)
        x = x + self.attn(fn.layernorm(x, self.norm1), B, H, W)
        return x + self.ffn(fn.layernorm(x, self.norm2), B, H, W)


def _conv3x3(t, conv, B, H, W):
    return fn.Conv3x3Fn.apply(t.view(B, H * W, -1), conv.weight, conv.bias, H, W).view(B * H * W, -1)


class DownSample(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(channels, channels // 2, kernel_size=3, padding=1, bias=False),
                                  nn.PixelUnshuffle(2))

    def forward(self, t, B, H, W):
        return fn.PixelUnshuffleFn.apply(_conv3x3(t, self.body[0], B, H, W), B, H // 2, W // 2)


class UpSample(nn.Module):
    """FFT amplitude / phase up-sampler (lines 161-188): FFT passes (csrc/fft.cu), the abs/angle, mag*exp(i pha),
    |ifft2| and LeakyReLU chains as one fused elementwise kernel each (csrc/spectral_ew.cu), 1x1 convs as GEMMs."""

    def __init__(self, channels, channel_red):
        super().__init__()
        self.amp_fuse = nn.Sequential(nn.Conv2d(channels, channels, 1, 1, 0), nn.LeakyReLU(0.1, inplace=False),
                                      nn.Conv2d(channels, channels, 1, 1, 0))
        self.pha_fuse = nn.Sequential(nn.Conv2d(channels, channels, 1, 1, 0), nn.LeakyReLU(0.1, inplace=False),
                                      nn.Conv2d(channels, channels, 1, 1, 0))
        self.post = nn.Conv2d(channels, channels // 2 if channel_red else channels, 1, 1, 0)

    @staticmethod
    def _mlp(seq, x2d):   # Conv1x1 -> LeakyReLU(0.1) -> Conv1x1 on tokens (tensor-core GEMMs)
        h = fn.LeakyReluFn.apply(fn.linear(x2d, _w2(seq[0]), seq[0].bias), 0.1, True)
        return fn.linear(h, _w2(seq[2]), seq[2].bias, rounded=True)

    def forward(self, t, B, H, W):   # tokens (B*H*W, C) -> tokens (B*2H*2W, C_out)
        C = t.shape[1]
        # fft2 on the shared-memory FFT passes (csrc/fft.cu), NHWC, no NCHW round trip
        f = fn.Fft2Fn.apply(t.view(B, H, W, C), B, H, W, C, False, False, 1.0)                 # (B, H, W, C, 2)
        mag0, pha0 = fn.PolarSplitFn.apply(f)
        mag = self._mlp(self.amp_fuse, mag0.view(B * H * W, C))
        pha = self._mlp(self.pha_fuse, pha0.view(B * H * W, C))
        z = fn.PolarJoinFn.apply(mag, pha).view(B, H, W, C, 2)
        # ifft2 of the (2, 2)-tiled spectrum at 2H x 2W == ifft2 of the spectrum at H x W written to the even
        # pixels, exact zeros elsewhere: a quarter of the transform work and no tile()
        small = fn.CAbsFn.apply(fn.Fft2Fn.apply(z, B, H, W, C, True, True, 1.0 / (H * W)))
        y = fn.linear(small.view(B * H * W, C), _w2(self.post), self.post.bias)          # post(|z|) at even pixels
        return fn.EvenScatterFn.apply(y, self.post.bias, B, H, W)                         # post(0) = bias elsewhere


class UpSample1(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(channels, channels * 2, kernel_size=3, padding=1, bias=False),
                                  nn.PixelShuffle(2))

    def forward(self, t, B, H, W):
        return fn.PixelShuffleFn.apply(_conv3x3(t, self.body[0], B, H, W), B, H, W)     # tokens at (2H, 2W)


class UpS(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.Fups = UpSample(channels, True)
        self.Sups = UpSample1(channels)
        self.reduce = nn.Conv2d(channels, channels // 2, kernel_size=1, bias=False)

    def forward(self, t, B, H, W):   # tokens at (H, W) -> tokens at (2H, 2W)
        cat = torch.cat([self.Fups(t, B, H, W), self.Sups(t, B, H, W)], dim=1)
        return fn.linear(cat, _w2(self.reduce))


class SpectralTransformer(nn.Module):
    def __init__(self, num_blocks=[2, 3, 3, 4], num_heads=[1, 2, 4, 8], channels=[16, 32, 64, 128], num_refinement=4,
                 expansion_factor=2.66, ch=[64, 32, 16, 64]):
        super().__init__()
        self.embed_conv_rgb = nn.Conv2d(3, channels[0], kernel_size=3, padding=1, bias=False)
        self.encoders = nn.ModuleList([nn.Sequential(*[TransformerBlock(c, a, expansion_factor) for _ in range(n)])
                                       for n, a, c in zip(num_blocks, num_heads, channels)])
        self.down1, self.down2, self.down3 = DownSample(channels[0]), DownSample(channels[1]), DownSample(channels[2])
        self.ups_1, self.ups_2, self.ups_3, self.ups_4 = UpS(128), UpS(64), UpS(32), UpS(3)
        self.ups1 = UpSample1(32)
        self.reduces2 = nn.Conv2d(64, 32, kernel_size=1, bias=False)
        self.reduces1 = nn.Conv2d(128, 64, kernel_size=1, bias=False)
        self.decoders = nn.ModuleList([nn.Sequential(*[TransformerBlock(channels[2], num_heads[2], expansion_factor)
                                                       for _ in range(num_blocks[2])])])
        self.decoders.append(nn.Sequential(*[TransformerBlock(channels[1], num_heads[1], expansion_factor)
                                             for _ in range(num_blocks[1])]))
        self.decoders.append(nn.Sequential(*[TransformerBlock(channels[1], num_heads[0], expansion_factor)
                                             for _ in range(num_blocks[0])]))
        self.refinement = nn.Sequential(*[TransformerBlock(channels[1], num_heads[0], expansion_factor)
                                          for _ in range(num_refinement)])
        self.output = nn.Conv2d(8, 3, kernel_size=3, padding=1, bias=False)
        self.output1 = nn.Conv2d(16, 8, kernel_size=3, padding=1, bias=False)
        self.ups2 = UpSample1(16)
        self.outputl = nn.Conv2d(32, 8, kernel_size=3, padding=1, bias=False)

    @staticmethod
    def _stage(blocks, t, B, H, W):
        for blk in blocks:
            t = blk(t, B, H, W)
        return t

    def forward(self, RGB_input):
        if not RGB_input.is_cuda:
            raise RuntimeError("uwr SpectralTransformer runs on CUDA (B200) only; there is no CPU fallback")
        x = RGB_input.contiguous().float()
        B, _, H, W = x.shape
        if H % 8 or W % 8:
            raise ValueError("SpectralTransformer needs H and W divisible by 8")
        f0 = fn.ConvImg2TokFn.apply(x, self.embed_conv_rgb.weight, None)     # 3 -> 16: thin direct conv (csrc/conv_small.cu)
        e1 = self._stage(self.encoders[0], f0, B, H, W)
        e2 = self._stage(self.encoders[1], self.down1(e1, B, H, W), B, H // 2, W // 2)
        e3 = self._stage(self.encoders[2], self.down2(e2, B, H // 2, W // 2), B, H // 4, W // 4)
        e4 = self._stage(self.encoders[3], self.down3(e3, B, H // 4, W // 4), B, H // 8, W // 8)
        d3 = fn.linear(torch.cat([self.ups_1(e4, B, H // 8, W // 8), e3], 1), _w2(self.reduces1))
        d3 = self._stage(self.decoders[0], d3, B, H // 4, W // 4)
        d2 = fn.linear(torch.cat([self.ups_2(d3, B, H // 4, W // 4), e2], 1), _w2(self.reduces2))
        d2 = self._stage(self.decoders[1], d2, B, H // 2, W // 2)
        fd = self._stage(self.decoders[2], torch.cat([self.ups_3(d2, B, H // 2, W // 2), e1], 1), B, H, W)
        fr = self._stage(self.refinement, fd, B, H, W)
        o = _conv3x3(fr, self.outputl, B, H, W)                               # 32 -> 8
        return fn.ConvTok2ImgFn.apply(o, self.output.weight, None, None, B, H, W)   # 8 -> 3: thin direct conv
