"""B200-native AST (Adaptive Sparse Transformer U-Net) behind the reference's nn.Module surface.

Drop-in for src/Models/AST.py::AST (AST.py:680-921): same constructor defaults, same module tree
(hence the same state_dict keys/shapes and — because parameters are created in the same order
with the same torch initialisers — bit-identical weights for a given torch.manual_seed), same
forward signature.  The forward/backward math runs entirely in the uwr CUDA kernels
(uwr.blocks); the nn.Linear / nn.Conv2d / nn.LayerNorm objects below are parameter containers
only and their own forward() is never called.
"""
import math

import torch
import torch.nn as nn

from . import blocks


def _trunc_normal_(t, std=0.02):
    # timm.layers.trunc_normal_ == nn.init.trunc_normal_ with absolute cut-offs [-2, 2]
    return nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)


class DropPathPool:
    """All DropPath scale vectors of one forward pass drawn by TWO kernels instead of two per residual branch.

    A training forward of AST asks for 36 per-sample scale vectors (18 blocks x attention / FFN branch), each a
    `bernoulli_` + `div_` on `batch` floats: 72 launches of ~2 us in the step's CUDA graph.  The pool learns the
    sequence of requests (module, keep probability) during the first training forward, and from then on draws the whole
    (requests, batch) matrix at the top of the forward -- `torch.bernoulli` of the per-row keep probabilities, divided by
    them -- and hands out its rows.  Same distribution per row (independent Bernoulli(keep_i) / keep_i, timm semantics);
    a request that does not match the learnt sequence (injected masks, a patched `scale`, another batch size) falls back
    to the per-call draw."""

    def __init__(self):
        self.order = None      # ids of the DropPath modules in request order
        self.keep = None       # (requests, 1) keep probabilities on the device
        self.rows = None
        self.i = 0
        self.recording = None

    def begin(self, batch, device):
        self.rows, self.i = None, 0
        if self.order is None:
            self.recording = []
        elif self.order:
            if self.keep.device != device:
                self.keep = self.keep.to(device)
            self.rows = torch.bernoulli(self.keep.expand(-1, batch)).div_(self.keep)

    def take(self, mod, batch, device):
        if self.rows is not None:
            if self.i < len(self.order) and self.order[self.i] == id(mod) and self.rows.shape[1] == batch:
                self.i += 1
                return self.rows[self.i - 1]
            self.rows = None   # out of sequence: the rest of this forward draws per call
        if self.recording is not None:
            self.recording.append((id(mod), 1.0 - mod.drop_prob))
        return None

    def end(self, device):
        if self.recording is not None:
            self.order = [m for m, _ in self.recording]
            if self.order:
                self.keep = torch.tensor([[k] for _, k in self.recording], dtype=torch.float32, device=device)
            self.recording = None
        self.rows = None


class DropPath(nn.Module):
    """timm DropPath semantics expressed as a per-sample scale vector that the
    residual GEMM epilogue consumes; `forced` lets parity tests inject the reference's masks."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob
        self.forced = None
        self.pool = None       # DropPathPool of the owning model (AST), if any

    def scale(self, batch, device):
        if self.forced is not None:
            return self.forced.to(device=device, dtype=torch.float32).reshape(batch).contiguous()
        if self.drop_prob == 0.0 or not self.training:
            return None
        if self.pool is not None:
            row = self.pool.take(self, batch, device)
            if row is not None:
                return row
        keep = 1.0 - self.drop_prob
        m = torch.empty(batch, device=device, dtype=torch.float32).bernoulli_(keep)
        return m.div_(keep)


class LinearProjection(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0, bias=True):
        super().__init__()
        inner = dim_head * heads
        self.heads = heads
        self.to_q = nn.Linear(dim, inner, bias=bias)
        self.to_kv = nn.Linear(dim, inner * 2, bias=bias)
        self.dim = dim
        self.inner_dim = inner


def _relative_position_index(win):
    # closed form of AST.py:160-169: (yi - yj + 7) * 15 + (xi - xj + 7)
    ys, xs = torch.meshgrid(torch.arange(win), torch.arange(win), indexing="ij")
    ys, xs = ys.flatten(), xs.flatten()
    return (ys[:, None] - ys[None, :] + win - 1) * (2 * win - 1) + (xs[:, None] - xs[None, :] + win - 1)


class WindowAttention_sparse(nn.Module):
    sparse = True

    def __init__(self, dim, win_size, num_heads, token_projection="linear", qkv_bias=True, qk_scale=None,
                 attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.dim = dim
        self.win_size = win_size
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.relative_position_bias_table = nn.Parameter(
            torch.zeros((2 * win_size[0] - 1) * (2 * win_size[1] - 1), num_heads))
        self.register_buffer("relative_position_index", _relative_position_index(win_size[0]))
        _trunc_normal_(self.relative_position_bias_table, std=0.02)
        if token_projection != "linear":
            raise Exception("Projection error!")
        self.qkv = LinearProjection(dim, num_heads, dim // num_heads, bias=qkv_bias)
        self.token_projection = token_projection
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.softmax = nn.Softmax(dim=-1)
        if self.sparse:
            self.relu = nn.ReLU()
            self.w = nn.Parameter(torch.ones(2))

    def extra_repr(self):
        return f"dim={self.dim}, win_size={self.win_size}, num_heads={self.num_heads}"


class WindowAttention(WindowAttention_sparse):
    """Plain softmax window attention (AST.py:68-140) = the w1 = 0 special case of the same kernel."""
    sparse = False


class _GELU(nn.GELU):
    pass


class LeFF(nn.Module):
    def __init__(self, dim=32, hidden_dim=128, act_layer=nn.GELU, drop=0.0, use_eca=False):
        super().__init__()
        self.linear1 = nn.Sequential(nn.Linear(dim, hidden_dim), act_layer())
        self.dwconv = nn.Sequential(
            nn.Conv2d(hidden_dim, hidden_dim, groups=hidden_dim, kernel_size=3, stride=1, padding=1), act_layer())
        self.linear2 = nn.Sequential(nn.Linear(hidden_dim, dim))
        self.dim = dim
        self.hidden_dim = hidden_dim
        self.eca = nn.Identity()


class Mlp(nn.Module):
    """plain MLP token mixer (AST.py:272-291; `token_mlp in ['ffn','mlp']`, not the registry default):
    LayerNorm, both Linears and the GELU on the uwr kernels; the residual add is a PyTorch elementwise op."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        self.in_features, self.hidden_features, self.out_features = in_features, hidden_features, out_features

    def block_forward(self, x, norm, dp_scale, H, W):
        from . import fn
        y = fn.linear(fn.layernorm(x, norm), self.fc1.weight, self.fc1.bias, rounded=True)
        y = fn.linear(fn.GeluFn.apply(y, True), self.fc2.weight, self.fc2.bias, rounded=True)
        if dp_scale is not None:
            y = y * dp_scale.view(-1, 1, 1)
        return x + y


class TransformerBlock(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, win_size=8, shift_size=0, mlp_ratio=4.0, qkv_bias=True,
                 qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 token_projection="linear", token_mlp="leff", att=True, sparseAtt=False):
        super().__init__()
        self.att = att
        self.sparseAtt = sparseAtt
        self.dim = dim
        self.input_resolution = input_resolution
        self.num_heads = num_heads
        self.win_size = win_size
        self.shift_size = shift_size
        self.mlp_ratio = mlp_ratio
        self.token_mlp = token_mlp
        if min(self.input_resolution) <= self.win_size:
            self.shift_size = 0
            self.win_size = min(self.input_resolution)
        assert 0 <= self.shift_size < self.win_size, "shift_size must in 0-win_size"
        if self.att:
            if self.win_size != 8:
                raise NotImplementedError("uwr window attention kernels are specialised to 8x8 windows")
            self.norm1 = norm_layer(dim)
            cls = WindowAttention_sparse if sparseAtt else WindowAttention
            self.attn = cls(dim, win_size=(self.win_size, self.win_size), num_heads=num_heads, qkv_bias=qkv_bias,
                            qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop,
                            token_projection=token_projection)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        hidden = int(dim * mlp_ratio)
        if token_mlp == "leff":
            self.mlp = LeFF(dim, hidden, act_layer=act_layer, drop=drop)
        elif token_mlp == "frfn":
            from .frfn import FRFN
            self.mlp = FRFN(dim, hidden, act_layer=act_layer, drop=drop)
        elif token_mlp in ("ffn", "mlp"):
            self.mlp = Mlp(in_features=dim, hidden_features=hidden, act_layer=act_layer, drop=drop)
        else:
            raise Exception("FFN error!")

    def extra_repr(self):
        return (f"dim={self.dim}, input_resolution={self.input_resolution}, num_heads={self.num_heads}, "
                f"win_size={self.win_size}, shift_size={self.shift_size}, mlp_ratio={self.mlp_ratio}")

    def _dp(self, batch, device):
        return self.drop_path.scale(batch, device) if isinstance(self.drop_path, DropPath) else None

    def forward(self, x, mask=None):
        return self.forward_linked(x, mask, None)[0]

    def forward_linked(self, x, mask=None, link_prev=None):
        """forward + the backward link to the NEXT block function (blocks.py: the LayerNorm backward of a function
        emits the rounded, DropPath-scaled gradient copy its predecessor's backward starts from)."""
        if mask is not None:
            raise NotImplementedError("input masks are never passed by the trainer (AST.py:558-565)")
        B, L, C = x.shape
        H = W = int(math.sqrt(L))
        m = self.mlp
        leff = isinstance(m, LeFF)
        grad = torch.is_grad_enabled()
        link_mid = {} if (self.att and leff and grad) else None
        link_next = {} if (leff and grad) else None
        if self.att:
            a = self.attn
            x = blocks.AttnBlockFn.apply(
                x, self.norm1.weight, self.norm1.bias, a.qkv.to_q.weight, a.qkv.to_q.bias, a.qkv.to_kv.weight,
                a.qkv.to_kv.bias, a.relative_position_bias_table, a.w if a.sparse else None, a.proj.weight,
                a.proj.bias, self._dp(B, x.device), H, W, self.num_heads, self.shift_size, float(a.scale),
                link_prev, link_mid)
        if leff:
            x = blocks.LeFFBlockFn.apply(
                x, self.norm2.weight, self.norm2.bias, m.linear1[0].weight, m.linear1[0].bias,
                m.dwconv[0].weight, m.dwconv[0].bias, m.linear2[0].weight, m.linear2[0].bias,
                self._dp(B, x.device), H, W, link_mid if self.att else link_prev, link_next)
        else:
            x = m.block_forward(x, self.norm2, self._dp(B, x.device), H, W)
        return x, link_next


class BasicASTLayer(nn.Module):
    def __init__(self, dim, output_dim, input_resolution, depth, num_heads, win_size, mlp_ratio=4.0, qkv_bias=True,
                 qk_scale=None, drop=0.0, attn_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm,
                 use_checkpoint=False, token_projection="linear", token_mlp="ffn", shift_flag=True, att=False,
                 sparseAtt=False):
        super().__init__()
        self.att = att
        self.sparseAtt = sparseAtt
        self.dim = dim
        self.input_resolution = input_resolution
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        self.blocks = nn.ModuleList([
            TransformerBlock(dim=dim, input_resolution=input_resolution, num_heads=num_heads, win_size=win_size,
                             shift_size=(0 if (i % 2 == 0) else win_size // 2) if shift_flag else 0,
                             mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop,
                             attn_drop=attn_drop,
                             drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                             norm_layer=norm_layer, token_projection=token_projection, token_mlp=token_mlp,
                             att=att, sparseAtt=sparseAtt)
            for i in range(depth)])

    def extra_repr(self):
        return f"dim={self.dim}, input_resolution={self.input_resolution}, depth={self.depth}"

    def forward(self, x, mask=None):
        link = None
        for blk in self.blocks:
            x, link = blk.forward_linked(x, mask, link)
        return x


class Downsample(nn.Module):
    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_channel, out_channel, kernel_size=4, stride=2, padding=1))
        self.in_channel = in_channel
        self.out_channel = out_channel

    def forward(self, x):
        H = W = int(math.sqrt(x.shape[1]))
        return blocks.DownsampleFn.apply(x, self.conv[0].weight, self.conv[0].bias, H, W)


class Upsample(nn.Module):
    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.deconv = nn.Sequential(nn.ConvTranspose2d(in_channel, out_channel, kernel_size=2, stride=2))
        self.in_channel = in_channel
        self.out_channel = out_channel

    def forward_cat(self, x, skip):
        """upsample(x) concatenated with the encoder skip along channels (AST.py:903-904)."""
        H = W = int(math.sqrt(x.shape[1]))
        return blocks.UpsampleCatFn.apply(x, self.deconv[0].weight, self.deconv[0].bias, skip, H, W)


class InputProj(nn.Module):
    def __init__(self, in_channel=3, out_channel=64, kernel_size=3, stride=1, norm_layer=None,
                 act_layer=nn.LeakyReLU):
        super().__init__()
        self.proj = nn.Sequential(
            nn.Conv2d(in_channel, out_channel, kernel_size=3, stride=stride, padding=kernel_size // 2),
            act_layer(inplace=True))
        self.norm = norm_layer(out_channel) if norm_layer is not None else None
        self.in_channel = in_channel
        self.out_channel = out_channel

    def forward(self, x):
        if self.norm is not None:
            raise NotImplementedError("InputProj norm_layer is unused by AST (AST.py:710-711)")
        slope = getattr(self.proj[1], "negative_slope", 0.01)
        return blocks.InputProjFn.apply(x, self.proj[0].weight, self.proj[0].bias, slope)


class OutputProj(nn.Module):
    def __init__(self, in_channel=64, out_channel=3, kernel_size=3, stride=1, norm_layer=None, act_layer=None):
        super().__init__()
        self.proj = nn.Sequential(
            nn.Conv2d(in_channel, out_channel, kernel_size=3, stride=stride, padding=kernel_size // 2))
        if act_layer is not None or norm_layer is not None:
            raise NotImplementedError("OutputProj act/norm are unused by AST (AST.py:712)")
        self.norm = None
        self.in_channel = in_channel
        self.out_channel = out_channel

    def forward(self, tokens, residual_img=None):
        H = W = int(math.sqrt(tokens.shape[1]))
        return blocks.OutputProjFn.apply(tokens, self.proj[0].weight, self.proj[0].bias, residual_img, H, W)


class AST(nn.Module):
    def __init__(self, img_size=256, in_chans=3, dd_in=3, embed_dim=32, depths=[2, 2, 2, 2, 2, 2, 2, 2, 2],
                 num_heads=[1, 2, 4, 8, 16, 16, 8, 4, 2], win_size=8, mlp_ratio=4.0, qkv_bias=True, qk_scale=None,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.1, norm_layer=nn.LayerNorm, patch_norm=True,
                 use_checkpoint=False, token_projection="linear", token_mlp="leff", dowsample=Downsample,
                 upsample=Upsample, shift_flag=True, **kwargs):
        super().__init__()
        if in_chans != 3:
            raise NotImplementedError("uwr output projection kernel is specialised to 3 output channels")
        self.num_enc_layers = len(depths) // 2
        self.num_dec_layers = len(depths) // 2
        self.embed_dim = embed_dim
        self.patch_norm = patch_norm
        self.mlp_ratio = mlp_ratio
        self.token_projection = token_projection
        self.mlp = token_mlp
        self.win_size = win_size
        self.reso = img_size
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.dd_in = dd_in

        enc_dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths[:self.num_enc_layers]))]
        conv_dpr = [drop_path_rate] * depths[4]
        dec_dpr = enc_dpr[::-1]

        self.input_proj = InputProj(in_channel=dd_in, out_channel=embed_dim, kernel_size=3, stride=1,
                                    act_layer=nn.LeakyReLU)
        self.output_proj = OutputProj(in_channel=2 * embed_dim, out_channel=in_chans, kernel_size=3, stride=1)

        def layer(mult, level, depth_idx, dpr, att):
            return BasicASTLayer(dim=embed_dim * mult, output_dim=embed_dim * mult,
                                 input_resolution=(img_size // (2 ** level), img_size // (2 ** level)),
                                 depth=depths[depth_idx], num_heads=num_heads[depth_idx], win_size=win_size,
                                 mlp_ratio=self.mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                                 attn_drop=attn_drop_rate, drop_path=dpr, norm_layer=norm_layer,
                                 use_checkpoint=use_checkpoint, token_projection=token_projection,
                                 token_mlp=token_mlp, shift_flag=shift_flag, att=att, sparseAtt=att)

        # registration order mirrors AST.py:715-861 (it fixes state_dict order and the init RNG stream)
        self.encoderlayer_0 = layer(1, 0, 0, enc_dpr[sum(depths[:0]):sum(depths[:1])], False)
        self.dowsample_0 = dowsample(embed_dim, embed_dim * 2)
        self.encoderlayer_1 = layer(2, 1, 1, enc_dpr[sum(depths[:1]):sum(depths[:2])], False)
        self.dowsample_1 = dowsample(embed_dim * 2, embed_dim * 4)
        self.encoderlayer_2 = layer(4, 2, 2, enc_dpr[sum(depths[:2]):sum(depths[:3])], False)
        self.dowsample_2 = dowsample(embed_dim * 4, embed_dim * 8)
        self.encoderlayer_3 = layer(8, 3, 3, enc_dpr[sum(depths[:3]):sum(depths[:4])], False)
        self.dowsample_3 = dowsample(embed_dim * 8, embed_dim * 16)
        self.conv = layer(16, 4, 4, conv_dpr, True)
        self.upsample_0 = upsample(embed_dim * 16, embed_dim * 8)
        self.decoderlayer_0 = layer(16, 3, 5, dec_dpr[:depths[5]], True)
        self.upsample_1 = upsample(embed_dim * 16, embed_dim * 4)
        self.decoderlayer_1 = layer(8, 2, 6, dec_dpr[sum(depths[5:6]):sum(depths[5:7])], True)
        self.upsample_2 = upsample(embed_dim * 8, embed_dim * 2)
        self.decoderlayer_2 = layer(4, 1, 7, dec_dpr[sum(depths[5:7]):sum(depths[5:8])], True)
        self.upsample_3 = upsample(embed_dim * 4, embed_dim)
        self.decoderlayer_3 = layer(2, 0, 8, dec_dpr[sum(depths[5:8]):sum(depths[5:9])], True)

        self.apply(self._init_weights)
        # one pool of DropPath draws per forward (a plain attribute: no parameters, nothing in the state_dict)
        object.__setattr__(self, "_dp_pool", DropPathPool())
        for m in self.modules():
            if isinstance(m, DropPath):
                m.pool = self._dp_pool

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"absolute_pos_embed"}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {"relative_position_bias_table"}

    def extra_repr(self):
        return (f"embed_dim={self.embed_dim}, token_projection={self.token_projection}, "
                f"token_mlp={self.mlp},win_size={self.win_size}")

    def adjacent_grad_pairs(self):
        """(to_q, to_kv) weight and bias pairs: their gradients are one packed-QKV GEMM / column sum, so
        uwr.train.GradBuckets lays their slots out back to back (ops.fused_grad_slot)."""
        pairs = []
        for m in self.modules():
            if isinstance(m, LinearProjection):
                pairs.append((m.to_q.weight, m.to_kv.weight))
                if m.to_q.bias is not None:
                    pairs.append((m.to_q.bias, m.to_kv.bias))
        return pairs

    def forward(self, x, mask=None):
        if not x.is_cuda:
            raise RuntimeError("uwr AST runs on CUDA (B200) only; there is no CPU fallback")
        x = x.contiguous().float()
        pool = self._dp_pool if self.training else None
        if pool is not None:
            pool.begin(x.shape[0], x.device)
        try:
            return self._forward(x, mask)
        finally:
            if pool is not None:
                pool.end(x.device)

    def _forward(self, x, mask):
        y = self.input_proj(x)
        conv0 = self.encoderlayer_0(y, mask=mask)
        pool0 = self.dowsample_0(conv0)
        conv1 = self.encoderlayer_1(pool0, mask=mask)
        pool1 = self.dowsample_1(conv1)
        conv2 = self.encoderlayer_2(pool1, mask=mask)
        pool2 = self.dowsample_2(conv2)
        conv3 = self.encoderlayer_3(pool2, mask=mask)
        pool3 = self.dowsample_3(conv3)
        conv4 = self.conv(pool3, mask=mask)
        deconv0 = self.decoderlayer_0(self.upsample_0.forward_cat(conv4, conv3), mask=mask)
        deconv1 = self.decoderlayer_1(self.upsample_1.forward_cat(deconv0, conv2), mask=mask)
        deconv2 = self.decoderlayer_2(self.upsample_2.forward_cat(deconv1, conv1), mask=mask)
        deconv3 = self.decoderlayer_3(self.upsample_3.forward_cat(deconv2, conv0), mask=mask)
        return self.output_proj(deconv3, x if self.dd_in == 3 else None)
