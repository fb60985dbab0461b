"""Seeded AST state_dict built with plain torch initialisers (no reference import, no uwr import).

TEST / BENCH INFRASTRUCTURE (see oracle/__init__.py).  Used by bench.py's baseline legs so that the
reference arm never loads the product library.  Shapes and init distributions follow the reference
constructor (src/Models/AST.py:680-872: trunc_normal_(std=.02) Linear weights, zero biases, LayerNorm
(1, 0), default Conv2d / ConvTranspose2d init, rel-pos table trunc_normal_(std=.02), w = ones(2));
the draw ORDER is not the reference's, so the values are not bit-identical to a seeded reference
module (parity tests use the product's module tree, which is — tests/test_oracle_golden.py).
"""
import math

import torch

WIN = 8


def _linear(sd, pre, cin, cout, g):
    w = torch.empty(cout, cin)
    torch.nn.init.trunc_normal_(w, mean=0.0, std=0.02, a=-2.0, b=2.0, generator=g)
    sd[pre + "weight"] = w
    sd[pre + "bias"] = torch.zeros(cout)


def _conv(sd, pre, shape, fan_in, nbias, g):
    bound = 1.0 / math.sqrt(fan_in)          # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    sd[pre + "weight"] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    sd[pre + "bias"] = (torch.rand(nbias, generator=g) * 2 - 1) * bound


def _rel_index():
    ys, xs = torch.meshgrid(torch.arange(WIN), torch.arange(WIN), indexing="ij")
    ys, xs = ys.flatten(), xs.flatten()
    return (ys[:, None] - ys[None, :] + WIN - 1) * (2 * WIN - 1) + (xs[:, None] - xs[None, :] + WIN - 1)


def _block(sd, pre, dim, heads, att, g):
    if att:
        sd[pre + "norm1.weight"], sd[pre + "norm1.bias"] = torch.ones(dim), torch.zeros(dim)
        t = torch.empty(225, heads)
        torch.nn.init.trunc_normal_(t, mean=0.0, std=0.02, a=-2.0, b=2.0, generator=g)
        sd[pre + "attn.relative_position_bias_table"] = t
        sd[pre + "attn.relative_position_index"] = _rel_index()
        _linear(sd, pre + "attn.qkv.to_q.", dim, dim, g)
        _linear(sd, pre + "attn.qkv.to_kv.", dim, 2 * dim, g)
        _linear(sd, pre + "attn.proj.", dim, dim, g)
        sd[pre + "attn.w"] = torch.ones(2)
    sd[pre + "norm2.weight"], sd[pre + "norm2.bias"] = torch.ones(dim), torch.zeros(dim)
    _linear(sd, pre + "mlp.linear1.0.", dim, 4 * dim, g)
    _conv(sd, pre + "mlp.dwconv.0.", (4 * dim, 1, 3, 3), 9, 4 * dim, g)
    _linear(sd, pre + "mlp.linear2.0.", 4 * dim, dim, g)


def ast_state_dict(seed=1234, embed_dim=32, num_heads=(1, 2, 4, 8, 16, 16, 8, 4, 2), depths=(2,) * 9):
    """All 274 entries of the default AST (AST.py:681-687 defaults)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    E = embed_dim
    _conv(sd, "input_proj.proj.0.", (E, 3, 3, 3), 27, E, g)
    _conv(sd, "output_proj.proj.0.", (3, 2 * E, 3, 3), 2 * E * 9, 3, g)
    dims = [E, 2 * E, 4 * E, 8 * E, 16 * E, 16 * E, 8 * E, 4 * E, 2 * E]
    names = ["encoderlayer_0", "encoderlayer_1", "encoderlayer_2", "encoderlayer_3", "conv",
             "decoderlayer_0", "decoderlayer_1", "decoderlayer_2", "decoderlayer_3"]
    for i, (name, dim) in enumerate(zip(names, dims)):
        for b in range(depths[i]):
            _block(sd, f"{name}.blocks.{b}.", dim, num_heads[i], i >= 4, g)
    for k in range(4):
        c = E * 2 ** k
        _conv(sd, f"dowsample_{k}.conv.0.", (2 * c, c, 4, 4), c * 16, 2 * c, g)
    for k, (cin, cout) in enumerate([(16 * E, 8 * E), (16 * E, 4 * E), (8 * E, 2 * E), (4 * E, E)]):
        _conv(sd, f"upsample_{k}.deconv.0.", (cin, cout, 2, 2), cout * 4, cout, g)
    return sd
