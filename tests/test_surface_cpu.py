"""CPU: the drop-in surface (registry, module tree, loss registry) and the C ABI (library loads,
exports every symbol that include/uwr_b200.h declares).  No kernel is launched here."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, PKG


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "uwr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uwr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(build_lib):
    lib = ctypes.CDLL(build_lib)
    names = _declared_symbols()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.uwr_abi_version.restype = ctypes.c_int
    assert lib.uwr_abi_version() == 1
    lib.uwr_last_error.restype = ctypes.c_char_p
    assert lib.uwr_last_error() is not None


def test_ctypes_table_matches_header(build_lib):
    from uwr import _lib
    declared = set(_declared_symbols())
    bound = set(_lib.SIGNATURES)
    assert bound <= declared, bound - declared
    # descriptor structs mirror the C layout: ask the C compiler (the header is plain C)
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "sz.c")
        open(src, "w").write('#include <stdio.h>\n#include "uwr_b200.h"\nint main(void){printf("%zu %zu %zu %zu", '
                             'sizeof(uwr_gemm_desc), sizeof(uwr_attn_desc), '
                             '__builtin_offsetof(uwr_gemm_desc, workspace_bytes), '
                             '__builtin_offsetof(uwr_attn_desc, scale));return 0;}\n')
        exe = os.path.join(td, "sz")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        g, a, go, ao = map(int, subprocess.check_output([exe]).split())
    assert ctypes.sizeof(_lib.GemmDesc) == g and _lib.GemmDesc.workspace_bytes.offset == go
    assert ctypes.sizeof(_lib.AttnDesc) == a and _lib.AttnDesc.scale.offset == ao


def test_error_paths_return_codes_not_crashes(build_lib):
    """argument validation happens before any launch, so it is testable without a GPU"""
    from uwr import _lib
    d = _lib.GemmDesc()
    rc = _lib.fn["uwr_gemm_tf32"](ctypes.byref(d), None)
    assert rc == -1 and b"null operand" in _lib.fn["uwr_last_error"]()
    assert _lib.fn["uwr_set_gemm_precision"](2) == -1
    assert _lib.fn["uwr_gemm_tcgen05_supported"](ctypes.byref(d)) == 0
    with pytest.raises(_lib.UwrError):
        _lib.check(_lib.fn["uwr_pixel_loss"](None, None, None, None, None, 0, 1, 3, 8, 8, 1, None), "uwr_pixel_loss")
    # entry points added later in the round: same contract (negative code + message, nothing launched)
    f = _lib.fn
    assert f["uwr_dft_hw_real"](None, None, None, 1, 16, 16, 32, 1.0, None) == -1
    assert b"null pointer" in f["uwr_last_error"]()
    assert f["uwr_dft_lc_real"](None, None, None, 1, 16, 16, 32, 1.0, None) == -1
    assert f["uwr_fft2_hw"](None, None, None, 1, 16, 16, 32, 0, 0, 1.0, None) == -1
    assert f["uwr_gelu_mul_fwd"](None, 8, None, 4, 4, None) == -1
    assert f["uwr_gelu_mul_bwd"](None, None, 8, None, 4, 4, None) == -1
    assert f["uwr_layernorm_bwd_ds_supported"](1024, 64) == 1 and f["uwr_layernorm_bwd_ds_supported"](1024, 48) == 0
    assert f["uwr_layernorm_bwd_ds_workspace_bytes"](1 << 20, 64) > 0
    assert f["uwr_layernorm_bwd_ds"](*([None] * 10), 1, None, None, None, 1024, 64, None) == -1
    assert f["uwr_dft_workspace_bytes"](2, 16, 16, 32) == 2 * 2 * 16 * 16 * 32 * 8
    assert f["uwr_set_attn_tcgen05"](0) == 0 and f["uwr_set_attn_tcgen05"](2) == 0     # back to the default (auto)


def test_registry_surface():
    import uwr
    assert uwr.get_names() == ["SpectralTransformer", "NewModel", "NewBigModel", "NewBigFRFNModel", "AST"]
    with pytest.raises(KeyError, match="Unknown model: nope"):
        uwr.init_model("nope")
    m = uwr.init_model("AST", use_dwt="Fourier")   # use_dwt is dropped (src/Models/__init__.py:25-29)
    sd = m.state_dict()
    assert len(sd) == 274
    assert sd["conv.blocks.0.attn.qkv.to_kv.weight"].shape == (1024, 512)
    assert sd["encoderlayer_0.blocks.0.mlp.dwconv.0.weight"].shape == (128, 1, 3, 3)
    assert sd["dowsample_0.conv.0.weight"].shape == (64, 32, 4, 4)
    assert sd["upsample_0.deconv.0.weight"].shape == (512, 256, 2, 2)
    assert sd["decoderlayer_3.blocks.1.attn.relative_position_index"].dtype == torch.int64
    assert "encoderlayer_0.blocks.0.norm1.weight" not in sd     # encoder stages have no attention
    assert sd["conv.blocks.0.attn.w"].tolist() == [1.0, 1.0]


def test_no_cpu_fallback():
    import uwr
    m = uwr.AST(img_size=128)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 128, 128))
    with pytest.raises(ValueError, match="Unsupported loss"):
        uwr.LossFunction("nope", "cpu").getloss(torch.zeros(1, 3, 8, 8), torch.zeros(1, 3, 8, 8))
    with pytest.raises(TypeError):
        from uwr import ops
        ops.layernorm_fwd(torch.zeros(4, 32), torch.ones(32), torch.zeros(32))


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under uwr/ may reference it"""
    for dirpath, _, files in os.walk(os.path.join(PKG, "uwr")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle's", ""), os.path.join(dirpath, f)


def test_droppath_semantics():
    from uwr.ast import DropPath
    dp = DropPath(0.25)
    dp.eval()
    assert dp.scale(8, "cpu") is None
    dp.train()
    torch.manual_seed(0)
    s = dp.scale(4096, "cpu")
    assert all(v == 0.0 or abs(v - 1.0 / 0.75) < 1e-6 for v in s.unique().tolist())
    assert abs((s > 0).float().mean().item() - 0.75) < 0.03
    assert DropPath(0.0).train().scale(4, "cpu") is None


def test_torch_library_ops_are_registered():
    """the kernels are exposed as torch.library ops `uwr::<name>` (CUDA-only: a CPU tensor is refused
    by the dispatcher, there is no fallback kernel)."""
    import torch
    import uwr  # noqa: F401
    from uwr import torchlib
    for name in torchlib.OPS:
        op = getattr(torch.ops.uwr, name)
        assert str(op.default._schema).startswith(f"uwr::{name}(")
    with pytest.raises(NotImplementedError):
        torch.ops.uwr.layernorm(torch.zeros(4, 4), torch.ones(4), torch.zeros(4), 1e-5)


def test_fflmix_refuses_to_run_without_vgg_weights(monkeypatch, tmp_path):
    """no silent random-init perceptual network — absent weights raise unless "random" is requested."""
    from uwr import fflmix
    monkeypatch.delenv("UWR_VGG16_WEIGHTS", raising=False)
    monkeypatch.setattr(torch.hub, "get_dir", lambda: str(tmp_path))
    with pytest.raises(RuntimeError, match="pretrained VGG16"):
        fflmix.resolve_vgg16_weights(None)
    assert fflmix.resolve_vgg16_weights("random") == ("random", None)
    monkeypatch.setenv("UWR_VGG16_WEIGHTS", "random")
    assert fflmix.resolve_vgg16_weights(None) == ("random", None)
    # a real checkpoint in the hub cache (or UWR_VGG16_WEIGHTS=path) is picked up and actually loaded
    import torchvision
    net = torchvision.models.vgg16(weights=None)
    with torch.no_grad():
        net.features[0].weight.fill_(0.125)
    ck = tmp_path / "checkpoints"
    ck.mkdir()
    torch.save(net.state_dict(), ck / fflmix.VGG_FILE)
    monkeypatch.delenv("UWR_VGG16_WEIGHTS")
    kind, path = fflmix.resolve_vgg16_weights(None)
    assert kind == "file" and path.endswith(fflmix.VGG_FILE)
    vgg = fflmix.VGGPerceptual(path)
    assert vgg.weights_source == path and float(vgg.blocks[0][0].weight.flatten()[0]) == 0.125


def test_direct_gradient_writes_are_scoped_to_trainstep():
    """overwrite-semantics gradient writes are only enabled inside TrainStep's own backward."""
    from uwr import ops
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.zeros(4)
    p._uwr_direct = True
    assert ops.grad_slot(p) is None or not p.grad.is_cuda      # outside the context: never a direct slot
    with ops.direct_grad_writes():
        pass
    assert ops._DIRECT_WRITES is False


def test_split_plan_fits_one_wave_and_conv_geometry_gate(build_lib):
    """Host logic of the tcgen05 GEMM that needs no GPU: (1) the split-contraction plan never creates more work items
    than SMs when fewer tiles than SMs exist (a rounded-up split count once ran 3 tiles x 50 splits = 150 items on 148
    SMs, two waves, on several AST weight-gradient shapes); (2) which convolution geometries the implicit-GEMM entry
    point serves; (3) the new descriptor structs mirror the C layout."""
    import subprocess, tempfile
    from uwr import _lib
    f = _lib.fn
    sms = f["uwr_device_sm_count"]()
    pick_bn = lambda n: 32 if n <= 32 else 64 if n <= 64 else 128 if n <= 128 else 256
    for M, N, K in [(64, 256, 1 << 20), (384, 128, 262144), (768, 256, 65536), (1536, 512, 16384), (256, 1024, 65536),
                    (2048, 512, 4096), (512, 512, 16384), (288, 32, 1 << 20), (32, 128, 1 << 20), (128, 128, 262144)]:
        tiles = -(-M // 128) * -(-N // pick_bn(N))
        nbytes = f["uwr_gemm_tcgen05_workspace_bytes"](M, N, K, 1)
        splits = max(1, nbytes // (M * N * 4))
        assert tiles * splits <= sms, (M, N, K, tiles, splits)
        assert tiles * (splits + 1) > 0.7 * sms or splits * 8 * 32 >= K // 2, (M, N, K, tiles, splits)   # and it does fill the chip
    assert f["uwr_gemm_tcgen05_workspace_bytes"](64, 256, 4096, 0) == 0

    def desc(mode, B, H, W, Cin, Cout, k, stride, pad):
        d = _lib.ConvGemmDesc()
        d.mode, d.x, d.ld_x, d.B, d.H, d.W, d.Cin, d.Cout = mode, 0x1000, Cin, B, H, W, Cin, Cout
        d.kh = d.kw = k
        d.stride, d.pad = stride, pad
        d.w = d.y = d.dy = d.dw = 0x2000
        d.ld_y = d.ld_dy = Cout
        return d
    ok = lambda *a: f["uwr_convgemm_tcgen05_supported"](ctypes.byref(desc(*a)))
    assert ok(0, 16, 256, 256, 32, 64, 4, 2, 1) == 1      # AST Downsample, forward
    assert ok(1, 16, 256, 256, 32, 64, 4, 2, 1) == 1      # ... weight gradient
    assert ok(2, 16, 256, 256, 32, 32, 3, 1, 1) == 1      # transposed weight gradient
    assert ok(0, 2, 16, 16, 32, 64, 3, 1, 1) == 1
    assert ok(0, 2, 16, 48, 32, 32, 3, 1, 1) == 0         # width neither a power of two nor a multiple of the box
    assert ok(0, 2, 16, 16, 24, 64, 3, 1, 1) == 0         # Cin % 32
    assert ok(0, 1, 8, 8, 32, 64, 3, 1, 1) == 0           # 64 output pixels < one 128-row tile
    assert ok(1, 1, 8, 8, 32, 64, 3, 1, 1) == 1           # ... but a multiple of the 32-pixel contraction chunk
    assert ok(3, 2, 16, 16, 32, 64, 3, 1, 1) == 0         # unknown mode
    d = desc(1, 16, 256, 256, 32, 64, 4, 2, 1)
    assert f["uwr_convgemm_tcgen05_workspace_bytes"](ctypes.byref(d)) % (64 * 512 * 4) == 0
    assert f["uwr_convgemm_tcgen05"](ctypes.byref(desc(0, 2, 16, 48, 32, 32, 3, 1, 1)), None) == -1
    assert b"unsupported geometry" in f["uwr_last_error"]()
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "sz.c")
        open(src, "w").write('#include <stdio.h>\n#include "uwr_b200.h"\nint main(void){printf("%zu %zu %zu %zu", '
                             'sizeof(uwr_convgemm_desc), __builtin_offsetof(uwr_convgemm_desc, workspace_bytes), '
                             'sizeof(uwr_attn_desc), __builtin_offsetof(uwr_attn_desc, dkv_colsum));return 0;}\n')
        exe = os.path.join(td, "sz")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        cs, co, asz, ao = map(int, subprocess.check_output([exe]).split())
    assert ctypes.sizeof(_lib.ConvGemmDesc) == cs and _lib.ConvGemmDesc.workspace_bytes.offset == co
    assert ctypes.sizeof(_lib.AttnDesc) == asz and _lib.AttnDesc.dkv_colsum.offset == ao


def test_persistent_grids_fit_one_resident_wave(build_lib):
    """The depthwise and attention-backward kernels are persistent: a fixed number of CTAs per channel group / head, each
    striding over that group's tiles.  The per-group count must be rounded DOWN -- rounded up, 16 channel groups got
    19 CTAs each (304 on 296 slots) and 8 heads 56 each (448 on 444), and the left-over CTAs ran their whole share after
    the first wave: up to twice the kernel time at the three deepest AST levels.  Host logic only, read back through the
    workspace-size entry points."""
    from uwr import _lib
    f = _lib.fn
    sms = f["uwr_device_sm_count"]()
    for B, H, Ch in [(16, 256, 128), (16, 256, 256), (16, 128, 512), (16, 64, 1024), (16, 32, 2048), (16, 16, 2048),
                     (8, 256, 96), (8, 256, 176), (8, 128, 352), (1, 256, 256), (1, 16, 2048), (2, 32, 10240)]:
        groups = -(-Ch // 32)
        tiles = B * (H // 16) ** 2
        ctas = f["uwr_dwconv_gelu_bwd_workspace_bytes"](B, H, H, Ch) // (11 * Ch * 4)
        assert 1 <= ctas <= tiles
        if groups <= 2 * sms:
            assert ctas * groups <= 2 * sms, (B, H, Ch, ctas, groups)          # two CTAs per SM: one resident wave
            if tiles * groups >= 2 * sms:
                assert (ctas + 1) * groups > 2 * sms, (B, H, Ch, ctas, groups)  # ... and the wave is as full as it can be
        else:
            assert ctas == 1
    for heads, hd, H in [(1, 32, 256), (2, 32, 256), (4, 32, 128), (8, 32, 64), (16, 32, 32), (32, 32, 16), (2, 16, 256),
                         (4, 8, 256), (3, 64, 64), (5, 128, 32)]:
        d = _lib.AttnDesc()
        d.B, d.H, d.W, d.heads, d.head_dim = 16, H, H, heads, hd
        per_head = f["uwr_window_attn_bwd_workspace_bytes"](ctypes.byref(d)) // (heads * (225 + 3 + 3 * hd) * 4)
        tiles = 16 * (H // 8) ** 2
        assert 1 <= per_head <= tiles
        # CTAs per head = floor(resident slots / heads): without a device the resident count per SM falls back to 1, on
        # a B200 it is what the runtime reports (3 for head_dim <= 32).  So heads x CTAs-per-head fills k x SMs slots
        # for a whole k, from below: never one CTA more than the slots, never a whole head's worth less.
        total = per_head * heads
        k = -(-total // sms)
        assert total <= k * sms and (total > k * sms - heads or per_head == tiles), (heads, hd, per_head, k)
        if not torch.cuda.is_available():
            assert k == 1


def test_droppath_pool_hands_out_independent_scaled_bernoulli_rows():
    """uwr.ast.DropPathPool (host logic, device-agnostic): after the request sequence of one training forward has been
    learnt, all scale vectors of a forward come from ONE (requests, batch) draw -- every row is 0 or 1 / keep_i for ITS
    module's keep probability, rows follow the request order, and anything out of sequence falls back to per-call draws."""
    from uwr.ast import AST, DropPath
    torch.manual_seed(0)
    m = AST(img_size=128).train()
    pool = m._dp_pool
    dps = [x for x in m.modules() if isinstance(x, DropPath)]
    assert dps and all(d.pool is pool for d in dps)
    assert "_dp_pool" not in m.state_dict() and len(m.state_dict()) == 274
    dev = torch.device("cpu")
    seq = [d for d in dps for _ in range(2)]                      # attention branch, then FFN branch, per block
    pool.begin(4, dev)
    first = [d.scale(4, dev) for d in seq]                        # learning pass: per-call draws
    pool.end(dev)
    assert pool.order == [id(d) for d in seq] and pool.keep.shape == (len(seq), 1)
    pool.begin(4, dev)
    rows = [d.scale(4, dev) for d in seq]
    assert pool.i == len(seq) and all(r.data_ptr() == pool.rows[i].data_ptr() for i, r in enumerate(rows))
    pool.end(dev)
    for d, r, f in zip(seq, rows, first):
        keep = 1.0 - d.drop_prob
        for v in (r, f):
            assert v.shape == (4,) and bool(((v == 0) | ((v - 1.0 / keep).abs() < 1e-6)).all())
    # statistics of a large draw: mean of every row ~ 1 (E[bernoulli(keep) / keep] = 1)
    pool.begin(20000, dev)
    big = pool.rows.clone()
    pool.end(dev)
    assert big.shape == (len(seq), 20000) and float((big.mean(dim=1) - 1.0).abs().max()) < 0.02
    # out of sequence (another module asks first): the pool steps aside for the rest of that forward
    pool.begin(4, dev)
    v = seq[-1].scale(4, dev)
    assert pool.rows is None and v.shape == (4,)
    pool.end(dev)
    # injected masks bypass the pool altogether
    seq[0].forced = torch.tensor([1.0, 0.0, 1.0, 0.0])
    pool.begin(4, dev)
    assert torch.equal(seq[0].scale(4, dev), seq[0].forced)
    pool.end(dev)
    m.eval()
    assert dps[-1].scale(4, dev) is None
