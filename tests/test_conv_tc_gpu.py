"""GPU tests of the tcgen05/TMEM/TMA implicit-GEMM conv3x3 (PU_MATH_TF32).

The kernel multiplies TF32 operands exactly and accumulates in fp32, so against a float64 convolution of the
SAME TF32-rounded inputs and weights the only differences are fp32 accumulation order and the final RN
rounding of the stored output to TF32 (2^-11 relative): tolerance 6e-4 of max|y| on outputs, and the
un-rounded comparison (dgrad/wgrad of the autograd path) within 2e-3."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def tf32_round(x):
    """round-to-nearest (ties away) to a 10-bit mantissa, like cvt.rna.tf32.f32"""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def test_tc_path_is_available_on_b200():
    from pu_b200 import _lib
    assert _lib.tc_available(), "tcgen05/TMA conv path must be usable on the B200 box"


SHAPES = [
    # B, C0, C1, Cout, H, W, relu, res
    (2, 8, 0, 8, 128, 128, True, False),    # inc.2 / up4.2
    (2, 8, 8, 8, 128, 128, True, False),    # up4.0: fused concat
    (3, 16, 0, 16, 64, 64, True, True),     # residual epilogue
    (2, 32, 32, 16, 32, 32, True, False),   # two chunks (one per source)
    (4, 64, 0, 64, 16, 16, True, False),    # N = 64, K chunks of 32
    (2, 128, 0, 32, 8, 8, False, False),    # 4 K chunks, tiny spatial
    (1, 16, 0, 24, 101, 101, True, True),   # odd sizes (UNetpRes@101), Cout = 24
    (2, 8, 8, 8, 25, 25, False, False),
    (1, 64, 64, 128, 12, 12, True, False),  # Cout > 64: two co blocks
    (1, 8, 0, 8, 6, 300, True, False),      # W > 248: x tiling
]


@pytest.mark.parametrize("B,C0,C1,Cout,H,W,relu,res", SHAPES)
def test_conv3x3_tc_forward(B, C0, C1, Cout, H, W, relu, res):
    from pu_b200 import ops
    g = torch.Generator().manual_seed(C0 * 7 + C1 + Cout + H)
    Cin = C0 + C1
    x0 = tf32_round(torch.randn(B, C0, H + 1, W + 2, generator=g))
    x1 = tf32_round(torch.randn(B, C1, H, W, generator=g)) if C1 else None
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    b = torch.randn(Cout, generator=g)
    r = torch.randn(B, Cout, H, W, generator=g) if res else None
    cat = x0[:, :, 1:, 1:W + 1].double()
    if C1:
        cat = torch.cat([cat, x1.double()], 1)
    yr = F.conv2d(cat, tf32_round(w).double(), b.double(), padding=1)
    if res:
        yr = yr + r.double()
    if relu:
        yr = F.relu(yr)
    with torch.no_grad():
        yo = ops.conv3x3(nhwc(x0).to(DEV), nhwc(x1).to(DEV) if C1 else None, w.to(DEV), b.to(DEV), nhwc(r).to(DEV) if res else None,
                         relu, H, W, 1, 1, 0, 0, ops.MATH_TF32)
    torch.cuda.synchronize()
    e = rel_err(nchw(yo), yr)
    assert e[0] < 6e-4, "max-rel %g l2-rel %g" % e
    # the stored output is exactly TF32-representable
    assert torch.equal(tf32_round(yo.cpu()), yo.cpu())


def unpack_mask(m, C):
    """uint8 [B,H,W,C/8] -> bool [B,H,W,C] (bit j of byte g = channel 8g+j)."""
    bits = (m.unsqueeze(-1).to(torch.int32) >> torch.arange(8, device=m.device, dtype=torch.int32)) & 1
    return bits.reshape(m.shape[:-1] + (C,)).bool()


@pytest.mark.parametrize("math_name", ["tf32", "fp32"])
@pytest.mark.parametrize("B,C0,C1,Cout,H,W,crop", [(2, 8, 0, 8, 128, 128, 0), (2, 8, 8, 8, 64, 64, 0), (3, 16, 16, 8, 37, 29, 2),
                                                    (2, 32, 32, 16, 32, 32, 0), (2, 64, 0, 64, 8, 8, 0), (1, 64, 64, 128, 12, 12, 1),
                                                    (64, 8, 8, 8, 128, 128, 0)])
def test_packed_relu_masks(math_name, B, C0, C1, Cout, H, W, crop):
    """Premasked-gradient protocol with PACKED masks: (1) the forward epilogue writes bit = (y > 0) for
    its own output; (2) a dgrad given the packed masks of its two destinations equals the unmasked dgrad times the
    unpacked masks, bit for bit — including the channel-split destinations and a cropped first source."""
    from pu_b200 import ops
    math = ops.MATH_TF32 if math_name == "tf32" else ops.MATH_FP32
    g = torch.Generator().manual_seed(B + C0 + Cout + H)
    x0 = tf32_round(torch.randn(B, H + crop, W + crop, C0, generator=g)).to(DEV)
    x1 = tf32_round(torch.randn(B, H, W, C1, generator=g)).to(DEV) if C1 else None
    w = tf32_round(torch.randn(Cout, C0 + C1, 3, 3, generator=g) / (3 * (C0 + C1) ** 0.5)).to(DEV)
    b = torch.randn(Cout, generator=g).to(DEV)
    with torch.no_grad():
        y, ym = ops.conv3x3_m(x0, x1, w, b, None, True, H, W, crop, crop, 0, 0, math, None, None, True, True)
        y_plain = ops.conv3x3(x0, x1, w, b, None, True, H, W, crop, crop, 0, 0, math)
        assert torch.equal(y, y_plain)  # emitting the mask does not change the output
        assert torch.equal(unpack_mask(ym, Cout), y > 0)
        frac = float((y > 0).float().mean())
        assert 0.2 < frac < 0.8
        # dgrad with packed masks of the sources (random bits; the masks have the geometry of the FULL source tensors)
        dy = tf32_round(torch.randn(B, H, W, Cout, generator=g)).to(DEV)
        m0 = torch.randint(0, 256, (B, H + crop, W + crop, C0 // 8), generator=g, dtype=torch.uint8).to(DEV)
        m1 = torch.randint(0, 256, (B, H, W, C1 // 8), generator=g, dtype=torch.uint8).to(DEV) if C1 else None
        _, dx0, dx1, _, _ = ops.conv3x3_bwd(dy, y, x0, x1, w, False, True, H, W, crop, crop, 0, 0, math, True, False, m0, m1, True)
        _, ex0, ex1, _, _ = ops.conv3x3_bwd(dy, y, x0, x1, w, False, True, H, W, crop, crop, 0, 0, math, True, False, None, None, True)
        assert float(ex0.abs().max()) > 0
        assert torch.equal(dx0, ex0 * unpack_mask(m0, C0))
        if C1:
            assert torch.equal(dx1, ex1 * unpack_mask(m1, C1))
        if crop:  # the border of a cropped source receives no gradient
            assert float(dx0[:, :crop].abs().max()) == 0 and float(dx0[:, :, :crop].abs().max()) == 0


@pytest.mark.parametrize("C0,C1,Cout,H,W", [(8, 8, 8, 64, 64), (16, 0, 16, 32, 32), (32, 32, 16, 16, 16), (64, 0, 64, 8, 8),
                                             (32, 0, 32, 32, 32), (64, 64, 32, 16, 16), (128, 0, 64, 12, 12), (64, 0, 128, 6, 6)])
def test_conv3x3_tc_autograd_vs_fp32_path(C0, C1, Cout, H, W):
    """Forward + dgrad (tcgen05) + wgrad through autograd in TF32 mode vs the strict-fp32 CUDA-core path."""
    from pu_b200 import ops
    g = torch.Generator().manual_seed(C0 + Cout + H)
    B = 2
    x0 = tf32_round(torch.randn(B, H, W, C0, generator=g))
    x1 = tf32_round(torch.randn(B, H, W, C1, generator=g)) if C1 else None
    w = tf32_round(torch.randn(Cout, C0 + C1, 3, 3, generator=g) / (3 * (C0 + C1) ** 0.5))
    b = torch.randn(Cout, generator=g)
    R = tf32_round(torch.randn(B, H, W, Cout, generator=g))
    outs = {}
    for math in (ops.MATH_FP32, ops.MATH_TF32):
        x0d = x0.to(DEV).requires_grad_(True)
        x1d = x1.to(DEV).requires_grad_(True) if C1 else None
        wd, bd = w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
        y = ops.conv3x3(x0d, x1d, wd, bd, None, True, H, W, 0, 0, 0, 0, math)
        (y * R.to(DEV)).sum().backward()
        outs[math] = (y.detach(), x0d.grad, x1d.grad if C1 else None, wd.grad, bd.grad)
    torch.cuda.synchronize()
    names = ["y", "dx0", "dx1", "dw", "db"]
    for n, a, bb in zip(names, outs[ops.MATH_TF32], outs[ops.MATH_FP32]):
        if a is None:
            continue
        e = rel_err(a, bb)
        assert e[0] < 2e-3, "%s: max-rel %g l2-rel %g" % (n, e[0], e[1])


def test_tf32_mode_whole_model_close_to_fp32():
    """UNetp in TF32 mode (tcgen05 convs) vs strict fp32: outputs and trace within the north-star 1e-3 (L2).
    Parameter gradients are held to 2e-2: measured 2e-3..8e-3 on B200, dominated by ReLU masks that flip where a
    pre-activation is within TF32 rounding of zero (inherent to any TF32 forward, cuDNN's included) — the strict
    1e-3 gradient parity is the fp32 mode's contract (tests/test_models_gpu.py)."""
    import contextlib
    import io
    from pu_b200 import UNetp
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        net = UNetp(1, 1, torch.device(DEV), rule="oja", nbf=64, batched=True)
    g = torch.Generator().manual_seed(1)
    x = torch.rand(4, 1, 64, 64, generator=g).to(DEV)
    hebb = (0.05 * torch.randn(64, 64, generator=g)).to(DEV)
    target = (torch.rand(4, 64, 64, generator=g) > 0.6).float().to(DEV)
    res = {}
    for mode in ("fp32", "tf32"):
        net.conv_math = mode
        net.zero_grad(set_to_none=True)
        out, hn = net(x, hebb)
        torch.nn.BCELoss()(out.reshape(-1), target.reshape(-1)).backward()
        res[mode] = (out.detach().clone(), hn.detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None})
    assert rel_err(res["tf32"][0], res["fp32"][0])[1] < 1e-3
    assert rel_err(res["tf32"][1], res["fp32"][1])[1] < 1e-3
    worst = max(rel_err(res["tf32"][2][k], res["fp32"][2][k])[1] for k in res["fp32"][2])
    assert worst < 2e-2, "worst parameter-gradient L2 error %g" % worst


FLAT_SHAPES = [
    # B, C0, C1, Cout, H, W
    (2, 8, 0, 8, 128, 128),     # narrow layers (N padded to 16), resident weights
    (2, 8, 8, 8, 64, 64),
    (2, 8, 0, 16, 64, 64),
    (2, 32, 32, 16, 32, 32),
    (1, 16, 0, 24, 37, 29),
    (2, 64, 0, 64, 32, 32),
    (1, 64, 64, 128, 24, 24),   # fused concat, two co blocks
    (2, 128, 0, 64, 16, 16),
    (1, 32, 32, 64, 40, 37),    # odd width
    (3, 64, 0, 64, 101, 101),   # several tiles per image, ragged edges
    (1, 256, 0, 256, 12, 12),   # 16 K chunks, four co blocks
    # tap-row split (kernel template KS: tiles of <= 2 blocks of a >= 32-channel flat layer — the deep levels of the U-Net)
    (64, 64, 0, 64, 8, 8),      # down4 @B=64: one block per tile, one tile per CTA, two accumulator buffers
    (64, 64, 0, 64, 16, 16),    # down3.1: two blocks per tile, ONE accumulator buffer (384 TMEM columns)
    (64, 32, 0, 64, 16, 16),    # down3.0
    (400, 64, 0, 64, 8, 8),     # several tiles per CTA: the two buffers alternate
    (300, 32, 0, 32, 16, 16),   # 32 columns, two blocks, several tiles per CTA
    (3, 32, 32, 64, 9, 7),      # concat pair inside a split tile, ragged
]


@pytest.fixture
def flat_mode(monkeypatch):
    """Force the flat (one MMA per tap, N = 64, 128-pixel blocks) variant of the tcgen05 conv — by default it is only chosen for
    the wide, tensor-bound layers with at least two waves of 512-pixel tiles (SURVEY.md §8d "TC demo")."""
    from pu_b200 import ops
    monkeypatch.setenv("PU_TC_FLAT", "1")
    ops._tc_flat.cache_clear()
    yield
    monkeypatch.delenv("PU_TC_FLAT")
    ops._tc_flat.cache_clear()


@pytest.mark.parametrize("B,C0,C1,Cout,H,W", FLAT_SHAPES)
def test_conv3x3_tc_flat_mode(flat_mode, B, C0, C1, Cout, H, W):
    """Flat-mode forward vs a float64 conv of the same TF32-rounded operands, and dgrad / wgrad through autograd vs the
    strict-fp32 CUDA-core path (same tolerances as the folded kernel)."""
    from pu_b200 import _lib, ops
    assert _lib.load().pu_conv3x3_tc_flat(B, H, W, C0, C1, Cout) == 1
    from pu_b200 import ops as _o
    _o._tc_resident.cache_clear()
    g = torch.Generator().manual_seed(C0 + Cout + H)
    Cin = C0 + C1
    x0 = tf32_round(torch.randn(B, H, W, C0, generator=g))
    x1 = tf32_round(torch.randn(B, H, W, C1, generator=g)) if C1 else None
    w = tf32_round(torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5))
    b = torch.randn(Cout, generator=g)
    R = tf32_round(torch.randn(B, H, W, Cout, generator=g))
    cat = nchw(x0).double() if not C1 else torch.cat([nchw(x0).double(), nchw(x1).double()], 1)
    yr = F.relu(F.conv2d(cat, w.double(), b.double(), padding=1))
    outs = {}
    for math in (ops.MATH_FP32, ops.MATH_TF32):
        x0d = x0.to(DEV).requires_grad_(True)
        x1d = x1.to(DEV).requires_grad_(True) if C1 else None
        wd, bd = w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
        y = ops.conv3x3(x0d, x1d, wd, bd, None, True, H, W, 0, 0, 0, 0, math)
        (y * R.to(DEV)).sum().backward()
        outs[math] = (y.detach(), x0d.grad, x1d.grad if C1 else None, wd.grad, bd.grad)
    torch.cuda.synchronize()
    e = rel_err(nchw(outs[ops.MATH_TF32][0]), yr)
    assert e[0] < 6e-4, "forward: max-rel %g l2-rel %g" % e
    for n, a, bb in zip(["y", "dx0", "dx1", "dw", "db"], outs[ops.MATH_TF32], outs[ops.MATH_FP32]):
        if a is not None:
            e = rel_err(a, bb)
            assert e[0] < 2e-3, "%s: max-rel %g l2-rel %g" % (n, e[0], e[1])
