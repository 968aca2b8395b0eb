"""GPU tests of the TMA-fed TF32 weight-gradient kernel (csrc/conv3x3_wgrad_tma.cu) through the C-ABI.

The kernel multiplies TF32 operands exactly (inputs are stored TF32-rounded by their producers) and accumulates in fp32
(mma.sync partial sums per warp, a shared-memory reduction per CTA, fp32 atomics across CTAs), so against a float64
weight gradient of the SAME operands only the summation order differs: tolerance 2e-5 of max|dw| (K up to 10^6 pixels).
Reference: the parameter gradients autograd computes for nn.Conv2d(k=3, p=1) (unet_p.py:105-116, unet_p_res.py:150-158)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def tf32_round(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


SHAPES = [
    # B, C0, C1, Cout, H, W, (oy0, ox0, extra rows/cols of source 0), with_bias
    (2, 8, 0, 8, 128, 128, (0, 0, 0), True),      # inc.2 / up4.2: one chunk, 1024-pixel tiles
    (3, 8, 8, 8, 128, 128, (0, 0, 0), True),      # up4.0: two sources, one chunk each
    (2, 16, 0, 16, 64, 64, (0, 0, 0), True),      # 2 chunks x 2 co tiles
    (2, 8, 0, 16, 64, 64, (0, 0, 0), False),
    (5, 16, 16, 8, 64, 64, (1, 2, 3), True),      # cropped skip connection (window inside a larger tensor), 4 chunks
    (2, 32, 32, 16, 32, 32, (0, 0, 0), True),     # grid.y = 2 chunk groups
    (3, 64, 0, 64, 16, 16, (0, 0, 0), True),      # 8 channel groups
    (7, 64, 64, 32, 16, 16, (0, 0, 0), True),
    (9, 64, 0, 64, 8, 8, (0, 0, 0), True),        # several images per tile, batch tail (9 % 4 != 0)
    (3, 8, 8, 8, 101, 101, (0, 0, 0), True),      # UNetpRes sizes: ragged tiles
    (2, 16, 16, 16, 50, 50, (0, 1, 1), True),
    (2, 32, 0, 32, 25, 25, (0, 0, 0), True),
    (5, 128, 0, 128, 6, 6, (0, 0, 0), True),
    (1, 24, 0, 40, 21, 37, (0, 0, 0), True),      # chunk counts 3 and 5: one chunk / one co tile per CTA
    (1, 8, 0, 8, 5, 5, (0, 0, 0), True),
]


def _ref(x, g):
    """float64 weight gradient and bias gradient; x: [B, H, W, Cin] (window), g: [B, H, W, Cout]"""
    x64 = x.double().permute(0, 3, 1, 2)
    g64 = g.double().permute(0, 3, 1, 2)
    dw = torch.nn.grad.conv2d_weight(x64, (g.shape[3], x.shape[3], 3, 3), g64, padding=1)
    return dw, g64.sum((0, 2, 3))


@pytest.mark.parametrize("B,C0,C1,Cout,H,W,crop,with_bias", SHAPES)
def test_wgrad_tma_vs_float64(B, C0, C1, Cout, H, W, crop, with_bias):
    from pu_b200 import _lib
    gen = torch.Generator().manual_seed(B + 3 * C0 + 5 * C1 + 7 * Cout + 11 * H)
    oy, ox, extra = crop
    H0, W0 = H + oy + extra, W + ox + extra
    x0 = tf32_round(torch.randn(B, H0, W0, C0, generator=gen)).to(DEV)
    x1 = tf32_round(torch.randn(B, H, W, C1, generator=gen)).to(DEV) if C1 else None
    g = tf32_round(torch.randn(B, H, W, Cout, generator=gen)).to(DEV)
    dw = torch.full((Cout, C0 + C1, 3, 3), float("nan"), device=DEV)
    db = torch.full((Cout,), float("nan"), device=DEV) if with_bias else None
    _lib.call("pu_conv3x3_wgrad", x0.data_ptr(), H0, W0, C0, oy, ox, x1.data_ptr() if C1 else None, H, W, C1, 0, 0,
              g.data_ptr(), dw.data_ptr(), db.data_ptr() if with_bias else None, B, H, W, Cout, 1, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    xw = x0[:, oy:oy + H, ox:ox + W, :]
    xcat = torch.cat([xw, x1], dim=3) if C1 else xw
    dw_ref, db_ref = _ref(xcat, g)
    err = float((dw.double() - dw_ref).abs().max() / dw_ref.abs().max())
    assert err < 2e-5, err
    if with_bias:
        errb = float((db.double() - db_ref).abs().max() / db_ref.abs().max())
        assert errb < 2e-5, errb


@pytest.mark.parametrize("B,Cout,H,W,pad", [(2, 8, 128, 128, 0), (3, 16, 64, 64, 0), (5, 8, 40, 36, 0), (2, 8, 30, 44, 4), (1, 16, 8, 8, 0),
                                            (2, 8, 37, 21, 0)])
def test_stem_wgrad_tf32(B, Cout, H, W, pad):
    """The one-input-channel stem in TF32 mode: TMA + MMA kernel (taps as the rows of ONE m16n8k8 tile per 8 pixels, ones row =
    bias gradient) for 16-byte aligned inputs, the streaming FFMA kernel otherwise (W = 21).  The input image is rounded to TF32
    inside the kernel, so the reference uses the rounded image."""
    from pu_b200 import _lib
    gen = torch.Generator().manual_seed(B + Cout + H + W)
    x = torch.randn(B, H + pad, W + pad, 1, generator=gen).to(DEV)   # the op's window starts at (pad, pad) of a larger tensor
    g = tf32_round(torch.randn(B, H, W, Cout, generator=gen)).to(DEV)
    dw = torch.full((Cout, 1, 3, 3), float("nan"), device=DEV)
    db = torch.full((Cout,), float("nan"), device=DEV)
    _lib.call("pu_conv3x3_wgrad", x.data_ptr(), H + pad, W + pad, 1, pad, pad, None, 0, 0, 0, 0, 0,
              g.data_ptr(), dw.data_ptr(), db.data_ptr(), B, H, W, Cout, 1, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    xw = x[:, pad:pad + H, pad:pad + W, :]
    aligned = W % 4 == 0 and (W + pad) % 4 == 0 and pad % 4 == 0
    dw_ref, db_ref = _ref(tf32_round(xw.cpu()).to(DEV) if aligned else xw, g)
    err = float((dw.double() - dw_ref).abs().max() / dw_ref.abs().max())
    errb = float((db.double() - db_ref).abs().max() / db_ref.abs().max())
    assert err < 2e-5 and errb < 2e-5, (err, errb)


def test_wgrad_tma_matches_first_kernel_bitwise_inputs():
    """A/B: the TMA-fed kernel and the cp.async-fed first version (PU_WGRAD_V=1, selected in a fresh process) see the same
    operands; both stay within the fp32 summation-order bound of the float64 result."""
    import subprocess
    import sys
    code = r"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.getcwd(), "plastic-unet_b200"))
from pu_b200 import _lib
gen = torch.Generator().manual_seed(5)
x = torch.randn(4, 64, 64, 16, generator=gen).cuda(); g = torch.randn(4, 64, 64, 16, generator=gen).cuda()
i = x.view(torch.int32); x = ((i + 0x1000) & ~0x1FFF).view(torch.float32)
i = g.view(torch.int32); g = ((i + 0x1000) & ~0x1FFF).view(torch.float32)
dw = torch.empty(16, 16, 3, 3, device="cuda"); db = torch.empty(16, device="cuda")
_lib.call("pu_conv3x3_wgrad", x.data_ptr(), 64, 64, 16, 0, 0, None, 0, 0, 0, 0, 0, g.data_ptr(), dw.data_ptr(), db.data_ptr(), 4, 64, 64, 16, 1,
          torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
ref = torch.nn.grad.conv2d_weight(x.double().permute(0, 3, 1, 2), (16, 16, 3, 3), g.double().permute(0, 3, 1, 2), padding=1)
print("ERR %.3e %.3e" % (float((dw.double() - ref).abs().max() / ref.abs().max()),
                         float((db.double() - g.double().sum((0, 1, 2))).abs().max() / g.double().sum((0, 1, 2)).abs().max())))
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for ver in ("1", "2"):
        env = dict(os.environ, PU_WGRAD_V=ver)
        out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        e = [float(v) for v in out.stdout.split("ERR")[1].split()]
        assert e[0] < 2e-5 and e[1] < 2e-5, (ver, e)
