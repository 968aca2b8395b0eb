"""CPU tests of the index algebra behind the tcgen05 kernels (no GPU, no extension calls): numpy restatements of
(a) the kx-folded implicit GEMM with overlapping 8-row groups and its shuffle realignment (csrc/conv3x3_tc.cu), and
(b) ConvTranspose2d(k=3, s=2, p=0) + crop as a zero-inserted stride-1 convolution with the flipped kernel
(ops.convT3x3s2_tc; reference unet_p_res.py:207,214-217)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "plastic-unet_b200"))


@pytest.mark.parametrize("TH,TW,Cin,Cout", [(5, 14, 8, 8), (3, 30, 16, 8), (7, 9, 8, 16)])
def test_kx_folded_blocks_equal_conv3x3(TH, TW, Cin, Cout):
    """E[row, (kx, co)] = sum_{ky, ci} A[pixel(row) + ky*PW, ci] * W[co, ci, ky, kx] over 128-row blocks whose 16 groups
    of 8 rows start 6 pixels apart; out[p] = E0[row] + E1[row+1] + E2[row+2] for rows with (row & 7) < 6.  Must equal the
    zero-padded 3x3 convolution on every valid pixel of the tile."""
    rng = np.random.default_rng(TH * 100 + TW)
    PW = TW + 2
    halo = rng.standard_normal(((TH + 2), PW, Cin))  # halo tile incl. the zero padding ring where the image ends
    w = rng.standard_normal((Cout, Cin, 3, 3))
    flat = halo.reshape(-1, Cin)
    nmb = -(-(TH * PW) // 96)
    rows_needed = (nmb - 1) * 96 + 98 + 2 * PW
    flat = np.concatenate([flat, rng.standard_normal((max(0, rows_needed - flat.shape[0]), Cin))])  # garbage tail rows
    out = np.full((TH, TW, Cout), np.nan)
    for mb in range(nmb):
        row_pixel = np.array([mb * 96 + (r >> 3) * 6 + (r & 7) for r in range(128)])
        E = np.zeros((128, 3, Cout))
        for ky in range(3):
            A = flat[row_pixel + ky * PW]  # descriptor start offset ky*PW rows, SBO = 6 rows
            for kx in range(3):
                E[:, kx, :] += A @ w[:, :, ky, kx].T
        for r in range(128):
            if (r & 7) >= 6:
                continue  # rows 6,7 of a group duplicate rows 0,1 of the next one
            p = row_pixel[r]
            yy, xx = divmod(p, PW)
            if yy < TH and xx < TW:
                out[yy, xx] = E[r, 0] + E[r + 1, 1] + E[r + 2, 2]
    ref = F.conv2d(torch.from_numpy(halo).permute(2, 0, 1)[None], torch.from_numpy(w))[0].permute(1, 2, 0).numpy()
    assert not np.isnan(out).any(), "every valid pixel of the tile must be produced by exactly one block row"
    np.testing.assert_allclose(out, ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("H,W", [(6, 6), (12, 12), (5, 7)])
def test_zero_inserted_conv_equals_cropped_conv_transpose(H, W):
    from pu_b200 import ops
    rng = np.random.default_rng(H * 10 + W)
    Cin, Cout = 3, 2
    x = torch.from_numpy(rng.standard_normal((1, Cin, H, W)))
    w = torch.from_numpy(rng.standard_normal((Cin, Cout, 3, 3)))
    full = F.conv_transpose2d(x, w, stride=2)  # (2H+1) x (2W+1)
    w_conv = w.flip(2, 3).permute(1, 0, 2, 3)  # w_conv[co][ci][k] = w[ci][co][2-k]
    checked = 0
    for oy in range(0, 4):
        for ox in range(0, 4):
            for ey in range(0, 3):
                for ex in range(0, 3):
                    Ho, Wo = 2 * H + 1 - oy - ey, 2 * W + 1 - ox - ex
                    z = torch.zeros(1, Cin, Ho, Wo, dtype=torch.float64)
                    for r in range(Ho):
                        for s in range(Wo):
                            fy, fx = r + oy, s + ox
                            if fy % 2 == 1 and fx % 2 == 1 and (fy - 1) // 2 < H and (fx - 1) // 2 < W:
                                z[0, :, r, s] = x[0, :, (fy - 1) // 2, (fx - 1) // 2]
                    got = F.conv2d(z, w_conv, padding=1)
                    same = torch.allclose(got, full[:, :, oy:oy + Ho, ox:ox + Wo], atol=1e-12)
                    ok = ops.zero_window_ok(oy, Ho, 2 * H + 1) and ops.zero_window_ok(ox, Wo, 2 * W + 1)
                    if ok:
                        assert same, (oy, ox, ey, ex)
                        checked += 1
                    # the reference's crops (0 or 1 from the start, none from the end) are always accepted
                    if oy <= 1 and ox <= 1 and ey == 0 and ex == 0:
                        assert ok
    assert checked > 0
