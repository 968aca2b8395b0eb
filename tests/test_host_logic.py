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


@pytest.mark.parametrize("C,H,W", [(8, 4, 5), (16, 3, 3)])
def test_convT2x2_as_three_pixel_gemms(C, H, W):
    """csrc/convT_mma.cu: ConvTranspose2d(C, C, 2, stride=2) (unet_p.py:155) as GEMMs over the P = H*W input pixels with
    Wm[ci, (a, c, co)] = w[ci, co, a, c]: forward Y' = X Wm + pixel shuffle; dgrad dX = sum_a dY_a Wm_a^T where dY_a is
    the output row 2h+a viewed as (c, co); wgrad dWm = X^T dY', db = column sums folded over (a, c)."""
    rng = np.random.default_rng(C + H)
    x = torch.from_numpy(rng.standard_normal((1, C, H, W))).requires_grad_(True)
    w = torch.from_numpy(rng.standard_normal((C, C, 2, 2))).requires_grad_(True)
    b = torch.from_numpy(rng.standard_normal(C)).requires_grad_(True)
    R = torch.from_numpy(rng.standard_normal((1, C, 2 * H, 2 * W)))
    y = F.conv_transpose2d(x, w, b, stride=2)
    (y * R).sum().backward()
    X = x.detach().permute(0, 2, 3, 1).reshape(H * W, C).numpy()                     # [P, Cin]
    Wm = w.detach().permute(0, 2, 3, 1).reshape(C, 4 * C).numpy()                    # [Cin, (a, c, co)]
    Yp = X @ Wm + np.tile(b.detach().numpy(), 4)                                     # [P, (a, c, co)]
    y_ours = Yp.reshape(H, W, 2, 2, C).transpose(0, 2, 1, 3, 4).reshape(2 * H, 2 * W, C)   # pixel shuffle
    np.testing.assert_allclose(y_ours, y.detach()[0].permute(1, 2, 0).numpy(), rtol=1e-10, atol=1e-10)
    dY = R[0].permute(1, 2, 0).numpy()                                               # [2H, 2W, Cout]
    dYp = dY.reshape(H, 2, W, 2, C).transpose(0, 2, 1, 3, 4).reshape(H * W, 4 * C)   # [P, (a, c, co)]
    dX = np.zeros((H * W, C))
    for a in range(2):
        dYa = dYp[:, a * 2 * C:(a + 1) * 2 * C]                                      # output row parity a: (c, co) contiguous
        dX += dYa @ Wm[:, a * 2 * C:(a + 1) * 2 * C].T
    np.testing.assert_allclose(dX.reshape(H, W, C), x.grad[0].permute(1, 2, 0).numpy(), rtol=1e-10, atol=1e-10)
    dWm = X.T @ dYp
    np.testing.assert_allclose(dWm.reshape(C, 2, 2, C).transpose(0, 3, 1, 2), w.grad.numpy(), rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(dYp.sum(0).reshape(4, C).sum(0), b.grad.numpy(), rtol=1e-10, atol=1e-10)


def test_wgrad_tap_pair_tiles():
    """conv3x3_wgrad_mma_kernel: dw[co, ci, tap] = sum_p x[p + tap] g[p, co] as five m16n8k8 tiles per 8-pixel group:
    tile t holds taps (2t, 2t+1) in its 16 rows (8 ci each); the spare rows of tile 4 are fed ones => column sums of g."""
    rng = np.random.default_rng(3)
    H, W, Ci, Co = 6, 16, 8, 8
    x = rng.standard_normal((H + 2, W + 2, Ci))          # halo tile
    g = rng.standard_normal((H, W, Co))
    acc = np.zeros((5, 16, Co))
    for yy in range(H):
        for xg in range(0, W, 8):
            B = g[yy, xg:xg + 8]                         # [8 pixels, Co]
            for t in range(5):
                A = np.ones((16, 8))
                for half in range(2):
                    tap = 2 * t + half
                    if tap <= 8:
                        ky, kx = divmod(tap, 3)
                        A[8 * half:8 * half + 8] = x[yy + ky, xg + kx:xg + kx + 8].T   # [ci, pixel]
                acc[t] += A @ B
    dw = np.zeros((Co, Ci, 9))
    for tap in range(9):
        dw[:, :, tap] = acc[tap // 2, 8 * (tap % 2):8 * (tap % 2) + 8].T
    xt = torch.from_numpy(x).permute(2, 0, 1)[None]
    gt = torch.from_numpy(g).permute(2, 0, 1)[None]
    ref = torch.nn.grad.conv2d_weight(xt, (Co, Ci, 3, 3), gt, padding=0).numpy().reshape(Co, Ci, 9)
    np.testing.assert_allclose(dw, ref, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(acc[4, 8], g.sum((0, 1)), rtol=1e-10, atol=1e-10)   # ones rows = bias gradient


def _plan(B, H, W, C0, C1, Cout, resident):
    import ctypes
    from pu_b200 import _lib
    out = (ctypes.c_int * 17)()
    lib = _lib.load()
    rc = lib.pu_conv3x3_tc_plan(B, H, W, C0, C1, Cout, 1 if resident else 0, ctypes.addressof(out))
    if rc != 0:
        return None
    keys = ["TH", "TW", "PW", "tilesX", "tilesY", "nmb", "cols", "n3", "a_bytes", "w_stage", "w_res", "tmem_cols", "nchunks",
            "ncoblk", "nstages", "smem", "fold"]
    return dict(zip(keys, list(out)))


def _model_conv_shapes():
    """(C0, C1, Cout, side, batch) of every tcgen05-eligible 3x3 conv (forward and its dgrad) of UNetp @128/@512,
    UNetpCoord @128 and UNetpRes n16/n8 @101 (reference unet_p.py:33-49, unet_p_res.py:36-66)."""
    shapes = set()
    for side, B in ((128, 64), (512, 4)):
        ch = [8, 16, 32, 64, 64]
        for lvl in range(5):
            s = side >> lvl
            c = ch[lvl]
            shapes.add((c, 0, c, s, B))                       # second conv of a double_conv (+ its dgrad: same shape)
            if lvl > 0:
                shapes.add((ch[lvl - 1], 0, c, s, B))         # first conv of down
                shapes.add((c, 0, ch[lvl - 1], s, B))         # its dgrad
        for cin, cout, lvl in ((64, 32, 3), (32, 16, 2), (16, 8, 1), (8, 8, 0)):
            s = side >> lvl
            shapes.add((cin, cin, cout, s, B))                # conv on cat[skip, up]
            shapes.add((cout, 0, 2 * cin, s, B))              # its dgrad (outputs split over the two sources)
    for n in (16, 8):
        sizes = [101, 50, 25, 12, 6]
        for lvl in range(5):
            c = n << lvl
            shapes.add((c, 0, c, sizes[lvl], 32))
            if lvl > 0:
                shapes.add((c >> 1, 0, c, sizes[lvl], 32))
                shapes.add((c, 0, c >> 1, sizes[lvl], 32))
                shapes.add((c >> 1, c >> 1, c >> 1, sizes[lvl - 1], 32))   # up: cat[up, skip] -> C/2
                shapes.add((c >> 1, 0, c, sizes[lvl - 1], 32))             # its dgrad
                shapes.add((c, 0, c >> 1, sizes[lvl - 1], 32))             # transposed conv as a conv on the canvas
    return sorted(t for t in shapes if t is not None)


def test_tc_conv_planner_invariants():
    """Host-only planner of csrc/conv3x3_tc.cu (pu_conv3x3_tc_plan): every conv shape of the models gets a plan that
    respects the hardware limits the kernel relies on."""
    shapes = _model_conv_shapes()
    assert len(shapes) > 40
    planned = 0
    for (C0, C1, Cout, side, B) in shapes:
        for resident in (True, False):
            p = _plan(B, side, side, C0, C1, Cout, resident)
            if p is None:
                assert resident, "streamed-weights plan must exist for %s" % ((C0, C1, Cout, side, B),)
                continue
            planned += 1
            ctx = ((C0, C1, Cout, side, B, resident), p)
            assert p["cols"] in (8, 16, 32, 64) and p["n3"] == -(-3 * p["cols"] // 16) * 16, ctx
            assert p["PW"] == p["TW"] + 2 and p["PW"] <= 256 and p["TH"] + 2 <= 256, ctx            # TMA box limits
            assert p["tilesX"] * p["TW"] >= side and p["tilesY"] * p["TH"] >= side, ctx               # tiles cover the image
            blk = 96 if p["fold"] else 128          # folded: kx taps in N, 96 output pixels per block; flat: 128
            nacc = p["n3"] if p["fold"] else max(16, p["cols"])  # TMEM columns per block
            assert p["nmb"] * blk >= p["TH"] * p["PW"], ctx                                          # blocks cover the tile rows
            assert p["nmb"] * nacc <= 256, ctx                                                       # one TMEM accumulator buffer
            assert p["fold"] or p["nmb"] <= (3 if p["cols"] >= 64 else 6), ctx                       # flat: blocks per MMA-issuing warp
            assert p["tmem_cols"] <= 512 and p["tmem_cols"] >= 2 * p["nmb"] * nacc, ctx
            assert p["tmem_cols"] & (p["tmem_cols"] - 1) == 0 and p["tmem_cols"] >= 32, ctx
            assert 2 <= p["nstages"] <= 4 and p["a_bytes"] % 1024 == 0 and p["w_stage"] % 1024 == 0, ctx
            assert p["smem"] <= 227 * 1024, ctx
            assert (p["w_res"] > 0) == resident and (p["w_stage"] == 0) == resident, ctx
            assert p["ncoblk"] == -(-Cout // 64) and 1 <= p["nchunks"] <= 64, ctx
            # shared memory accounting: stages + resident weights + bookkeeping
            assert p["smem"] >= p["nstages"] * (p["a_bytes"] + p["w_stage"]) + p["w_res"], ctx
    assert planned > len(shapes)  # most shapes have both a resident and a streamed plan


def test_wgrad_tma_plan_invariants():
    """Host planner of the TMA-fed weight-gradient kernel (pu_conv3x3_wgrad_plan, no device): for every conv shape of the five
    BASELINE configurations the tile covers 16 / NCI strips of 8x8 pixels, the chunk / co-tile split divides the channel counts,
    the grid fits the 148 SMs and the ring + reduction buffer fit in shared memory."""
    import ctypes
    from pu_b200 import _lib
    lib = _lib.load()
    shapes = []
    for B, s0 in ((64, 128), (32, 101), (8, 512), (1, 21), (3, 37)):
        s, c = s0, 8
        for _ in range(6):
            if s < 1:
                break
            shapes += [(B, s, s, c, 0, c), (B, s, s, c, c, c), (B, s, s, c, 0, 2 * c), (B, s, s, 2 * c, 2 * c, c)]
            s //= 2
            c *= 2
    shapes += [(1, 21, 37, 24, 0, 40), (2, 6, 300, 8, 8, 8), (64, 128, 128, 8, 8, 8)]
    out = (ctypes.c_int * 12)()
    for (B, H, W, C0, C1, Cout) in shapes:
        assert lib.pu_conv3x3_wgrad_plan(B, H, W, C0, C1, Cout, out) == 0, (B, H, W, C0, C1, Cout)
        nci, nco, TW, TH, NB, tX, tY, tB, stages, gx, gy, smem = list(out)
        nchunks, ncot = (C0 + C1) // 8, Cout // 8
        assert nci in (1, 2, 4) and nco in (1, 2) and nchunks % nci == 0 and ncot % nco == 0
        assert TW % 8 == 0 and TH % 8 == 0 and (TW // 8) * (TH // 8) * NB == 16 // nci
        assert not (C0 == 8 or C1 == 8) or 8 * (TW + 2) <= 256  # (channel, x)-merged box rows of the 8-channel sources
        assert Cout != 8 or 8 * TW <= 256
        assert tX * TW >= W and tY * TH >= H and tB * NB >= B and (tX - 1) * TW < W and (tY - 1) * TH < H and (tB - 1) * NB < B
        assert gy == (nchunks // nci) * (ncot // nco)
        assert 1 <= gx <= max(1, 148 // gy) and gx <= tX * tY * tB
        assert 1 <= stages <= 4 and smem <= 227 * 1024
        stage = nci * (((TW + 2) * (TH + 2) * NB * 32 + 127) // 128 * 128) + nco * ((TW * TH * NB * 32 + 127) // 128 * 128)
        assert smem >= max(stages * stage, 16 * 32 * 20 * nco * 4)
    assert lib.pu_conv3x3_wgrad_plan(1, 8, 8, 4, 0, 8, out) != 0  # channel counts must be multiples of 8


def test_wgrad_tma_strip_walk():
    """conv3x3_wgrad_tma_kernel's index algebra in numpy: a tile of NB x TH x TW pixels is cut into 8x8 strips, a warp walks
    down its strip keeping halo rows (yy, yy+1, yy+2) of the [pixel][8] x plane as the A fragments of taps (ky, kx) — element
    (row ci, k = pixel j) = plane[yy + ky][xs + j + kx][ci] — against B = g plane[yy][xs + j][co]; five tap-pair tiles per
    group, the spare rows of the fifth fed ones (bias gradient).  Zero fill outside the image = TMA OOB fill."""
    rng = np.random.default_rng(11)
    B, H, W, Ci, Co = 3, 13, 21, 8, 8
    TW, TH, NB = 16, 8, 2                              # 2 x 1 strips x 2 images = 4 strips (NCI = 4)
    x = rng.standard_normal((B, H, W, Ci))
    g = rng.standard_normal((B, H, W, Co))
    acc = np.zeros((5, 16, Co))
    for tb in range(-(-B // NB)):
        for ty in range(-(-H // TH)):
            for tx in range(-(-W // TW)):
                xp = np.zeros((NB, TH + 2, TW + 2, Ci))  # the staged halo plane (box start (x0 - 1, y0 - 1, b0), OOB -> 0)
                gp = np.zeros((NB, TH, TW, Co))
                for nb in range(NB):
                    b = tb * NB + nb
                    if b >= B:
                        continue
                    for hy in range(TH + 2):
                        for hx in range(TW + 2):
                            yy, xx = ty * TH + hy - 1, tx * TW + hx - 1
                            if 0 <= yy < H and 0 <= xx < W:
                                xp[nb, hy, hx] = x[b, yy, xx]
                    for hy in range(TH):
                        for hx in range(TW):
                            yy, xx = ty * TH + hy, tx * TW + hx
                            if yy < H and xx < W:
                                gp[nb, hy, hx] = g[b, yy, xx]
                for ps in range(NB * (TH // 8) * (TW // 8)):
                    tws, ths = TW // 8, TH // 8
                    sx, sy, nb = ps % tws, (ps // tws) % ths, ps // (tws * ths)
                    for yy in range(8):
                        Bf = gp[nb, sy * 8 + yy, sx * 8:sx * 8 + 8]                     # [8 pixels, Co]
                        for t in range(5):
                            A = np.ones((16, 8))
                            for half in range(2):
                                tap = 2 * t + half
                                if tap <= 8:
                                    ky, kx = divmod(tap, 3)
                                    A[8 * half:8 * half + 8] = xp[nb, sy * 8 + yy + ky, sx * 8 + kx:sx * 8 + kx + 8].T
                            acc[t] += A @ Bf
    dw = np.zeros((Co, Ci, 9))
    for tap in range(9):
        dw[:, :, tap] = acc[tap // 2, 8 * (tap % 2):8 * (tap % 2) + 8].T
    ref = torch.nn.grad.conv2d_weight(torch.from_numpy(x).permute(0, 3, 1, 2), (Co, Ci, 3, 3), torch.from_numpy(g).permute(0, 3, 1, 2),
                                      padding=1).numpy().reshape(Co, Ci, 9)
    np.testing.assert_allclose(dw, ref, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(acc[4, 8], g.sum((0, 1, 2)), rtol=1e-10, atol=1e-10)


# --------------------------------------------------------------------------------------------------
# fused training-step head (csrc/head.cu, head_bce_fused_kernel) and the pooling arg-max code (csrc/pool.cu)
# --------------------------------------------------------------------------------------------------
def _hf_swz(r):
    return ((r & 3) << 3) | (r & 4)


def test_head_weff_swizzle_serves_both_gemms_without_bank_conflicts():
    """Weff sits in shared memory once, [128][128] floats, column index XOR-swizzled by s(r) = 8*(r&3) + 4*((r>>2)&1).  The B
    fragments of an mma.sync.m16n8k8 (lane = 4*g + t: b0 = B[k=t][n=g], b1 = B[k=t+4][n=g]) must hit 32 distinct banks both for
    phase 1 (B[k][n] = Weff[k][n]) and for phase 2 (B[k][n] = Weff[n][k]); float4 staging stores must stay whole and distinct."""
    lanes = [(l >> 2, l & 3) for l in range(32)]
    for kk in range(0, 128, 8):
        for n0 in range(0, 128, 8):
            for dk in (0, 4):
                p1 = {((kk + t + dk) * 128 + ((n0 + g) ^ _hf_swz(kk + t + dk))) % 32 for g, t in lanes}
                p2 = {((n0 + g) * 128 + ((kk + t + dk) ^ _hf_swz(n0 + g))) % 32 for g, t in lanes}
                assert len(p1) == 32 and len(p2) == 32, (kk, n0, dk)
    for r in range(128):
        cols = [(4 * q) ^ _hf_swz(r) for q in range(32)]
        assert all(c % 4 == 0 for c in cols) and len(set(cols)) == 32  # a swizzled quad is still a 16-byte aligned quad
    # the A fragments come from the [64][132] X / gA tile: rows g (and g + 8), columns k + t (and + 4)
    assert len({((g * 132) + t) % 32 for g, t in lanes}) == 32


def test_three_term_tf32_split_keeps_fp32_level_products():
    """hi = the top 19 bits of x (a valid TF32 operand), lo = the exact residual truncated to TF32; x*y ~ hi*hi' + hi*lo' + lo*hi'.
    Emulated bit for bit in numpy: every product is within 2^-19 of the exact one, and a K = 128 dot product of head-like operands
    matches float64 to ~1e-6 — while the plain TF32 product (phase 2, gX) is at the 1e-3 level."""
    rng = np.random.default_rng(0)

    def split(x):
        xi = x.astype(np.float32).view(np.uint32)
        hi = (xi & np.uint32(0xffffe000)).view(np.float32)
        lo = ((x.astype(np.float32) - hi).view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)
        return hi.astype(np.float64), lo.astype(np.float64)

    x = rng.standard_normal(100000).astype(np.float32)
    y = (0.05 * rng.standard_normal(100000)).astype(np.float32)
    xh, xl = split(x)
    yh, yl = split(y)
    exact = x.astype(np.float64) * y.astype(np.float64)
    three = xh * yh + xh * yl + xl * yh
    assert np.max(np.abs(three - exact) / np.abs(exact)) < 2.0 ** -19
    one = xh * yh
    assert np.max(np.abs(one - exact) / np.abs(exact)) > 2.0 ** -12  # why the logits need the split
    X = rng.standard_normal((64, 128)).astype(np.float32)
    W = (0.02 * rng.standard_normal((128, 128))).astype(np.float32)
    Xh, Xl = split(X)
    Wh, Wl = split(W)
    ref = X.astype(np.float64) @ W.astype(np.float64)
    z3 = Xh @ Wh + Xh @ Wl + Xl @ Wh
    assert np.max(np.abs(z3 - ref)) / np.max(np.abs(ref)) < 2e-6
    assert np.max(np.abs(Xh @ Wh - ref)) / np.max(np.abs(ref)) > 1e-4


@pytest.mark.parametrize("relu", [False, True])
def test_pool_argmax_code_routes_like_max_pool2d(relu):
    """The forward pooling kernel records, per pooled element, bits 0-1 = position of the maximum in the 2x2 window (first element
    in row-major order that is strictly greater, or NaN — ATen's rule) and bit 2 = (maximum > 0); the backward kernel routes from
    that byte alone.  numpy restatement against torch's own max_pool2d indices and autograd, including windows of four zeros."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 5, 8, 6, generator=g)  # [B, C, H, W]
    if relu:
        x = torch.relu(x)
    xn = x.numpy()
    B, C, H, W = xn.shape
    m = np.full((B, C, H // 2, W // 2), -np.inf, dtype=np.float32)
    arg = np.zeros((B, C, H // 2, W // 2), dtype=np.uint8)
    for k in range(4):
        v = xn[:, :, (k >> 1)::2, (k & 1)::2][:, :, :H // 2, :W // 2]
        take = (v > m) | np.isnan(v)
        m = np.where(take, v, m)
        arg = np.where(take, k, arg).astype(np.uint8)
    code = arg | ((m > 0).astype(np.uint8) << 2)
    y, idx = F.max_pool2d(x, 2, return_indices=True)
    iy, ix = (idx // W).numpy(), (idx % W).numpy()
    oy, ox = np.meshgrid(np.arange(H // 2), np.arange(W // 2), indexing="ij")
    assert np.array_equal((code & 3), ((iy - 2 * oy) * 2 + (ix - 2 * ox)).astype(np.uint8))
    assert np.array_equal(m, y.numpy())
    # routing (mask_in = the ReLU mask of x): dx = dy at the recorded position, zeroed where the maximum is not > 0
    dy_ = torch.randn(y.shape, generator=g)
    xr = x.clone().requires_grad_(True)
    (F.max_pool2d(xr, 2) * dy_).sum().backward()
    want = xr.grad.numpy() * ((xn > 0) if relu else 1.0)
    got = np.zeros_like(xn)
    for k in range(4):
        sel = ((code & 3) == k) & (((code >> 2) & 1).astype(bool) | (not relu))
        got[:, :, (k >> 1)::2, (k & 1)::2][:, :, :H // 2, :W // 2] = np.where(sel, dy_.numpy(), 0.0)
    assert np.array_equal(got, want.astype(np.float32))
