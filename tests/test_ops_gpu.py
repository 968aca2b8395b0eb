"""GPU parity tests, one custom op at a time: every pu_b200 op (forward and autograd) against the stock
PyTorch formula of the reference call site it replaces, evaluated on the CPU in float64.
Tolerance for the strict-fp32 kernels: 2e-5 relative to the output's max magnitude (fp32 re-association);
the north-star bound for the model is 1e-3."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-5
DEV = "cuda"


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def leaf(t, dev=DEV, dtype=torch.float32):
    return t.detach().to(device=dev, dtype=dtype).requires_grad_(True)


def check(a, b, tol=TOL, what=""):
    e = rel_err(a, b)
    assert e[0] < tol, "%s: max-rel %g l2-rel %g" % (what, e[0], e[1])


@pytest.fixture(autouse=True)
def _ops():
    from pu_b200 import ops
    return ops


CONV_SHAPES = [
    # B, Cin, Cout, H, W, relu, res, bias
    (2, 1, 8, 32, 32, True, False, True),     # stem: C_in = 1
    (2, 8, 8, 40, 33, True, True, True),      # ragged tile edges, residual
    (1, 3, 5, 9, 7, False, False, True),      # ragged channel counts, W <= 8 config
    (3, 32, 16, 16, 16, True, False, False),  # W <= 16 config, no bias
    (2, 64, 64, 8, 8, True, True, True),      # bottleneck
    (1, 16, 24, 101, 50, False, True, True),  # odd sizes of the residual net
    (2, 12, 8, 6, 6, True, False, True),      # 6x6 bottleneck of UNetpRes@101
]


@pytest.mark.parametrize("B,Cin,Cout,H,W,relu,res,bias", CONV_SHAPES)
def test_conv3x3_single_source(B, Cin, Cout, H, W, relu, res, bias):
    from pu_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + Cin * 10 + Cout)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (3 * Cin ** 0.5)
    b = torch.randn(Cout, generator=g) if bias else None
    r = torch.randn(B, Cout, H, W, generator=g) if res else None
    R = torch.randn(B, Cout, H, W, generator=g)
    # reference (float64, CPU): unet_p.py:105-107 / unet_p_res.py:150-158,186-189
    xr, wr = leaf(x, "cpu", torch.float64), leaf(w, "cpu", torch.float64)
    br = leaf(b, "cpu", torch.float64) if bias else None
    rr = leaf(r, "cpu", torch.float64) if res else None
    yr = F.conv2d(xr, wr, br, padding=1)
    if res:
        yr = yr + rr
    if relu:
        yr = F.relu(yr)
    (yr * R.double()).sum().backward()
    # ours
    xo, wo = leaf(nhwc(x)), leaf(w)
    bo = leaf(b) if bias else None
    ro = leaf(nhwc(r)) if res else None
    yo = ops.conv3x3(xo, None, wo, bo, ro, relu, H, W, 0, 0, 0, 0, ops.MATH_FP32)
    (yo * nhwc(R).to(DEV)).sum().backward()
    check(nchw(yo), yr, what="y")
    check(nchw(xo.grad), xr.grad, what="dx")
    check(wo.grad, wr.grad, 5 * TOL, what="dw")
    if bias:
        check(bo.grad, br.grad, 5 * TOL, what="db")
    if res:
        check(nchw(ro.grad), rr.grad, what="dres")


@pytest.mark.parametrize("Cout,H,W,B", [(8, 128, 128, 2), (16, 37, 21, 3)])
def test_conv3x3_stem_wgrad_with_bias(Cout, H, W, B):
    """C_in = 1 stem (unet_p.py:124-132): the streaming wgrad kernel also produces the bias gradient when the incoming
    gradient is already masked (premasked protocol).  fp32 math: 1e-3 relative."""
    from pu_b200 import ops
    g = torch.Generator().manual_seed(Cout + H)
    x = torch.randn(B, H, W, 1, generator=g).to(DEV)
    w = torch.randn(Cout, 1, 3, 3, generator=g).to(DEV)
    dy = torch.randn(B, H, W, Cout, generator=g).to(DEV)
    y = torch.ones_like(dy)
    _, _, _, dw, db = ops.conv3x3_bwd(dy, y, x, None, w, True, True, H, W, 0, 0, 0, 0, ops.MATH_FP32, False, True, None, None, True)
    xr = x.cpu().double().permute(0, 3, 1, 2)
    dyr = dy.cpu().double().permute(0, 3, 1, 2)
    dwr = torch.nn.grad.conv2d_weight(xr, (Cout, 1, 3, 3), dyr, padding=1)
    check(dw, dwr, 5 * TOL, what="dw")
    check(db, dyr.sum(dim=(0, 2, 3)), 5 * TOL, what="db")


@pytest.mark.parametrize("C0,C1,H,W,c0,c1", [(8, 8, 32, 32, (0, 0), (0, 0)), (16, 16, 12, 12, (0, 0), (1, 1)),
                                              (4, 12, 25, 25, (3, 2), (0, 0)), (64, 64, 16, 16, (0, 0), (0, 0))])
def test_conv3x3_two_sources_with_crop(C0, C1, H, W, c0, c1):
    """cat + negative F.pad (crop) + conv fused: unet_p.py:161-166, unet_p_res.py:215-219."""
    from pu_b200 import ops
    g = torch.Generator().manual_seed(C0 + C1 + H)
    B, Cout = 2, 8
    H0, W0, H1, W1 = H + c0[0] + 1, W + c0[1] + 2, H + c1[0], W + c1[1] + 1
    x0 = torch.randn(B, C0, H0, W0, generator=g)
    x1 = torch.randn(B, C1, H1, W1, generator=g)
    w = torch.randn(Cout, C0 + C1, 3, 3, generator=g) / (3 * (C0 + C1) ** 0.5)
    b = torch.randn(Cout, generator=g)
    R = torch.randn(B, Cout, H, W, generator=g)
    x0r, x1r, wr, br = (leaf(t, "cpu", torch.float64) for t in (x0, x1, w, b))
    cat = torch.cat([x0r[:, :, c0[0]:c0[0] + H, c0[1]:c0[1] + W], x1r[:, :, c1[0]:c1[0] + H, c1[1]:c1[1] + W]], 1)
    yr = F.relu(F.conv2d(cat, wr, br, padding=1))
    (yr * R.double()).sum().backward()
    x0o, x1o, wo, bo = leaf(nhwc(x0)), leaf(nhwc(x1)), leaf(w), leaf(b)
    yo = ops.conv3x3(x0o, x1o, wo, bo, None, True, H, W, c0[0], c0[1], c1[0], c1[1], ops.MATH_FP32)
    (yo * nhwc(R).to(DEV)).sum().backward()
    check(nchw(yo), yr, what="y")
    check(nchw(x0o.grad), x0r.grad, what="dx0")
    check(nchw(x1o.grad), x1r.grad, what="dx1")
    check(wo.grad, wr.grad, 5 * TOL, what="dw")
    check(bo.grad, br.grad, 5 * TOL, what="db")


@pytest.mark.parametrize("Cin,Cout,coords,relu", [(8, 1, 0, False), (16, 1, 0, False), (1, 8, 2, True), (1, 8, 3, True), (3, 4, 2, True)])
def test_conv1x1(Cin, Cout, coords, relu):
    """outc (unet_p.py:173) and the CoordConv stem (coord_conv_script.py:69-96,153)."""
    from pu_b200 import ops
    import plastic_unet_oracle as orc
    g = torch.Generator().manual_seed(Cin + Cout + coords)
    B, H, W = 2, 21, 21
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin + coords, generator=g)
    b = torch.randn(Cout, generator=g)
    R = torch.randn(B, Cout, H, W, generator=g)
    xr, wr, br = (leaf(t, "cpu", torch.float64) for t in (x, w, b))
    xin = xr
    if coords:
        xin = orc.add_coords(xr, with_r=(coords == 3))
    yr = F.conv2d(xin, wr.view(Cout, Cin + coords, 1, 1), br)
    if relu:
        yr = F.relu(yr)
    (yr * R.double()).sum().backward()
    xo, wo, bo = leaf(nhwc(x)), leaf(w), leaf(b)
    yo = ops.conv1x1(xo, wo, bo, coords, relu)
    (yo * nhwc(R).to(DEV)).sum().backward()
    check(nchw(yo), yr, what="y")
    check(nchw(xo.grad), xr.grad, what="dx")
    check(wo.grad, wr.grad, 5 * TOL, what="dw")
    check(bo.grad, br.grad, 5 * TOL, what="db")


@pytest.mark.parametrize("Cin,Cout,H,W", [(8, 8, 16, 16), (64, 64, 8, 8), (128, 64, 4, 4), (6, 3, 5, 7)])
def test_convT2x2s2(Cin, Cout, H, W):
    """nn.ConvTranspose2d(C, C, 2, stride=2): unet_p.py:155."""
    from pu_b200 import ops
    g = torch.Generator().manual_seed(Cin + H)
    B = 2
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cin, Cout, 2, 2, generator=g) / Cin ** 0.5
    b = torch.randn(Cout, generator=g)
    R = torch.randn(B, Cout, 2 * H, 2 * W, generator=g)
    xr, wr, br = (leaf(t, "cpu", torch.float64) for t in (x, w, b))
    yr = F.conv_transpose2d(xr, wr, br, stride=2)
    (yr * R.double()).sum().backward()
    xo, wo, bo = leaf(nhwc(x)), leaf(w), leaf(b)
    yo = ops.convT2x2s2(xo, wo, bo)
    (yo * nhwc(R).to(DEV)).sum().backward()
    check(nchw(yo), yr, what="y")
    check(nchw(xo.grad), xr.grad, what="dx")
    check(wo.grad, wr.grad, 5 * TOL, what="dw")
    check(bo.grad, br.grad, 5 * TOL, what="db")


@pytest.mark.parametrize("C,H,W,B,mask", [(8, 64, 64, 3, False), (16, 32, 32, 2, True), (32, 16, 16, 2, False), (64, 8, 8, 5, True),
                                          (8, 5, 7, 1, False), (64, 3, 3, 1, False)])
def test_convT2x2s2_tf32_mma(C, H, W, B, mask):
    """TF32 tensor-core path (convT_mma.cu: forward, dgrad with the producer's ReLU mask, wgrad + bias gradient) of
    nn.ConvTranspose2d(C, C, 2, stride=2), unet_p.py:155.  Operands are rounded to TF32 (10-bit mantissa, RN) and
    accumulated in fp32, so the bound is 2e-3 of the tensor's max instead of the fp32 path's 1e-3 relative."""
    from pu_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(B, C, H, W, generator=g)
    if mask:
        x = x.relu()
    w = torch.randn(C, C, 2, 2, generator=g) / C ** 0.5
    b = torch.randn(C, generator=g)
    R = torch.randn(B, C, 2 * H, 2 * W, generator=g)
    xr, wr, br = (leaf(t, "cpu", torch.float64) for t in (x, w, b))
    yr = F.conv_transpose2d(xr, wr, br, stride=2)
    (yr * R.double()).sum().backward()
    dx_ref = xr.grad * (x > 0).double() if mask else xr.grad
    xo, wo, bo = leaf(nhwc(x)), leaf(w), leaf(b)
    yo = ops.convT2x2s2(xo, wo, bo, True, mask)
    (yo * nhwc(R).to(DEV)).sum().backward()

    def close(got, ref, what):
        err = (got.detach().cpu().double() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
        assert err < 2e-3, "%s: max err / max |ref| = %g" % (what, err)

    close(nchw(yo), yr, "y")
    close(nchw(xo.grad), dx_ref, "dx")
    close(wo.grad, wr.grad, "dw")
    close(bo.grad, br.grad, "db")
    # the stored output is exactly representable in TF32 (PU_FLAG_ROUND_TF32)
    bits = yo.detach().view(torch.int32)
    assert int((bits & 0x1FFF).abs().max()) == 0


@pytest.mark.parametrize("Cin,Cout,H,W,crop,scaled", [(16, 8, 6, 6, (1, 1), False), (32, 16, 12, 12, (0, 0), True),
                                                       (8, 4, 25, 25, (1, 1), True), (64, 32, 3, 3, (0, 1), False)])
def test_convT3x3s2_cropped(Cin, Cout, H, W, crop, scaled):
    """nn.ConvTranspose2d(in, out, 3, stride=2) + negative-pad crop (+ Dropout2d scale): unet_p_res.py:207,214-217."""
    from pu_b200 import ops
    g = torch.Generator().manual_seed(Cin + H)
    B = 2
    Ho, Wo = 2 * H + 1 - (1 if crop[0] else 0), 2 * W + 1 - (1 if crop[1] else 0)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cin, Cout, 3, 3, generator=g) / Cin ** 0.5
    b = torch.randn(Cout, generator=g)
    s = (torch.rand(B, Cout, generator=g) > 0.5).float() * 2.0 if scaled else None
    R = torch.randn(B, Cout, Ho, Wo, generator=g)
    xr, wr, br = (leaf(t, "cpu", torch.float64) for t in (x, w, b))
    yr = F.conv_transpose2d(xr, wr, br, stride=2)[:, :, crop[0]:crop[0] + Ho, crop[1]:crop[1] + Wo]
    if scaled:
        yr = yr * s.double().view(B, Cout, 1, 1)
    (yr * R.double()).sum().backward()
    xo, wo, bo = leaf(nhwc(x)), leaf(w), leaf(b)
    yo = ops.convT3x3s2(xo, wo, bo, s.to(DEV) if scaled else None, Ho, Wo, crop[0], crop[1])
    (yo * nhwc(R).to(DEV)).sum().backward()
    check(nchw(yo), yr, what="y")
    check(nchw(xo.grad), xr.grad, what="dx")
    check(wo.grad, wr.grad, 5 * TOL, what="dw")
    check(bo.grad, br.grad, 5 * TOL, what="db")


@pytest.mark.parametrize("Cin,Cout,H,W,crop", [(32, 16, 50, 50, (0, 0)), (64, 32, 25, 25, (1, 1)), (256, 128, 6, 6, (1, 1)),
                                               (16, 8, 12, 12, (0, 1))])
def test_convT3x3s2_tc_path(Cin, Cout, H, W, crop):
    """TF32 path of nn.ConvTranspose2d(in, out, 3, stride=2) + crop (unet_p_res.py:207,214-217): zero insertion + tcgen05
    conv3x3 with the flipped kernel.  TF32 operands, fp32 accumulation: 2e-3 of the tensor's max."""
    from pu_b200 import ops
    g = torch.Generator().manual_seed(Cin + H)
    B = 2
    oy, ox = crop
    Ho, Wo = 2 * H + 1 - oy, 2 * W + 1 - ox
    assert ops.convT3x3s2_tc_ok(Cin, Cout, H, W, Ho, Wo, oy, ox)
    x = torch.randn(B, Cin, H, W, generator=g)
    x = ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)  # activations reach this op TF32-rounded by their producer
    w = torch.randn(Cin, Cout, 3, 3, generator=g) / Cin ** 0.5
    b = torch.randn(Cout, generator=g)
    R = torch.randn(B, Cout, Ho, Wo, generator=g)
    xr, wr, br = (leaf(t, "cpu", torch.float64) for t in (x, w, b))
    yr = F.conv_transpose2d(xr, wr, br, stride=2)[:, :, oy:oy + Ho, ox:ox + Wo]
    (yr * R.double()).sum().backward()
    xo, wo, bo = leaf(nhwc(x)), leaf(w), leaf(b)
    yo = ops.convT3x3s2_tc(xo, wo, bo, Ho, Wo, oy, ox)
    (yo * nhwc(R).to(DEV)).sum().backward()

    def close(got, ref, what):
        err = (got.detach().cpu().double() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
        assert err < 2e-3, "%s: max err / max |ref| = %g" % (what, err)

    close(nchw(yo), yr, "y")
    close(nchw(xo.grad), xr.grad, "dx")
    close(wo.grad, wr.grad, "dw")
    close(bo.grad, br.grad, "db")


@pytest.mark.parametrize("C,H,W,scaled", [(8, 32, 32, False), (16, 101, 101, True), (3, 25, 13, False), (64, 12, 12, True)])
def test_maxpool2_floor_and_ties(C, H, W, scaled):
    """nn.MaxPool2d(2) floor mode incl. ATen's first-max tie-break on ReLU zeros (unet_p.py:139, unet_p_res.py:247-248)."""
    from pu_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    B = 2
    x = F.relu(torch.randn(B, C, H, W, generator=g))  # ~half exact zeros => many 4-way ties
    s = (torch.rand(B, C, generator=g) > 0.25).float() / 0.75 if scaled else None
    R = torch.randn(B, C, H // 2, W // 2, generator=g)
    xr = leaf(x, "cpu", torch.float64)
    yr = F.max_pool2d(xr, 2)
    if scaled:
        yr = yr * s.double().view(B, C, 1, 1)
    (yr * R.double()).sum().backward()
    xo = leaf(nhwc(x))
    yo = ops.maxpool2(xo, s.to(DEV) if scaled else None)
    (yo * nhwc(R).to(DEV)).sum().backward()
    assert torch.equal(nchw(yo).cpu().double(), yr.detach()) or rel_err(nchw(yo), yr)[0] < 1e-7
    check(nchw(xo.grad), xr.grad, 1e-7, what="dx (tie-break)")


def test_bilinear2x():
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True): unet_p.py:153."""
    from pu_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 6, 9, 7, generator=g)
    R = torch.randn(2, 6, 18, 14, generator=g)
    xr = leaf(x, "cpu", torch.float64)
    yr = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True)
    (yr * R.double()).sum().backward()
    xo = leaf(nhwc(x))
    yo = ops.bilinear2x(xo)
    (yo * nhwc(R).to(DEV)).sum().backward()
    check(nchw(yo), yr, what="y")
    check(nchw(xo.grad), xr.grad, what="dx")


def test_concat_scale():
    from pu_b200 import ops
    g = torch.Generator().manual_seed(4)
    B, C0, C1, H, W = 2, 8, 4, 12, 12
    x0 = torch.randn(B, C0, 13, 13, generator=g)
    x1 = torch.randn(B, C1, H, W, generator=g)
    s = (torch.rand(B, C0 + C1, generator=g) > 0.5).float() * 2
    R = torch.randn(B, C0 + C1, H, W, generator=g)
    x0r, x1r = leaf(x0, "cpu", torch.float64), leaf(x1, "cpu", torch.float64)
    yr = torch.cat([x0r[:, :, 1:, 1:], x1r], 1) * s.double().view(B, -1, 1, 1)
    (yr * R.double()).sum().backward()
    x0o, x1o = leaf(nhwc(x0)), leaf(nhwc(x1))
    yo = ops.concat_scale(x0o, x1o, s.to(DEV), H, W, 1, 1, 0, 0)
    (yo * nhwc(R).to(DEV)).sum().backward()
    check(nchw(yo), yr, 1e-7)
    check(nchw(x0o.grad), x0r.grad, 1e-7)
    check(nchw(x1o.grad), x1r.grad, 1e-7)


@pytest.mark.parametrize("train,relu,C", [(True, True, 8), (True, False, 16), (False, True, 6), (False, False, 64)])
def test_batchnorm(train, relu, C):
    """nn.BatchNorm2d (+ReLU): unet_p.py:106-107, unet_p_res.py:151,175."""
    from pu_b200 import ops
    g = torch.Generator().manual_seed(C)
    B, H, W = 3, 10, 9
    x = torch.randn(B, C, H, W, generator=g) * 2 + 0.5
    ga, be = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    R = torch.randn(B, C, H, W, generator=g)
    xr, gr, br = (leaf(t, "cpu", torch.float64) for t in (x, ga, be))
    rmr, rvr = rm.double().clone(), rv.double().clone()
    yr = F.batch_norm(xr, rmr, rvr, gr, br, train, 0.1, 1e-5)
    if relu:
        yr = F.relu(yr)
    (yr * R.double()).sum().backward()
    xo, go, bo = leaf(nhwc(x)), leaf(ga), leaf(be)
    rmo, rvo = rm.to(DEV).clone(), rv.to(DEV).clone()
    yo, mean, invstd = ops.batchnorm(xo, go, bo, rmo, rvo, train, 0.1, 1e-5, relu)
    if train:
        ops.bn_update_running(mean.detach(), invstd.detach(), rmo, rvo, 0.1, 1e-5, B * H * W)
    (yo * nhwc(R).to(DEV)).sum().backward()
    check(nchw(yo), yr, 5 * TOL, "y")
    check(nchw(xo.grad), xr.grad, 1e-4, "dx")
    check(go.grad, gr.grad, 1e-4, "dgamma")
    check(bo.grad, br.grad, 1e-4, "dbeta")
    check(rmo, rmr, 1e-5, "running_mean")
    check(rvo, rvr, 1e-4, "running_var")


def test_layout_roundtrip():
    from pu_b200 import ops
    x = torch.randn(2, 5, 9, 11, device=DEV)
    y = ops.nchw_to_nhwc(x)
    assert torch.equal(y, x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(ops.nhwc_to_nchw(y), x)


@pytest.mark.parametrize("N,B", [(32, 1), (101, 1), (128, 4), (21, 3)])
def test_plastic_head(N, B):
    """unet_p.py:70-79 and its autograd (closed forms of SURVEY.md §8a row 8)."""
    from pu_b200 import ops
    g = torch.Generator().manual_seed(N + B)
    X = torch.randn(B * N, N, generator=g)
    w, al, hb = 0.1 * torch.randn(N, N, generator=g), 0.1 * torch.rand(N, N, generator=g), 0.5 * torch.randn(N, N, generator=g)
    R = torch.randn(B * N, N, generator=g)
    Xr, wr, ar, hr = (leaf(t, "cpu", torch.float64) for t in (X, w, al, hb))
    Sr = torch.sigmoid(Xr.mm(wr + torch.mul(ar, hr)))
    (Sr * R.double()).sum().backward()
    Xo, wo, ao, ho = leaf(X), leaf(w), leaf(al), leaf(hb)
    So, weff = ops.plastic_head(Xo, wo, ao, ho)
    (So * R.to(DEV)).sum().backward()
    check(So, Sr, what="S")
    check(weff, (wr + ar * hr), what="weff")
    check(Xo.grad, Xr.grad, what="gX")
    check(wo.grad, wr.grad, 5 * TOL, what="gw")
    check(ao.grad, ar.grad, 5 * TOL, what="galpha")
    check(ho.grad, hr.grad, 5 * TOL, what="ghebb")


@pytest.mark.parametrize("N,B,scale", [(32, 1, 1.0), (101, 3, 1.0), (128, 64, 1.0), (21, 3, 1.0), (64, 8, 0.25), (128, 5, 1.0)])
def test_plastic_head_bce_fused(N, B, scale):
    """The training-step form of the head (TrainStep, TF32 mode): head (unet_p.py:70-79) + nn.BCELoss (train.py:100-103) +
    the backward of both in one launch, 3xTF32 tensor-core GEMMs — against the float64 formula.  Ragged N (101, 21: scalar
    path, partial MMA tiles), batch tails (rows beyond B*N in the last 64-row tile), a non-unit loss seed."""
    from pu_b200 import ops
    g = torch.Generator().manual_seed(7 * N + B)
    X = torch.randn(B * N, N, generator=g)
    w, al, hb = 0.1 * torch.randn(N, N, generator=g), 0.1 * torch.rand(N, N, generator=g), 0.5 * torch.randn(N, N, generator=g)
    T = (torch.rand(B, N, N, generator=g) > 0.6).float()
    Xr, wr, ar, hr = (leaf(t, "cpu", torch.float64) for t in (X, w, al, hb))
    Sr = torch.sigmoid(Xr.mm(wr + torch.mul(ar, hr)))
    loss_r = F.binary_cross_entropy(Sr.view(B, N, N), T.double())
    (loss_r * scale).backward()
    Xo, wo, ao, ho = leaf(X), leaf(w), leaf(al), leaf(hb)
    So, loss_o, gA, gX = ops.plastic_head_bce(Xo, wo, ao, ho, T.to(DEV), True)
    seed = torch.full((1,), scale, device=DEV)
    ops.UNIT_GRAD = seed if scale == 1.0 else None
    try:
        loss_o.backward(seed)
    finally:
        ops.UNIT_GRAD = None
    check(So, Sr, what="S")  # the logits are 3xTF32: fp32 level
    assert abs(float(loss_o.detach()) - float(loss_r)) < 1e-5 * abs(float(loss_r)), (float(loss_o.detach()), float(loss_r))
    # gX is a plain TF32 product (it feeds the TF32 data-gradient convs): bound = TF32 rounding of both operands
    check(Xo.grad, Xr.grad, 2e-3, what="gX")
    assert rel_err(Xo.grad, Xr.grad)[1] < 1e-3
    wtol = 5 * TOL if ops.HEAD_WGRAD_TERMS == 3 else 2e-3
    check(wo.grad, wr.grad, wtol, what="gw")
    check(ao.grad, ar.grad, wtol, what="galpha")
    check(ho.grad, hr.grad, wtol, what="ghebb")
    # the error-compensated parameter gradients (terms = 3) are at fp32 level
    gA_r = (Sr.detach() - T.double().view(B * N, N)) / (B * N * N) * scale
    saved = ops.HEAD_WGRAD_TERMS
    ops.HEAD_WGRAD_TERMS = 3
    try:
        gw3, _, _ = ops.plastic_head_wgrad(Xo.detach(), gA_r.float().to(DEV).contiguous(), ao.detach(), ho.detach(), False, False)
    finally:
        ops.HEAD_WGRAD_TERMS = saved
    check(gw3, Xr.detach().t().mm(gA_r), 5 * TOL, what="gw (3xTF32)")
    # Weff computed beforehand (TrainStep does that off the critical path) gives bit-identical results
    with torch.no_grad():
        wf = ops.head_weff(wo.detach(), ao.detach(), ho.detach())
        S3, loss3, gA3, gX3 = ops.plastic_head_bce(Xo.detach(), wo.detach(), ao.detach(), ho.detach(), T.to(DEV), True, wf)
    assert torch.equal(S3, So) and torch.equal(loss3, loss_o.detach()) and torch.equal(gA3, gA.detach()) and torch.equal(gX3, gX.detach())
    # the strict-fp32 head + pu_bce_fwd_bwd (the path the fp32 mode keeps) agrees to fp32 level
    X2, w2, a2, h2 = leaf(X), leaf(w), leaf(al), leaf(hb)
    S2, _ = ops.plastic_head(X2, w2, a2, h2)
    assert float((S2 - So).abs().max()) < 2e-6


@pytest.mark.parametrize("rule", ["hebb", "oja"])
@pytest.mark.parametrize("N,K", [(32, 1), (101, 1), (64, 5)])
def test_trace_update(rule, N, K):
    """unet_p.py:81-84 (row-0 semantics) + closed-form backward; K>1 = mean of per-sample updates."""
    from pu_b200 import ops
    import plastic_unet_oracle as orc
    g = torch.Generator().manual_seed(N + K)
    X = torch.randn(K, N, N, generator=g)
    S = torch.sigmoid(torch.randn(K, N, N, generator=g))
    hb = 0.3 * torch.randn(N, N, generator=g)
    eta = torch.tensor([0.07])
    R = torch.randn(N, N, generator=g)
    Xr, Sr, hr, er = (leaf(t, "cpu", torch.float64) for t in (X, S, hb, eta))
    out_r = orc.trace_update_batched(hr, Xr, Sr, er, rule)
    (out_r * R.double()).sum().backward()
    Xo, So, ho, eo = leaf(X.view(K * N, N)), leaf(S.view(K * N, N)), leaf(hb), leaf(eta)
    out_o = ops.trace_update(ho, Xo, So, eo, ops.RULE_HEBB if rule == "hebb" else ops.RULE_OJA, N * N, K)
    (out_o * R.to(DEV)).sum().backward()
    check(out_o, out_r, what="hebb'")
    check(ho.grad, hr.grad, what="ghebb")
    check(Xo.grad.view(K, N, N), Xr.grad, what="gpre")
    check(So.grad.view(K, N, N), Sr.grad, what="gpost")
    check(eo.grad, er.grad, 5 * TOL, what="geta")
    # data-parallel split form == fused form
    dq = ops.trace_delta(Xo.detach(), So.detach(), N, N * N, K)
    out_s = ops.trace_apply(ho.detach(), dq, eo.detach(), ops.RULE_HEBB if rule == "hebb" else ops.RULE_OJA, K)
    check(out_s, out_o, 1e-6, what="split form")


def test_bce_and_adam_tail():
    """pu_bce_fwd_bwd == nn.BCELoss + autograd; pu_adam_step == torch.optim.Adam (train.py:66-70,101-112)."""
    from pu_b200 import _lib
    g = torch.Generator().manual_seed(8)
    n = 5000
    s = torch.sigmoid(3 * torch.randn(n, generator=g))
    s[0], s[1] = 0.0, 1.0  # log clamp at -100
    t = (torch.rand(n, generator=g) > 0.5).float()
    sr = leaf(s, "cpu", torch.float64)
    lr_ = F.binary_cross_entropy(sr, t.double())
    lr_.backward()
    so, to = s.to(DEV), t.to(DEV)
    loss, gs = torch.zeros(1, device=DEV), torch.empty(n, device=DEV)
    _lib.call("pu_bce_fwd_bwd", so.data_ptr(), to.data_ptr(), loss.data_ptr(), gs.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    assert abs(float(loss) - float(lr_)) < 1e-4 * float(lr_)
    check(gs[2:], sr.grad[2:], 1e-5, "gS")
    # Adam, 3 steps
    p0 = torch.randn(1000, generator=g)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-2)
    po, m, v = p0.to(DEV).clone(), torch.zeros(1000, device=DEV), torch.zeros(1000, device=DEV)
    step, lr_t = torch.zeros(1, device=DEV), torch.tensor([1e-2], device=DEV)
    for i in range(3):
        gr = torch.randn(1000, generator=g)
        pr.grad = gr.clone()
        opt.step()
        gro = gr.to(DEV)
        _lib.call("pu_adam_step", po.data_ptr(), gro.data_ptr(), m.data_ptr(), v.data_ptr(), step.data_ptr(), lr_t.data_ptr(),
                  0.9, 0.999, 1e-8, 1.0, 1000, torch.cuda.current_stream().cuda_stream)
    check(po, pr, 1e-5, "adam params")
    assert float(step) == 3.0


@pytest.mark.parametrize("C,H,W,mask_in", [(8, 16, 16, False), (8, 128, 128, True), (16, 37, 21, False), (3, 9, 8, False), (3, 9, 8, True),
                                           (64, 16, 16, True)])
def test_pool_skip_accumulates_both_gradients(C, H, W, mask_in):
    """ops.pool_skip(x) -> (maxpool2(x), x): one backward call that sums the pooled-path and the skip-path gradients
    inside the pooling kernel == autograd's separate accumulation (unet_p.py:59-66 dataflow)."""
    from pu_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(2, H, W, C, generator=g).to(DEV)
    if mask_in:
        x = torch.relu(x)  # a ReLU output: windows of four zeros (ties), masked maxima
    R1 = torch.randn(2, H // 2, W // 2, C, generator=g).to(DEV)
    R2 = torch.randn(2, H, W, C, generator=g).to(DEV)
    assert ops.POOL_CODE  # pool_skip routes its backward from the arg-max code of the forward pass, maxpool2 re-reads x
    xa = x.clone().requires_grad_(True)
    p, s = ops.pool_skip(xa, mask_in)
    ((p * R1).sum() + (s * R2).sum()).backward()
    xb = x.clone().requires_grad_(True)
    pb = ops.maxpool2(xb, None, mask_in)
    ((pb * R1).sum() + (xb * R2).sum()).backward()
    assert torch.equal(p, pb) and torch.equal(s, x)
    check(xa.grad, xb.grad, 1e-6, what="dx")
    # only one of the two outputs used
    xc = x.clone().requires_grad_(True)
    p, s = ops.pool_skip(xc, mask_in)
    (s * R2).sum().backward()
    check(xc.grad, R2, 1e-7, what="skip only")


@pytest.mark.parametrize("rule", ["hebb", "oja"])
@pytest.mark.parametrize("N,B", [(128, 64), (101, 3), (32, 1), (64, 8)])
def test_trace_rows_all_on_tensor_cores(rule, N, B):
    """Opt-in rows='all' contraction (K = B*N pairs) on mma.sync with the 3xTF32 split vs the same update in float64:
    fp32-level accuracy (the plain fp32 CUDA-core contraction of the row-0 mode is the comparison point)."""
    from pu_b200 import ops
    g = torch.Generator().manual_seed(N + B)
    X = torch.randn(B * N, N, generator=g)
    S = torch.sigmoid(torch.randn(B * N, N, generator=g))
    hebb = 0.05 * torch.randn(N, N, generator=g)
    eta = torch.tensor([0.03])
    K = B * N
    delta = X.double().t() @ S.double()
    q = (S.double() ** 2).sum(0)
    e = float(eta)
    if rule == "hebb":
        ref = (1 - e) * hebb.double() + e * delta / K
    else:
        ref = hebb.double() * (1 - e * q / K)[None, :] + e * delta / K
    dq = ops.trace_delta_tc(X.to(DEV), S.to(DEV), N, N, K)
    out = ops.trace_apply(hebb.to(DEV), dq, eta.to(DEV), ops.RULE_HEBB if rule == "hebb" else ops.RULE_OJA, K)
    check(dq[:N * N].view(N, N), delta, 2e-6, what="delta (3xTF32)")
    check(dq[N * N:], q, 2e-6, what="q")
    check(out, ref, 2e-6, what="trace")
    # the CUDA-core contraction gives the same payload
    dq2 = ops.trace_delta(X.to(DEV), S.to(DEV), N, N, K)
    check(dq, dq2, 5e-6, what="tensor-core vs CUDA-core payload")


def test_model_trace_rows_all_mode():
    """UNetp(trace_rows='all'): outputs unchanged, trace = mean over ALL rows of all maps of the reference's per-row outer
    products (what unet_p.py:82's bmm holds before [0]); default 'row0' stays the reference."""
    import pu_b200
    import pu_b200.ops as ops_mod
    from conftest import quiet
    torch.manual_seed(1)
    net = quiet(pu_b200.UNetp, 1, 1, torch.device(DEV), rule="hebb", nbf=32, batched=True)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(3, 1, 32, 32, generator=g).to(DEV)
    hebb = (0.05 * torch.randn(32, 32, generator=g)).to(DEV)
    captured = {}
    real = ops_mod.trace_delta_tc

    def spy(pre, post, N, ld, K):
        captured["X"], captured["S"], captured["K"] = pre.clone(), post.clone(), K
        return real(pre, post, N, ld, K)

    with torch.no_grad():
        out0, h0 = net(x, hebb)
        net.trace_rows = "all"
        ops_mod.trace_delta_tc = spy
        try:
            out1, h1 = net(x, hebb)
        finally:
            ops_mod.trace_delta_tc = real
        with pytest.raises(ValueError):
            net.trace_rows = "bogus"
            net(x, hebb)
    assert torch.equal(out0, out1) and not torch.allclose(h0, h1)
    # the reference's bmm over every row, bmm(activin.unsqueeze(2), activout.unsqueeze(1)) -> [N, N, N] per map (unet_p.py:82
    # before the [0]); 'all' = (1 - eta) * hebb + eta * mean over all rows of all maps of those outer products
    eta = float(net.eta)
    X, S = captured["X"].cpu().double().view(3, 32, 32), captured["S"].cpu().double().view(3, 32, 32)
    assert captured["K"] == 3 * 32 and torch.equal(captured["S"].view(3, 32, 32), out1)
    outer = torch.stack([torch.bmm(X[b].unsqueeze(2), S[b].unsqueeze(1)) for b in range(3)])  # [B, N(rows), N, N]
    ref = (1 - eta) * hebb.cpu().double() + eta * outer.mean(dim=(0, 1))
    check(h1, ref, 2e-6, what="rows='all' trace")


@pytest.mark.parametrize("momentum,track", [(None, True), (0.3, True), (0.1, False)])
def test_bn_module_semantics_momentum_none_and_no_running_stats(momentum, track):
    """modules._bn follows nn.BatchNorm2d for momentum=None (cumulative moving average) and track_running_stats=False
    (batch statistics in eval mode too, no buffers)."""
    from pu_b200 import modules as M
    C = 8
    g = torch.Generator().manual_seed(3)
    ref = torch.nn.BatchNorm2d(C, momentum=momentum, track_running_stats=track)
    ours = torch.nn.BatchNorm2d(C, momentum=momentum, track_running_stats=track).to(DEV)
    with torch.no_grad():
        w, b = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
        ref.weight.copy_(w); ref.bias.copy_(b); ours.weight.copy_(w); ours.bias.copy_(b)
    for step in range(3):
        x = torch.randn(2, C, 9, 7, generator=g)
        yr = ref(x)
        yo = M._bn(nhwc(x).to(DEV), ours, False)
        check(nchw(yo), yr, 2e-5, what="train y step %d" % step)
    if track:
        check(ours.running_mean, ref.running_mean, 2e-5, what="running_mean")
        check(ours.running_var, ref.running_var, 2e-5, what="running_var")
        assert int(ours.num_batches_tracked) == int(ref.num_batches_tracked) == 3
    ref.eval(); ours.eval()
    x = torch.randn(2, C, 9, 7, generator=g)
    check(nchw(M._bn(nhwc(x).to(DEV), ours, False)), ref(x), 2e-5, what="eval y")
