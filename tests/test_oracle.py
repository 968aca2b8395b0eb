"""CPU tests of the oracle (oracle/plastic_unet_oracle.py) against the golden vectors generated from the
real reference (oracle/make_golden.py), plus the pinning of the un-referenced extensions by composition."""
import os

import numpy as np
import pytest
import torch

import plastic_unet_oracle as orc
from conftest import FWD_CASES, TRAIN_CASES, Case, quiet, rel_err

# the golden vectors were produced by torch CPU fp32 on the build container; another host CPU may pick other
# mkldnn kernels, so the oracle-vs-golden check allows a few fp32 ulps of re-association, not bit equality
TOL = 2e-5


def _run_oracle(c, dtype=torch.float32):
    sd = orc.leaf_state(c.state_dict(), dtype=dtype)
    x = c.t("x", dtype=dtype).requires_grad_(True)
    hebb = c.t("hebb", dtype=dtype).requires_grad_(True)
    kw = c.body_kw()
    masks = c.masks()
    if masks:
        kw["masks"] = [m.to(dtype) for m in masks]
    activ, out, hebb_new = orc.forward(c.kind, sd, x, hebb, rule=c.rule, alfa_type=c.ctor_kw.get("alfa_type", "free"), **kw)
    loss = orc.bce_mean(out.view(-1), c.t("target", dtype=dtype)) + (hebb_new * c.t("R", dtype=dtype)).sum()
    loss.backward()
    return sd, x, hebb, activ, out, hebb_new, loss


@pytest.mark.parametrize("name", FWD_CASES)
def test_oracle_matches_golden(name):
    torch.set_num_threads(1)
    c = Case(name)
    sd, x, hebb, activ, out, hebb_new, loss = _run_oracle(c)
    assert rel_err(out, c.t("activout"))[0] < TOL
    assert rel_err(activ, c.t("activ"))[0] < TOL
    assert rel_err(hebb_new, c.t("hebb_new"))[0] < TOL
    assert abs(float(loss) - float(c.z["loss"])) < 1e-4
    assert rel_err(x.grad, c.t("grad_x"))[0] < 50 * TOL
    assert rel_err(hebb.grad, c.t("grad_hebb"))[0] < 50 * TOL
    keys = [str(k) for k in c.z["grad_keys"]]
    for k, l2, s in zip(keys, c.z["grad_l2"], c.z["grad_sum"]):
        g = sd[k].grad
        assert g is not None, k
        assert abs(float(g.double().norm()) - l2) <= 1e-3 * max(l2, 1e-12), k
    for k in c.z.files:
        if k.startswith("grad::"):
            assert rel_err(sd[k[6:]].grad, c.t(k))[0] < 50 * TOL, k


@pytest.mark.parametrize("name", ["unetp_hebb_n32", "unetpres_oja_n21_dropout"])
def test_oracle_fp64_bounds_fp32_rounding(name):
    """The oracle's own fp32 rounding (vs the same graph in fp64) is far below the 1e-3 parity tolerance."""
    c = Case(name)
    r32 = _run_oracle(c)
    r64 = _run_oracle(c, torch.float64)
    assert rel_err(r32[4], r64[4])[0] < 1e-5
    assert rel_err(r32[5], r64[5])[0] < 1e-5
    assert rel_err(r32[1].grad, r64[1].grad)[1] < 1e-4


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_oracle_training_loop_matches_golden(name):
    torch.set_num_threads(1)
    c = Case(name)
    sd = orc.leaf_state(c.state_dict())
    kw = {"dropout_ratio": 0.0} if c.kind == "unetpres" else {}
    losses, hebb = orc.train_steps(c.kind, sd, c.t("imgs"), c.t("masks"), c.rule, lr=float(c.z["lr"]), gamma=0.5, steplr=2, **kw)
    assert np.allclose(losses, c.z["losses"], rtol=0, atol=2e-6)
    assert rel_err(hebb, c.t("hebb_final"))[0] < 1e-4
    assert rel_err(sd["w"], c.t("final::w"))[0] < 1e-4


@pytest.mark.parametrize("rule", ["hebb", "oja"])
def test_head_numpy_restatement(rule):
    """numpy fp64 closed forms == the torch restatement of unet_p.py:70-84 (row-0 trace semantics)."""
    g = torch.Generator().manual_seed(5)
    N = 17
    X = torch.randn(N, N, generator=g)
    w, alpha, hebb = 0.1 * torch.randn(N, N, generator=g), 0.1 * torch.rand(N, N, generator=g), 0.05 * torch.randn(N, N, generator=g)
    eta = torch.tensor([0.03])
    activ, out = orc.plastic_head(X, w, alpha, hebb)
    hn = orc.trace_update(hebb, X, out, eta, rule)
    A, S, H = orc.head_numpy(X.numpy(), w.numpy(), alpha.numpy(), hebb.numpy(), 0.03, rule)
    assert np.allclose(activ.numpy(), A, atol=1e-5)
    assert np.allclose(out.numpy(), S, atol=1e-6)
    assert np.allclose(hn.numpy(), H, atol=1e-6)
    # all-rows contraction is NOT the reference
    allrows = (1 - 0.03) * hebb.numpy() + 0.03 * X.numpy().T @ S
    if rule == "hebb":
        assert not np.allclose(H, allrows, atol=1e-4)


def test_batched_extension_is_mean_of_reference_updates():
    """B>1 oracle == per-sample reference forward from the shared trace; trace = mean of per-sample traces."""
    c = Case("unetp_oja_n32")
    sd = c.state_dict()
    g = torch.Generator().manual_seed(9)
    x = torch.rand(3, 1, 32, 32, generator=g)
    hebb = c.t("hebb")
    activ, out, hn = orc.forward("unetp", sd, x, hebb, rule="oja")
    singles = [orc.forward("unetp", sd, x[b:b + 1], hebb, rule="oja") for b in range(3)]
    for b in range(3):
        assert rel_err(out[b], singles[b][1])[0] < 1e-5
    assert rel_err(hn, torch.stack([s[2] for s in singles]).mean(0))[0] < 1e-5
    # closed batched form used by the CUDA kernel: hebb*(1 - eta*mean(s0^2)) + eta*mean outer(x0, s0)
    maps = orc.unetp_body(sd, x).view(3, 32, 32)
    x0, s0, eta = maps[:, 0, :], out[:, 0, :], sd["eta"]
    closed = hebb * (1 - eta * (s0 ** 2).mean(0))[None, :] + eta * torch.einsum("bi,bj->ij", x0, s0) / 3
    assert rel_err(closed, hn)[0] < 1e-5


def test_add_coords_analytic():
    """coord_conv_script.py:69-96: xx varies along width in [-1,1], yy along height; corners are +-1."""
    x = torch.zeros(2, 1, 5, 7)
    y = orc.add_coords(x, with_r=True)
    assert y.shape == (2, 4, 5, 7)
    xx, yy, rr = y[0, 1], y[0, 2], y[0, 3]
    assert float(xx[0, 0]) == -1 and float(xx[0, -1]) == 1 and float(yy[0, 0]) == -1 and float(yy[-1, 0]) == 1
    assert torch.all(xx[:, 1:] > xx[:, :-1]) and torch.all(yy[1:, :] > yy[:-1, :])
    assert torch.allclose(xx[0], xx[3]) and torch.allclose(yy[:, 0], yy[:, 4])
    assert torch.allclose(rr, torch.sqrt((xx - 0.5) ** 2 + (yy - 0.5) ** 2))


def test_pad_101_to_128():
    x = torch.ones(1, 1, 101, 101)
    y = orc.pad_101_to_128(x)
    assert y.shape == (1, 1, 128, 128) and float(y.sum()) == 101 * 101
    assert float(y[0, 0, 12].sum()) == 0 and float(y[0, 0, 13, 13]) == 1 and float(y[0, 0, 113, 113]) == 1 and float(y[0, 0, 114].sum()) == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/unet"), reason="reference only exists in the build container")
def test_oracle_is_bit_exact_with_live_reference():
    """Build container only: import the real reference and compare bit for bit (same check as make_golden.py)."""
    import subprocess
    import sys
    code = r'''
import sys, io, contextlib, torch
sys.path.insert(0, "/root/reference/src"); sys.path.insert(0, "%s")
import plastic_unet_oracle as orc
from unet import UNetpRes
torch.set_num_threads(1); torch.manual_seed(21)
with contextlib.redirect_stdout(io.StringIO()):
    net = UNetpRes(1, 1, torch.device("cpu"), neurons=2, dropout_ratio=0.0, rule="oja", nbf=21)
x = torch.rand(1, 1, 21, 21); hebb = 0.05 * torch.randn(21, 21)
out_r, hebb_r = net(x, hebb)
_, out_o, hebb_o = orc.forward("unetpres", net.state_dict(), x, hebb, rule="oja", dropout_ratio=0.0)
assert torch.equal(out_r, out_o) and torch.equal(hebb_r, hebb_o)
print("BITEXACT")
''' % os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "BITEXACT" in r.stdout, r.stderr[-2000:]
