"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every declared symbol, argument
validation fails loudly without a GPU, and the Python surface mirrors the reference's `unet` package
(constructor signatures, attributes, state_dict keys, parameter order, RNG consumption)."""
import ctypes
import inspect

import numpy as np
import pytest
import torch

from conftest import Case, quiet
from pu_b200 import UNetp, UNetpCoord, UNetpRes, _lib


def test_library_exports_every_declared_symbol():
    protos = _lib.parse_header()
    assert len(protos) >= 35
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in protos:
        assert hasattr(lib, name), "libpu_b200.so does not export %s declared in include/plastic_unet_b200.h" % name
    assert _lib.version() >= 100


def test_bad_arguments_fail_loudly_without_gpu():
    lib = _lib.load()
    # null pointers / non-positive dims are rejected before any CUDA call
    with pytest.raises(_lib.PuError, match="bad"):
        _lib.call("pu_pack_w3x3", None, None, 8, 8, 0, 0, 8, None)
    with pytest.raises(_lib.PuError):
        _lib.call("pu_maxpool2_fwd", None, None, None, 1, 4, 4, 8, None)
    with pytest.raises(_lib.PuError, match="rule"):
        _lib.call("pu_trace_update_fwd", 16, 16, 16, 4, 1, 16, 7, 16, 4, None)
    assert b"rule" in lib.pu_last_error()
    # the entry points added for the fused training-step head, the pooling arg-max code and the step's edges
    with pytest.raises(_lib.PuError, match="bad argument"):
        _lib.call("pu_plastic_head_bce", None, None, None, None, None, None, None, None, None, None, None, 1, 8, None)
    with pytest.raises(_lib.PuError, match="N=256"):
        _lib.call("pu_plastic_head_bce", 16, 16, 16, 16, None, 16, 16, 16, 16, None, None, 1, 256, None)
    with pytest.raises(_lib.PuError, match="terms"):
        _lib.call("pu_plastic_head_wgrad_tc", 16, 16, 16, 16, 16, None, None, 1, 8, 2, None)
    with pytest.raises(_lib.PuError, match="code is NULL"):
        _lib.call("pu_maxpool2_fwd_code", 16, None, 16, None, 1, 4, 4, 8, None)
    with pytest.raises(_lib.PuError, match="multiples of 4"):
        _lib.call("pu_copy2", 16, 16, 6, 16, 16, 8, None)
    with pytest.raises(_lib.PuError):
        _lib.call("pu_adam_table_step", None, 0, None, None, None, None, None, 0.9, 0.999, 1e-8, 1.0, None)


def test_ops_refuse_cpu_tensors():
    from pu_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.maxpool2(torch.zeros(1, 4, 4, 8), None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.plastic_head(torch.zeros(4, 4), torch.zeros(4, 4), torch.zeros(4, 4), torch.zeros(4, 4))


# the reference's constructor signatures (src/unet/unet_p.py:9, src/unet/unet_p_res.py:10), in order
REF_UNETP = ["n_channels", "n_classes", "device", ("alfa_type", "free"), ("rule", "hebb"), ("nbf", 128), ("batch_norm", False),
             ("bilinear_upsample", False)]
REF_UNETPRES = ["n_channels", "n_classes", "device", ("neurons", 16), ("dropout_ratio", 0.5), ("alfa_type", "free"), ("rule", "hebb"),
                ("nbf", 128), ("batch_norm", False), ("bilinear_upsample", False)]


@pytest.mark.parametrize("cls,ref", [(UNetp, REF_UNETP), (UNetpRes, REF_UNETPRES)])
def test_constructor_signature_is_a_superset_of_the_reference(cls, ref):
    params = list(inspect.signature(cls.__init__).parameters.values())[1:]
    for p, r in zip(params, ref):
        if isinstance(r, tuple):
            assert p.name == r[0] and p.default == r[1]
        else:
            assert p.name == r and p.default is inspect.Parameter.empty
    for p in params[len(ref):]:  # extensions must be optional
        assert p.default is not inspect.Parameter.empty


@pytest.mark.parametrize("name,seed", [("unetp_hebb_n32", 0), ("unetpres_hebb_n21", 7), ("unetp_bn_bilinear_n32", 4),
                                       ("unetpres_bn_n32_train", 12)])
def test_state_dict_and_init_match_reference(name, seed):
    """Same seed => same keys in the same order, same shapes AND same initial values as the reference module
    (the weights files were written by the reference's own constructors)."""
    c = Case(name)
    cls = UNetp if c.kind == "unetp" else UNetpRes
    torch.manual_seed(seed)
    net = quiet(cls, 1, 1, torch.device("cpu"), **c.ctor_kw)
    sd = net.state_dict()
    assert list(sd.keys()) == list(c.weights.files)
    for k in c.weights.files:
        assert tuple(sd[k].shape) == tuple(c.weights[k].shape), k
        assert np.array_equal(sd[k].numpy(), c.weights[k]), k
    assert [n for n, _ in net.named_parameters()][:3] == ["w", "alpha", "eta"]
    # checkpoints interchange
    net.load_state_dict(c.state_dict(), strict=True)


def test_attributes_and_zero_hebb():
    net = quiet(UNetpRes, 1, 1, torch.device("cpu"), neurons=2, nbf=21, rule="oja")
    for attr in ("w", "alpha", "eta", "nbf", "rule", "alfa_type", "torch_dev", "n_classes", "n_channels"):
        assert hasattr(net, attr)
    assert isinstance(net.w, torch.nn.Parameter) and net.w.shape == (21, 21) and net.eta.shape == (1,)
    h = net.initialZeroHebb()
    assert h.shape == (21, 21) and h.dtype == torch.float32 and float(h.abs().sum()) == 0
    assert net.eta.data.cpu().numpy().shape == (1,)  # train.py:146


def test_batch_size_one_contract():
    net = quiet(UNetp, 1, 1, torch.device("cpu"), nbf=32)
    with pytest.raises(ValueError, match="Only batch size: 1 is supported, but was: 2"):  # unet_p.py:55-56
        net(torch.zeros(2, 1, 32, 32), net.initialZeroHebb())
    res = quiet(UNetpRes, 1, 1, torch.device("cpu"), neurons=2, nbf=21)
    with pytest.raises(RuntimeError, match="invalid for input of size"):  # view() failure of unet_p_res.py:116
        res(torch.zeros(2, 1, 21, 21), res.initialZeroHebb())


def test_prints_the_reference_banner(capsys):
    UNetp(1, 1, torch.device("cpu"), rule="oja", nbf=16)
    assert "UNet plastic model with plastic rule [oja] initialized" in capsys.readouterr().out  # unet_p.py:52


def test_dropin_package_exports():
    import unet
    assert unet.UNetp is UNetp and unet.UNetpRes is UNetpRes and unet.UNetpCoord is UNetpCoord


def test_scaled_variants_construct():
    n5 = quiet(UNetp, 1, 1, torch.device("cpu"), nbf=64, depth=5, base=16)
    assert hasattr(n5, "down5") and hasattr(n5, "up5")
    assert n5.inc.conv.conv[0].weight.shape[0] == 16 and n5.down5.mpconv[1].conv[0].weight.shape == (256, 256, 3, 3)
    r5 = quiet(UNetpRes, 1, 1, torch.device("cpu"), neurons=4, nbf=64, depth=5)
    assert hasattr(r5, "conv5") and hasattr(r5, "uconv5") and r5.mid.mconv[0].weight.shape[:2] == (128, 64)
    c = quiet(UNetpCoord, 1, 1, torch.device("cpu"), nbf=32, with_r=True)
    assert c.stem.conv.weight.shape == (8, 4, 1, 1)
