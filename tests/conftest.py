"""pytest configuration: the `gpu` marker, import paths, golden-fixture helpers."""
import ast
import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "plastic-unet_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


class Case:
    """One golden case: inputs, reference outputs/gradients, and the shared weights file."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.kind = str(self.z["meta_kind"])
        self.rule = str(self.z["meta_rule"])
        self.ctor_kw = ast.literal_eval(str(self.z["ctor_kw"]))
        wfile = str(self.z["weights_file"])
        self.weights = np.load(os.path.join(GOLDEN, wfile + ".npz"), allow_pickle=False) if wfile else None

    def t(self, key, device="cpu", dtype=torch.float32):
        return torch.from_numpy(np.asarray(self.z[key])).to(device=device, dtype=dtype)

    def state_dict(self, device="cpu"):
        out = {}
        if self.weights is None:
            # big cases: the reference's seeded initial weights are re-created with the drop-in constructor (same RNG
            # consumption, tests/test_boundary.py) and proven identical through the stored per-key fingerprint
            import pu_b200
            cls = pu_b200.UNetp if self.kind == "unetp" else pu_b200.UNetpRes
            torch.manual_seed(int(self.z["weights_seed"]))
            sd = quiet(cls, 1, 1, torch.device("cpu"), **self.ctor_kw).state_dict()
            assert [str(k) for k in self.z["w_keys"]] == list(sd.keys())
            for k, s, l2 in zip(sd.keys(), self.z["w_sum"], self.z["w_l2"]):
                v = sd[k].double()
                assert float(v.sum()) == float(s) and float(v.norm()) == float(l2), "seeded weights differ from the reference's: " + k
                out[k] = sd[k].detach().clone().to(device)
            return out
        for k in self.weights.files:
            out[k] = torch.from_numpy(self.weights[k]).to(device)
        return out

    def masks(self, device="cpu"):
        keys = sorted(k for k in self.z.files if k.startswith("mask::"))
        return [torch.from_numpy(self.z[k]).to(device) for k in keys]

    def body_kw(self):
        kw = {}
        if self.kind == "unetpres":
            kw["dropout_ratio"] = self.ctor_kw.get("dropout_ratio", 0.5)
            kw["batch_norm"] = self.ctor_kw.get("batch_norm", False)
            kw["training"] = bool(int(self.z["meta_train"])) if "meta_train" in self.z.files else True
        elif self.kind == "unetp":
            kw["batch_norm"] = self.ctor_kw.get("batch_norm", False)
            kw["bilinear"] = self.ctor_kw.get("bilinear_upsample", False)
            kw["training"] = bool(int(self.z["meta_train"])) if "meta_train" in self.z.files else True
        return kw


FWD_CASES = ["unetp_hebb_n32", "unetp_oja_n32", "unetp_crop_n32_in37", "unetp_bn_bilinear_n32", "unetpres_hebb_n21",
             "unetpres_oja_n21_dropout", "unetpres_bn_n21_eval", "unetpres_bn_n32_train",
             # BASELINE sizes: configs[1] (UNetp @128) and configs[3] (unet_p_res_script.py variant, neurons 8 @101)
             "unetp_oja_n128", "unetpres8_hebb_n101_dropout", "unetpres8_oja_n101_eval"]
BIG_CASES = ["unetp_oja_n128", "unetpres8_hebb_n101_dropout", "unetpres8_oja_n101_eval"]
MARGIN_CASES = ["margin_unetp_oja_n32"]
TRAIN_CASES = ["train_unetp_hebb_n32", "train_unetpres_oja_n21"]


def rel_err(a, b):
    """(max-abs error / max|b|, L2 error / ||b||) in float64."""
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return (float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30),
            float((a - b).norm()) / max(float(b.norm()), 1e-30))
