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
        self.weights = np.load(os.path.join(GOLDEN, str(self.z["weights_file"]) + ".npz"), allow_pickle=False)

    def t(self, key, device="cpu", dtype=torch.float32):
        return torch.from_numpy(np.asarray(self.z[key])).to(device=device, dtype=dtype)

    def state_dict(self, device="cpu"):
        out = {}
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
             "unetpres_oja_n21_dropout", "unetpres_bn_n21_eval", "unetpres_bn_n32_train"]
TRAIN_CASES = ["train_unetp_hebb_n32", "train_unetpres_oja_n21"]


def rel_err(a, b):
    """(max-abs error / max|b|, L2 error / ||b||) in float64."""
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return (float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30),
            float((a - b).norm()) / max(float(b.norm()), 1e-30))
