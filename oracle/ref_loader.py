"""Test infrastructure: import the reference's driver modules (train / eval / infer / utils) with a chosen `unet` package.

The reference drivers import h5py, skimage, matplotlib and seaborn, none of which is installed: they are
replaced by MagicMock stubs — the functions exercised here (train.train, eval.eval_net, eval.score_model_best_iou,
infer.inference, utils.iou_metric.*, utils.rle_encode.encode) only touch h5py, through a mocked file object.

    ref = load(ref_src)                      # everything from the reference (source tree or oracle/_ref bytecode)
    ours = load(ref_src, unet_first=DROPIN)  # same drivers, but `from unet import ...` resolves to the drop-in package
"""
import importlib
import sys
import types
from unittest.mock import MagicMock

STUBS = ["h5py", "skimage", "skimage.io", "skimage.transform", "skimage.util", "matplotlib", "matplotlib.pyplot",
         "matplotlib.gridspec", "seaborn"]
NAMES = ["unet", "utils", "eval", "train", "infer"]


_HOOKED = False


def _install_pyb_hook():
    """Let the import system find `X.pyb` (bytecode written by oracle/build_ref.py) the way it finds a sourceless X.pyc."""
    global _HOOKED
    if _HOOKED:
        return
    from importlib import machinery
    details = list(importlib._bootstrap_external._get_supported_file_loaders()) + [(machinery.SourcelessFileLoader, [".pyb"])]
    sys.path_hooks.insert(0, machinery.FileFinder.path_hook(*details))
    sys.path_importer_cache.clear()
    _HOOKED = True


def load(ref_src, unet_first=None):
    _install_pyb_hook()
    for m in STUBS:
        if m not in sys.modules or not isinstance(sys.modules[m], MagicMock):
            try:
                importlib.import_module(m)
            except Exception:
                sys.modules[m] = MagicMock()
    saved_path = list(sys.path)
    saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in NAMES}
    for k in list(saved_mods):
        del sys.modules[k]
    try:
        sys.path[:0] = ([unet_first] if unet_first else []) + [ref_src]
        ns = types.SimpleNamespace()
        for n in NAMES:
            setattr(ns, "eval_" if n == "eval" else n, importlib.import_module(n))
        ns.iou_metric = importlib.import_module("utils.iou_metric")
        ns.rle_encode = importlib.import_module("utils.rle_encode")
        return ns
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k.split(".")[0] in NAMES]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
