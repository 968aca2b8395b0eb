"""The reference's OWN drivers — train.train (train.py:29-211), eval.eval_net (eval.py:66-103),
eval.score_model_best_iou (eval.py:20-64), infer.inference (infer.py:28-48) — run UNCHANGED, once with the reference's
`unet` package on the CPU and once with the drop-in `unet` package (plastic-unet_b200/unet) on the B200, on the same
arrays; the results must agree.  The drivers come from oracle/_ref (bytecode of the unmodified reference built by
oracle/build_ref.py in the build container — /root/reference does not exist on the GPU box); the four absent
third-party packages are MagicMock stubs."""
import os
import tempfile

import numpy as np
import pytest
import torch

import build_ref
import ref_loader
from conftest import ROOT, Case, quiet, rel_err

pytestmark = pytest.mark.gpu
REF_SRC = os.path.join(ROOT, "oracle", "_ref", "src")
DROPIN = os.path.join(ROOT, "plastic-unet_b200")

if not build_ref.available():
    pytest.skip("oracle/_ref not built (run oracle/build_ref.py in the build container)", allow_module_level=True)


@pytest.fixture(scope="module")
def drivers():
    ref = ref_loader.load(REF_SRC)
    ours = ref_loader.load(REF_SRC, unet_first=DROPIN)
    assert ref.train.UNetp.__module__ == "unet.unet_p"  # the reference's class
    assert ours.train.UNetp.__module__ == "pu_b200.modules"  # `from unet import UNetp` in train.py resolved to the drop-in
    assert ours.eval_.eval_net.__code__.co_code == ref.eval_.eval_net.__code__.co_code  # same driver code on both sides
    return ref, ours


def discs(n, size, seed):
    """Bright discs on a dim background, float64 arrays shaped like load_train_dataset's output (utils/data_set.py:43-44)."""
    g = torch.Generator().manual_seed(seed)
    X = np.zeros((n, 1, size, size))
    y = np.zeros((n, size, size))
    yy, xx = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    for i in range(n):
        cy, cx = (torch.rand(2, generator=g) * (size - 12) + 6).tolist()
        r = float(torch.rand(1, generator=g)) * 5 + 7
        m = ((yy - cy) ** 2 + (xx - cx) ** 2) < r * r
        X[i, 0] = 0.15 * torch.rand(size, size, generator=g).numpy() + 0.8 * m
        y[i] = m
    return X, y


def test_train_py_runs_unchanged_on_the_dropin(drivers):
    ref, ours = drivers
    c = Case("train_unetp_hebb_n32")
    X, y = discs(7, 32, 5)
    nets = {}
    for tag, drv, dev in (("ref", ref, torch.device("cpu")), ("ours", ours, torch.device("cuda"))):
        net = quiet(drv.train.UNetp, 1, 1, dev, rule="hebb", nbf=32)
        net.load_state_dict(c.state_dict())
        with tempfile.TemporaryDirectory() as out_dir:
            params = dict(lr=1e-3, gamma=0.5, steplr=4, stop_time=0, epochs=2, debug=False, device=dev, val_every=1, save_every=1,
                          rollout=100, out_dir=out_dir)
            quiet(drv.train.train, net, X[:5], X[5:], y[:5], y[5:], params)  # train.py:29
            sd_file = torch.load(os.path.join(out_dir, "train_net.pth"), map_location="cpu")  # train.py:203
        nets[tag] = (net, sd_file)
    sd_r, sd_o = nets["ref"][0].state_dict(), nets["ours"][0].state_dict()
    assert list(sd_r.keys()) == list(sd_o.keys()) == list(nets["ours"][1].keys())
    for k in sd_r:
        d = sd_o[k].cpu().double() - sd_r[k].double()
        upd = sd_r[k].double() - c.state_dict()[k].double()
        assert float(d.norm()) <= 2e-3 * float(upd.norm()) + 1e-7, k  # error of the 10-step UPDATE, per tensor
    # a checkpoint written through the drop-in loads into the reference module and vice versa (train.py:293-296)
    nets["ref"][0].load_state_dict(nets["ours"][1])
    nets["ours"][0].load_state_dict(nets["ref"][1])


def test_eval_and_infer_run_unchanged_on_the_dropin(drivers):
    ref, ours = drivers
    c = Case("margin_unetp_oja_n32")  # trained weights: decisions have margin
    X, y = discs(6, 32, 9)
    res = {}
    for tag, drv, dev in (("ref", ref, torch.device("cpu")), ("ours", ours, torch.device("cuda"))):
        net = quiet(drv.train.UNetp, 1, 1, dev, rule="oja", nbf=32)
        net.load_state_dict(c.state_dict())
        acc, loss = drv.eval_.eval_net(net, X_val=X, y_val=y, device=dev, criterion=torch.nn.BCELoss())  # eval.py:66
        thr, iou = drv.eval_.score_model_best_iou(net, X, y, dev)  # eval.py:20
        masks = [drv.infer.inference(net, X[i], dev) for i in range(len(X))]  # infer.py:28
        res[tag] = (acc, loss, thr, iou, masks)
    (acc_r, loss_r, thr_r, iou_r, m_r), (acc_o, loss_o, thr_o, iou_o, m_o) = res["ref"], res["ours"]
    print("\n[eval.py on the drop-in] acc %.6f / %.6f  loss %.6f / %.6f  best thr %.4f / %.4f  iou %.4f / %.4f"
          % (acc_o, acc_r, loss_o, loss_r, thr_o, thr_r, iou_o, iou_r))
    assert abs(loss_o - loss_r) < 1e-5 * max(1.0, abs(loss_r))
    assert acc_o == acc_r and thr_o == thr_r and iou_o == iou_r
    for a, b in zip(m_o, m_r):
        assert a.shape == b.shape == (32, 32) and a.dtype == b.dtype
        assert float(np.abs(a - b).max()) < 1e-5
        assert np.array_equal(a > 0.5, b > 0.5)  # infer.py:81


def test_gpu_inference_tail_matches_the_reference_drivers(drivers):
    """InferStep (batched forward) + pu_b200.infer_tail == infer.predict's per-image loop + RLE (infer.py:73-99) and
    eval.score_model_best_iou (eval.py:48-62), byte for byte."""
    ref, _ = drivers
    import pu_b200
    from pu_b200 import infer_tail as it
    from pu_b200.trainer import InferStep
    c = Case("margin_unetp_oja_n32")
    X, y = discs(8, 32, 13)
    cpu = torch.device("cpu")
    net_r = quiet(ref.train.UNetp, 1, 1, cpu, rule="oja", nbf=32)
    net_r.load_state_dict(c.state_dict())
    masks_r = [ref.infer.inference(net_r, X[i], cpu) for i in range(len(X))]
    rle_r = [ref.rle_encode.encode(np.round(m > 0.5)) for m in masks_r]  # infer.py:99
    thr_r, iou_r = ref.eval_.score_model_best_iou(net_r, X, y, cpu)
    net = quiet(pu_b200.UNetp, 1, 1, torch.device("cuda"), rule="oja", nbf=32, batched=True)
    net.load_state_dict(c.state_dict())
    step = InferStep(net, len(X), 32).capture()
    out = step.step(torch.from_numpy(X.astype(np.float32)).cuda())
    assert it.rle_encode_batch(out, 0.5) == rle_r
    assert step.predict_rle(torch.from_numpy(X.astype(np.float32)).cuda(), 0.5) == rle_r  # the batched infer.predict
    thr, iou, _ = it.score_best_iou(out, torch.from_numpy(y.astype(np.float32)).cuda())
    assert thr == thr_r and iou == iou_r
