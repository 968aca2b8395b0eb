"""Inference tail: threshold sweep + IoU + mask threshold + RLE.

CPU part: the oracle restatement (oracle/infer_tail_oracle.py) against tests/golden/infer_tail.npz, which
oracle/make_golden.py produced by calling the REFERENCE's own eval.score_model_best_iou, utils.iou_metric and
utils.rle_encode functions; and the product's host-side score arithmetic fed with numpy-computed counts.
GPU part (-m gpu): the CUDA kernels through the public pu_b200.infer_tail API — scores bit-identical, RLE strings
byte-identical, masks equal, including empty / full / ragged / checkerboard images."""
import os

import numpy as np
import pytest
import torch

import infer_tail_oracle as ito
from conftest import GOLDEN

Z = np.load(os.path.join(GOLDEN, "infer_tail.npz"), allow_pickle=False)
TAGS = ["a", "b"]
MASK_THRESHOLDS = [0.5, 0.35, 0.62]


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_matches_reference_golden(tag):
    preds, labels = Z[tag + "_preds"], Z[tag + "_labels"].astype(np.float64)
    thr, iou, ious = ito.sweep_best_iou(labels, [p for p in preds])
    assert np.array_equal(ious, Z[tag + "_ious"]) and ious.dtype == np.float32
    assert thr == float(Z[tag + "_thr_best"]) and iou == Z[tag + "_iou_best"]
    fast = np.array([ito.fast_iou_metric(labels[b].astype(np.float32).reshape(-1), preds[b].reshape(-1)) for b in range(preds.shape[0])])
    assert np.array_equal(fast, Z[tag + "_fast_iou"])
    for mt in MASK_THRESHOLDS:
        enc = [ito.rle_encode(np.round(preds[b] > mt)) for b in range(preds.shape[0])]
        assert enc == [str(s) for s in Z["%s_rle_%g" % (tag, mt)]]
        assert np.array_equal(np.stack([ito.threshold_mask(p, mt) for p in preds]), Z["%s_mask_%g" % (tag, mt)])


def _numpy_counts(preds, labels, thresholds, mode):
    B = preds.shape[0]
    out = np.zeros((B, len(thresholds), 6), dtype=np.int32)
    for b in range(B):
        l = labels[b].reshape(-1)
        cls = np.where((l >= 0) & (l < 0.5), 0, np.where((l >= 0.5) & (l <= 1), 1, 2)) if mode == 0 else (l > 0).astype(int)
        for j, t in enumerate(thresholds):
            p = preds[b].reshape(-1).astype(np.float64) > t
            out[b, j] = [np.sum((cls == 0) & ~p), np.sum((cls == 0) & p), np.sum((cls == 1) & ~p), np.sum((cls == 1) & p), np.sum(p), p.size]
    return out


@pytest.mark.parametrize("tag", TAGS)
def test_host_score_arithmetic_from_counts(tag):
    """pu_b200.infer_tail's float arithmetic (counts -> IoU scores) reproduces the reference bit for bit."""
    from pu_b200 import infer_tail as it
    preds, labels = Z[tag + "_preds"], Z[tag + "_labels"]
    thr = it.sweep_thresholds()
    assert np.array_equal(thr, ito.sweep_thresholds())
    counts = _numpy_counts(preds, labels, thr, 0)
    ious = np.array([it.iou_metric_batch_from_counts(counts[:, j]) for j in range(len(thr))])
    assert np.array_equal(ious, Z[tag + "_ious"]) and ious.dtype == np.float32


def test_c_abi_rejects_bad_arguments_without_a_gpu():
    from pu_b200 import _lib
    lib = _lib.load()
    assert lib.pu_threshold_counts(None, None, None, 4, 0, 1, 16, None, None, None) != 0
    assert b"null" in lib.pu_last_error()
    assert lib.pu_mask_rle(None, 0.5, 1, 4, 4, None, None, 18, None, None) != 0
    assert lib.pu_mask_rle_smem_bytes(128, 128) == 4 * ((128 * 128 + 32) // 32 + 1 + 33)


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_sweep_and_iou_bit_exact(tag):
    from pu_b200 import infer_tail as it
    dev = torch.device("cuda")
    preds = torch.from_numpy(Z[tag + "_preds"]).to(dev)
    labels = torch.from_numpy(Z[tag + "_labels"]).to(dev)
    thr, iou, ious = it.score_best_iou(preds, labels)
    assert np.array_equal(ious, Z[tag + "_ious"]) and ious.dtype == np.float32  # eval.py:52
    assert thr == float(Z[tag + "_thr_best"]) and iou == Z[tag + "_iou_best"]  # eval.py:57-59
    assert np.array_equal(it.fast_iou_metric(preds, labels), Z[tag + "_fast_iou"])  # eval.py:100
    # integer counts against numpy, unsorted thresholds, both label modes
    ths = [0.62, -0.3, 0.5, 0.35, 0.999, 0.5000001]
    for mode in (0, 1):
        got = it.sweep_counts(preds, labels, ths, mode)
        assert np.array_equal(got, _numpy_counts(Z[tag + "_preds"], Z[tag + "_labels"], ths, mode))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_mask_and_rle_byte_exact(tag):
    from pu_b200 import infer_tail as it
    preds = torch.from_numpy(Z[tag + "_preds"]).cuda()
    for mt in MASK_THRESHOLDS:
        enc, mask = it.rle_encode_batch(preds, mt, want_mask=True)
        assert enc == [str(s) for s in Z["%s_rle_%g" % (tag, mt)]]  # infer.py:99 + utils/rle_encode.py:6-17
        assert np.array_equal(mask.cpu().numpy(), Z["%s_mask_%g" % (tag, mt)])  # infer.py:88


@pytest.mark.gpu
@pytest.mark.parametrize("R,C,B", [(1, 1, 2), (5, 3, 3), (32, 32, 2), (33, 65, 2), (128, 128, 64), (512, 512, 2), (7, 300, 1)])
def test_gpu_rle_ragged_shapes_vs_oracle(R, C, B):
    """Ragged / non-square / maximal sizes against the pinned oracle; round trip: decoded runs rebuild the mask."""
    from pu_b200 import infer_tail as it
    g = torch.Generator().manual_seed(R * 1000 + C)
    preds = torch.rand(B, R, C, generator=g)
    preds[0] = (preds[0] > 0.5).float()  # dense transitions
    if B > 1:
        preds[1].fill_(1.0)  # one run covering everything
    enc, mask = it.rle_encode_batch(preds.cuda(), 0.5, want_mask=True)
    for b in range(B):
        ref_mask = ito.threshold_mask(preds[b].numpy(), np.float32(0.5))
        assert enc[b] == ito.rle_encode(ref_mask), (R, C, b)
        assert np.array_equal(mask[b].cpu().numpy(), ref_mask)
        v = [int(t) for t in enc[b].split()]
        flat = np.zeros(R * C, dtype=np.uint8)
        for s, n in zip(v[::2], v[1::2]):
            flat[s - 1:s - 1 + n] = 1
        assert np.array_equal(flat.reshape(C, R).T, ref_mask)
