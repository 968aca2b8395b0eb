"""Inference tail on the GPU: threshold sweep + IoU metrics + mask threshold + run-length encoding.

Mirrors the reference functions the drivers call after the forward pass (paths relative to the reference's src/):

    score_best_iou(preds, labels)        eval.py:48-62 (31 thresholds) + utils/iou_metric.py:26-87 (iou_metric_batch)
    fast_iou_metric(preds, labels)       utils/iou_metric.py:22-24 exactly as eval.py:100 calls it (flattened arrays)
    threshold_mask / rle_encode_batch    infer.py:81,88,99 + utils/rle_encode.py:6-17

The kernels (csrc/infer_tail.cu) produce only INTEGERS — confusion counts for every (image, threshold) in ONE pass over
the predictions (the reference re-reads them once per threshold), mask bytes and (start, length) run pairs — and the
few flops per image that turn counts into scores are done here in numpy float64, operation for operation as the
reference does them, so the scores are bit-identical and the RLE strings byte-identical (tests/test_infer_tail_gpu.py
against goldens produced by the reference's own functions).  No CPU fallback: tensors must be CUDA tensors.
"""
import numpy as np
import torch

from . import _lib

LABEL_HIST = 0  # np.histogram bins [0, 0.5, 1] (iou_metric.py:34-36)
LABEL_GT0 = 1   # `A > 0` (iou_metric.py:10)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _flat2(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32):
        raise RuntimeError("pu_b200.infer_tail: %s must be a float32 CUDA tensor (there is no CPU fallback)" % what)
    t = t.contiguous()
    return t.view(t.shape[0], -1)


def sweep_thresholds():
    """eval.py:48-50: np.linspace(0.3, 0.7, 31) through the inverse sigmoid (float64)."""
    t = np.linspace(0.3, 0.7, 31)
    return np.log(t / (1 - t))


def sweep_counts(preds, labels, thresholds, label_mode=LABEL_HIST):
    """-> int32 ndarray [B, T, 6] = {c00, c01, c10, c11, pred_ones, npix} for every image and threshold, one pass.
    `thresholds`: float64 values in any order (compared as (double)pred > thr, like numpy against a float64 array)."""
    P, L = _flat2(preds, "preds"), _flat2(labels, "labels")
    if P.shape != L.shape:
        raise RuntimeError("pu_b200.infer_tail: preds %s and labels %s differ in shape" % (tuple(P.shape), tuple(L.shape)))
    thr = np.asarray(thresholds, dtype=np.float64).reshape(-1)
    order = np.argsort(thr, kind="stable")
    B, npix = P.shape
    T = thr.shape[0]
    thr_dev = torch.from_numpy(thr[order].copy()).to(P.device)
    ws = torch.empty(B * 3 * (T + 1), dtype=torch.int32, device=P.device)
    counts = torch.empty((B, T, 6), dtype=torch.int32, device=P.device)
    _lib.call("pu_threshold_counts", P.data_ptr(), L.data_ptr(), thr_dev.data_ptr(), T, int(label_mode), B, npix,
              ws.data_ptr(), counts.data_ptr(), _stream())
    out = counts.cpu().numpy()
    inv = np.empty_like(order)
    inv[order] = np.arange(T)
    return out[:, inv, :]


_PREC_THRESHOLDS = np.arange(0.5, 1.0, 0.05)  # iou_metric.py:69 (and :15)


def iou_metric_from_counts(c):
    """utils/iou_metric.py:26-79 from the confusion counts of ONE (image, threshold): foreground intersection c11 and
    union = area_true1 + area_pred1 - c11 (:34-43, background row/column dropped :46-50, zeros -> 1e-9), then the mean
    over the ten IoU thresholds of tp / (tp + fp + fn) (:56-77) — with one foreground object that is 1.0 or 0.0."""
    c = np.asarray(c, dtype=np.int64)
    inter = np.float64(c[3])
    union = np.float64((c[2] + c[3]) + c[4] - c[3])
    inter = np.float64(1e-9) if inter == 0 else inter
    union = np.float64(1e-9) if union == 0 else union
    iou = inter / union
    prec = []
    for t in _PREC_THRESHOLDS:
        match = bool(iou > t)
        tp, miss = (1, 0) if match else (0, 1)
        prec.append(tp / (tp + miss + miss))
    return np.mean(prec)


def iou_metric_batch_from_counts(counts_t):
    """utils/iou_metric.py:81-87: float32(mean over the batch).  counts_t: [B, 6] for one threshold."""
    return np.array(np.mean([iou_metric_from_counts(c) for c in counts_t]), dtype=np.float32)


def score_best_iou(preds, labels):
    """eval.py:48-62 on device-resident predictions [B, ...] and labels: -> (threshold_best, iou_best, ious[31])."""
    thresholds = sweep_thresholds()
    counts = sweep_counts(preds, labels, thresholds, LABEL_HIST)
    ious = np.array([iou_metric_batch_from_counts(counts[:, j]) for j in range(thresholds.shape[0])])
    k = np.argmax(ious)
    return thresholds[k], ious[k], ious


def fast_iou_metric(preds, labels):
    """Per image, what eval.py:100 gets from utils/iou_metric.fast_iou_metric on the FLATTENED prediction/target: the loop
    of iou_metric.py:8-18 then runs over pixels, every pixel scores 1.0 iff (target > 0) == (pred > 0.5), so the value is
    the pixel accuracy = agreeing pixels / pixels (np.mean of 0.0/1.0 values is exact).  -> float64 ndarray [B]."""
    c = sweep_counts(preds, labels, [0.5], LABEL_GT0)[:, 0, :].astype(np.int64)
    agree = c[:, 0] + c[:, 3]
    return agree.astype(np.float64) / c[:, 5].astype(np.float64)


def _runs(preds, thr64, want_mask):
    if not (isinstance(preds, torch.Tensor) and preds.is_cuda and preds.dtype == torch.float32 and preds.dim() == 3):
        raise RuntimeError("pu_b200.infer_tail: preds must be a float32 CUDA tensor [B, R, C]")
    P = preds.contiguous()
    B, R, C = P.shape
    cap = R * C + 2
    runs = torch.empty((B, cap), dtype=torch.int32, device=P.device)
    count = torch.empty(B, dtype=torch.int32, device=P.device)
    mask = torch.empty((B, R, C), dtype=torch.uint8, device=P.device) if want_mask else None
    _lib.call("pu_mask_rle", P.data_ptr(), float(thr64), B, R, C, mask.data_ptr() if want_mask else None,
              runs.data_ptr(), cap, count.data_ptr(), _stream())
    cnt = count.cpu().numpy()
    if (cnt < 0).any():
        raise RuntimeError("pu_mask_rle: run buffer too small")
    m = int(cnt.max()) if B else 0
    host = runs[:, :max(m, 1)].contiguous().cpu().numpy()
    return host, cnt, mask


def rle_encode_batch(preds, mask_threshold=0.5, want_mask=False):
    """infer.py:99 for a batch: encode(np.round(pred > mask_threshold)) per image -> list of 'start length ...' strings
    (utils/rle_encode.py:6-17), optionally with the uint8 masks of infer.py:88 (device tensor [B, R, C]).
    mask_threshold is a python float in the reference (numpy then compares in float32): rounded to float32 here."""
    thr = float(np.float32(mask_threshold))
    host, cnt, mask = _runs(preds, thr, want_mask)
    out = [' '.join(str(int(v)) for v in host[b, :cnt[b]]) for b in range(host.shape[0])]
    return (out, mask) if want_mask else out


def threshold_mask(preds, mask_threshold=0.5):
    """infer.py:81,88: (mask > mask_threshold).astype(np.uint8) for a batch, on the device."""
    return rle_encode_batch(preds, mask_threshold, want_mask=True)[1]
