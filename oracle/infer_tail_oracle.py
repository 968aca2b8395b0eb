"""ORACLE (inference tail) — test infrastructure only; never imported by the product path.

CPU / numpy restatement of the reference's threshold sweep, IoU metrics and run-length encoding
(paths relative to /root/reference/src/):

    fast_iou_metric    utils/iou_metric.py:6-24     (as eval.py:100 calls it: on FLATTENED arrays, so every pixel is a "batch" item)
    iou_metric(_batch) utils/iou_metric.py:26-87    (two-class histogram2d form)
    sweep_best_iou     eval.py:48-62                (31 thresholds, logit-transformed, argmax of the batch IoU)
    rle_encode         utils/rle_encode.py:6-17     (column-major, 1-based "start length" pairs)
    threshold_mask     infer.py:81,88,99            (mask > mask_threshold)

Pinned: oracle/make_golden.py imports the reference's own functions, asserts these restatements return IDENTICAL values
(floats bit-for-bit, strings byte-for-byte) on random and edge-case inputs, and stores tests/golden/infer_tail.npz.
"""
import numpy as np


def fast_iou_metric(y_true, y_pred):
    """utils/iou_metric.py:22-24 -> get_iou_vector(A, B = y_pred > 0.5) (:6-20): loops over A.shape[0].  For each item:
    iou = (|t & p| + 1e-10) / (|t | p| + 1e-10) with t = A > 0, p = B > 0; score = mean over thresholds 0.5..0.95 of (iou > thr)."""
    A, B = np.asarray(y_true), np.asarray(y_pred) > 0.5
    thresholds = np.arange(0.5, 1, 0.05)
    scores = []
    for k in range(A.shape[0]):
        t, p = A[k] > 0, B[k] > 0
        iou = (np.sum(np.logical_and(t, p) > 0) + 1e-10) / (np.sum(np.logical_or(t, p) > 0) + 1e-10)
        scores.append(np.mean([iou > thr for thr in thresholds]))
    return np.mean(scores)


def iou_metric(labels, y_pred):
    """utils/iou_metric.py:26-79 for the two-class case it is written for: 2x2 histogram over bins [0,0.5,1] of (label, pred),
    foreground IoU = c11 / (area_true1 + area_pred1 - c11) with zeros replaced by 1e-9, then the mean over the ten
    thresholds 0.5..0.95 of tp / (tp + fp + fn) — which for a single foreground object is 1.0 if iou > thr else 0.0."""
    labels = np.asarray(labels, dtype=np.float64).ravel()
    pred = np.asarray(y_pred, dtype=np.float64).ravel()
    edges = [0, 0.5, 1]
    inter = np.histogram2d(labels, pred, bins=(edges, edges))[0]
    area_true = np.histogram(labels, bins=edges)[0]
    area_pred = np.histogram(pred, bins=edges)[0]
    union = area_true[:, None] + area_pred[None, :] - inter
    i11, u11 = inter[1, 1], union[1, 1]
    i11 = 1e-9 if i11 == 0 else i11
    u11 = 1e-9 if u11 == 0 else u11
    iou = i11 / u11
    prec = []
    for thr in np.arange(0.5, 1.0, 0.05):
        match = iou > thr
        tp, fp, fn = int(match), int(not match), int(not match)
        prec.append(tp / (tp + fp + fn))
    return np.mean(prec)


def iou_metric_batch(y_true, y_pred):
    """utils/iou_metric.py:81-87."""
    return np.array(np.mean([iou_metric(y_true[b], y_pred[b]) for b in range(y_true.shape[0])]), dtype=np.float32)


def sweep_thresholds():
    """eval.py:48-50: linspace(0.3, 0.7, 31) pushed through the inverse sigmoid."""
    t = np.linspace(0.3, 0.7, 31)
    return np.log(t / (1 - t))


def sweep_best_iou(y_valid, preds_valid):
    """eval.py:52-59 -> (threshold_best, iou_best, ious)."""
    thresholds = sweep_thresholds()
    preds = np.asarray(preds_valid)
    ious = np.array([iou_metric_batch(y_valid, preds > thr) for thr in thresholds])
    k = np.argmax(ious)
    return thresholds[k], ious[k], ious


def rle_encode(im):
    """utils/rle_encode.py:6-17: flatten column-major, pad a zero on both sides, positions (1-based) where the value changes;
    every second position becomes a run length."""
    px = np.asarray(im).flatten(order='F')
    px = np.concatenate([[0], px, [0]])
    change = np.where(px[1:] != px[:-1])[0] + 1
    change[1::2] -= change[::2]
    return ' '.join(str(v) for v in change)


def threshold_mask(pred, mask_threshold):
    """infer.py:81,88: (mask > mask_threshold).astype(np.uint8) with a python-float threshold (numpy compares in float32)."""
    return (np.asarray(pred) > mask_threshold).astype(np.uint8)
