"""Frame-level detection metrics with the reference's names and results (``utils/metric_utils.py:4-37``).

``calculate_metrics`` sweeps the 21 thresholds ``np.arange(0, 1.05, 0.05)`` with a strict ``>``, counts a hit where
``2 * target - prediction == 1``, defines recall / precision as 1 when their denominator is empty and integrates
precision over recall with the rectangle rule -- evaluated here for all thresholds at once.  NumPy arrays or torch
tensors (any device) are accepted; the arithmetic is NumPy float64 as in the reference, so the numbers are identical
(tests/test_host_logic.py pins them against the reference's own outputs in tests/golden/metrics_reference.npz).
"""
import numpy as np

THRESHOLDS = np.arange(0.00, 1.05, 0.05)


def _np(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def calculate_metrics(output, target):
    """``(recall[21], precision[21], AP)`` of probabilities ``output`` against ``target`` (first common frames)."""
    output, target = _np(output), _np(target)
    n = min(output.shape[0], target.shape[0])
    o, t = output[:n], target[:n]
    shape = (-1,) + (1,) * o.ndim
    pred = np.where(o[None] > THRESHOLDS.reshape(shape), 1, 0)            # (21, frames, classes)
    axes = tuple(range(1, pred.ndim))
    tp = ((2 * t[None] - pred) == 1).sum(axis=axes).astype(np.float64)
    n_pos = pred.sum(axis=axes).astype(np.float64)
    n_gt = float(t.sum())
    recall = tp / n_gt if n_gt > 0 else np.ones_like(tp)
    precision = np.where(n_pos > 0, tp / np.where(n_pos > 0, n_pos, 1.0), 1.0)
    ap = np.sum(precision[:-1] * (recall[:-1] - recall[1:]))
    return recall, precision, ap


def f_score(recll, precision, beta=1):
    return (1 + beta ** 2) * recll * precision / (beta ** 2 * recll + precision + 1e-9)
