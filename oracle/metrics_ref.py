"""NumPy restatement of the reference's frame metrics, ``utils/metric_utils.py:4-37`` (TEST INFRASTRUCTURE ONLY).

21-threshold recall/precision sweep with a strict ``>`` comparison, rectangle-rule AP, F-beta.  Pinned against the
verbatim reference module by ``tests/golden/make_golden.py`` / ``tests/test_oracle_metrics.py``.
"""
import numpy as np


def recall_precision(pred, target):
    tp = ((2 * target - pred) == 1).sum()
    n_gt, n_pos = target.sum(), pred.sum()
    return (float(tp) / float(n_gt) if n_gt > 0 else 1), (float(tp) / float(n_pos) if n_pos > 0 else 1)


def calculate_metrics(output, target):
    ths = np.arange(0.00, 1.05, 0.05)
    n = min(output.shape[0], target.shape[0])
    t, o = target[:n], output[:n]
    rec, prec = [], []
    for th in ths:
        r, p = recall_precision(np.where(o > th, 1, 0), t)
        rec.append(r)
        prec.append(p)
    rec, prec = np.array(rec), np.array(prec)
    ap = np.sum(prec[:-1] * (rec[:-1] - rec[1:]))
    return rec, prec, ap


def f_score(recall, precision, beta=1):
    return (1 + beta ** 2) * recall * precision / (beta ** 2 * recall + precision + 1e-9)


def create_event_matrix(frames_num, start_times, end_times, frames_per_second=3, classes_num=1):
    """dataset/spectogram/spectograms_dataset.py:205-218."""
    m = np.zeros((frames_num, classes_num))
    for s, e in zip(start_times, end_times):
        m[int(round(s * frames_per_second)): int(round(e * frames_per_second)) + 1] = 1
    return m
