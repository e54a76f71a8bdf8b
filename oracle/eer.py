"""ORACLE (test infrastructure, never the product path): numpy restatement of the metric half
of the hot path -- EER, confusion counts, min-max blend, ensemble mean.

Follows /root/reference/scripts/evaluation.py:7-56 (== src/evaluation.py:12-48),
/root/reference/src/predict_hybrid.py:81-85,149-151 and /root/reference/src/ensemble.py:121
statement by statement.  The only degree of freedom added is ``kind``: the reference calls
``np.argsort`` with the default (unstable, SIMD-dependent) kind; ``kind="stable"`` is the
contract the device radix sort implements.  On tie-free scores both give
the same permutation.  Pinned by tests/golden/eer_cases.npz (outputs of the unmodified
reference function run in the build container).
"""
from __future__ import annotations

import numpy as np

THRESHOLD_EPSILON = 1e-6  # scripts/evaluation.py:31


def eer_details(scores, labels, kind=None):
    """Returns dict(eer, threshold, eer_idx, perm) -- the reference returns only the first two."""
    scores_np = np.array(scores)                                   # :8
    labels_np = np.array(labels)                                   # :9
    order = np.argsort(scores_np) if kind is None else np.argsort(scores_np, kind=kind)   # :11
    sorted_scores = scores_np[order]                               # :12
    sorted_labels = labels_np[order]                               # :13
    n_bonafide = np.sum(labels_np)                                 # :15
    n_spoof = len(labels_np) - n_bonafide                          # :16
    if n_bonafide == 0 or n_spoof == 0:                            # :18-19
        return dict(eer=0.0, threshold=0.0, eer_idx=-1, perm=order)
    far = np.concatenate([[1.0], (n_spoof - np.cumsum(sorted_labels == 0)) / n_spoof])    # :21-23
    frr = np.concatenate([[0.0], np.cumsum(sorted_labels == 1) / n_bonafide])             # :24-26
    eer_idx = int(np.argmin(np.abs(far - frr)))                    # :28 (first minimum)
    eer = (far[eer_idx] + frr[eer_idx]) / 2.0                      # :29
    if eer_idx == 0:                                               # :32-37
        threshold = sorted_scores[0] - THRESHOLD_EPSILON
    elif eer_idx == len(sorted_scores):
        threshold = sorted_scores[-1] + THRESHOLD_EPSILON
    else:
        threshold = sorted_scores[eer_idx - 1]
    return dict(eer=float(eer), threshold=float(threshold), eer_idx=eer_idx, perm=order)


def calculate_eer(scores, labels, kind=None):
    d = eer_details(scores, labels, kind)
    return d["eer"], d["threshold"]


def confusion_at_threshold(scores, labels, threshold):
    """scripts/evaluation.py:42-56."""
    scores_np = np.array(scores)
    labels_np = np.array(labels).astype(int)
    pred = (scores_np > threshold).astype(int)
    tp = int(np.sum((pred == 1) & (labels_np == 1)))
    fn = int(np.sum((pred == 0) & (labels_np == 1)))
    fp = int(np.sum((pred == 1) & (labels_np == 0)))
    tn = int(np.sum((pred == 0) & (labels_np == 0)))
    far = fp / (fp + tn) if (fp + tn) > 0 else 0.0
    frr = fn / (tp + fn) if (tp + fn) > 0 else 0.0
    return tp, fp, tn, fn, float(far), float(frr)


def normalise_01(scores):
    """src/predict_hybrid.py:81-85 (== src/hybrid_ensemble.py:64-69), float64."""
    scores = np.asarray(scores, dtype=np.float64)
    lo, hi = scores.min(), scores.max()
    if hi - lo < 1e-12:
        return np.zeros_like(scores)
    return (scores - lo) / (hi - lo)


def hybrid_blend(sup_scores, cae_scores, alpha=0.80):
    """src/predict_hybrid.py:149-151: alpha*minmax(sup) + (1-alpha)*minmax(cae)."""
    return alpha * normalise_01(sup_scores) + (1 - alpha) * normalise_01(cae_scores)


def ensemble_mean(all_scores):
    """src/ensemble.py:121: np.mean over the model axis (float64 pairwise add then divide)."""
    return np.mean([np.asarray(s, dtype=np.float64) for s in all_scores], axis=0)
