"""Drop-in for the scoring/metric part of ``src/hybrid_ensemble.py``: ``normalise_scores`` and the alpha sweep
(/root/reference/src/hybrid_ensemble.py:64-69,127-151), with the 21 blends and their EERs evaluated on the device
without the score vectors returning to the host between alphas."""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from dfs_b200.metrics import alpha_sweep as _alpha_sweep  # noqa: E402
from dfs_b200.metrics import calculate_eer, normalise_01  # noqa: E402,F401


def normalise_scores(scores):
    """(s - min) / (max - min), zeros when the range is below 1e-12 (hybrid_ensemble.py:64-69)."""
    return normalise_01(scores)


def alpha_sweep(sup_scores, cae_scores, labels, alpha_steps=21, verbose=False):
    """Returns (best_alpha, best_eer, table) where table rows are (alpha, eer, threshold) -- the loop of
    hybrid_ensemble.py:131-151 (``alpha = 1`` is 100 % supervised)."""
    res = _alpha_sweep(sup_scores, cae_scores, labels, alpha_steps=alpha_steps)
    table = list(zip(res["alphas"].tolist(), res["eer"].tolist(), res["threshold"].tolist()))
    if verbose:
        print(f"\n{'alpha':>6s}  {'EER':>10s}")
        print("-" * 20)
        best = 1.0
        for a, e, _ in table:
            marker = " *" if e < best else ""
            best = min(best, e)
            print(f"  {a:.2f}    {e:.6f}{marker}")
    return res["best_alpha"], res["best_eer"], table
