"""Drop-in for the scoring/metric part of ``src/hybrid_ensemble.py``: ``normalise_scores`` and the alpha sweep
(/root/reference/src/hybrid_ensemble.py:64-69,127-151), with the 21 blends and their EERs evaluated on the device
without the score vectors returning to the host between alphas."""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

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


def parse_args(argv=None):
    import argparse
    p = argparse.ArgumentParser(description="Hybrid supervised + CAE ensemble evaluation.")
    p.add_argument("--sup-checkpoint", required=True, help="Path to supervised model checkpoint")
    p.add_argument("--sup-arch", default="cnn2d", choices=["cnn2d"])
    p.add_argument("--cae-checkpoint", required=True, help="Path to CAE checkpoint")
    p.add_argument("--cae-normalizer", required=True, help="Path to CAE normalizer.pt")
    p.add_argument("--dev-features", default="data/dev/features.pkl")
    p.add_argument("--dev-labels", default="data/dev/labels.pkl")
    p.add_argument("--batch-size", type=int, default=32)
    p.add_argument("--device", default=None)
    p.add_argument("--alpha-steps", type=int, default=21, help="Number of alpha values to sweep (0 to 1)")
    return p.parse_args(argv)


def main(argv=None):
    """The reference's hybrid_ensemble.py CLI (:96-160): supervised-only EER, CAE-only EER (+MSE), alpha sweep, summary."""
    import numpy as np
    import pandas as pd
    from dataset_cae import FeatureNormalizer
    from ingest import load_feature_table, merge_labels
    from model import CNN2D
    from model_cae import ConvAutoencoder
    from predict import load_checkpoint_into, resolve_device
    from scoring import score_models_once

    args = parse_args(argv)
    device = resolve_device(args.device)
    sup_model = load_checkpoint_into(CNN2D(in_features=180, dropout=0.2).to(device), args.sup_checkpoint, device)
    cae_normalizer = FeatureNormalizer.load(args.cae_normalizer)
    cae_model = load_checkpoint_into(ConvAutoencoder().to(device), args.cae_checkpoint, device)
    table = load_feature_table(args.dev_features)
    idx, labels = merge_labels(table, pd.read_pickle(args.dev_labels))
    if len(idx) != len(table):
        table = table.take(idx)
    labels = labels.astype(np.float64)
    sup_scores, cae_scores = score_models_once([sup_model, cae_model], table, device, [None, cae_normalizer])   # one upload (hybrid_ensemble.py:119-130 makes two passes)
    sup_eer, _ = calculate_eer(sup_scores.tolist(), labels.tolist())
    print(f"Supervised-only  EER = {sup_eer:.6f}")
    cae_eer, _ = calculate_eer(cae_scores.tolist(), labels.tolist())
    print(f"CAE-only         EER = {cae_eer:.6f}")
    best_alpha, best_eer, table_rows = alpha_sweep(sup_scores, cae_scores, labels, alpha_steps=args.alpha_steps, verbose=True)
    print(f"\n{'=' * 60}")
    print("Hybrid Ensemble Results")
    print(f"  Supervised-only EER: {sup_eer:.6f}")
    print(f"  CAE-only EER:        {cae_eer:.6f}")
    print(f"  Best hybrid EER:     {best_eer:.6f}  (alpha={best_alpha:.2f})")
    print("  alpha=1.0 means 100% supervised, alpha=0.0 means 100% CAE")
    print(f"{'=' * 60}")
    return dict(sup_eer=sup_eer, cae_eer=cae_eer, best_alpha=best_alpha, best_eer=best_eer, sweep=table_rows)


if __name__ == "__main__":
    main()
