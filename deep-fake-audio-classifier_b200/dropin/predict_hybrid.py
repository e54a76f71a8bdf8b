"""Drop-in for ``src/predict_hybrid.py``: hybrid (2D-CNN + CAE reconstruction error) predictions with the
reference's flags and output (/root/reference/src/predict_hybrid.py:100-207).  Both models score the same pinned
slab through ONE upload (scoring.score_models_once -> dfs_group_score_host); min-max normalisation and the alpha blend run in float64 on the device (``dfs_blend_f64``), bit-identical to
``alpha * normalise_01(sup) + (1 - alpha) * normalise_01(cae)`` in numpy.
"""
import argparse
import os
import pickle
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (_HERE, os.path.dirname(_HERE)):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from dataset_cae import FeatureNormalizer  # noqa: E402
from ingest import load_feature_table  # noqa: E402
from model import CNN2D  # noqa: E402
from model_cae import ConvAutoencoder  # noqa: E402
from predict import load_checkpoint_into, resolve_device  # noqa: E402
from scoring import (get_cae_scores, get_supervised_scores, hybrid_blend, normalise_01, score_models_once,  # noqa: E402,F401
                     write_predictions)


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--sup-checkpoint", required=True)
    p.add_argument("--cae-checkpoint", required=True)
    p.add_argument("--cae-normalizer", required=True)
    p.add_argument("--test-features", required=True, help="Path to final test features.pkl")
    p.add_argument("--existing-submission", default=None, help="Path to existing .pkl submission for comparison")
    p.add_argument("--alpha", type=float, default=0.80, help="Hybrid weight: alpha*supervised + (1-alpha)*cae")
    p.add_argument("--out", default="prediction_hybrid.pkl", help="Output prediction pkl")
    p.add_argument("--batch-size", type=int, default=32)
    p.add_argument("--device", default=None)
    return p.parse_args(argv)


def print_distribution(name, scores):
    print(f"\n  {name}")
    print(f"    min={scores.min():.6f}  max={scores.max():.6f}")
    print(f"    mean={scores.mean():.6f}  median={np.median(scores):.6f}")
    print(f"    std={scores.std():.6f}")
    print(f"    est real (>0.5): {(scores > 0.5).sum()}  est fake (<=0.5): {(scores <= 0.5).sum()}")


def compare_with_existing(pred_df, existing, out=print):
    """The --existing-submission report of predict_hybrid.py:166-207: the new HYBRID predictions against an older
    prediction file (a bare DataFrame or a submission dict holding one under "predictions"), rows aligned by uttid.
    Returns dict(n, diff, agree, disagreements) for tests."""
    import pandas as pd
    old_df = existing["predictions"] if isinstance(existing, dict) and "predictions" in existing else existing
    print_distribution("Existing submission", old_df["predictions"].values)
    both = pd.merge(pred_df, old_df, on="uttid", suffixes=("_new", "_old"))
    new, old = both["predictions_new"].values, both["predictions_old"].values
    diff = new - old
    out("\n  Per-sample diff (new - old):")
    out(f"    mean={diff.mean():.6f}  std={diff.std():.6f}")
    out(f"    min={diff.min():.6f}  max={diff.max():.6f}")
    new_cls, old_cls = (new > 0.5).astype(int), (old > 0.5).astype(int)
    agree = int((new_cls == old_cls).sum())
    out(f"    class agreement: {agree}/{len(both)} ({100 * agree / len(both):.1f}%)")
    where = np.where(new_cls != old_cls)[0]
    shown = where if len(where) <= 20 else where[:10]
    if 0 < len(where) <= 20:
        out("    disagreements:")
    elif len(where) > 20:
        out(f"    {len(where)} disagreements (showing first 10):")
    for i in shown:
        row = both.iloc[i]
        out(f"      {row['uttid']}: old={row['predictions_old']:.4f} new={row['predictions_new']:.4f}")
    return dict(n=len(both), diff=diff, agree=agree, disagreements=[both.iloc[i]["uttid"] for i in where])


def main(argv=None):
    args = parse_args(argv)
    device = resolve_device(args.device)
    table = load_feature_table(args.test_features)
    print(f"Test set: {len(table)} samples")
    sup_model = load_checkpoint_into(CNN2D(in_features=180, dropout=0.2).to(device), args.sup_checkpoint, device)
    cae_norm = FeatureNormalizer.load(args.cae_normalizer)
    cae_model = load_checkpoint_into(ConvAutoencoder().to(device), args.cae_checkpoint, device)
    # predict_hybrid.py:142-145 runs the two models over the table one after the other; here both score every slab of ONE upload
    print("Running supervised + CAE inference (one upload)...")
    sup_scores, cae_scores = score_models_once([sup_model, cae_model], table, device, [None, cae_norm])
    hybrid = hybrid_blend(sup_scores, cae_scores, args.alpha)            # predict_hybrid.py:149-151, float64 on the device
    pred_df = write_predictions(table.uttids, hybrid, args.out)
    print(f"\nSaved hybrid predictions to {args.out}")
    print(f"\n{'=' * 60}")
    print("Distribution Comparison")
    print_distribution("Supervised-only (sigmoid)", sup_scores)
    print_distribution("CAE-only (raw MSE, higher=real)", cae_scores)
    print_distribution(f"Hybrid (alpha={args.alpha})", hybrid)
    if args.existing_submission:                                          # predict_hybrid.py:166-207
        with open(args.existing_submission, "rb") as f:
            compare_with_existing(pred_df, pickle.load(f))
    print(f"\n{'=' * 60}")
    return pred_df


if __name__ == "__main__":
    main()
