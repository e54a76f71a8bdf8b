"""Drop-in for ``src/predict_hybrid.py``: hybrid (2D-CNN + CAE reconstruction error) predictions with the
reference's flags and output (/root/reference/src/predict_hybrid.py:100-196).  Both models score the same pinned
slab; min-max normalisation and the alpha blend run in float64 on the device (``dfs_blend_f64``), bit-identical to
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
from scoring import get_cae_scores, get_supervised_scores, hybrid_blend, normalise_01, write_predictions  # noqa: E402,F401


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


def main(argv=None):
    args = parse_args(argv)
    device = resolve_device(args.device)
    table = load_feature_table(args.test_features)
    print(f"Test set: {len(table)} samples")
    sup_model = load_checkpoint_into(CNN2D(in_features=180, dropout=0.2).to(device), args.sup_checkpoint, device)
    cae_norm = FeatureNormalizer.load(args.cae_normalizer)
    cae_model = load_checkpoint_into(ConvAutoencoder().to(device), args.cae_checkpoint, device)
    print("Running supervised inference...")
    sup_scores = get_supervised_scores(sup_model, table, device, args.batch_size)
    print("Running CAE inference...")
    cae_scores = get_cae_scores(cae_model, table, cae_norm, device, args.batch_size)
    hybrid = hybrid_blend(sup_scores, cae_scores, args.alpha)            # predict_hybrid.py:149-151, float64 on the device
    pred_df = write_predictions(table.uttids, hybrid, args.out)
    print(f"\nSaved hybrid predictions to {args.out}")
    print_distribution("Supervised (sigmoid)", sup_scores)
    print_distribution("CAE MSE", cae_scores)
    print_distribution(f"Hybrid (alpha={args.alpha})", hybrid)
    if args.existing_submission:                                          # predict_hybrid.py:166-192
        with open(args.existing_submission, "rb") as f:
            old = pickle.load(f)
        old_preds = old["predictions"]["predictions"].values if isinstance(old, dict) else old["predictions"].values
        flips = int(((old_preds > 0.5) != (sup_scores > 0.5)).sum())
        print(f"\n  decisions changed vs existing submission: {flips}")
    return pred_df


if __name__ == "__main__":
    main()
