"""Drop-in for ``src/evaluation_cae.py``: CAE reconstruction-error scores, EER under both score conventions and the
reference's report (/root/reference/src/evaluation_cae.py:28-146).  The per-utterance MSE comes from the fused native path
(features ingested once, reconstruction never materialised); the two EERs run on the device."""
import argparse
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (_HERE, os.path.dirname(_HERE)):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402

from dataset_cae import FeatureNormalizer  # noqa: E402
from dfs_b200.metrics import calculate_eer  # noqa: E402
from ingest import load_feature_table, merge_labels  # noqa: E402
from model_cae import ConvAutoencoder  # noqa: E402
from predict import load_checkpoint_into, resolve_device  # noqa: E402
from scoring import get_cae_scores  # noqa: E402


def cae_metrics(all_mse, labels):
    """The metrics dict of evaluate_cae (evaluation_cae.py:58-88) from per-sample MSE scores and labels."""
    all_mse = np.asarray(all_mse, dtype=np.float64)
    labels = np.asarray(labels)
    eer_neg, thr_neg = calculate_eer((-all_mse).tolist(), labels.tolist())     # convention A: -MSE (fakes have MORE error)
    eer_pos, thr_pos = calculate_eer(all_mse.tolist(), labels.tolist())        # convention B: +MSE
    if eer_neg <= eer_pos:
        eer, threshold_mse, convention = eer_neg, -thr_neg, "standard (-MSE: fakes have higher error)"
    else:
        eer, threshold_mse, convention = eer_pos, thr_pos, "inverted (+MSE: fakes have lower error)"
    return {"avg_mse": float(np.mean(all_mse)), "avg_mse_bonafide": float(np.mean(all_mse[labels == 1])),
            "avg_mse_spoof": float(np.mean(all_mse[labels == 0])), "eer": eer, "eer_neg": eer_neg, "eer_pos": eer_pos,
            "threshold_mse": threshold_mse, "convention": convention}


def evaluate_cae(model, table, labels, normalizer, device):
    """(metrics, mse_scores, labels) like the reference's evaluate_cae, for an ingest.FeatureTable."""
    mse = get_cae_scores(model, table, normalizer, device)
    labels = np.asarray(labels, dtype=np.float64)
    return cae_metrics(mse, labels), mse, labels


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Evaluate CAE checkpoint.")
    p.add_argument("--features", required=True, help="Path to features.pkl")
    p.add_argument("--labels", required=True, help="Path to labels.pkl")
    p.add_argument("--checkpoint", required=True, help="Path to CAE checkpoint")
    p.add_argument("--normalizer", required=True, help="Path to normalizer.pt")
    p.add_argument("--batch-size", type=int, default=32)
    p.add_argument("--base-channels", type=int, default=32)
    p.add_argument("--device", default=None)
    return p.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    device = resolve_device(args.device)
    normalizer = FeatureNormalizer.load(args.normalizer)
    print(f"Loaded normalizer from {args.normalizer}")
    model = load_checkpoint_into(ConvAutoencoder(base_channels=args.base_channels).to(device), args.checkpoint, device)
    print(f"Loaded checkpoint from {args.checkpoint}")
    table = load_feature_table(args.features)
    idx, labels = merge_labels(table, pd.read_pickle(args.labels))
    if len(idx) != len(table):
        table = table.take(idx)
    print(f"Evaluating on {len(table)} samples...")
    metrics, mse_scores, labels = evaluate_cae(model, table, labels, normalizer, device)
    print(f"\n{'=' * 60}")
    print("CAE Anomaly Detection Results")
    print(f"  Avg MSE (all):      {metrics['avg_mse']:.6f}")
    print(f"  Avg MSE (bonafide): {metrics['avg_mse_bonafide']:.6f}")
    print(f"  Avg MSE (spoof):    {metrics['avg_mse_spoof']:.6f}")
    print(f"  MSE ratio (spoof/bonafide): {metrics['avg_mse_spoof'] / metrics['avg_mse_bonafide']:.2f}x")
    print(f"  EER (-MSE):         {metrics['eer_neg']:.6f}")
    print(f"  EER (+MSE):         {metrics['eer_pos']:.6f}")
    print(f"  Best EER:           {metrics['eer']:.6f}  ({metrics['convention']})")
    print(f"  Threshold (MSE):    {metrics['threshold_mse']:.6f}")
    print(f"{'=' * 60}")
    return metrics, mse_scores, labels


if __name__ == "__main__":
    main()
