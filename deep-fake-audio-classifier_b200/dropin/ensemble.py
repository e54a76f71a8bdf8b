"""Drop-in for ``src/ensemble.py``: ensemble-average the sigmoid scores of several checkpoints on a labelled set and
report the per-model and ensemble EER, with the reference's flags and printout (/root/reference/src/ensemble.py:66-128).
The features are ingested once into a pinned slab, every slab is uploaded ONCE and scored by all checkpoints (dfs_group_score_host), the mean
(``np.mean(all_scores, axis=0)``, :121) and the EERs are computed on the device."""
import argparse
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (_HERE, os.path.dirname(_HERE)):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402

from dfs_b200.metrics import calculate_eer, ensemble_mean  # noqa: E402
from ingest import load_feature_table, merge_labels  # noqa: E402
from model import CNN2D  # noqa: E402
from model_cnn1d import CNN1D  # noqa: E402
from predict import load_checkpoint_into, resolve_device, score_table  # noqa: E402
from scoring import collect_scores, score_models_once  # noqa: E402,F401  (collect_scores: per-batch variant, same name as the reference helper)


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Ensemble-average sigmoid scores from multiple checkpoints.")
    p.add_argument("--checkpoints", nargs="+", required=True, help="arch:path pairs, e.g. cnn2d:checkpoints/final_robust/cnn2d_best.pt")
    p.add_argument("--dev-features", default="data/dev/features.pkl")
    p.add_argument("--dev-labels", default="data/dev/labels.pkl")
    p.add_argument("--batch-size", type=int, default=32)
    p.add_argument("--device", default=None)
    p.add_argument("--in-features", type=int, default=180)
    p.add_argument("--dropout", type=float, default=0.2)
    p.add_argument("--swap-tf", action="store_true", default=True)
    p.add_argument("--no-swap-tf", dest="swap_tf", action="store_false")
    return p.parse_args(argv)


def load_model(arch, checkpoint_path, device, in_features=180, dropout=0.2):
    """Instantiate a model, load checkpoint weights, set to eval mode (ensemble.py:31-49)."""
    kwargs = {"in_features": in_features, "dropout": dropout}
    model = CNN1D(**kwargs) if arch == "cnn1d" else CNN2D(**kwargs)
    load_checkpoint_into(model, checkpoint_path, device)
    model.to(device)
    model.eval()
    return model


def main(argv=None):
    args = parse_args(argv)
    device = resolve_device(args.device)
    table = load_feature_table(args.dev_features)
    idx, labels = merge_labels(table, pd.read_pickle(args.dev_labels))      # make_loader's inner merge on uttid, feature order
    if len(idx) != len(table):
        table = table.take(idx)
    labels = labels.astype(np.float64).tolist()                             # the loader yields float labels (dataset.py:54)
    if not args.swap_tf:
        raise ValueError("--no-swap-tf: stored features are [180,321]; the models take (B,321,180)")
    specs = [spec.split(":", 1) for spec in args.checkpoints]
    models = [load_model(arch, path, device, in_features=args.in_features, dropout=args.dropout) for arch, path in specs]
    # ensemble.py:105-122 scores the dev set once per checkpoint; here all of them score every slab of ONE upload
    all_scores = score_models_once(models, table, device, apply_sigmoid=True)
    results = []
    for (arch, path), scores in zip(specs, all_scores):
        eer, thr = calculate_eer(scores.tolist(), labels)
        results.append((arch, path, eer, thr))
        print(f"  {arch:6s}  {path}")
        print(f"         EER={eer:.6f}  threshold={thr:.6f}")
    ensemble_scores = ensemble_mean(all_scores)                             # float64 mean over models, on the device
    eer, thr = calculate_eer(ensemble_scores.tolist(), labels)
    print(f"\n{'=' * 60}")
    print(f"Ensemble of {len(all_scores)} models")
    print(f"  EER      = {eer:.6f}")
    print(f"  threshold= {thr:.6f}")
    print(f"{'=' * 60}")
    return dict(models=results, ensemble_scores=ensemble_scores, eer=eer, threshold=thr)


if __name__ == "__main__":
    main()
