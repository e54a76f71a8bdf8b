"""Drop-in for ``src/predict.py`` (the reference's prediction CLI): same flags, same checkpoint handling, same
``prediction.pkl`` (/root/reference/src/predict.py:11-123) -- but features.pkl is repacked once into a pinned slab
(ingest.py) and scored through the chunked host pipeline of libdfs_b200.so instead of a bs-32 DataLoader loop.

    python deep-fake-audio-classifier_b200/dropin/predict.py --features features.pkl --checkpoint best.pt \\
        --model cnn2d --out prediction.pkl [--device cuda:0] [--no-apply-sigmoid] [--no-swap-tf]

``--batch-size`` / ``--num-workers`` are accepted for command-line compatibility and ignored (there is no
DataLoader).  ``--device`` must be a CUDA device: the scoring path has no CPU fallback.
"""
import argparse
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (_HERE, os.path.dirname(_HERE)):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from ingest import load_feature_table  # noqa: E402
from model import CNN2D  # noqa: E402
from model_cnn1d import CNN1D  # noqa: E402
from scoring import write_predictions  # noqa: E402


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="Generate prediction.pkl from a model checkpoint.")
    parser.add_argument("--features", required=True, help="Path to features.pkl")
    parser.add_argument("--checkpoint", required=True, help="Path to model checkpoint")
    parser.add_argument("--model", required=True, choices=["cnn2d", "cnn1d"])
    parser.add_argument("--out", required=True, help="Output path for prediction.pkl")
    parser.add_argument("--batch-size", type=int, default=32)
    parser.add_argument("--num-workers", type=int, default=2)
    parser.add_argument("--device", default=None, help="cuda[:i] (the native path has no mps / cpu fallback)")
    parser.add_argument("--in-features", type=int, default=180)
    parser.add_argument("--dropout", type=float, default=0.3)
    parser.add_argument("--apply-sigmoid", action="store_true", default=True)
    parser.add_argument("--no-apply-sigmoid", action="store_true", default=False)
    swap_group = parser.add_mutually_exclusive_group()
    swap_group.add_argument("--swap-tf", dest="swap_tf", action="store_true", help="swap time and feature dimensions (T <-> F) (default)")
    swap_group.add_argument("--no-swap-tf", dest="swap_tf", action="store_false", help="disable time/feature swap")
    parser.set_defaults(swap_tf=True)
    return parser.parse_args(argv)


def resolve_device(device_arg):
    if device_arg:
        return device_arg
    if torch.cuda.is_available():
        return "cuda"
    raise RuntimeError("dfs_b200 predict: no CUDA device (the native scoring path has no mps / cpu fallback)")


def load_checkpoint_into(model, path, device):
    """{'model_state': sd, ...} or a bare state dict (src/predict.py:78-85)."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    try:
        ckpt = torch.load(path, map_location=device, weights_only=True)
    except TypeError:
        ckpt = torch.load(path, map_location=device)
    model.load_state_dict(ckpt["model_state"] if isinstance(ckpt, dict) and "model_state" in ckpt else ckpt)
    return model


def score_table(model, table, device, apply_sigmoid=True, swap_tf=True):
    """All rows of a FeatureTable through the model's native scorer; float64 numpy like ``.tolist()`` + DataFrame."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("dfs_b200 predict: --device must be a CUDA device (no CPU fallback)")
    if not swap_tf:
        # the reference would feed (B,180,321) maps to a model built for (B,321,180): shapes do not fit its layers either
        raise ValueError("--no-swap-tf: stored features are [180,321]; the models take (B,321,180)")
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    scorer = model.native(torch.device("cuda", index))
    scores = scorer.score_host(table.view(), int(bool(apply_sigmoid)))
    return np.array(scores.tolist())


def main(argv=None):
    args = parse_args(argv)
    device = resolve_device(args.device)
    apply_sigmoid = False if args.no_apply_sigmoid else args.apply_sigmoid
    kwargs = {"in_features": args.in_features, "dropout": args.dropout}
    model = (CNN1D(**kwargs) if args.model == "cnn1d" else CNN2D(**kwargs)).to(device)
    load_checkpoint_into(model, args.checkpoint, device)
    model.eval()
    table = load_feature_table(args.features)
    predictions = score_table(model, table, device, apply_sigmoid, args.swap_tf)
    if len(predictions) != len(table):
        raise ValueError("Number of predictions does not match number of rows in features.pkl")
    return write_predictions(table.uttids, predictions, args.out)


if __name__ == "__main__":
    main()
