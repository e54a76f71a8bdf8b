"""Drop-in for the scoring helpers of ``src/predict_hybrid.py`` / ``src/ensemble.py`` /
``src/predict.py``: same names, arguments and return types, but the whole feature table is scored
through one pinned slab and the engine's chunked host pipeline instead of a bs-32 DataLoader loop.

    get_supervised_scores(model, features_df, device, batch_size=32)           predict_hybrid.py:52-63
    get_cae_scores(model, features_df, normalizer, device, batch_size=32)      predict_hybrid.py:66-78
    score_models_once(models, features_df, device, normalizers)               both of the above (and ensemble.py:105-122) over ONE upload
    normalise_01(scores)                                                      predict_hybrid.py:81-85
    write_predictions(uttids, scores, path)                                   predict.py:116-122
"""
import os
import sys

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

import torch  # noqa: E402

from dfs_b200.metrics import ensemble_mean, hybrid_blend, normalise_01  # noqa: E402,F401


def features_slab(features_df):
    """features.pkl rows are torch.Tensor[180,321] (README.md:41-48): pack once into a pinned
    [N,180,321] fp32 slab (ingest.pack_features); the engine reads it through strides as the (N,321,180) view the
    reference builds with .transpose(1,2) (predict.py:103-105).  Accepts a DataFrame or an ingest.FeatureTable."""
    from ingest import FeatureTable, pack_features
    if isinstance(features_df, FeatureTable):
        return features_df.slab
    return pack_features(features_df["features"].reset_index(drop=True))


def _device_index(device):
    d = torch.device(device)
    if d.type != "cuda":
        raise RuntimeError("dfs_b200 scoring runs on CUDA only (no CPU fallback)")
    return d.index if d.index is not None else torch.cuda.current_device()


@torch.no_grad()
def get_supervised_scores(model, features_df, device, batch_size=32):
    """np.ndarray float64 of sigmoid scores in row order (batch_size is accepted for signature parity)."""
    model.eval()
    model.to(device)
    slab = features_slab(features_df)                       # (N,180,321)
    scorer = model.native(torch.device("cuda", _device_index(device)))
    scores = scorer.score_host(slab.transpose(1, 2), 1)     # strided (N,321,180) view, H2D inside
    return np.array(scores.tolist())                        # fp32 -> python floats -> float64, like .tolist() + np.array


@torch.no_grad()
def get_cae_scores(model, features_df, normalizer, device, batch_size=32):
    """Per-utterance MSE of the normalised input vs its reconstruction, fused on the device."""
    model.eval()
    model.to(device)
    if normalizer is not None:
        model.set_normalizer(normalizer.mean, normalizer.std)
    slab = features_slab(features_df)
    scorer = model.native(torch.device("cuda", _device_index(device)))
    mse = scorer.score_host(slab.transpose(1, 2), int(normalizer is not None))
    return np.array(mse.tolist())


@torch.no_grad()
def score_models_once(models, features_df, device, normalizers=None, apply_sigmoid=True):
    """Several models over ONE upload of the feature table (dfs_group_score_host): the reference makes one DataLoader pass
    per model (src/ensemble.py:105-122, src/predict_hybrid.py:142-145); here every staged slab crosses PCIe once and all
    models score it.  ``models``: drop-in CNN2D / CNN1D / ConvAutoencoder instances; ``normalizers[i]``: FeatureNormalizer
    (or None) for autoencoder i.  Returns one float64 numpy vector per model, in row order -- sigmoid scores
    (``apply_sigmoid``) for the classifiers, reconstruction MSE for autoencoders -- equal bit for bit to
    get_supervised_scores / get_cae_scores run one after the other."""
    from dfs_b200 import ScorerGroup
    dev = torch.device("cuda", _device_index(device))
    normalizers = list(normalizers) if normalizers is not None else [None] * len(models)
    scorers, flags = [], []
    for model, norm in zip(models, normalizers):
        model.eval()
        model.to(device)
        is_cae = hasattr(model, "set_normalizer")
        if is_cae and norm is not None:
            model.set_normalizer(norm.mean, norm.std)
        scorers.append(model.native(dev))
        flags.append(int(norm is not None) if is_cae else int(bool(apply_sigmoid)))
    slab = features_slab(features_df)
    group = ScorerGroup(scorers)
    try:
        outs = group.score_host(slab.transpose(1, 2), flags)
    finally:
        group.close()
    return [np.array(o.tolist()) for o in outs]


def collect_scores(model, dataloader, device, swap_tf=True):
    """src/ensemble.py:52-63 on top of the drop-in model (per-batch path)."""
    scores = []
    model.eval()
    with torch.no_grad():
        for features, _ in dataloader:
            features = features.to(device)
            if swap_tf:
                features = features.transpose(1, 2)
            scores.extend(torch.sigmoid(model(features).squeeze(-1)).cpu().tolist())
    return scores


def write_predictions(uttids, scores, path):
    """prediction.pkl = DataFrame{uttid: object, predictions: float64}, RangeIndex (predict.py:116-122;
    pandas 3 would write uttid as 'str' dtype, the shipped goldens are 'object' -- SURVEY.md §8c)."""
    import pandas as pd
    scores = np.asarray(scores, dtype=np.float64)
    if len(scores) != len(uttids):
        raise ValueError("Number of predictions does not match number of rows in features.pkl")
    df = pd.DataFrame({"uttid": pd.Series(list(uttids), dtype=object), "predictions": scores})
    df.to_pickle(path)
    return df
