"""Seeded input files for the CLI parity tests: features.pkl / labels.pkl / checkpoints / normaliser in the
reference's on-disk formats, rebuilt bit-identically from dfs_b200.synthetic wherever the
tests run (the goldens in tests/golden/cli_cases.npz were produced from these exact files by the reference)."""
import os

import numpy as np

N_UTTS = 12


def uttids():
    # deliberately not sorted, so an accidental re-ordering shows up
    return [f"utt_{(7 * i + 3) % 100:04d}" for i in range(N_UTTS)]


def labels():
    return np.array([1, 0, 0, 1, 1, 0, 1, 0, 0, 1, 0, 1], dtype=np.int64)


def write_fixture_files(directory):
    import pandas as pd
    import torch
    from dfs_b200 import synthetic as syn

    x = syn.features(N_UTTS, seed=1234)                                   # (N,321,180)
    feats = [torch.from_numpy(np.ascontiguousarray(x[i].T)) for i in range(N_UTTS)]   # rows are [180,321] (README.md:45-48)
    paths = {k: os.path.join(directory, v) for k, v in
             dict(features="features.pkl", labels="labels.pkl", cnn2d="cnn2d.pt", cnn1d="cnn1d.pt", cae="cae.pt", normalizer="cae_norm.pt").items()}
    pd.DataFrame({"uttid": uttids(), "features": feats}).to_pickle(paths["features"])
    # labels in a different row order than the features: the merge is on uttid
    order = np.array([5, 0, 11, 3, 8, 1, 10, 2, 7, 4, 9, 6])
    pd.DataFrame({"uttid": [uttids()[i] for i in order], "label": labels()[order]}).to_pickle(paths["labels"])

    def t(sd):
        return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}

    torch.save({"model_state": t(syn.cnn2d_state(0)), "epoch": 3, "config": {"model": "cnn2d"}}, paths["cnn2d"])   # training/checkpoint.py:59-66
    torch.save(t(syn.cnn1d_state(0)), paths["cnn1d"])                                                              # bare state dict
    torch.save({"model_state": t(syn.cae_state(0))}, paths["cae"])
    mean, std = syn.normalizer_stats(1)
    torch.save({"mean": torch.from_numpy(mean), "std": torch.from_numpy(std)}, paths["normalizer"])               # dataset_cae.py:43-52
    return paths
