"""GPU parity of the HYBRID path end to end against the reference's own functions (tests/golden/hybrid_wide.npz, written by
make_golden.py::make_hybrid_wide from the unmodified src/predict_hybrid.py: get_supervised_scores, get_cae_scores,
normalise_01 and the alpha blend of main(), lines 52-85 and 142-151) on 1,024 heterogeneous, heavy-tailed utterances stored
the way features.pkl stores them ((180,321) rows of a DataFrame).  The drop-in helpers of the same names score that table
through the C ABI: one upload, both models, float64 min-max + blend and the EER on the device."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from conftest import GOLDEN, PKG, ROOT  # noqa: E402

sys.path.insert(0, os.path.join(PKG, "dropin"))
import dfs_b200 as D  # noqa: E402
from dfs_b200 import synthetic as syn  # noqa: E402
import dataset_cae  # noqa: E402
import model as m2  # noqa: E402
import model_cae as mc  # noqa: E402
import scoring  # noqa: E402

H = np.load(os.path.join(GOLDEN, "hybrid_wide.npz"))
T = np.load(os.path.join(GOLDEN, "trained.npz"))
REL = 1e-3       # north_star: per-utterance scores within 1e-3 relative
EER_ABS = 1e-4   # north_star: EER within 0.01 percentage points


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)))


def _record(name, payload):
    """Measured parity figures next to the assertions (gpurun_out/ is brought back from the GPU box)."""
    import json
    path = os.path.join(ROOT, "gpurun_out", "parity_round2.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    data = {}
    if os.path.exists(path):
        with open(path) as f:
            data = json.load(f)
    data[name] = payload
    with open(path, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)


def _t(sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


@pytest.fixture(scope="module")
def table():
    import pandas as pd
    n = int(H["n"])
    x = syn.features_structured(n, seed=int(H["seed"]))
    assert syn.state_digest([x[:64]]) == str(T["features_sha256_first64"])
    return pd.DataFrame({"uttid": [f"utt_{i:05d}" for i in range(n)],
                         "features": [torch.from_numpy(np.ascontiguousarray(x[i].T)) for i in range(n)]})


def _models(precision):
    sd2 = syn.cnn2d_state(0, logit_scale=float(T["cnn2d_scale"]), classifier_bias=float(T["cnn2d_bias"]))
    sdc = syn.cae_state(0)
    mean, std = syn.normalizer_stats(1)
    assert syn.state_digest(sdc) == str(H["cae_sha256"]) and syn.state_digest([mean, std]) == str(H["cae_norm_sha256"])
    sup = m2.CNN2D(in_features=180, dropout=0.2)
    sup.load_state_dict(_t(sd2))
    sup.precision = precision
    cae = mc.ConvAutoencoder()
    cae.load_state_dict(_t(sdc))
    norm = dataset_cae.FeatureNormalizer()
    norm.mean, norm.std = torch.from_numpy(mean), torch.from_numpy(std)
    return sup, cae, norm


# what each mode guarantees on an unsaturated sigmoid in this regime (tests/test_gpu_round2.py::SIGMOID_REL)
SUP_REL = {"fp16": 1.5e-2, "split": REL}


@pytest.mark.parametrize("precision", ["fp16", "split"])
def test_hybrid_scores_and_eer_against_the_reference_functions(table, precision):
    sup, cae, norm = _models(precision)
    sup_scores, cae_mse = scoring.score_models_once([sup, cae], table, "cuda", [None, norm])
    assert sup_scores.dtype == np.float64 and cae_mse.dtype == np.float64 and len(cae_mse) == int(H["n"])
    # the two reference-named helpers, one table pass each, return the same bits as the single upload
    if precision == "fp16":
        np.testing.assert_array_equal(scoring.get_cae_scores(cae, table, norm, "cuda"), cae_mse)
        np.testing.assert_array_equal(scoring.get_supervised_scores(sup, table, "cuda"), sup_scores)
    open_ = (H["sup_scores"] > 1e-6) & (H["sup_scores"] < 1 - 1e-6)
    rec = dict(n=int(H["n"]), cae_mse_range=[float(H["cae_mse"].min()), float(H["cae_mse"].max())],
               max_rel_cae_mse_err=_rel(cae_mse, H["cae_mse"]), max_rel_sup_err_unsaturated=_rel(sup_scores[open_], H["sup_scores"][open_]))
    assert rec["max_rel_cae_mse_err"] <= REL, rec
    assert rec["max_rel_sup_err_unsaturated"] <= SUP_REL[precision], rec

    alpha = float(H["alpha"])
    hybrid = scoring.hybrid_blend(sup_scores, cae_mse, alpha)            # predict_hybrid.py:148-150 on the device, float64
    # the device blend IS the reference's arithmetic: on the reference's own columns it returns the reference's bits
    np.testing.assert_array_equal(scoring.hybrid_blend(H["sup_scores"], H["cae_mse"], alpha), H["hybrid"])
    np.testing.assert_array_equal(scoring.normalise_01(H["cae_mse"]), H["cae_norm"])
    rec["max_abs_hybrid_err"] = float(np.max(np.abs(hybrid - H["hybrid"])))
    assert rec["max_abs_hybrid_err"] <= SUP_REL[precision], rec

    lab = H["labels"]
    for name, scores, want in (("sup", sup_scores, H["eer_thr_sup"]), ("cae", scoring.normalise_01(cae_mse), H["eer_thr_cae"]),
                               ("hybrid", hybrid, H["eer_thr_hybrid"])):
        eer, thr = D.calculate_eer(scores, lab)
        rec[f"eer_{name}_ref"], rec[f"eer_{name}_dev"] = float(want[0]), float(eer)
        assert abs(eer - want[0]) <= EER_ABS, (name, rec)
    # and on the reference's columns the device EER is the reference's, bit for bit
    assert D.calculate_eer(H["hybrid"], lab) == tuple(H["eer_thr_hybrid"])
    _record(f"hybrid_wide/{precision}", rec)
