"""CPU gate for SURVEY.md §8(f) rows 1-2: feature ingestion (features.pkl -> one pinned slab + uttid index), the
prediction.pkl writer, the CLI flag surface of predict.py / predict_hybrid.py, their error behaviour, and the
committed CLI goldens themselves (produced by the unmodified reference, tests/golden/make_golden_cli.py)."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, PKG

torch = pytest.importorskip("torch")
pd = pytest.importorskip("pandas")
sys.path.insert(0, os.path.join(PKG, "dropin"))

import cli_fixtures as fx  # noqa: E402
import ingest  # noqa: E402
import predict as dpredict  # noqa: E402
import predict_hybrid as dhybrid  # noqa: E402
import scoring  # noqa: E402
from oracle import eer as oeer  # noqa: E402

CLI = np.load(os.path.join(GOLDEN, "cli_cases.npz"), allow_pickle=False)


def test_pack_features_matches_row_by_row_float_cast(tmp_path):
    paths = fx.write_fixture_files(str(tmp_path))
    df = pd.read_pickle(paths["features"])
    table = ingest.load_feature_table(paths["features"])
    assert len(table) == fx.N_UTTS and table.slab.dtype == torch.float32 and tuple(table.slab.shape) == (fx.N_UTTS, 180, 321)
    assert list(table.uttids) == fx.uttids() and table.uttids.dtype == object
    for i in range(fx.N_UTTS):                                            # what FeatureOnlyDataset.__getitem__ returns (predict.py:62-63)
        assert torch.equal(table.slab[i], df["features"].iloc[i].float())
    v = table.view()                                                      # the .transpose(1, 2) view, no copy
    assert tuple(v.shape) == (fx.N_UTTS, 321, 180) and v.stride() == (180 * 321, 1, 321) and v.data_ptr() == table.slab.data_ptr()
    # non-fp32 rows are cast like .float(); numpy rows are accepted
    mixed = [df["features"].iloc[0].double(), df["features"].iloc[1].half(), df["features"].iloc[2].numpy()]
    slab = ingest.pack_features(mixed)
    assert torch.equal(slab[0], mixed[0].float()) and torch.equal(slab[1], mixed[1].float()) and torch.equal(slab[2], df["features"].iloc[2])
    t16 = ingest.load_feature_table(paths["features"], dtype=torch.float16)        # half-width slab for dfs_score_host_f16
    assert t16.slab.dtype == torch.float16 and torch.equal(t16.slab, table.slab.half()) and t16.take([2]).slab.dtype == torch.float16
    sub = table.take([3, 1])
    assert list(sub.uttids) == [fx.uttids()[3], fx.uttids()[1]] and torch.equal(sub.slab[0], table.slab[3])


def test_ingest_errors_mirror_the_reference():
    with pytest.raises(ValueError, match="uttid"):                         # predict.py:89-90
        ingest.load_feature_table(pd.DataFrame({"features": [torch.zeros(180, 321)]}))
    with pytest.raises(ValueError, match="shape"):
        ingest.pack_features([torch.zeros(321, 180)])
    with pytest.raises(ValueError, match="no rows"):
        ingest.pack_features([])
    t = ingest.FeatureTable(np.array(["a"], dtype=object), torch.zeros(1, 180, 321))
    with pytest.raises(ValueError, match="label"):                         # scripts/evaluation.py:75-76
        ingest.merge_labels(t, pd.DataFrame({"uttid": ["a"]}))


def test_merge_labels_is_the_inner_merge_of_the_reference(tmp_path):
    paths = fx.write_fixture_files(str(tmp_path))
    table = ingest.load_feature_table(paths["features"])
    labels_df = pd.read_pickle(paths["labels"])
    idx, lab = ingest.merge_labels(table, labels_df)
    ref = pd.merge(pd.read_pickle(paths["features"]), labels_df, on="uttid", how="inner").reset_index(drop=True)   # dataset.py:29-33
    assert [table.uttids[i] for i in idx] == list(ref["uttid"].values)
    assert np.array_equal(lab, ref["label"].to_numpy().astype(np.uint8))
    assert np.array_equal(lab, fx.labels().astype(np.uint8))              # features order
    # rows without a label are dropped, extra labels ignored
    part = pd.concat([labels_df.iloc[:5], pd.DataFrame({"uttid": ["zzz"], "label": [1]})])
    idx2, _ = ingest.merge_labels(table, part)
    assert len(idx2) == 5


def test_cli_flag_surface_matches_the_reference():
    a = dpredict.parse_args(["--features", "f", "--checkpoint", "c", "--model", "cnn1d", "--out", "o"])
    assert (a.batch_size, a.num_workers, a.device, a.in_features, a.dropout, a.apply_sigmoid, a.no_apply_sigmoid, a.swap_tf) == \
           (32, 2, None, 180, 0.3, True, False, True)                       # predict.py:11-38 defaults
    a = dpredict.parse_args(["--features", "f", "--checkpoint", "c", "--model", "cnn2d", "--out", "o", "--no-apply-sigmoid", "--no-swap-tf"])
    assert a.no_apply_sigmoid and not a.swap_tf
    with pytest.raises(SystemExit):
        dpredict.parse_args(["--features", "f", "--checkpoint", "c", "--model", "cae", "--out", "o"])
    h = dhybrid.parse_args(["--sup-checkpoint", "s", "--cae-checkpoint", "c", "--cae-normalizer", "n", "--test-features", "t"])
    assert (h.alpha, h.out, h.batch_size, h.device, h.existing_submission) == (0.80, "prediction_hybrid.pkl", 32, None, None)
    import evaluation_cae as dcae
    import hybrid_ensemble as dhe
    e = dhe.parse_args(["--sup-checkpoint", "s", "--cae-checkpoint", "c", "--cae-normalizer", "n"])   # hybrid_ensemble.py:96-109
    assert (e.sup_arch, e.dev_features, e.dev_labels, e.batch_size, e.device, e.alpha_steps) == \
           ("cnn2d", "data/dev/features.pkl", "data/dev/labels.pkl", 32, None, 21)
    with pytest.raises(SystemExit):
        dhe.parse_args(["--sup-checkpoint", "s", "--cae-checkpoint", "c", "--cae-normalizer", "n", "--sup-arch", "cnn1d"])
    c = dcae.parse_args(["--features", "f", "--labels", "l", "--checkpoint", "c", "--normalizer", "n"])  # evaluation_cae.py:94-104
    assert (c.batch_size, c.base_channels, c.device) == (32, 32, None)


def test_cli_refuses_to_run_without_cuda(tmp_path):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    paths = fx.write_fixture_files(str(tmp_path))
    with pytest.raises(RuntimeError, match="no mps / cpu fallback"):
        dpredict.main(["--features", paths["features"], "--checkpoint", paths["cnn2d"], "--model", "cnn2d", "--out", str(tmp_path / "p.pkl")])
    with pytest.raises(RuntimeError, match="CUDA"):
        dpredict.main(["--features", paths["features"], "--checkpoint", paths["cnn2d"], "--model", "cnn2d", "--out", str(tmp_path / "p.pkl"),
                       "--device", "cpu"])
    with pytest.raises(FileNotFoundError):
        dpredict.load_checkpoint_into(dpredict.CNN2D(), str(tmp_path / "missing.pt"), "cpu")


def test_prediction_pkl_format_and_submission_validator(tmp_path):
    """prediction.pkl as scripts/generate_submission.py:20-36 validates it: exactly the two columns, float predictions,
    uttid set equal to the features' -- and the dtype facts of the shipped examples/prediction.pkl."""
    facts = json.load(open(os.path.join(GOLDEN, "prediction_format.json")))
    scores = CLI["predict_cnn2d_sigmoid"]
    p = tmp_path / "prediction.pkl"
    df = scoring.write_predictions(fx.uttids(), scores.tolist(), str(p))
    back = pd.read_pickle(p)
    assert list(back.columns) == facts["columns"] == ["uttid", "predictions"]
    assert {c: str(t) for c, t in back.dtypes.items()} == facts["dtypes"]
    assert type(back.index).__name__ == facts["index_type"]
    assert all(isinstance(x, (float, np.floating)) for x in back["predictions"].values)
    assert set(back["uttid"].values) == set(fx.uttids()) and list(back["uttid"].values) == fx.uttids()
    assert np.array_equal(back["predictions"].to_numpy(), scores) and df.equals(back)
    with pytest.raises(ValueError, match="does not match"):               # predict.py:113-114
        scoring.write_predictions(fx.uttids(), scores[:-1], str(p))


def test_cli_goldens_are_consistent_with_the_oracle():
    """The reference CLI outputs pinned in cli_cases.npz agree with the oracle restatements (so the GPU tests that
    compare against them and against the oracle check the same thing)."""
    from dfs_b200 import synthetic as syn
    from oracle import models_np as onp
    x = syn.features(fx.N_UTTS, seed=1234)
    z2 = onp.cnn2d_forward(syn.cnn2d_state(0), x)[:, 0]
    np.testing.assert_allclose(CLI["predict_cnn2d_logits"], z2, rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(CLI["predict_cnn2d_sigmoid"], onp.sigmoid(z2), rtol=1e-5)
    np.testing.assert_allclose(CLI["predict_cnn1d_sigmoid"], onp.sigmoid(onp.cnn1d_forward(syn.cnn1d_state(0), x)[:, 0]), rtol=1e-5)
    mean, std = syn.normalizer_stats(1)
    np.testing.assert_allclose(CLI["cae_scores"], onp.cae_mse_scores(syn.cae_state(0), x, mean, std), rtol=1e-5)
    assert np.array_equal(CLI["predict_hybrid"], oeer.hybrid_blend(CLI["sup_scores"], CLI["cae_scores"], 0.8))   # bit-exact blend
    for a, (eer, thr) in zip(CLI["alpha_sweep_alphas"], CLI["alpha_sweep_eer_thr"]):
        comb = a * oeer.normalise_01(CLI["sup_scores"]) + (1 - a) * oeer.normalise_01(CLI["cae_scores"])
        assert oeer.calculate_eer(comb.tolist(), fx.labels().tolist()) == (eer, thr)
    # scripts/evaluation.py's printout of the cnn2d predictions
    eer, thr = oeer.calculate_eer(CLI["predict_cnn2d_sigmoid"], fx.labels())
    text = str(CLI["evaluation_stdout"])
    assert f"EER: {eer:.6f}" in text and f"Threshold: {thr:.6f}" in text


def test_cae_metrics_follow_the_reference_report():
    """evaluation_cae.py:58-88 on the reference's own CAE scores: the printed numbers of the reference CLI, recomputed with
    the oracle EER (the drop-in's cae_metrics needs the device; its arithmetic outside the two EERs is this)."""
    text = str(CLI["evaluation_cae_stdout"])
    mse, lab = CLI["cae_scores"], fx.labels()
    e_neg, t_neg = oeer.calculate_eer((-mse).tolist(), lab.tolist())
    e_pos, t_pos = oeer.calculate_eer(mse.tolist(), lab.tolist())
    assert f"Avg MSE (all):      {np.mean(mse):.6f}" in text
    assert f"Avg MSE (bonafide): {np.mean(mse[lab == 1]):.6f}" in text and f"Avg MSE (spoof):    {np.mean(mse[lab == 0]):.6f}" in text
    assert f"EER (-MSE):         {e_neg:.6f}" in text and f"EER (+MSE):         {e_pos:.6f}" in text
    thr = -t_neg if e_neg <= e_pos else t_pos
    assert f"Threshold (MSE):    {thr:.6f}" in text
    # hybrid_ensemble.py's printout: the sweep table is the pinned sweep
    htext = str(CLI["hybrid_ensemble_stdout"])
    for a, (eer, _) in zip(CLI["alpha_sweep_alphas"], CLI["alpha_sweep_eer_thr"]):
        assert f"  {a:.2f}    {eer:.6f}" in htext
