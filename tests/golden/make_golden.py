"""Generate the golden fixtures by running the UNMODIFIED reference in the build container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

/root/reference is imported read-only via sys.path (never copied); it does not exist on the
GPU box, so the tests only ever read the .npz files written here.  What is pinned:

* models.npz    -- reference CNN2D / CNN1D / ConvAutoencoder outputs (logits, sigmoid, embedding
                   digest, per-utterance MSE, latent digest) on dfs_b200.synthetic weights + inputs,
                   plus sha256 digests of those weights/inputs so a drifting factory fails loudly.
* eer_cases.npz -- scripts/evaluation.py::calculate_eer + confusion_at_threshold outputs on the
                   README semantics cases, seeded tie-free vectors and tie-heavy vectors (with the
                   argsort permutation the reference's numpy produced here).
* blend_known_answer.npz -- the one bit-exact known-answer triple the reference ships
                   (results/prediction_{final_test,cae_only_final,hybrid_final}.pkl).
* prediction_format.json -- dtype/shape facts of examples/prediction.pkl.
* trained.npz   -- trained-like regime: classifier scale / bias calibrated so that the REFERENCE's logits are centred and span
                   +-20 on 2,048 heterogeneous utterances (dfs_b200.synthetic.features_structured, heavy tails to -61 / +86);
                   reference logits, sigmoids, labels and the reference's own EER on them.
* dlq.npz       -- reference DeepfakeDetector (src/dlqueen_model.py) logits on seeded weights, full-length and ragged batches.
* hybrid_wide.npz -- the reference's OWN hybrid scoring functions (src/predict_hybrid.py: get_supervised_scores, get_cae_scores,
                   normalise_01, the alpha blend of main()) on the first 1,024 structured utterances, stored the way the reference
                   stores them ((180,321) rows in a DataFrame), with the calibrated 2D-CNN of trained.npz and the CAE + normaliser
                   of models.npz; the reference's calculate_eer on the supervised, CAE and hybrid columns.
"""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, "deep-fake-audio-classifier_b200"))
sys.path.insert(0, os.path.join(REF, "src"))

import torch  # noqa: E402

from dfs_b200 import synthetic as syn  # noqa: E402


def _load_ref_eval():
    spec = importlib.util.spec_from_file_location("ref_scripts_evaluation", os.path.join(REF, "scripts", "evaluation.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _to_torch(sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def make_models():
    from model import CNN2D
    from model_cae import ConvAutoencoder
    from model_cnn1d import CNN1D
    import torch.nn as nn

    torch.set_num_threads(8)
    n = 12
    x_np = syn.features(n, seed=1234)
    x = torch.from_numpy(x_np)
    out = {"n": n, "features_sha256": syn.state_digest([x_np])}

    for scale, tag in ((1.0, "init"), (2000.0, "trained")):
        sd = syn.cnn2d_state(0, logit_scale=scale)
        m = CNN2D(in_features=180, dropout=0.2)
        m.load_state_dict(_to_torch(sd))
        m.eval()
        with torch.no_grad():
            # the reference hands the model a transposed, non-contiguous view (predict.py:103-105)
            xt = x.transpose(1, 2).contiguous().transpose(1, 2)
            logits, emb = m(xt, return_embedding=True)
        out[f"cnn2d_{tag}_sha256"] = syn.state_digest(sd)
        out[f"cnn2d_{tag}_logits"] = logits.squeeze(-1).numpy()
        out[f"cnn2d_{tag}_sigmoid"] = torch.sigmoid(logits.squeeze(-1)).numpy()
        if tag == "init":
            out["cnn2d_init_embedding_head"] = emb[:, :512].numpy()
            out["cnn2d_init_embedding_sum"] = emb.double().sum(1).numpy()

        sd1 = syn.cnn1d_state(0, logit_scale=scale / 20.0 if scale > 1 else 1.0)
        m1 = CNN1D(in_features=180, dropout=0.2)
        m1.load_state_dict(_to_torch(sd1))
        m1.eval()
        with torch.no_grad():
            l1 = m1(x).squeeze(-1)
        out[f"cnn1d_{tag}_sha256"] = syn.state_digest(sd1)
        out[f"cnn1d_{tag}_logits"] = l1.numpy()
        out[f"cnn1d_{tag}_sigmoid"] = torch.sigmoid(l1).numpy()

    sdc = syn.cae_state(0)
    mean, std = syn.normalizer_stats(1)
    mc = ConvAutoencoder()
    mc.load_state_dict(_to_torch(sdc))
    mc.eval()
    with torch.no_grad():
        xn = (x - torch.from_numpy(mean)) / torch.from_numpy(std)          # dataset_cae.py:37-41
        recon, latent = mc(xn)
        mse = nn.MSELoss(reduction="none")(recon, xn).view(n, -1).mean(1)  # predict_hybrid.py:76
    out["cae_sha256"] = syn.state_digest(sdc)
    out["cae_norm_sha256"] = syn.state_digest([mean, std])
    out["cae_mse"] = mse.numpy()
    out["cae_latent_sum"] = latent.double().sum((1, 2, 3)).numpy()
    out["cae_recon_row0"] = recon[:, 0, :].numpy()
    out["cae_recon_last_row_absmax"] = recon[:, 320, :].abs().max().numpy()
    out["cae_recon_shape"] = np.array(recon.shape)
    out["cae_latent_shape"] = np.array(latent.shape)
    np.savez_compressed(os.path.join(HERE, "models.npz"), **out)
    print("models.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


def make_trained():
    """Trained-like regime on heterogeneous inputs (trained.npz).  The random-init classifiers put every logit near one value
    (2D-CNN: 0.06 +- 0.016), so sigmoids never leave 0.51-0.52 and a x2000 classifier only moves all of them to +130 (sigmoid
    exactly 1).  Here the classifier is CALIBRATED against the unmodified reference: scale and bias are chosen on the first 256
    structured utterances so that the reference's logits are centred on 0 and span +-20, then the reference scores all 2,048
    utterances.  Pins: the calibration constants, the reference logits / sigmoids, labels drawn from the reference logits,
    and the reference's own calculate_eer on its scores and on its logits."""
    from model import CNN2D
    from model_cnn1d import CNN1D
    ev = _load_ref_eval()
    torch.set_num_threads(os.cpu_count() or 8)
    n, n_cal, seed = 2048, 256, 4321
    x_np = syn.features_structured(n, seed=seed)
    out = {"n": n, "n_cal": n_cal, "seed": seed, "features_sha256_first64": syn.state_digest([x_np[:64]])}

    def run(model, xs):
        outs = []
        with torch.no_grad():
            for i in range(0, len(xs), 32):                                   # predict.py:100-111 batch loop, bs 32
                b = torch.from_numpy(xs[i:i + 32])
                b = b.transpose(1, 2).contiguous().transpose(1, 2)            # the reference's non-contiguous view (predict.py:103-105)
                outs.append(model(b).squeeze(-1))
        return torch.cat(outs).numpy()

    for tag, cls, factory in (("cnn2d", CNN2D, syn.cnn2d_state), ("cnn1d", CNN1D, syn.cnn1d_state)):
        m = cls(in_features=180, dropout=0.2)
        m.load_state_dict(_to_torch(factory(0)))
        m.eval()
        cal = run(m, x_np[:n_cal]).astype(np.float64)
        scale = 20.0 / float(np.max(np.abs(cal - cal.mean())))
        b0 = float(factory(0)["classifier.bias"][0])
        bias = scale * (b0 - float(cal.mean()))
        sd = factory(0, logit_scale=scale, classifier_bias=bias)
        m.load_state_dict(_to_torch(sd))
        logits = run(m, x_np)
        sig = torch.sigmoid(torch.from_numpy(logits)).numpy()
        rng = np.random.Generator(np.random.PCG64(99))
        lab = (rng.random(n) < 1.0 / (1.0 + np.exp(-logits.astype(np.float64) / 4.0))).astype(np.int64)
        out[f"{tag}_scale"], out[f"{tag}_bias"] = np.float64(scale), np.float64(bias)
        out[f"{tag}_sha256"] = syn.state_digest(sd)
        out[f"{tag}_logits"], out[f"{tag}_sigmoid"], out[f"{tag}_labels"] = logits, sig, lab
        out[f"{tag}_eer_thr_scores"] = np.array(ev.calculate_eer(np.array(sig.tolist()), lab), dtype=np.float64)   # float64 column like .tolist()
        out[f"{tag}_eer_thr_logits"] = np.array(ev.calculate_eer(np.array(logits.tolist()), lab), dtype=np.float64)
        print(tag, "scale", scale, "bias", bias, "logit range", logits.min(), logits.max(), "EER", out[f"{tag}_eer_thr_scores"],
              "unsaturated", int(((sig > 1e-6) & (sig < 1 - 1e-6)).sum()))
    np.savez_compressed(os.path.join(HERE, "trained.npz"), **out)


def make_hybrid_wide():
    """The hybrid path end to end through the reference's unmodified functions (src/predict_hybrid.py:52-85, 142-151)."""
    import pandas as pd
    import predict_hybrid as ph                                       # the unmodified module
    from dataset_cae import FeatureNormalizer
    from model import CNN2D
    from model_cae import ConvAutoencoder
    ev = _load_ref_eval()
    torch.set_num_threads(os.cpu_count() or 8)
    tr = np.load(os.path.join(HERE, "trained.npz"))
    n, seed, alpha = 1024, int(tr["seed"]), 0.8
    x_np = syn.features_structured(n, seed=seed)
    assert syn.state_digest([x_np[:64]]) == str(tr["features_sha256_first64"])
    # features.pkl rows are (180, 321) tensors (src/predict_hybrid.py:46-47 transposes them back)
    df = pd.DataFrame({"uttid": [f"utt_{i:05d}" for i in range(n)],
                       "features": [torch.from_numpy(np.ascontiguousarray(x_np[i].T)) for i in range(n)]})
    sd2 = syn.cnn2d_state(0, logit_scale=float(tr["cnn2d_scale"]), classifier_bias=float(tr["cnn2d_bias"]))
    assert syn.state_digest(sd2) == str(tr["cnn2d_sha256"])
    sup = CNN2D(in_features=180, dropout=0.2)
    sup.load_state_dict(_to_torch(sd2))
    sdc = syn.cae_state(0)
    mean, std = syn.normalizer_stats(1)
    cae = ConvAutoencoder()
    cae.load_state_dict(_to_torch(sdc))
    norm = FeatureNormalizer()
    norm.mean, norm.std = torch.from_numpy(mean), torch.from_numpy(std)
    sup_scores = ph.get_supervised_scores(sup, df, "cpu", batch_size=32)
    cae_mse = ph.get_cae_scores(cae, df, norm, "cpu", batch_size=32)
    sup_norm = ph.normalise_01(sup_scores)                            # predict_hybrid.py:148-150, verbatim semantics
    cae_norm = ph.normalise_01(cae_mse)
    hybrid = alpha * sup_norm + (1 - alpha) * cae_norm
    lab = tr["cnn2d_labels"][:n]
    out = dict(n=n, seed=seed, alpha=alpha, cae_sha256=syn.state_digest(sdc), cae_norm_sha256=syn.state_digest([mean, std]),
               sup_scores=sup_scores, cae_mse=cae_mse, sup_norm=sup_norm, cae_norm=cae_norm, hybrid=hybrid, labels=lab,
               eer_thr_sup=np.array(ev.calculate_eer(sup_scores, lab), dtype=np.float64),
               eer_thr_cae=np.array(ev.calculate_eer(cae_norm, lab), dtype=np.float64),
               eer_thr_hybrid=np.array(ev.calculate_eer(hybrid, lab), dtype=np.float64))
    assert np.array_equal(sup_scores.astype(np.float32), tr["cnn2d_sigmoid"][:n])   # same reference, two call paths
    np.savez_compressed(os.path.join(HERE, "hybrid_wide.npz"), **out)
    print("hybrid_wide.npz: cae_mse", cae_mse.min(), cae_mse.max(), "EER sup / cae / hybrid", out["eer_thr_sup"], out["eer_thr_cae"],
          out["eer_thr_hybrid"])


def make_eer():
    ev = _load_ref_eval()
    cases = {}

    def add(name, scores, labels):
        scores = np.asarray(scores)
        labels = np.asarray(labels)
        eer, thr = ev.calculate_eer(scores, labels)
        tp, fp, tn, fn, far, frr = ev.confusion_at_threshold(scores, labels, thr)
        cases[name + "/scores"] = scores
        cases[name + "/labels"] = labels
        cases[name + "/eer_thr"] = np.array([eer, thr], dtype=np.float64)
        cases[name + "/confusion"] = np.array([tp, fp, tn, fn], dtype=np.int64)
        cases[name + "/far_frr"] = np.array([far, frr], dtype=np.float64)
        cases[name + "/ref_argsort"] = np.argsort(scores)   # the unstable order numpy gave HERE

    # README.md:110-115 semantics + single class (scripts/evaluation.py:18-19)
    add("perfect", [0.1, 0.2, 0.8, 0.9], [0, 0, 1, 1])
    add("inverted", [0.1, 0.2, 0.8, 0.9], [1, 1, 0, 0])
    add("single_class_pos", [0.3, 0.4, 0.5], [1, 1, 1])
    add("single_class_neg", [0.3, 0.4, 0.5], [0, 0, 0])
    add("one_each", [0.7, 0.2], [1, 0])
    add("idx0_edge", [0.9, 0.8, 0.1], [0, 0, 1])
    for n, seed in ((257, 1), (4096, 2), (100000, 3)):
        s, l = syn.tie_free_scores(n, seed)
        add(f"tiefree_f32_{n}", s, l)
        add(f"tiefree_f64_{n}", s.astype(np.float64), l.astype(np.int64))
    # float labels as evaluate() passes them (src/evaluation.py:92)
    s, l = syn.tie_free_scores(1000, 9)
    add("float_labels_1000", s.astype(np.float64), l.astype(np.float64))
    # tie-heavy: saturated sigmoid-like scores; single-label tie groups are order invariant
    rng = np.random.Generator(np.random.PCG64(77))
    lab = (rng.random(2000) < 0.4).astype(np.int64)
    sc = np.where(lab == 1, 1.0, rng.random(2000) * 0.5).astype(np.float32)
    sc[lab == 0][:50] = 0.0
    add("ties_pure_groups_2000", sc, lab)
    # tie-heavy with mixed-label groups: the reference order is numpy-SIMD dependent
    sc2 = np.round(rng.random(2000) * 20).astype(np.float32) / 20.0
    lab2 = (rng.random(2000) < 0.3 + 0.4 * sc2).astype(np.int64)
    add("ties_mixed_groups_2000", sc2, lab2)
    # negative scores / logits (evaluate() default apply_sigmoid=False)
    lg = (rng.standard_normal(3000) * 5).astype(np.float32)
    lb = (rng.random(3000) < 1 / (1 + np.exp(-lg))).astype(np.int64)
    add("logits_3000", lg, lb)
    np.savez_compressed(os.path.join(HERE, "eer_cases.npz"), **cases)
    print("eer_cases.npz:", sorted({k.split("/")[0] for k in cases}))


def make_blend():
    import pandas as pd
    a = pd.read_pickle(os.path.join(REF, "results", "prediction_final_test.pkl"))
    b = pd.read_pickle(os.path.join(REF, "results", "prediction_cae_only_final.pkl"))
    h = pd.read_pickle(os.path.join(REF, "results", "prediction_hybrid_final.pkl"))
    assert (a["uttid"].values == b["uttid"].values).all() and (a["uttid"].values == h["uttid"].values).all()
    np.savez_compressed(os.path.join(HERE, "blend_known_answer.npz"),
                        sup=a["predictions"].values.astype(np.float64),
                        cae_minmaxed=b["predictions"].values.astype(np.float64),
                        hybrid=h["predictions"].values.astype(np.float64),
                        alpha=np.float64(0.80))
    ex = pd.read_pickle(os.path.join(REF, "examples", "prediction.pkl"))
    facts = {"columns": list(ex.columns), "dtypes": {c: str(t) for c, t in ex.dtypes.items()},
             "index_type": type(ex.index).__name__, "rows": int(len(ex)),
             "first_uttid": str(ex["uttid"].iloc[0])}
    with open(os.path.join(HERE, "prediction_format.json"), "w") as f:
        json.dump(facts, f, indent=1)
    print("blend_known_answer.npz, prediction_format.json:", facts)


def make_dlq():
    """DeepfakeDetector (src/dlqueen_model.py:156-173, imported unmodified) on seeded weights / inputs: full-length batch and a
    ragged one (frames beyond the length zeroed, as pad_sequence leaves them)."""
    spec = importlib.util.spec_from_file_location("ref_dlqueen_model", os.path.join(REF, "src", "dlqueen_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.set_num_threads(8)
    n = 12
    x = syn.features(n, seed=1234)
    lengths = np.array([321, 300, 321, 123, 64, 321, 1, 200, 321, 319, 8, 250], dtype=np.int64)
    xz = x.copy()
    for i, l in enumerate(lengths):
        xz[i, l:, :] = 0
    out = {"n": n, "features_sha256": syn.state_digest([x]), "lengths": lengths}
    for scale, tag in ((1.0, "init"), (300.0, "trained")):
        sd = syn.dlq_state(0, logit_scale=scale)
        m = mod.DeepfakeDetector(in_ch=180, hidden=256, dropout=0.3)
        assert list(m.state_dict().keys()) == list(sd.keys())
        m.load_state_dict(_to_torch(sd))
        m.eval()
        with torch.no_grad():
            full = m(torch.from_numpy(x).transpose(1, 2).contiguous(), torch.full((n,), 321, dtype=torch.long))
            ragged = m(torch.from_numpy(xz).transpose(1, 2).contiguous(), torch.from_numpy(lengths))
        out[f"dlq_{tag}_sha256"] = syn.state_digest(sd)
        out[f"dlq_{tag}_logits_full"] = full.numpy()
        out[f"dlq_{tag}_logits_ragged"] = ragged.numpy()
    out["state_dict_keys"] = np.array(list(sd.keys()))
    np.savez_compressed(os.path.join(HERE, "dlq.npz"), **out)
    print("dlq.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    which = sys.argv[1:] or ["models", "eer", "blend", "dlq", "trained"]
    if "models" in which:
        make_models()
    if "eer" in which:
        make_eer()
    if "blend" in which:
        make_blend()
    if "dlq" in which:
        make_dlq()
    if "trained" in which:
        make_trained()
    if "hybrid_wide" in which:
        make_hybrid_wide()
