"""Golden outputs of the UNMODIFIED reference command-line tools, produced in the
build container on CPU:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_cli.py      -> tests/golden/cli_cases.npz

* src/predict.py --model cnn2d|cnn1d [--no-apply-sigmoid]     (prediction.pkl: uttid order, dtypes, values)
* src/predict_hybrid.py --alpha 0.8                            (hybrid prediction.pkl)
* scripts/evaluation.py prediction.pkl labels.pkl              (printed EER / threshold / confusion)
* src/ensemble.py --checkpoints cnn2d:... cnn1d:...            (per-model and ensemble EER printout)
* src/hybrid_ensemble.py's alpha sweep loop (lines 139-151)    (EER per alpha on the dev scores)
* src/evaluation.py::evaluate with nn.BCEWithLogitsLoss        (avg_loss, eer, threshold on logits)

The input pickles are rebuilt from seeds by tests/cli_fixtures.py on the GPU box (same bytes: numpy PCG64), so
nothing under /root/reference is needed at test time.
"""
import importlib.util
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, "deep-fake-audio-classifier_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pandas as pd  # noqa: E402
import torch  # noqa: E402

import cli_fixtures as fx  # noqa: E402


def run(cmd, cwd):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", OMP_NUM_THREADS="8")
    r = subprocess.run([sys.executable] + cmd, cwd=cwd, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{cmd} failed:\n{r.stdout}\n{r.stderr}")
    return r.stdout


def main():
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        paths = fx.write_fixture_files(tmp)
        out["uttids"] = np.array(fx.uttids(), dtype=object).astype(str)
        for model in ("cnn2d", "cnn1d"):
            for flag, tag in (([], "sigmoid"), (["--no-apply-sigmoid"], "logits")):
                dst = os.path.join(tmp, f"pred_{model}_{tag}.pkl")
                run([os.path.join(REF, "src", "predict.py"), "--features", paths["features"], "--checkpoint", paths[model], "--model", model,
                     "--out", dst, "--device", "cpu", "--num-workers", "0", "--dropout", "0.2"] + flag, tmp)
                df = pd.read_pickle(dst)
                assert list(df["uttid"].values) == fx.uttids()
                out[f"predict_{model}_{tag}"] = df["predictions"].to_numpy(dtype=np.float64)
                out[f"predict_{model}_{tag}_dtype"] = str(df["predictions"].dtype)
        dst = os.path.join(tmp, "pred_hybrid.pkl")
        run([os.path.join(REF, "src", "predict_hybrid.py"), "--sup-checkpoint", paths["cnn2d"], "--cae-checkpoint", paths["cae"],
             "--cae-normalizer", paths["normalizer"], "--test-features", paths["features"], "--alpha", "0.8", "--out", dst, "--device", "cpu"], tmp)
        df = pd.read_pickle(dst)
        out["predict_hybrid"] = df["predictions"].to_numpy(dtype=np.float64)
        # scripts/evaluation.py on the cnn2d sigmoid predictions
        text = run([os.path.join(REF, "scripts", "evaluation.py"), os.path.join(tmp, "pred_cnn2d_sigmoid.pkl"), paths["labels"]], tmp)
        out["evaluation_stdout"] = np.array(text)
        # src/ensemble.py on the two supervised checkpoints (per-model and ensemble EER printout)
        text = run([os.path.join(REF, "src", "ensemble.py"), "--checkpoints", "cnn2d:" + paths["cnn2d"], "cnn1d:" + paths["cnn1d"],
                    "--dev-features", paths["features"], "--dev-labels", paths["labels"], "--device", "cpu"], os.path.join(REF, "src"))
        out["ensemble_stdout"] = np.array(text)
        # src/hybrid_ensemble.py and src/evaluation_cae.py CLIs (printouts)
        out["hybrid_ensemble_stdout"] = np.array(run([os.path.join(REF, "src", "hybrid_ensemble.py"), "--sup-checkpoint", paths["cnn2d"],
                                                      "--cae-checkpoint", paths["cae"], "--cae-normalizer", paths["normalizer"],
                                                      "--dev-features", paths["features"], "--dev-labels", paths["labels"], "--device", "cpu"],
                                                     os.path.join(REF, "src")))
        out["evaluation_cae_stdout"] = np.array(run([os.path.join(REF, "src", "evaluation_cae.py"), "--features", paths["features"], "--labels",
                                                     paths["labels"], "--checkpoint", paths["cae"], "--normalizer", paths["normalizer"], "--device", "cpu"],
                                                    os.path.join(REF, "src")))
        # the alpha sweep of src/hybrid_ensemble.py:131-151 on (sup, cae) score vectors of the reference itself
        spec = importlib.util.spec_from_file_location("ref_scripts_evaluation", os.path.join(REF, "scripts", "evaluation.py"))
        ev = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ev)
        sys.path.insert(0, os.path.join(REF, "src"))
        import hybrid_ensemble as he                                             # noqa: E402  (normalise_scores)
        from predict_hybrid import get_cae_scores, get_supervised_scores        # noqa: E402
        from dataset_cae import FeatureNormalizer                                # noqa: E402
        from model import CNN2D                                                  # noqa: E402
        from model_cae import ConvAutoencoder                                    # noqa: E402
        fdf = pd.read_pickle(paths["features"])
        labels = pd.merge(fdf[["uttid"]], pd.read_pickle(paths["labels"]), on="uttid", how="inner")["label"].to_numpy()   # features order
        assert np.array_equal(labels, fx.labels())
        sup = CNN2D(in_features=180, dropout=0.2)
        sup.load_state_dict(torch.load(paths["cnn2d"])["model_state"])
        cae = ConvAutoencoder()
        cae.load_state_dict(torch.load(paths["cae"])["model_state"])
        sup_scores = get_supervised_scores(sup, fdf, "cpu", 32)
        cae_scores = get_cae_scores(cae, fdf, FeatureNormalizer.load(paths["normalizer"]), "cpu", 32)
        out["sup_scores"], out["cae_scores"] = sup_scores, cae_scores
        sup_n, cae_n = he.normalise_scores(sup_scores), he.normalise_scores(cae_scores)
        alphas = np.linspace(0.0, 1.0, 21)
        sweep = []
        for a in alphas:
            combined = a * sup_n + (1 - a) * cae_n
            sweep.append(ev.calculate_eer(combined.tolist(), labels.tolist()))
        out["alpha_sweep_alphas"] = alphas
        out["alpha_sweep_eer_thr"] = np.array(sweep, dtype=np.float64)
        # src/evaluation.py::evaluate with BCEWithLogitsLoss on a (features, label) loader, logits (apply_sigmoid=False)
        import evaluation as ref_eval                                            # noqa: E402
        from dataset import AudioDeepfakeDataset                                 # noqa: E402
        from torch.utils.data import DataLoader                                  # noqa: E402
        ds = AudioDeepfakeDataset(paths["features"], paths["labels"])
        metrics, scores, labs = ref_eval.evaluate(sup, DataLoader(ds, batch_size=5, shuffle=False), criterion=torch.nn.BCEWithLogitsLoss(),
                                                  device="cpu", swap_tf=True)
        out["evaluate_metrics"] = np.array([metrics["avg_loss"], metrics["eer"], metrics["threshold"]], dtype=np.float64)
        out["evaluate_scores"] = np.array(scores, dtype=np.float64)
        out["evaluate_labels"] = np.array(labs, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "cli_cases.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
