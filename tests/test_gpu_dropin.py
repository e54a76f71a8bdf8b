"""GPU: the drop-in classes / helpers give the reference's outputs through the native path."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from conftest import GOLDEN, PKG  # noqa: E402

sys.path.insert(0, os.path.join(PKG, "dropin"))
from dfs_b200 import synthetic as syn  # noqa: E402
import evaluation as dev_eval  # noqa: E402
import model as m2  # noqa: E402
import model_cae as mc  # noqa: E402
import model_cnn1d as m1  # noqa: E402
import scoring  # noqa: E402
from oracle import eer as oeer  # noqa: E402

G = np.load(os.path.join(GOLDEN, "models.npz"))


def _t(sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


def test_predict_py_inner_loop_with_dropin_models():
    """src/predict.py:100-111 verbatim semantics: (B,180,321) storage -> .to(device) -> transpose(1,2) -> model -> sigmoid."""
    x = torch.from_numpy(syn.features(12, seed=1234))
    stored = x.transpose(1, 2).contiguous()                  # features.pkl rows are [180,321]
    for cls, sd, key in ((m2.CNN2D, syn.cnn2d_state(0), "cnn2d_init_sigmoid"), (m1.CNN1D, syn.cnn1d_state(0), "cnn1d_init_sigmoid")):
        model = cls(in_features=180, dropout=0.3).to("cuda")
        model.load_state_dict(_t(sd))
        model.eval()
        preds = []
        with torch.no_grad():
            for i in range(0, 12, 5):
                feats = stored[i:i + 5].to("cuda").transpose(1, 2)
                preds.extend(torch.sigmoid(model(feats).squeeze(-1)).detach().cpu().tolist())
        assert _rel(np.array(preds), G[key]) <= 1e-3
        # weights changed in place -> the native handle is rebuilt
        with torch.no_grad():
            model.classifier.bias.add_(1.0)
            z = model(stored[:2].to("cuda").transpose(1, 2)).squeeze(-1).cpu().numpy()
        ref = np.log(G[key][:2] / (1 - G[key][:2])) + 1.0
        np.testing.assert_allclose(z, ref, atol=2e-3)


def test_embedding_and_cae_forward_contracts():
    x = torch.from_numpy(syn.features(4, seed=1234)).cuda()
    model = m2.CNN2D().cuda().eval()
    model.load_state_dict(_t(syn.cnn2d_state(0)))
    logits, emb = model(x, return_embedding=True)
    assert tuple(logits.shape) == (4, 1) and tuple(emb.shape) == (4, 23040)
    cae = mc.ConvAutoencoder().cuda().eval()
    cae.load_state_dict(_t(syn.cae_state(0)))
    mean, std = syn.normalizer_stats(1)
    xn = (x - torch.from_numpy(mean).cuda()) / torch.from_numpy(std).cuda()
    recon, latent = cae(xn)
    mse = torch.nn.MSELoss(reduction="none")(recon, xn).view(4, -1).mean(1)       # predict_hybrid.py:76 on the compat path
    assert _rel(mse.cpu().numpy(), G["cae_mse"][:4]) <= 1e-3
    assert _rel(cae.score_mse(xn, apply_normalizer=False).cpu().numpy(), G["cae_mse"][:4]) <= 1e-3


def test_predict_hybrid_flow(tmp_path):
    """get_supervised_scores + get_cae_scores + normalise_01 + alpha blend + prediction.pkl (predict_hybrid.py:142-158)."""
    import pandas as pd
    n = 12
    x = syn.features(n, seed=1234)
    df = pd.DataFrame({"uttid": [f"utt_{i}" for i in range(n)], "features": [torch.from_numpy(x[i].T.copy()) for i in range(n)]})
    sup = m2.CNN2D(in_features=180, dropout=0.2)
    sup.load_state_dict(_t(syn.cnn2d_state(0)))
    cae = mc.ConvAutoencoder()
    cae.load_state_dict(_t(syn.cae_state(0)))
    mean, std = syn.normalizer_stats(1)

    class Norm:
        pass
    norm = Norm()
    norm.mean, norm.std = torch.from_numpy(mean), torch.from_numpy(std)
    s = scoring.get_supervised_scores(sup, df, "cuda")
    c = scoring.get_cae_scores(cae, df, norm, "cuda")
    assert s.dtype == np.float64 and c.dtype == np.float64
    assert _rel(s, G["cnn2d_init_sigmoid"].astype(np.float64)) <= 1e-3
    assert _rel(c, G["cae_mse"].astype(np.float64)) <= 1e-3
    hyb = 0.8 * scoring.normalise_01(s) + (1 - 0.8) * scoring.normalise_01(c)
    assert np.array_equal(hyb, oeer.hybrid_blend(s, c, 0.8))
    out = scoring.write_predictions(df["uttid"].values, hyb, tmp_path / "prediction_hybrid.pkl")
    assert out["predictions"].dtype == np.float64
    labels = (np.arange(n) % 2).tolist()
    assert dev_eval.calculate_eer(hyb.tolist(), labels) == oeer.calculate_eer(hyb.tolist(), labels, kind="stable")


def test_precision_fp32_on_the_dropin_classes_tracks_the_reference():
    """`model.precision = "fp32"` (or DFS_B200_PRECISION=fp32) routes the drop-in CNN2D / CNN1D / ConvAutoencoder through the full-fp32
    CUDA-core kernels: logits / MSE agree with the unmodified reference (goldens) to fp32 round-off."""
    x = torch.from_numpy(syn.features(12, seed=1234)).cuda()
    for cls, state, key in ((m2.CNN2D, syn.cnn2d_state(0), "cnn2d_init_logits"), (m1.CNN1D, syn.cnn1d_state(0), "cnn1d_init_logits")):
        net = cls(in_features=180, dropout=0.2).cuda().eval()
        net.load_state_dict(_t(state))
        net.precision = "fp32"
        with torch.no_grad():
            got = net(x).squeeze(-1).cpu().numpy()
        np.testing.assert_allclose(got, G[key], rtol=0, atol=2e-6)
        net.precision = None                       # back on the tensor-core path: the fp16-operand result, within the score gate
        with torch.no_grad():
            fast = net(x).squeeze(-1).cpu().numpy()
        assert not np.array_equal(fast, got) and np.max(np.abs(fast - G[key])) <= 1e-3
    cae = mc.ConvAutoencoder().cuda().eval()
    cae.load_state_dict(_t(syn.cae_state(0)))
    cae.precision = "fp32"
    mean, std = syn.normalizer_stats(1)
    xn = (x - torch.from_numpy(mean).cuda()) / torch.from_numpy(std).cuda()
    with torch.no_grad():
        recon, latent = cae(xn)
    assert tuple(recon.shape) == (12, 321, 180) and tuple(latent.shape) == (12, 256, 20, 11)
    mse = torch.nn.MSELoss(reduction="none")(recon, xn).view(12, -1).mean(1).cpu().numpy()
    assert _rel(mse, G["cae_mse"]) <= 5e-6
    assert _rel(cae.score_mse(xn, apply_normalizer=False).cpu().numpy(), G["cae_mse"]) <= 5e-6
