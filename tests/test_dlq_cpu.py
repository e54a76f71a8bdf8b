"""CPU gate for the StatsPool detector: oracle restatement vs the golden logits of the unmodified
reference class, the drop-in's state-dict / constructor contract and its train-mode forward."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, PKG

torch = pytest.importorskip("torch")
sys.path.insert(0, os.path.join(PKG, "dropin"))

from dfs_b200 import synthetic as syn  # noqa: E402
import dlqueen_model as dq  # noqa: E402
from oracle import models_np as onp  # noqa: E402

G = np.load(os.path.join(GOLDEN, "dlq.npz"))


def _inputs():
    x = syn.features(int(G["n"]), seed=1234)
    assert syn.state_digest([x]) == str(G["features_sha256"])
    lengths = G["lengths"]
    xz = x.copy()
    for i, l in enumerate(lengths):
        xz[i, l:, :] = 0
    return x, xz, lengths


@pytest.mark.parametrize("tag,scale", [("init", 1.0), ("trained", 300.0)])
def test_oracle_matches_reference_golden(tag, scale):
    x, xz, lengths = _inputs()
    sd = syn.dlq_state(0, logit_scale=scale)
    assert syn.state_digest(sd) == str(G[f"dlq_{tag}_sha256"])
    np.testing.assert_allclose(onp.dlq_forward(sd, x), G[f"dlq_{tag}_logits_full"], rtol=2e-4, atol=2e-6 * scale)
    np.testing.assert_allclose(onp.dlq_forward(sd, xz, lengths), G[f"dlq_{tag}_logits_ragged"], rtol=2e-4, atol=2e-6 * scale)


def test_torch_oracle_matches_reference_golden():
    from oracle import models_torch as ot
    x, xz, lengths = _inputs()
    sd = syn.dlq_state(0)
    np.testing.assert_allclose(ot.dlq_forward(sd, torch.from_numpy(x)).numpy(), G["dlq_init_logits_full"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ot.dlq_forward(sd, torch.from_numpy(xz), lengths).numpy(), G["dlq_init_logits_ragged"], rtol=1e-5, atol=1e-7)


def test_dropin_contract_and_train_mode_forward():
    sd = syn.dlq_state(0)
    model = dq.DeepfakeDetector(in_ch=180, hidden=256, dropout=0.3)                # dlqueen_model.py:340,417
    assert list(model.state_dict().keys()) == list(G["state_dict_keys"]) == list(sd.keys())
    model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()})
    x, xz, lengths = _inputs()
    xb = torch.from_numpy(xz).transpose(1, 2).contiguous()                          # (B, C, T) as collate_fn builds it
    model.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(xb, torch.from_numpy(lengths))
    # the PyTorch layers (train-mode path) carry the reference's eval semantics when run through functional eval
    model.train()
    for mod in model.modules():
        if isinstance(mod, (torch.nn.BatchNorm1d, torch.nn.Dropout)):
            mod.eval()
    with torch.no_grad():
        got = model(xb, torch.from_numpy(lengths)).numpy()
    np.testing.assert_allclose(got, G["dlq_init_logits_ragged"], rtol=1e-5, atol=1e-7)
