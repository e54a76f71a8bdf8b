"""CPU gate for the drop-in boundary: constructors, state-dict keys/shapes, checkpoint
shapes, train-mode fall-through, eval-mode-on-CPU refusal, prediction.pkl format."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, PKG

torch = pytest.importorskip("torch")
sys.path.insert(0, os.path.join(PKG, "dropin"))

from dfs_b200 import synthetic as syn  # noqa: E402
import model as m2  # noqa: E402
import model_cae as mc  # noqa: E402
import model_cnn1d as m1  # noqa: E402
import scoring  # noqa: E402


def _t(sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


@pytest.mark.parametrize("cls,factory,kwargs", [
    (m2.CNN2D, syn.cnn2d_state, {"in_features": 180, "dropout": 0.3}),     # predict.py:51-52,76
    (m1.CNN1D, syn.cnn1d_state, {"in_features": 180, "dropout": 0.2}),     # ensemble.py:35-39
    (mc.ConvAutoencoder, syn.cae_state, {}),                               # predict_hybrid.py:133
])
def test_state_dict_contract(cls, factory, kwargs, tmp_path):
    model = cls(**kwargs)
    sd = factory(0)
    own = model.state_dict()
    assert list(own.keys()) == list(sd.keys())                            # the reference's key names, in order
    assert all(tuple(own[k].shape) == tuple(sd[k].shape) for k in sd)
    model.load_state_dict(_t(sd))                                          # strict
    # both checkpoint shapes the scoring scripts accept (predict.py:82-85), loadable with weights_only=True
    for payload in ({"model_state": model.state_dict(), "epoch": 3}, model.state_dict()):
        p = tmp_path / "ckpt.pt"
        torch.save(payload, p)
        ck = torch.load(p, map_location="cpu", weights_only=True)
        cls(**kwargs).load_state_dict(ck["model_state"] if "model_state" in ck else ck)


def test_train_mode_uses_torch_layers_and_eval_cpu_refuses():
    x = torch.randn(2, 321, 180)
    model = m2.CNN2D()
    model.train()
    logits, emb = model(x, return_embedding=True)
    assert tuple(logits.shape) == (2, 1) and tuple(emb.shape) == (2, 23040)          # model.py:45-49 smoke shapes
    assert tuple(m1.CNN1D().train()(x).shape) == (2, 1)
    recon, latent = mc.ConvAutoencoder().train()(x)
    assert tuple(recon.shape) == (2, 321, 180) and tuple(latent.shape) == (2, 256, 20, 11)
    model.eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(x)


def test_train_mode_matches_oracle_in_eval_semantics():
    """The torch layer tree of the drop-in is the reference architecture: with BN in eval it reproduces the golden logits."""
    G = np.load(os.path.join(GOLDEN, "models.npz"))
    x = torch.from_numpy(syn.features(2, seed=1234))
    model = m2.CNN2D()
    model.load_state_dict(_t(syn.cnn2d_state(0)))
    model.train()
    for mod in model.modules():
        if isinstance(mod, (torch.nn.BatchNorm2d, torch.nn.Dropout)):
            mod.eval()
    with torch.no_grad():
        logits = model(x).squeeze(-1).numpy()
    np.testing.assert_allclose(logits, G["cnn2d_init_logits"][:2], rtol=1e-5, atol=1e-6)


def test_prediction_pickle_format(tmp_path):
    import pandas as pd
    with open(os.path.join(GOLDEN, "prediction_format.json")) as f:
        facts = json.load(f)
    df = scoring.write_predictions(["raw_1", "raw_2", "raw_3"], np.array([0.1, 0.5, 0.9], dtype=np.float32), tmp_path / "prediction.pkl")
    back = pd.read_pickle(tmp_path / "prediction.pkl")
    assert list(back.columns) == facts["columns"]
    assert {c: str(t) for c, t in back.dtypes.items()} == facts["dtypes"]
    assert type(back.index).__name__ == facts["index_type"]
    assert back["predictions"].iloc[0] == float(np.float32(0.1))                     # fp32 score widened exactly
    with pytest.raises(ValueError):
        scoring.write_predictions(["a"], [0.1, 0.2], tmp_path / "x.pkl")              # predict.py:113-114
    assert len(df) == 3


def test_precision_selection_on_the_dropin_classes(monkeypatch):
    """`precision` is an instance / class attribute with the DFS_B200_PRECISION environment variable as the default; it is part of
    the scorer cache key (changing it rebuilds the native handle) and never part of the state dict."""
    monkeypatch.delenv("DFS_B200_PRECISION", raising=False)
    for cls in (m2.CNN2D, m1.CNN1D, mc.ConvAutoencoder):
        net = cls()
        assert net._precision() == "fp16"
        dev = torch.device("cuda", 0)
        k16 = net._weights_key(dev) + (net._precision(),)
        monkeypatch.setenv("DFS_B200_PRECISION", "fp32")
        assert net._precision() == "fp32"
        net.precision = "fp16"                                            # the attribute wins over the environment
        assert net._precision() == "fp16"
        net.precision = "fp32"
        assert net._weights_key(dev) + (net._precision(),) != k16
        assert not any("precision" in k for k in net.state_dict())
        net.precision = "split"                                           # the 2D-CNN's accurate tensor-core mode; the others take fp32
        assert net._precision() == ("split" if cls is m2.CNN2D else "fp32")
        monkeypatch.delenv("DFS_B200_PRECISION")
    from dfs_b200 import engine
    with pytest.raises(ValueError, match="precision must be"):
        engine._check_precision("bf16")
    with pytest.raises(ValueError, match="precision must be"):
        engine._check_precision("split")                                  # only where a scorer lists it (Cnn2dScorer)
    engine._check_precision("split", ("fp16", "fp32", "split"))


def test_hostmem_helpers():
    """dfs_b200.hostmem: sysfs cpulist parsing, and numa_local() is a no-op context manager on hosts without a second node or
    without a known GPU node (this container, and the KVM guests of the GPU pool: profiles/r02h_h2d_bw_n8.txt)."""
    from dfs_b200 import hostmem as H
    assert H.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11] and H.parse_cpulist("") == []
    assert isinstance(H.host_nodes(), list)
    import os as _os
    before = _os.sched_getaffinity(0)
    with H.numa_local(0) as ctx:
        assert set(ctx.applied) == {"node", "mempolicy", "cpus"}
    assert _os.sched_getaffinity(0) == before


def test_host_slab_validates_dtype_device_and_layout():
    """engine._host_slab is the gate in front of dfs_score_host / dfs_group_score_host: fp32 and fp16 pass through in place, every
    other dtype is converted to fp32 (its bytes must never be read as fp32), non-dense layouts and wrong shapes raise."""
    from dfs_b200 import engine as E
    x = torch.randn(3, 321, 180)
    s = E._host_slab(x)
    assert (s.n, s.f16, s.time_major, s.strides, s.ptr) == (3, False, 0, (57780, 180, 1), x.data_ptr())
    s64 = E._host_slab(x.double())                                         # float64 -> converted copy, not reinterpreted
    assert not s64.f16 and s64.keep.dtype == torch.float32 and torch.equal(s64.keep, x)
    sbf = E._host_slab(x.bfloat16())
    assert sbf.keep.dtype == torch.float32 and not sbf.f16
    assert E._host_slab(np.zeros((2, 321, 180), np.float64)).keep.dtype == np.float32
    rows = torch.randn(2, 180, 321)                                        # the reference's storage, viewed as (B,321,180)
    sv = E._host_slab(rows.transpose(1, 2))
    assert (sv.time_major, sv.strides) == (1, (57780, 1, 321))
    s16 = E._host_slab(np.zeros((2, 321, 180), np.float16))
    assert s16.f16 and s16.strides == (57780, 180, 1)
    with pytest.raises(ValueError, match="shape"):
        E._host_slab(torch.zeros(3, 180, 321))
    with pytest.raises(ValueError, match="dense"):
        E._host_slab(torch.zeros(6, 321, 180)[::2])
    with pytest.raises(ValueError, match="dense"):
        E._host_slab(torch.zeros(2, 321, 360)[:, :, ::2])
    if torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="HOST features"):
            E._host_slab(torch.zeros(1, 321, 180, device="cuda"))
    else:
        with pytest.raises(RuntimeError, match="HOST features"):
            E._host_slab(torch.zeros(1, 321, 180, device="meta"))


def test_compare_with_existing_aligns_rows_by_uttid():
    """predict_hybrid.py:166-207: the report compares the new HYBRID predictions with an older file, rows matched by uttid
    (a re-ordered or shorter old file must not shift the comparison); both a bare DataFrame and a submission dict are accepted."""
    import pandas as pd
    import predict_hybrid as ph
    new = pd.DataFrame({"uttid": ["a", "b", "c", "d"], "predictions": [0.9, 0.2, 0.6, 0.4]})
    old = pd.DataFrame({"uttid": ["d", "c", "a"], "predictions": [0.7, 0.1, 0.8]})
    lines = []
    for existing in (old, {"student_id": 1, "predictions": old}):
        lines.clear()
        r = ph.compare_with_existing(new, existing, out=lines.append)
        assert r["n"] == 3 and r["agree"] == 1 and sorted(r["disagreements"]) == ["c", "d"]
        np.testing.assert_allclose(sorted(r["diff"]), sorted([0.9 - 0.8, 0.6 - 0.1, 0.4 - 0.7]))
        assert any("class agreement: 1/3 (33.3%)" in ln for ln in lines)
        assert any("c: old=0.1000 new=0.6000" in ln for ln in lines)
