"""GPU parity for the StatsPool detector (src/dlqueen_model.py DeepfakeDetector) through the C ABI: golden logits of the
unmodified reference class, the float64 oracle on fresh seeds, ragged lengths, chunking, and the drop-in module."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from conftest import GOLDEN, PKG  # noqa: E402

sys.path.insert(0, os.path.join(PKG, "dropin"))
import dfs_b200 as D  # noqa: E402
from dfs_b200 import synthetic as syn  # noqa: E402
import dlqueen_model as dq  # noqa: E402
from oracle import models_np as onp  # noqa: E402

G = np.load(os.path.join(GOLDEN, "dlq.npz"))
REL = 1e-3


def _inputs():
    x = syn.features(int(G["n"]), seed=1234)
    lengths = G["lengths"]
    xz = x.copy()
    for i, l in enumerate(lengths):
        xz[i, l:, :] = 0
    return x, xz, lengths


def _sig(z):
    return 1.0 / (1.0 + np.exp(-z))


@pytest.mark.parametrize("tag,scale", [("init", 1.0), ("trained", 300.0)])
def test_dlq_matches_reference_golden(tag, scale):
    x, xz, lengths = _inputs()
    sc = D.DlqScorer(syn.dlq_state(0, logit_scale=scale), max_chunk=5)               # 12 utterances through ragged passes of 5
    full = sc.score(torch.from_numpy(x).cuda()).cpu().numpy()
    ragged = sc.score(torch.from_numpy(xz).cuda(), lengths).cpu().numpy()
    for got, key in ((full, "logits_full"), (ragged, "logits_ragged")):
        ref = G[f"dlq_{tag}_{key}"]
        assert np.max(np.abs(got - ref)) <= 2e-3 * max(1.0, np.abs(ref).max()), (tag, key, got, ref)
        assert np.max(np.abs(_sig(got) - _sig(ref)) / _sig(ref)) <= REL
    # sigmoid output and the reference's own storage order (B, 180, T) read through strides
    xt = torch.from_numpy(xz).cuda().transpose(1, 2).contiguous().transpose(1, 2)
    np.testing.assert_allclose(sc.score(xt, lengths).cpu().numpy(), ragged, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(sc.score(torch.from_numpy(x).cuda(), apply_sigmoid=True).cpu().numpy(), _sig(full), rtol=1e-5)


def test_dlq_vs_oracle_fresh_seed_and_one_pass():
    x = syn.features(40, seed=77)
    sd = syn.dlq_state(3)
    ref = onp.dlq_forward(sd, x)
    a = D.DlqScorer(sd).score(torch.from_numpy(x).cuda()).cpu().numpy()               # one pass
    b = D.DlqScorer(sd, max_chunk=16).score(torch.from_numpy(x).cuda()).cpu().numpy()  # three passes
    np.testing.assert_array_equal(a, b)
    assert np.max(np.abs(_sig(a) - _sig(ref)) / _sig(ref)) <= REL
    assert np.max(np.abs(a - ref)) <= 2e-3


def test_dlq_dropin_module_forward():
    x, xz, lengths = _inputs()
    model = dq.DeepfakeDetector(in_ch=180, hidden=256, dropout=0.3).cuda()
    model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in syn.dlq_state(0).items()})
    model.eval()
    xb = torch.from_numpy(xz).transpose(1, 2).contiguous().cuda()                    # (B, C, T) like collate_fn (dlqueen_model.py:98-103)
    with torch.no_grad():
        logits = model(xb, torch.from_numpy(lengths).cuda())                         # run_inference loop body, :222-224
    ref = G["dlq_init_logits_ragged"]
    assert tuple(logits.shape) == (12,)
    assert np.max(np.abs(logits.cpu().numpy() - ref)) <= 2e-3
    eer = D.calculate_eer(logits.cpu().numpy(), (np.arange(12) % 2))                  # evaluate_eer, :247-251
    assert 0.0 <= eer[0] <= 1.0


def test_dlq_layer1_on_cta_pairs_is_bit_identical():
    """Option "pair_mma" (default 1): layer 1 as tcgen05 cta_group::2 MMAs (2 groups of N = 128, 64 weight rows per CTA) against 4
    groups of N = 64 on single CTAs: same K order, so the logits must agree bit for bit.  40 utterances = 3 column tiles (odd: the
    last pair has a padding unit), ragged lengths, then passes of 16."""
    x = torch.from_numpy(syn.features(40, seed=78)).cuda()
    lengths = torch.tensor([321 - (7 * i) % 200 for i in range(40)], dtype=torch.int32).cuda()
    for chunk in (0, 16):
        sc = D.DlqScorer(syn.dlq_state(4), max_chunk=chunk)
        sc.set_option("pair_mma", 0)
        a, al = sc.score(x), sc.score(x, lengths=lengths)
        sc.set_option("pair_mma", 1)
        assert torch.equal(sc.score(x), a) and torch.equal(sc.score(x, lengths=lengths), al)
