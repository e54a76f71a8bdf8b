"""GPU parity for the metric half: device radix sort + FAR/FRR sweep vs the reference's goldens
(bit-exact on tie-free scores) and vs the kind="stable" oracle on tie-heavy scores; blend known
answer; confusion counts."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from conftest import GOLDEN  # noqa: E402
import dfs_b200 as D  # noqa: E402
from dfs_b200 import synthetic as syn  # noqa: E402
from oracle import eer as oeer  # noqa: E402

E = np.load(os.path.join(GOLDEN, "eer_cases.npz"))
CASES = sorted({k.split("/")[0] for k in E.files})


@pytest.mark.parametrize("case", CASES)
def test_eer_goldens(case):
    s, l = E[case + "/scores"], E[case + "/labels"]
    ref_eer, ref_thr = E[case + "/eer_thr"]
    d = D.eer_details(s, l, want_perm=True, want_sorted=True)
    stable = oeer.eer_details(s, l, kind="stable")
    # the device contract: identical to the reference algorithm with a stable argsort
    assert (d["eer"], d["threshold"]) == (stable["eer"], stable["threshold"])
    assert d["eer_idx"] == stable["eer_idx"]
    assert np.array_equal(d["perm"].cpu().numpy().astype(np.int64), stable["perm"])
    assert np.array_equal(d["sorted"].cpu().numpy(), np.array(s)[stable["perm"]])
    if not case.startswith("ties_mixed"):
        # tie-free (or single-label tie groups): bit-exact with the unmodified reference itself
        assert (d["eer"], d["threshold"]) == (ref_eer, ref_thr)
    if case.startswith("tiefree"):
        assert np.array_equal(d["perm"].cpu().numpy().astype(np.int64), E[case + "/ref_argsort"])
    # the radix-select path must agree with the sort path bit for bit on every golden
    for method in ("select", "sort"):
        e = D.eer_details(s, l, method=method)
        assert (e["eer"], e["threshold"], e["eer_idx"], e["n_bonafide"], e["n_spoof"]) == \
               (d["eer"], d["threshold"], d["eer_idx"], d["n_bonafide"], d["n_spoof"]), method
    thr = d["threshold"]
    assert D.confusion_at_threshold(s, l, thr) == oeer.confusion_at_threshold(s, l, thr)
    if not case.startswith("ties_mixed"):
        assert tuple(D.confusion_at_threshold(s, l, ref_thr)[:4]) == tuple(E[case + "/confusion"])


def test_eer_accepts_lists_and_tensors():
    assert D.calculate_eer([0.1, 0.2, 0.8, 0.9], [0, 0, 1, 1]) == (0.0, 0.2)
    assert D.calculate_eer([0.1, 0.2, 0.8, 0.9], [1, 1, 0, 0])[0] == 1.0
    assert D.calculate_eer([0.3, 0.4], [1, 1]) == (0.0, 0.0)
    s = torch.tensor([0.1, 0.2, 0.8, 0.9], device="cuda")
    lab = torch.tensor([0.0, 0.0, 1.0, 1.0], device="cuda")          # float labels as evaluate() passes them
    eer, thr = D.calculate_eer(s, lab)
    assert eer == 0.0 and thr == float(np.float32(0.2))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_eer_threshold_edges(dtype):
    # eer_idx == 0 and eer_idx == n use +-1e-6 in the score dtype (evaluation.py:31-35)
    for s, l in (([0.9, 0.8, 0.1], [0, 0, 1]), ([0.1, 0.2, 0.3, 0.9], [1, 0, 0, 0])):
        s = np.array(s, dtype=dtype)
        assert D.calculate_eer(s, l) == oeer.calculate_eer(s, l)


@pytest.mark.parametrize("n", [1, 2, 4095, 4096, 4097, 1_000_003])
def test_eer_sizes_and_negative_scores(n):
    rng = np.random.default_rng(n)
    s = rng.standard_normal(n).astype(np.float32) * 4
    s[::7] = -s[::7]
    l = (rng.random(n) < 0.45).astype(np.uint8)
    d = D.eer_details(s, l, want_perm=True)
    o = oeer.eer_details(s, l, kind="stable")
    assert (d["eer"], d["threshold"], d["eer_idx"]) == (o["eer"], o["threshold"], o["eer_idx"])
    assert np.array_equal(d["perm"].cpu().numpy().astype(np.int64), o["perm"])
    e = D.eer_details(s, l, method="select")
    assert (e["eer"], e["threshold"], e["eer_idx"]) == (o["eer"], o["threshold"], o["eer_idx"])


def test_eer_ten_million_tie_free_bit_exact():
    s, l = syn.tie_free_scores(10_000_000, seed=5)
    o = oeer.eer_details(s, l)            # the reference's own (unstable) argsort: tie-free => same order
    for method in ("sort", "select"):
        d = D.eer_details(s, l, method=method)
        assert (d["eer"], d["threshold"], d["eer_idx"]) == (o["eer"], o["threshold"], o["eer_idx"]), method
    assert 0.05 < d["eer"] < 0.45


def _select_vs_oracle(s, l):
    o = oeer.eer_details(s, l, kind="stable")
    for method in ("select", "sort"):
        d = D.eer_details(s, l, method=method)
        assert (d["eer"], d["threshold"], d["eer_idx"]) == (o["eer"], o["threshold"], o["eer_idx"]), (method, d, o)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("levels", [2, 3, 17, 1000])
def test_eer_select_tie_groups(dtype, levels):
    """Heavily tied scores: the crossing falls inside a mixed-label tie group, whose stable (index) order decides the
    argmin -- the compaction path of dfs_eer_select."""
    rng = np.random.default_rng(levels)
    for n in (50, 4097, 300_001):
        s = (rng.integers(0, levels, n) / levels - 0.3).astype(dtype)
        l = (rng.random(n) < 0.2 + 0.6 * (s - s.min()) / (np.ptp(s) + 1e-9)).astype(np.uint8)
        if l.min() == l.max():
            l[0], l[1] = 0, 1
        _select_vs_oracle(s, l)
    # every score equal: one group holds the whole curve
    s = np.full(70_001, 0.25, dtype=dtype)
    l = (rng.random(s.size) < 0.5).astype(np.uint8)
    _select_vs_oracle(s, l)


def test_eer_select_crossing_at_group_boundaries():
    """The chosen curve point is the START of the group that holds the first negative difference: the threshold is the
    largest score below that group (predecessor pass), incl. eer_idx = 0 and eer_idx = n."""
    _select_vs_oracle(np.array([0.1, 0.2, 0.8, 0.9], dtype=np.float32), np.array([0, 0, 1, 1], dtype=np.uint8))
    _select_vs_oracle(np.array([0.1, 0.2, 0.8, 0.9], dtype=np.float64), np.array([1, 1, 0, 0], dtype=np.uint8))
    _select_vs_oracle(np.array([0.9, 0.8, 0.1], dtype=np.float32), np.array([0, 0, 1], dtype=np.uint8))
    _select_vs_oracle(np.array([0.1, 0.2, 0.3, 0.9], dtype=np.float64), np.array([1, 0, 0, 0], dtype=np.uint8))
    _select_vs_oracle(np.array([-1.5, -0.0, 0.0, 2.0, 2.0, 7.0], dtype=np.float32), np.array([0, 1, 0, 1, 0, 1], dtype=np.uint8))
    rng = np.random.default_rng(11)
    for n in (3, 10, 257, 65_537):
        for _ in range(6):
            s = rng.standard_normal(n).astype(np.float32)
            l = (rng.random(n) < 0.5).astype(np.uint8)
            if l.min() == l.max():
                l[0], l[-1] = 0, 1
            _select_vs_oracle(s, l)
    # logits-like scores sharing most key bytes (constant bytes are skipped via key AND / OR)
    s = (0.5 + rng.random(200_000) * 1e-3).astype(np.float32)
    l = (rng.random(s.size) < 0.5).astype(np.uint8)
    _select_vs_oracle(s, l)


def test_eer_select_misaligned_device_pointers():
    rng = np.random.default_rng(3)
    n = 100_003
    s = torch.from_numpy(rng.standard_normal(n + 1).astype(np.float32)).cuda()[1:]      # 4-byte aligned only
    l = torch.from_numpy((rng.random(n + 3) < 0.4).astype(np.uint8)).cuda()[3:]         # 1-byte aligned only
    o = oeer.eer_details(s.cpu().numpy(), l.cpu().numpy(), kind="stable")
    d = D.eer_details(s, l, method="select")
    assert (d["eer"], d["threshold"], d["eer_idx"]) == (o["eer"], o["threshold"], o["eer_idx"])


def test_eer_select_single_class():
    for lab in (0, 1):
        d = D.eer_details(np.linspace(0, 1, 1000, dtype=np.float32), np.full(1000, lab, dtype=np.uint8), method="select")
        assert (d["eer"], d["threshold"], d["eer_idx"]) == (0.0, 0.0, -1)


def test_eer_hundred_million_properties():
    """BASELINE config 5 size: checked through size-independent properties (the CPU oracle needs ~40 s
    and 10 GB here): sortedness, permutation checksum, and FAR/FRR consistency at the reported index."""
    n = 100_000_000
    s, l = syn.tie_free_scores(n, seed=6)
    sd, ld = torch.from_numpy(s).cuda(), torch.from_numpy(l).cuda()
    d = D.eer_details(sd, ld, want_perm=True, want_sorted=True)
    srt, perm = d["sorted"], d["perm"].long()
    assert bool((srt[1:] > srt[:-1]).all())                              # strictly sorted (tie-free)
    assert int(perm.sum()) == n * (n - 1) // 2                           # a permutation of 0..n-1
    assert bool((sd[perm[:1000]] == srt[:1000]).all()) and bool((sd[perm[-1000:]] == srt[-1000:]).all())
    k = d["eer_idx"]
    c1 = int(ld[perm[:k]].sum())
    far = (d["n_spoof"] - (k - c1)) / d["n_spoof"]
    frr = c1 / d["n_bonafide"]
    assert d["eer"] == (far + frr) / 2.0
    assert d["threshold"] == float(srt[k - 1])
    assert abs(far - frr) < 1e-6
    e = D.eer_details(sd, ld, method="select")                           # the radix-select path, same answer
    assert (e["eer"], e["threshold"], e["eer_idx"]) == (d["eer"], d["threshold"], d["eer_idx"])


def test_blend_known_answer_bit_exact():
    B = np.load(os.path.join(GOLDEN, "blend_known_answer.npz"))
    alpha = float(B["alpha"])
    got = D.blend([B["sup"], B["cae_minmaxed"]], [alpha, 1 - alpha], [1, 0])
    assert np.array_equal(got, B["hybrid"])                              # the reference's shipped result, bit for bit
    assert np.array_equal(D.normalise_01(B["sup"]), oeer.normalise_01(B["sup"]))
    assert np.array_equal(D.hybrid_blend(B["sup"], B["cae_minmaxed"], alpha), oeer.hybrid_blend(B["sup"], B["cae_minmaxed"], alpha))
    assert np.array_equal(D.normalise_01(np.full(17, 0.25)), np.zeros(17))   # flat input -> zeros (predict_hybrid.py:83-84)


def test_ensemble_mean_and_fp32_inputs():
    rng = np.random.default_rng(3)
    a, b, c = (rng.random(5001).astype(np.float32) for _ in range(3))
    ref = oeer.ensemble_mean([a.astype(np.float64), b.astype(np.float64), c.astype(np.float64)])
    assert np.array_equal(D.ensemble_mean([a, b, c]), ref)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    assert np.array_equal(D.hybrid_blend(ta, tb, 0.8), oeer.hybrid_blend(a, b, 0.8))


def test_eer_select_tma_ring_matches_direct_loads():
    """The cp.async.bulk-fed histogram kernel and the direct-load one must pick the same crossing (sizes around the
    stage size 4096 / 2048 and the flush interval, fp32 and fp64)."""
    rng = np.random.default_rng(21)
    try:
        for dtype in (np.float32, np.float64):
            for n in (1, 2047, 4096, 4097, 14 * 4096 * 3 + 5, 2_000_003):
                s = torch.from_numpy((rng.standard_normal(n) * 3).astype(dtype)).cuda()
                l = torch.from_numpy((rng.random(n) < 0.5).astype(np.uint8)).cuda()
                if n == 1:
                    continue
                l[0], l[1] = 0, 1
                out = []
                for tma in (1, 0):
                    D._native.set_global_option("eer_select_tma", tma)
                    d = D.eer_details(s, l, method="select")
                    out.append((d["eer"], d["threshold"], d["eer_idx"], d["n_bonafide"]))
                d = D.eer_details(s, l, method="sort")
                assert out[0] == out[1] == (d["eer"], d["threshold"], d["eer_idx"], d["n_bonafide"]), (dtype, n)
    finally:
        D._native.set_global_option("eer_select_tma", 1)


def _sort_forms_agree(s, l, tag):
    """dfs_eer under every form of the radix passes (and with / without the speculative first pass, "eer_sort_overlap"): identical
    permutation, sorted scores and result."""
    out = []
    try:
        for form, overlap in ((0, 1), (1, 1), (2, 1), (3, 1), (4, 1), (5, 1), (1, 0), (2, 0), (4, 0)):
            D._native.set_global_option("eer_sort_onesweep", form)
            D._native.set_global_option("eer_sort_overlap", overlap)
            d = D.eer_details(s, l, want_perm=True, want_sorted=True)
            out.append(d)
    finally:
        D._native.set_global_option("eer_sort_onesweep", 1)
        D._native.set_global_option("eer_sort_overlap", 1)
    base = out[0]
    for form, d in enumerate(out[1:], start=1):
        assert torch.equal(d["perm"], base["perm"]), (tag, form)
        assert torch.equal(d["sorted"].view(torch.uint8), base["sorted"].view(torch.uint8)), (tag, form)   # bytes: NaN == NaN
        assert (d["eer"], d["threshold"], d["eer_idx"], d["n_bonafide"]) == \
               (base["eer"], base["threshold"], base["eer_idx"], base["n_bonafide"]) or np.isnan(d["threshold"]), (tag, form)
    return base


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_eer_sort_one_sweep_matches_super_tile_form_and_oracle(dtype):
    """The one-sweep passes (ticketed tiles + decoupled look-back; histogram by shared-memory atomics, by a kernel of its own or by
    ballots) against the count / scan / scatter form and the stable oracle: sizes around the tile (8,192 fp32 / 4,096 fp64 keys),
    many tiles (a long look-back chain), ties, NaN, a constant low key byte (the first pass is not byte 0) and negative scores."""
    rng = np.random.default_rng(77)
    tile = 8192 if dtype == np.float32 else 4096
    for n in (1, 2, 31, tile - 1, tile, tile + 1, 5 * tile + 3, 1_000_003, 600 * tile + 17):
        s = (rng.standard_normal(n) * 3).astype(dtype)
        l = (rng.random(n) < 0.4).astype(np.uint8)
        if n >= 2:
            l[0], l[1] = 0, 1
        base = _sort_forms_agree(torch.from_numpy(s).cuda(), torch.from_numpy(l).cuda(), ("normal", n))
        if n <= 1_000_003:
            o = oeer.eer_details(s, l, kind="stable")
            assert np.array_equal(base["perm"].cpu().numpy().astype(np.int64), o["perm"]), n
            assert (base["eer"], base["threshold"], base["eer_idx"]) == (o["eer"], o["threshold"], o["eer_idx"]), n
    n = 300_001
    l = (rng.random(n) < 0.5).astype(np.uint8)
    # heavy ties (17 levels), NaN scattered in, -0.0 / +0.0
    s = rng.integers(0, 17, n).astype(dtype) / 16
    s[::1001] = np.nan
    s[5::997] = -0.0
    base = _sort_forms_agree(torch.from_numpy(s).cuda(), torch.from_numpy(l).cuda(), "ties+nan")
    o = oeer.eer_details(s, l, kind="stable")
    assert np.array_equal(base["perm"].cpu().numpy().astype(np.int64), o["perm"])
    # low mantissa byte constant: the prep kernel's histogram of byte 0 is not the first pass's
    bits = np.dtype(dtype).itemsize * 8
    u = (rng.standard_normal(n) * 2).astype(dtype).view(np.uint32 if bits == 32 else np.uint64)
    s = (u & ~np.array(0xff, dtype=u.dtype)).view(dtype)
    base = _sort_forms_agree(torch.from_numpy(s).cuda(), torch.from_numpy(l).cuda(), "byte0 constant")
    o = oeer.eer_details(s, l, kind="stable")
    assert np.array_equal(base["perm"].cpu().numpy().astype(np.int64), o["perm"])
    # one value only: every pass is the identity
    s = np.full(n, 0.25, dtype=dtype)
    base = _sort_forms_agree(torch.from_numpy(s).cuda(), torch.from_numpy(l).cuda(), "constant")
    assert np.array_equal(base["perm"].cpu().numpy(), np.arange(n))
