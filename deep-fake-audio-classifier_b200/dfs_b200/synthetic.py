"""Seeded synthetic weights, features and labels for the scoring path.

The reference ships no trained weights, no features and no labels, so
every parity test and the benchmark run on *random-init weights of the named
architecture* and synthetic ``[N, 321, 180]`` LFCC-like maps (BASELINE.json north_star).

Everything here is pure numpy on a PCG64 stream so the same (seed) gives the same
arrays in the build container and on the GPU box.  State-dict key names and tensor
shapes follow the reference exactly:

* CNN2D  – /root/reference/src/model.py:14-31   (conv.{0,5,10}, BN conv.{1,6,11}, classifier)
* CNN1D  – /root/reference/src/model_cnn1d.py:14-35 (conv.{0,4,8}, BN conv.{1,5,9}, classifier)
* CAE    – /root/reference/src/model_cae.py:32-81 (encoder.{0,4,8,12}, BN encoder.{1,5,9,13},
           decoder.{0,3,6,9}, BN decoder.{1,4,7})

BatchNorm running stats / affine are randomised (gamma~U(.5,1.5), beta,mu~N(0,.1),
var~U(.5,1.5)): the default init makes BN an identity and would hide folding bugs
(SURVEY.md §7.1 step 1).
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict

import numpy as np

T_FRAMES = 321
N_FEATS = 180
FEATURE_STD = 3.2  # real data: std 3.19 (results/archive/20260206_final_prep/model_prediction_report.md:28)


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def _uniform_fan_in(rng, shape, fan_in):
    bound = 1.0 / np.sqrt(float(fan_in))
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def _bn(rng, sd, prefix, c):
    sd[prefix + ".weight"] = rng.uniform(0.5, 1.5, size=c).astype(np.float32)
    sd[prefix + ".bias"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
    sd[prefix + ".running_mean"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
    sd[prefix + ".running_var"] = rng.uniform(0.5, 1.5, size=c).astype(np.float32)
    sd[prefix + ".num_batches_tracked"] = np.array(100, dtype=np.int64)


def cnn2d_state(seed: int = 0, in_features: int = N_FEATS, base_channels: int = 32,
                logit_scale: float = 1.0, classifier_bias: float | None = None) -> "OrderedDict[str, np.ndarray]":
    """Random-init CNN2D state dict. ``logit_scale`` rescales the classifier weights and ``classifier_bias`` replaces the
    bias ("trained-like" regime of SURVEY.md §7.2 #5: logits centred on 0 and spanning +-20 instead of 0.06 +- 0.01; the
    pair is calibrated against the unmodified reference by tests/golden/make_golden.py and stored in trained.npz)."""
    rng = _rng(1000 + seed)
    bc = base_channels
    sd = OrderedDict()
    chans = [(1, bc), (bc, 2 * bc), (2 * bc, 4 * bc)]
    for (ci, co), conv_i, bn_i in zip(chans, (0, 5, 10), (1, 6, 11)):
        sd[f"conv.{conv_i}.weight"] = _uniform_fan_in(rng, (co, ci, 3, 3), ci * 9)
        sd[f"conv.{conv_i}.bias"] = _uniform_fan_in(rng, (co,), ci * 9)
        _bn(rng, sd, f"conv.{bn_i}", co)
    fan = 4 * bc * in_features
    sd["classifier.weight"] = _uniform_fan_in(rng, (1, fan), fan) * np.float32(logit_scale)
    sd["classifier.bias"] = _uniform_fan_in(rng, (1,), fan)
    if classifier_bias is not None:
        sd["classifier.bias"] = np.array([classifier_bias], dtype=np.float32)
    return sd


def cnn1d_state(seed: int = 0, in_features: int = N_FEATS, base_channels: int = 32,
                logit_scale: float = 1.0, classifier_bias: float | None = None) -> "OrderedDict[str, np.ndarray]":
    rng = _rng(2000 + seed)
    bc = base_channels
    sd = OrderedDict()
    chans = [(in_features, bc), (bc, 2 * bc), (2 * bc, 4 * bc)]
    for (ci, co), conv_i, bn_i in zip(chans, (0, 4, 8), (1, 5, 9)):
        sd[f"conv.{conv_i}.weight"] = _uniform_fan_in(rng, (co, ci, 3), ci * 3)
        sd[f"conv.{conv_i}.bias"] = _uniform_fan_in(rng, (co,), ci * 3)
        _bn(rng, sd, f"conv.{bn_i}", co)
    sd["classifier.weight"] = _uniform_fan_in(rng, (1, 4 * bc), 4 * bc) * np.float32(logit_scale)
    sd["classifier.bias"] = _uniform_fan_in(rng, (1,), 4 * bc)
    if classifier_bias is not None:
        sd["classifier.bias"] = np.array([classifier_bias], dtype=np.float32)
    return sd


def dlq_state(seed: int = 0, in_ch: int = N_FEATS, hidden: int = 256, logit_scale: float = 1.0) -> "OrderedDict[str, np.ndarray]":
    """Random-init DeepfakeDetector state dict (/root/reference/src/dlqueen_model.py:132-173): enc.net.{0,4,8} Conv1d
    (k = 5, 3, 3), enc.net.{1,5,9} BatchNorm1d, head.{0,3} Linear."""
    rng = _rng(4000 + seed)
    sd = OrderedDict()
    for (ci, k), conv_i, bn_i in zip(((in_ch, 5), (hidden, 3), (hidden, 3)), (0, 4, 8), (1, 5, 9)):
        sd[f"enc.net.{conv_i}.weight"] = _uniform_fan_in(rng, (hidden, ci, k), ci * k)
        sd[f"enc.net.{conv_i}.bias"] = _uniform_fan_in(rng, (hidden,), ci * k)
        _bn(rng, sd, f"enc.net.{bn_i}", hidden)
    sd["head.0.weight"] = _uniform_fan_in(rng, (hidden, 2 * hidden), 2 * hidden)
    sd["head.0.bias"] = _uniform_fan_in(rng, (hidden,), 2 * hidden)
    sd["head.3.weight"] = _uniform_fan_in(rng, (1, hidden), hidden) * np.float32(logit_scale)
    sd["head.3.bias"] = _uniform_fan_in(rng, (1,), hidden)
    return sd


def cae_state(seed: int = 0, base_channels: int = 32) -> "OrderedDict[str, np.ndarray]":
    rng = _rng(3000 + seed)
    bc = base_channels
    sd = OrderedDict()
    enc = [(1, bc), (bc, 2 * bc), (2 * bc, 4 * bc), (4 * bc, 8 * bc)]
    for (ci, co), conv_i, bn_i in zip(enc, (0, 4, 8, 12), (1, 5, 9, 13)):
        sd[f"encoder.{conv_i}.weight"] = _uniform_fan_in(rng, (co, ci, 3, 3), ci * 9)
        sd[f"encoder.{conv_i}.bias"] = _uniform_fan_in(rng, (co,), ci * 9)
        _bn(rng, sd, f"encoder.{bn_i}", co)
    dec = [(8 * bc, 4 * bc), (4 * bc, 2 * bc), (2 * bc, bc), (bc, 1)]
    for (ci, co), conv_i, bn_i in zip(dec, (0, 3, 6, 9), (1, 4, 7, None)):
        # ConvTranspose2d weight is (Cin, Cout, 2, 2); torch's fan_in for it is Cout*kh*kw
        sd[f"decoder.{conv_i}.weight"] = _uniform_fan_in(rng, (ci, co, 2, 2), co * 4)
        sd[f"decoder.{conv_i}.bias"] = _uniform_fan_in(rng, (co,), co * 4)
        if bn_i is not None:
            _bn(rng, sd, f"decoder.{bn_i}", co)
    return sd


def normalizer_stats(seed: int = 1, n_feats: int = N_FEATS):
    """mean180 ~ N(0, .1), std180 ~ U(2, 4) -> z-scored input ~ N(0,1)."""
    rng = _rng(4000 + seed)
    mean = (0.1 * rng.standard_normal(n_feats)).astype(np.float32)
    std = rng.uniform(2.0, 4.0, size=n_feats).astype(np.float32)
    return mean, std


def features(n: int, seed: int = 1234, start: int = 0) -> np.ndarray:
    """``[n, 321, 180]`` fp32 ~ N(0, 3.2^2); utterance ``i`` depends only on (seed, start+i)."""
    out = np.empty((n, T_FRAMES, N_FEATS), dtype=np.float32)
    for i in range(n):
        rng = np.random.Generator(np.random.PCG64([seed, start + i]))
        out[i] = (FEATURE_STD * rng.standard_normal((T_FRAMES, N_FEATS), dtype=np.float32))
    return out


def features_structured(n: int, seed: int = 4321, start: int = 0, heavy_tail: bool = True) -> np.ndarray:
    """``[n, 321, 180]`` fp32 utterances that DIFFER from one another the way real LFCC maps do -- unlike ``features()``,
    whose i.i.d. noise gives every utterance nearly the same embedding.  Per utterance (seeded by (seed, start+i)):
    a gain (lognormal around the real std 3.19), a spectral tilt over the 180 features, a slow temporal envelope, a
    per-feature offset pattern, and (``heavy_tail``) a dozen outliers reaching the real data's range -61 ... +86
    (results/archive/20260206_final_prep/model_prediction_report.md:24-29)."""
    out = np.empty((n, T_FRAMES, N_FEATS), dtype=np.float32)
    f = np.arange(N_FEATS, dtype=np.float32) / np.float32(N_FEATS - 1)
    t = np.arange(T_FRAMES, dtype=np.float32) / np.float32(T_FRAMES)
    for i in range(n):
        rng = np.random.Generator(np.random.PCG64([seed, start + i]))
        gain = np.float32(np.clip(FEATURE_STD * np.exp(0.5 * rng.standard_normal()), 1.0, 8.0))
        tilt = (1.0 + rng.uniform(-1.0, 1.0) * (f - 0.5)).astype(np.float32)
        env = (1.0 + 0.5 * np.sin(2.0 * np.pi * (rng.uniform(0.5, 3.0) * t + rng.random()))).astype(np.float32)
        offs = (rng.standard_normal() * np.cos(2.0 * np.pi * rng.uniform(1.0, 6.0) * f)).astype(np.float32)
        x = gain * rng.standard_normal((T_FRAMES, N_FEATS), dtype=np.float32) * tilt[None, :] * env[:, None] + offs[None, :]
        if heavy_tail:
            k = 12
            tt, ff = rng.integers(0, T_FRAMES, k), rng.integers(0, N_FEATS, k)
            mag = rng.uniform(30.0, 86.0, k)
            sign = np.where(rng.random(k) < 0.5, -1.0, 1.0)
            x[tt, ff] = np.where(sign < 0, -np.minimum(mag, 61.0), mag).astype(np.float32)
        out[i] = x
    return out


def labels(n: int, seed: int = 42, p_bonafide: float = 0.4) -> np.ndarray:
    return (_rng(5000 + seed).random(n) < p_bonafide).astype(np.uint8)


def tie_free_scores(n: int, seed: int = 0, k: float = 6.0):
    """Distinct fp32 scores + labels with an EER strictly between 0 and 0.5.

    Scores are distinct fp32 *bit patterns* in a positive range ([0.25, 0.75) when n fits,
    else [2^-20, 1)), assigned through a seeded affine bijection ``rank = (a*i + b) mod n``
    (gcd(a, n) = 1), so no two scores tie -- drawing ``rand()`` would collide massively at
    100 M.  Labels are Bernoulli(sigmoid(k * (rank/n - 0.5))).
    """
    import math
    rng = _rng(6000 + seed)
    lo, hi = (int(np.float32(v).view(np.uint32)) for v in (0.25, 0.75))
    if n > hi - lo:
        lo, hi = (int(np.float32(v).view(np.uint32)) for v in (2.0 ** -20, 1.0))
    if n > hi - lo:
        raise ValueError("n too large for distinct positive fp32 scores")
    step = (hi - lo) // max(n, 1)
    a = int(rng.integers(n // 3 + 1, max(n // 2, n // 3 + 2))) | 1
    while math.gcd(a, max(n, 1)) != 1:
        a += 2
    b = int(rng.integers(0, max(n, 1)))
    idx = np.arange(n, dtype=np.uint64)
    ranks = (idx * np.uint64(a) + np.uint64(b)) % np.uint64(max(n, 1))   # a*n < 2^64 for n < 2^32
    scores = (np.uint64(lo) + ranks * np.uint64(step)).astype(np.uint32).view(np.float32)
    p = 1.0 / (1.0 + np.exp(-k * (ranks.astype(np.float64) / max(n, 1) - 0.5)))
    lab = (rng.random(n) < p).astype(np.uint8)
    return scores, lab


def state_digest(sd) -> str:
    """sha256 over the arrays of a state dict / list of arrays (goldens pin the factory)."""
    h = hashlib.sha256()
    items = sd.items() if hasattr(sd, "items") else enumerate(sd)
    for k, v in items:
        h.update(str(k).encode())
        h.update(np.ascontiguousarray(v).tobytes())
    return h.hexdigest()
