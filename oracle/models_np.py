"""ORACLE (test infrastructure, never the product path): numpy restatement of the three scorers.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package (the shipped path is the CUDA extension and
fails loudly without it).

The arithmetic of the reference's models lives in a third-party dependency that is not under
/root/reference and is unpinned there ("No dependency lockfile is provided", AGENTS.md:34):
PyTorch's ``nn.Conv2d / Conv1d / ConvTranspose2d / BatchNorm / AvgPool / Linear`` (this image:
torch 2.11.0+cu128, CPU kernels).  This file restates their *published* eval-mode definitions in
plain numpy, in float64 by default (so it is an independent, higher-precision anchor), following
the reference's call sites line by line.  It is pinned by ``tests/golden/*.npz``: outputs of the
unmodified reference classes run in the build container on the same seeded weights and inputs
(``tests/golden/make_golden.py``).

All functions take a state dict of numpy arrays with the reference's key names.
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-5  # torch.nn.BatchNorm default, never overridden by the reference


def _bn_eval(x, sd, prefix, dtype):
    """y = (x - running_mean) / sqrt(running_var + eps) * weight + bias over channel axis 1."""
    shape = [1, -1] + [1] * (x.ndim - 2)
    mean = sd[prefix + ".running_mean"].astype(dtype).reshape(shape)
    var = sd[prefix + ".running_var"].astype(dtype).reshape(shape)
    g = sd[prefix + ".weight"].astype(dtype).reshape(shape)
    b = sd[prefix + ".bias"].astype(dtype).reshape(shape)
    return (x - mean) / np.sqrt(var + dtype(BN_EPS)) * g + b


def _conv2d_3x3_same(x, w, b):
    """nn.Conv2d(kernel_size=3, padding=1): x (B,Ci,H,W), w (Co,Ci,3,3), b (Co,) -> (B,Co,H,W)."""
    B, Ci, H, W = x.shape
    Co = w.shape[0]
    xp = np.zeros((B, Ci, H + 2, W + 2), dtype=x.dtype)
    xp[:, :, 1:-1, 1:-1] = x
    out = np.zeros((B, Co, H, W), dtype=x.dtype)
    for kh in range(3):
        for kw in range(3):
            patch = xp[:, :, kh:kh + H, kw:kw + W]                    # (B,Ci,H,W)
            out += np.einsum("bchw,oc->bohw", patch, w[:, :, kh, kw], optimize=True)
    return out + b.reshape(1, -1, 1, 1)


def _conv1d_k3_same(x, w, b):
    """nn.Conv1d(kernel_size=3, padding=1): x (B,Ci,L), w (Co,Ci,3) -> (B,Co,L)."""
    B, Ci, L = x.shape
    xp = np.zeros((B, Ci, L + 2), dtype=x.dtype)
    xp[:, :, 1:-1] = x
    out = np.zeros((B, w.shape[0], L), dtype=x.dtype)
    for k in range(3):
        out += np.einsum("bcl,oc->bol", xp[:, :, k:k + L], w[:, :, k], optimize=True)
    return out + b.reshape(1, -1, 1)


def _avgpool(x, kh, kw):
    """nn.AvgPool2d((kh,kw)) with floor: trailing rows/cols that do not fill a window are dropped."""
    B, C, H, W = x.shape
    Ho, Wo = H // kh, W // kw
    x = x[:, :, :Ho * kh, :Wo * kw].reshape(B, C, Ho, kh, Wo, kw)
    return x.mean(axis=(3, 5))


def _convT_k2s2(x, w, b, out_pad=(0, 0)):
    """nn.ConvTranspose2d(kernel_size=2, stride=2, output_padding=out_pad).

    out[n,co,2i+a,2j+b] = bias[co] + sum_ci x[n,ci,i,j] * W[ci,co,a,b]; rows/cols added by
    output_padding receive the bias only (measured exact vs torch)."""
    B, Ci, H, W = x.shape
    Co = w.shape[1]
    out = np.zeros((B, Co, 2 * H + out_pad[0], 2 * W + out_pad[1]), dtype=x.dtype)
    for a in range(2):
        for c in range(2):
            out[:, :, a:2 * H:2, c:2 * W:2] = np.einsum("bchw,co->bohw", x, w[:, :, a, c], optimize=True)
    return out + b.reshape(1, -1, 1, 1)


def _relu(x):
    return np.maximum(x, 0)


# --------------------------------------------------------------------------------------
# CNN2D  (/root/reference/src/model.py:12-42)
# --------------------------------------------------------------------------------------
def cnn2d_forward(sd, x, dtype=np.float64, return_embedding=False):
    """x (B, T=321, F=180) -> logits (B,1) [, embedding (B, 128*F)]."""
    dt = np.dtype(dtype).type
    h = np.asarray(x, dtype=dtype)[:, None, :, :]                                   # model.py:34
    for conv_i, bn_i, pool in ((0, 1, True), (5, 6, True), (10, 11, False)):         # model.py:14-30
        h = _conv2d_3x3_same(h, sd[f"conv.{conv_i}.weight"].astype(dtype), sd[f"conv.{conv_i}.bias"].astype(dtype))
        h = _relu(_bn_eval(h, sd, f"conv.{bn_i}", dt))
        if pool:
            h = _avgpool(h, 2, 1)                                                    # AvgPool2d((2,1)); Dropout = id in eval
    h = h.mean(axis=2)                                                               # model.py:37
    emb = h.reshape(h.shape[0], -1)                                                  # model.py:38  (index = c*F + f)
    logits = emb @ sd["classifier.weight"].astype(dtype).T + sd["classifier.bias"].astype(dtype)  # model.py:39
    if return_embedding:
        return logits, emb
    return logits


# --------------------------------------------------------------------------------------
# CNN1D  (/root/reference/src/model_cnn1d.py:12-46)
# --------------------------------------------------------------------------------------
def cnn1d_forward(sd, x, dtype=np.float64):
    dt = np.dtype(dtype).type
    h = np.asarray(x, dtype=dtype).transpose(0, 2, 1)                                # model_cnn1d.py:40
    for conv_i, bn_i in ((0, 1), (4, 5), (8, 9)):                                    # model_cnn1d.py:14-32
        h = _conv1d_k3_same(h, sd[f"conv.{conv_i}.weight"].astype(dtype), sd[f"conv.{conv_i}.bias"].astype(dtype))
        h = _relu(_bn_eval(h, sd, f"conv.{bn_i}", dt))
    h = h.mean(axis=2)                                                               # AdaptiveAvgPool1d(1) + flatten :43-44
    return h @ sd["classifier.weight"].astype(dtype).T + sd["classifier.bias"].astype(dtype)


# --------------------------------------------------------------------------------------
# ConvAutoencoder  (/root/reference/src/model_cae.py:23-125)
# --------------------------------------------------------------------------------------
def cae_forward(sd, x, dtype=np.float64):
    """x (B,321,180) *already normalised* -> (recon (B,321,180), latent (B,256,20,11))."""
    dt = np.dtype(dtype).type
    h = np.asarray(x, dtype=dtype)[:, None, :, :]                                    # model_cae.py:107
    for conv_i, bn_i in ((0, 1), (4, 5), (8, 9), (12, 13)):                          # model_cae.py:32-56
        h = _conv2d_3x3_same(h, sd[f"encoder.{conv_i}.weight"].astype(dtype), sd[f"encoder.{conv_i}.bias"].astype(dtype))
        h = _avgpool(_relu(_bn_eval(h, sd, f"encoder.{bn_i}", dt)), 2, 2)
    latent = h
    for conv_i, bn_i, opad in ((0, 1, (0, 0)), (3, 4, (0, 1)), (6, 7, (0, 0))):      # model_cae.py:61-76
        h = _convT_k2s2(h, sd[f"decoder.{conv_i}.weight"].astype(dtype), sd[f"decoder.{conv_i}.bias"].astype(dtype), opad)
        h = _relu(_bn_eval(h, sd, f"decoder.{bn_i}", dt))
    h = _convT_k2s2(h, sd["decoder.9.weight"].astype(dtype), sd["decoder.9.bias"].astype(dtype))  # :79, no BN/act
    T = x.shape[1]
    if h.shape[2] < T:                                                               # model_cae.py:116-119
        pad = np.zeros((h.shape[0], 1, T - h.shape[2], h.shape[3]), dtype=h.dtype)
        h = np.concatenate([h, pad], axis=2)
    elif h.shape[2] > T:
        h = h[:, :, :T, :]
    return h[:, 0], latent


def normalizer_transform(x, mean, std, dtype=np.float64):
    """FeatureNormalizer.transform (/root/reference/src/dataset_cae.py:37-41)."""
    return (np.asarray(x, dtype=dtype) - mean.astype(dtype)) / std.astype(dtype)


def cae_mse_scores(sd, x, mean=None, std=None, dtype=np.float64):
    """get_cae_scores inner loop (/root/reference/src/predict_hybrid.py:73-77):
    MSELoss(reduction='none')(recon, x).view(B,-1).mean(1), x normalised first if mean/std given."""
    if mean is not None:
        x = normalizer_transform(x, mean, std, dtype)
    x = np.asarray(x, dtype=dtype)
    recon, _ = cae_forward(sd, x, dtype)
    return ((recon - x) ** 2).reshape(x.shape[0], -1).mean(axis=1)


# --------------------------------------------------------------------------------------
# DeepfakeDetector / StatsPool  (/root/reference/src/dlqueen_model.py:115-173)
# --------------------------------------------------------------------------------------
def _conv1d_same(x, w, b):
    """nn.Conv1d(kernel_size=K, padding=K//2): x (B,Ci,L), w (Co,Ci,K) -> (B,Co,L)."""
    B, Ci, L = x.shape
    K = w.shape[2]
    xp = np.zeros((B, Ci, L + K - 1), dtype=x.dtype)
    xp[:, :, K // 2:K // 2 + L] = x
    out = np.zeros((B, w.shape[0], L), dtype=x.dtype)
    for k in range(K):
        out += np.einsum("bcl,oc->bol", xp[:, :, k:k + L], w[:, :, k], optimize=True)
    return out + b.reshape(1, -1, 1)


def _gelu(x):
    """nn.GELU() (approximate='none'): 0.5 x (1 + erf(x / sqrt 2))."""
    from math import erf
    return 0.5 * x * (1.0 + np.vectorize(erf, otypes=[x.dtype])(x / np.sqrt(x.dtype.type(2.0))))


def dlq_forward(sd, x, lengths=None, dtype=np.float64):
    """x (B,321,180) as every scorer here takes it (the reference holds (B,180,T), dlqueen_model.py:102); lengths (B,)
    valid frame counts (None = all T; frames beyond a length must be zero in x, as pad_sequence leaves them) -> logits (B,)."""
    h = np.asarray(x, dtype=dtype).transpose(0, 2, 1)                                # (B, C, T)
    B, _, T = h.shape
    for conv_i, bn_i in ((0, 1), (4, 5), (8, 9)):                                    # ConvEncoder :135-150 (Dropout = identity)
        h = _conv1d_same(h, sd[f"enc.net.{conv_i}.weight"].astype(dtype), sd[f"enc.net.{conv_i}.bias"].astype(dtype))
        h = _gelu(_bn_eval(h, sd, f"enc.net.{bn_i}", np.dtype(dtype).type))
    lengths = np.full(B, T) if lengths is None else np.asarray(lengths)
    mask = (np.arange(T)[None, :] < lengths[:, None]).astype(dtype)[:, None, :]      # StatsPool :121-122
    denom = np.maximum(mask.sum(axis=2), 1.0)                                        # :124
    mean = (h * mask).sum(axis=2) / denom                                            # :125
    var = (mask * (h - mean[:, :, None]) ** 2).sum(axis=2) / denom                   # :127
    z = np.concatenate([mean, np.sqrt(np.maximum(var, 1e-6))], axis=1)               # :128-129
    a = _gelu(z @ sd["head.0.weight"].astype(dtype).T + sd["head.0.bias"].astype(dtype))   # head :161-166
    return (a @ sd["head.3.weight"].astype(dtype).T + sd["head.3.bias"].astype(dtype))[:, 0]


def sigmoid(z):
    z = np.asarray(z)
    return 1.0 / (1.0 + np.exp(-z))
