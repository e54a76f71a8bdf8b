"""ORACLE (test infrastructure, never the product path): the three scorers restated on the
reference's own third-party dependency -- torch CPU fp32 functional ops (oneDNN kernels) --
so that the CPU baseline in bench.py times the same library kernels the reference's
``predict.py --device cpu`` loop reaches (``aten::mkldnn_convolution``).

Each function mirrors one reference ``forward`` (file:line in the comments) but takes a plain
state dict instead of an ``nn.Module``; eval-mode only (BN running stats, Dropout = identity).
``reference_loop_*`` mirror the host loops of predict.py / predict_hybrid.py (bs 32, no_grad).
Pinned by tests/golden (bit-for-bit equal to the unmodified reference classes on CPU).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))


def _bn(x, sd, p):
    return F.batch_norm(x, _t(sd, p + ".running_mean"), _t(sd, p + ".running_var"),
                        _t(sd, p + ".weight"), _t(sd, p + ".bias"), training=False, eps=1e-5)


@torch.no_grad()
def cnn2d_forward(sd, x, return_embedding=False):
    """src/model.py:33-42."""
    h = x.unsqueeze(1)
    for ci, bi, pool in ((0, 1, True), (5, 6, True), (10, 11, False)):
        h = F.relu(_bn(F.conv2d(h, _t(sd, f"conv.{ci}.weight"), _t(sd, f"conv.{ci}.bias"), padding=1), sd, f"conv.{bi}"))
        if pool:
            h = F.avg_pool2d(h, kernel_size=(2, 1))
    h = h.mean(dim=2)
    emb = h.flatten(1)
    logits = F.linear(emb, _t(sd, "classifier.weight"), _t(sd, "classifier.bias"))
    return (logits, emb) if return_embedding else logits


@torch.no_grad()
def cnn1d_forward(sd, x):
    """src/model_cnn1d.py:37-46."""
    h = x.transpose(1, 2)
    for ci, bi in ((0, 1), (4, 5), (8, 9)):
        h = F.relu(_bn(F.conv1d(h, _t(sd, f"conv.{ci}.weight"), _t(sd, f"conv.{ci}.bias"), padding=1), sd, f"conv.{bi}"))
    h = F.adaptive_avg_pool1d(h, 1).flatten(1)
    return F.linear(h, _t(sd, "classifier.weight"), _t(sd, "classifier.bias"))


@torch.no_grad()
def cae_forward(sd, x):
    """src/model_cae.py:83-125."""
    h = x.unsqueeze(1)
    for ci, bi in ((0, 1), (4, 5), (8, 9), (12, 13)):
        h = F.conv2d(h, _t(sd, f"encoder.{ci}.weight"), _t(sd, f"encoder.{ci}.bias"), padding=1)
        h = F.avg_pool2d(F.relu(_bn(h, sd, f"encoder.{bi}")), kernel_size=2)
    latent = h
    for ci, bi, opad in ((0, 1, (0, 0)), (3, 4, (0, 1)), (6, 7, (0, 0))):
        h = F.conv_transpose2d(h, _t(sd, f"decoder.{ci}.weight"), _t(sd, f"decoder.{ci}.bias"), stride=2, output_padding=opad)
        h = F.relu(_bn(h, sd, f"decoder.{bi}"))
    h = F.conv_transpose2d(h, _t(sd, "decoder.9.weight"), _t(sd, "decoder.9.bias"), stride=2)
    T, Tr = x.size(1), h.size(2)
    if Tr < T:
        h = F.pad(h, (0, 0, 0, T - Tr))
    elif Tr > T:
        h = h[:, :, :T, :]
    return h.squeeze(1), latent


@torch.no_grad()
def dlq_forward(sd, x, lengths=None):
    """src/dlqueen_model.py:115-173 (DeepfakeDetector, eval): x (B,321,180) -> logits (B,)."""
    h = x.transpose(1, 2)                                                              # (B, C, T)
    for conv_i, bn_i, pad in ((0, 1, 2), (4, 5, 1), (8, 9, 1)):
        h = F.gelu(_bn(F.conv1d(h, _t(sd, f"enc.net.{conv_i}.weight"), _t(sd, f"enc.net.{conv_i}.bias"), padding=pad), sd, f"enc.net.{bn_i}"))
    B, _, T = h.shape
    lengths = torch.full((B,), T, dtype=torch.long) if lengths is None else torch.as_tensor(lengths)
    mask = (torch.arange(T).unsqueeze(0) < lengths.unsqueeze(1)).unsqueeze(1).float()
    denom = mask.sum(dim=2).clamp(min=1.0)
    mean = (h * mask).sum(dim=2) / denom
    var = (mask * (h - mean.unsqueeze(-1)) ** 2).sum(dim=2) / denom
    z = torch.cat([mean, torch.sqrt(var.clamp(min=1e-6))], dim=1)
    a = F.gelu(F.linear(z, _t(sd, "head.0.weight"), _t(sd, "head.0.bias")))
    return F.linear(a, _t(sd, "head.3.weight"), _t(sd, "head.3.bias")).squeeze(1)


@torch.no_grad()
def reference_loop_supervised(forward, sd, feats, batch_size=32, apply_sigmoid=True):
    """predict.py:100-111 / predict_hybrid.py:52-63 on an in-memory (N,321,180) fp32 tensor."""
    out = []
    for i in range(0, feats.shape[0], batch_size):
        logits = forward(sd, feats[i:i + batch_size]).squeeze(-1)
        s = torch.sigmoid(logits) if apply_sigmoid else logits
        out.extend(s.cpu().tolist())
    return np.array(out)


@torch.no_grad()
def reference_loop_cae(sd, feats, mean=None, std=None, batch_size=32):
    """predict_hybrid.py:66-78 (+ the per-sample normaliser of :45-49 when mean/std given)."""
    out = []
    for i in range(0, feats.shape[0], batch_size):
        x = feats[i:i + batch_size]
        if mean is not None:
            x = (x - mean) / std
        recon, _ = cae_forward(sd, x)
        mse = F.mse_loss(recon, x, reduction="none").view(x.size(0), -1).mean(1)
        out.extend(mse.cpu().tolist())
    return np.array(out)
