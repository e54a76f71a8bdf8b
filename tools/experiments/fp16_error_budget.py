"""Which fp16 rounding dominates the 2D-CNN's logit error in the trained-like regime (tests/golden/trained.npz)?
CPU experiment in float64 with the rounding points of the tensor-core path switched on one at a time:
  x (input image), w1 / w2 / w3 (BN-folded conv weights), a1 / a2 (pooled activations stored between the layers).
Prints the max |logit - exact| over 32 structured utterances for each subset.  Run here (no GPU):
    python tools/experiments/fp16_error_budget.py
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "deep-fake-audio-classifier_b200"))
from dfs_b200 import synthetic as syn  # noqa: E402

T = np.load(os.path.join(ROOT, "tests", "golden", "trained.npz"))
sd = syn.cnn2d_state(int(os.environ.get("SEED", 0)), logit_scale=float(T["cnn2d_scale"]), classifier_bias=float(T["cnn2d_bias"]))
n = int(os.environ.get("N", 32))
x = torch.from_numpy(syn.features_structured(n, seed=int(T["seed"]))).double()
torch.set_num_threads(os.cpu_count() or 8)


def fold(conv, bn):
    w, b = (torch.from_numpy(sd[f"conv.{conv}.{k}"]).double() for k in ("weight", "bias"))
    g, beta, mu, var = (torch.from_numpy(sd[f"conv.{bn}.{k}"]).double() for k in ("weight", "bias", "running_mean", "running_var"))
    s = g / torch.sqrt(var + 1e-5)
    return w * s[:, None, None, None], (b - mu) * s + beta


def h(t, on):
    return t.half().double() if on else t


def forward(rx=False, rw1=False, rw2=False, rw3=False, ra1=False, ra2=False):
    (w1, b1), (w2, b2), (w3, b3) = fold(0, 1), fold(5, 6), fold(10, 11)
    a = F.relu(F.conv2d(h(x, rx).unsqueeze(1), h(w1, rw1), b1, padding=1))
    a = h(F.avg_pool2d(a, (2, 1)), ra1)
    a = F.relu(F.conv2d(a, h(w2, rw2), b2, padding=1))
    a = h(F.avg_pool2d(a, (2, 1)), ra2)
    a = F.relu(F.conv2d(a, h(w3, rw3), b3, padding=1))
    emb = a.mean(dim=2).flatten(1)
    return emb @ torch.from_numpy(sd["classifier.weight"]).double().t() + torch.from_numpy(sd["classifier.bias"]).double()


exact = forward()
print(f"n = {n}; exact logits span {float(exact.min()):.2f} .. {float(exact.max()):.2f}; reference fp32 logits agree to "
      f"{float((exact[:, 0] - torch.from_numpy(T['cnn2d_logits'][:n]).double()).abs().max()):.2e}")
cases = {"x": dict(rx=True), "w1": dict(rw1=True), "w2": dict(rw2=True), "w3": dict(rw3=True), "a1": dict(ra1=True), "a2": dict(ra2=True),
         "all weights": dict(rw1=True, rw2=True, rw3=True), "x + activations": dict(rx=True, ra1=True, ra2=True),
         "everything (the tensor-core path)": dict(rx=True, rw1=True, rw2=True, rw3=True, ra1=True, ra2=True)}
for name, kw in cases.items():
    d = (forward(**kw) - exact).abs()
    print(f"fp16 rounding of {name:36s} max |dlogit| = {float(d.max()):.3e}   mean = {float(d.mean()):.3e}")


# ---- error-diffusion rounding of the folded weights: the residual of each rounding is carried into the next weight of the
# same output channel (order: input channel, then the 3x3 taps), so that the rounded weights of a 3x3 stencil -- which multiply
# neighbouring, strongly correlated activations -- sum to the exact stencil sum up to one ulp
def diffuse(w):
    w = w.clone()
    co = w.shape[0]
    flat = w.reshape(co, -1)
    out = torch.empty_like(flat)
    carry = torch.zeros(co, dtype=torch.float64)
    for i in range(flat.shape[1]):
        tgt = flat[:, i] + carry
        q = tgt.half().double()
        out[:, i] = q
        carry = tgt - q
    return out.reshape(w.shape)


def forward_diffused(which=(1, 2, 3)):
    (w1, b1), (w2, b2), (w3, b3) = fold(0, 1), fold(5, 6), fold(10, 11)
    w1, w2, w3 = (diffuse(w) if i + 1 in which else w for i, w in enumerate((w1, w2, w3)))
    a = F.relu(F.conv2d(x.unsqueeze(1), w1, b1, padding=1))
    a = F.avg_pool2d(a, (2, 1))
    a = F.relu(F.conv2d(a, w2, b2, padding=1))
    a = F.avg_pool2d(a, (2, 1))
    a = F.relu(F.conv2d(a, w3, b3, padding=1))
    emb = a.mean(dim=2).flatten(1)
    return emb @ torch.from_numpy(sd["classifier.weight"]).double().t() + torch.from_numpy(sd["classifier.bias"]).double()


for which in ((1,), (2,), (3,), (1, 2, 3)):
    d = (forward_diffused(which) - exact).abs()
    print(f"error-diffused fp16 weights of layers {which}: max |dlogit| = {float(d.max()):.3e}   mean = {float(d.mean()):.3e}")
