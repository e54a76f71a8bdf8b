"""Smallest program that runs every kernel of the path at its benchmark shape: 2D-CNN (416 utterances), CAE (592), 1D-CNN in one
kernel (4,736), EER of 100 M scores by sort and by select.  Target of the ncu launch lists (per-kernel time shares)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deep-fake-audio-classifier_b200"))
import torch  # noqa: E402

import dfs_b200 as D  # noqa: E402
from dfs_b200 import synthetic as syn  # noqa: E402

reps = int(os.environ.get("REPS", 3))
x = D.fill_features(4736)
c2 = D.Cnn2dScorer(syn.cnn2d_state(0))
mean, std = syn.normalizer_stats(1)
ca = D.CaeScorer(syn.cae_state(0), mean, std)
c1 = D.Cnn1dScorer(syn.cnn1d_state(0))
c1.set_option("fused", int(os.environ.get("C1D_FUSED", 1)))
for _ in range(reps):
    a = c2.score(x[:416], True)
for _ in range(reps):
    b = ca.score(x[:592])
for _ in range(reps):
    c = c1.score(x, True)
torch.cuda.synchronize()
n = int(os.environ.get("EER_N", 100_000_000))
if n > 0:
    sc, lab = syn.tie_free_scores(n, seed=6)
    s, l = torch.from_numpy(sc).cuda(), torch.from_numpy(lab).cuda()
    for m in ("sort", "select"):
        for _ in range(2):
            r = D.eer_details(s, l, method=m)
    print("eer", r["eer"])
print("ok", float(a.sum()), float(b.sum()), float(c.sum()))
