"""Device-resident scoring rates of the three scorers + blend + EER (utterances/s, scores/s) on cuda:0."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-fake-audio-classifier_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import dfs_b200 as D  # noqa: E402
from dfs_b200 import synthetic as syn  # noqa: E402


def rate(fn, n, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t0)


n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x = D.fill_features(n)
mean, std = syn.normalizer_stats(1)
c2 = D.Cnn2dScorer(syn.cnn2d_state(0))
c1 = D.Cnn1dScorer(syn.cnn1d_state(0))
ca = D.CaeScorer(syn.cae_state(0), mean, std)
dq = D.DlqScorer(syn.dlq_state(0))
print(f"cnn2d  {rate(lambda: c2.score(x, True), n):12.0f} utt/s")
print(f"cnn1d  {rate(lambda: c1.score(x, True), n):12.0f} utt/s")
print(f"cae    {rate(lambda: ca.score(x), n):12.0f} utt/s")
print(f"dlq    {rate(lambda: dq.score(x, apply_sigmoid=True), n):12.0f} utt/s")
s2, s1, m = c2.score(x, True), c1.score(x, True), ca.score(x)
print(f"hybrid blend+eer on {n}: {rate(lambda: D.calculate_eer(D.hybrid_blend(D.ensemble_mean([s2, s1], as_numpy=False), m, 0.8, as_numpy=False), syn.labels(n)), n):12.0f} scores/s")
for big in (1_000_000, 100_000_000):
    sc, lab = syn.tie_free_scores(big, seed=6)
    sd, ld = torch.from_numpy(sc).cuda(), torch.from_numpy(lab).cuda()
    print(f"eer n={big}: {rate(lambda: D.eer_details(sd, ld), big) / 1e6:10.1f} Mscores/s")
