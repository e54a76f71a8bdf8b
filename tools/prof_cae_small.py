"""Smallest program that runs the CAE's kernels at their benchmark shape (passes of 592 utterances): target of ncu --set full (tools/gpu_r02zi.sh)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deep-fake-audio-classifier_b200"))
import torch  # noqa: E402

import dfs_b200 as D  # noqa: E402
from dfs_b200 import synthetic as syn  # noqa: E402

x = D.fill_features(592)
mean, std = syn.normalizer_stats(1)
ca = D.CaeScorer(syn.cae_state(0), mean, std)
for _ in range(3):
    s = ca.score(x)
torch.cuda.synchronize()
print("ok", float(s.sum()))
