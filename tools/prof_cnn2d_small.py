"""Smallest program that runs the 2D-CNN's kernels at their benchmark shape (one pass of 416 utterances, three times): the target
of the ncu --set full captures (tools/gpu_r02b.sh)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deep-fake-audio-classifier_b200"))
import torch  # noqa: E402

import dfs_b200 as D  # noqa: E402
from dfs_b200 import synthetic as syn  # noqa: E402

x = D.fill_features(416)
c2 = D.Cnn2dScorer(syn.cnn2d_state(0))
for _ in range(3):
    s = c2.score(x, True)
torch.cuda.synchronize()
print("ok", float(s.sum()))
