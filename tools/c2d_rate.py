import os, sys, time
sys.path.insert(0, os.path.join(os.getcwd(), "deep-fake-audio-classifier_b200"))
import torch
import dfs_b200 as D
from dfs_b200 import synthetic as syn
n = 16640
x = D.fill_features(n)
c2 = D.Cnn2dScorer(syn.cnn2d_state(0))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for fused in (0, 1, 0, 1):
    c2.set_option("conv12_fused", fused)
    for _ in range(2):
        c2.score(x, True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        s = c2.score(x, True)
    torch.cuda.synchronize()
    print("conv12_fused", fused, "cnn2d utt/s", n * reps / (time.perf_counter() - t0), float(s.sum()))
