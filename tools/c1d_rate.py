import os, sys, time
sys.path.insert(0, os.path.join(os.getcwd(), "deep-fake-audio-classifier_b200"))
import torch
import dfs_b200 as D
from dfs_b200 import synthetic as syn
n = 9472
x = D.fill_features(n)
c1 = D.Cnn1dScorer(syn.cnn1d_state(0))
for _ in range(3):
    c1.score(x, True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    c1.score(x, True)
torch.cuda.synchronize()
print("cnn1d utt/s", n * 5 / (time.perf_counter() - t0))
