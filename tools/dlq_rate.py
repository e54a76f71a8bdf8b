import sys, time
sys.path.insert(0, "deep-fake-audio-classifier_b200")
import torch
import dfs_b200 as D
from dfs_b200 import synthetic as syn
n = 9472
x = D.fill_features(n)
dq = D.DlqScorer(syn.dlq_state(0))
def rate(reps=10):
    dq.score(x, apply_sigmoid=True); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): dq.score(x, apply_sigmoid=True)
    torch.cuda.synchronize(); return n * reps / (time.perf_counter() - t0)
for v in (0, 1, 0, 1):
    dq.set_option("pair_mma", v); print("pair_mma", v, round(rate()))
