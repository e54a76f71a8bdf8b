"""1D-CNN scoring rate, one kernel per layer (option fused = 0) against the whole network in one kernel (fused = 1).
Run on the GPU box:  python tools/c1d_rate.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deep-fake-audio-classifier_b200"))
import torch  # noqa: E402

import dfs_b200 as D  # noqa: E402
from dfs_b200 import synthetic as syn  # noqa: E402

n = int(os.environ.get("N_UTT", 47360))                       # 10 passes of 4,736 = 10.9 GB of fp32 features
x = D.fill_features(n)
c1 = D.Cnn1dScorer(syn.cnn1d_state(0))
base = None
for fused in (0, 1):
    c1.set_option("fused", fused)
    for _ in range(3):
        s = c1.score(x, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        s = c1.score(x, True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rate = n / (ms * 1e-3)
    diff = 0.0 if base is None else float((s - base).abs().max())
    base = s if base is None else base
    print(f"cnn1d fused={fused}: {rate / 1e6:6.2f} M utt/s  {rate * 231120 / 1e12:5.2f} TB/s of fp32 input  ({ms:.3f} ms per {n})  max |score diff| vs fused=0: {diff:.2e}")
