"""2D-CNN scoring rate and per-kernel time by precision mode (fp16 | split | fp32) on cuda:0, device-resident pool."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-fake-audio-classifier_b200"))
import torch  # noqa: E402

import dfs_b200 as D  # noqa: E402
from dfs_b200 import synthetic as syn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8320
x = D.fill_features(n)
for prec in ("fp16", "split", "fp32"):
    m = 512 if prec == "fp32" else n
    sc = D.Cnn2dScorer(syn.cnn2d_state(0), precision=prec)
    sc.score(x[:m], True)
    torch.cuda.synchronize()
    reps = 1 if prec == "fp32" else 6
    t0 = time.perf_counter()
    for _ in range(reps):
        sc.score(x[:m], True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    line = f"cnn2d precision={prec:5s} {m * reps / dt:10.0f} utt/s"
    if prec != "fp32":
        sc.set_option("profile", 1)
        sc.score(x[:m], True)
        ms, cnt = sc.profile(4)
        line += "   per pass (us): " + "  ".join(f"{k} {1e3 * a / max(c, 1):.0f}" for k, a, c in zip(("conv1", "conv2", "conv3", "head"), ms, cnt))
    print(line, flush=True)
    del sc
