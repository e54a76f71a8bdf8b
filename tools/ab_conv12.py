"""A/B under sustained load: 2D-CNN with blocks 1 + 2 fused (default) vs one kernel per block, alternating legs of ~1.5 s each on
the same device-resident pool (the power cap decides the clock, so short runs flatter whatever saves no energy)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-fake-audio-classifier_b200"))
import torch  # noqa: E402

import dfs_b200 as D  # noqa: E402
from dfs_b200 import synthetic as syn  # noqa: E402

n = 16640
x = D.fill_features(n)
sc = D.Cnn2dScorer(syn.cnn2d_state(0))
sc.score(x, True)
for leg in range(6):
    fused = 1 - (leg & 1)
    sc.set_option("conv12_fused", fused)
    sc.score(x, True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < 1.5:
        sc.score(x, True)
        torch.cuda.synchronize()
        reps += 1
    dt = time.perf_counter() - t0
    sc.set_option("profile", 1)
    sc.score(x, True)
    ms, cnt = sc.profile(4)
    sc.set_option("profile", 0)
    print(f"conv12_fused={fused}: {n * reps / dt:9.0f} utt/s   per pass (us): " + "  ".join(f"{k} {1e3 * a / max(c, 1):.0f}" for k, a, c in zip(("k0", "k1", "conv3", "head"), ms, cnt)), flush=True)
