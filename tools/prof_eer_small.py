"""EER of 20 M tie-free scores by the sort path, twice: target of the ncu --set full capture of radix_downsweep_kernel."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deep-fake-audio-classifier_b200"))
import torch  # noqa: E402

import dfs_b200 as D  # noqa: E402
from dfs_b200 import synthetic as syn  # noqa: E402

n = int(os.environ.get("EER_N", 20_000_000))
if "EER_FORM" in os.environ:   # dfs_set_global_option("eer_sort_onesweep"): 0 = super-tile passes, 1..3 = one-sweep passes
    D._native.set_global_option("eer_sort_onesweep", int(os.environ["EER_FORM"]))
sc, lab = syn.tie_free_scores(n, seed=6)
s, l = torch.from_numpy(sc).cuda(), torch.from_numpy(lab).cuda()
for _ in range(2):
    r = D.eer_details(s, l, method="sort")
print("eer", r["eer"])
