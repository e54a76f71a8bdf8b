"""Time dfs_eer (sort path) at 100 M scores under every form of the radix passes (dfs_set_global_option "eer_sort_onesweep"):
0 = count / scan / scatter over super-tiles, 1..4 = the one-sweep forms listed in include/dfs_b200.h.  Prints ms per call (CUDA events, after warm-up) per input distribution and checks that every form
returns the same permutation.  Usage: python tools/eer_forms.py [n] [forms...]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "deep-fake-audio-classifier_b200"))
import torch  # noqa: E402

import dfs_b200 as D  # noqa: E402
from dfs_b200 import synthetic as syn  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    forms = sys.argv[2:] or ["0", "1", "2", "3", "4", "5", "1n"]      # "1n": form 1 without the speculative first pass
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    sc, lab = syn.tie_free_scores(n, seed=6)
    inputs = {"affine": (torch.from_numpy(sc).to(dev), torch.from_numpy(lab).to(dev))}
    logit = 4.0 * torch.randn(n, device=dev, generator=g)
    inputs["sigmoid"] = (torch.sigmoid(logit).contiguous(), (torch.rand(n, device=dev, generator=g) < torch.sigmoid(0.5 * logit)).to(torch.uint8))
    del logit
    for name, (s, l) in inputs.items():
        ref = None
        for form in forms:
            D._native.set_global_option("eer_sort_onesweep", int(form.rstrip("n")))
            D._native.set_global_option("eer_sort_overlap", 0 if form.endswith("n") else 1)
            d = D.eer_details(s, l, want_perm=True)
            if ref is None:
                ref = d
            same = bool(torch.equal(d["perm"], ref["perm"])) and (d["eer"], d["eer_idx"]) == (ref["eer"], ref["eer_idx"])
            del d
            for _ in range(3):
                D.eer_details(s, l, method="sort")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps):
                D.eer_details(s, l, method="sort")
            e1.record()
            torch.cuda.synchronize()
            print(f"{name:8s} n={n} form={form} sort_ms={e0.elapsed_time(e1) / reps:.3f} same_perm_as_form_{forms[0]}={same}", flush=True)
        ref = None
    D._native.set_global_option("eer_sort_onesweep", 1)
    D._native.set_global_option("eer_sort_overlap", 1)


if __name__ == "__main__":
    main()
