"""How reproducible is the reference against ITSELF at the spacing of random-init scores?

Runs the oracle port of the predict.py loop (torch CPU fp32, the reference's own dependency) on the same n synthetic
utterances twice -- batch size 32 with all threads, and batch size 7 with one thread (different oneDNN blocking /
summation order) -- and reports the rank swaps and the EER difference between the two runs under the bench's
rank-correlated labels.  This bounds what "EER within 0.01 pp end to end" can mean on random-init weights: any arithmetic
that is not bit-identical to one particular CPU configuration moves the EER by whole quanta of 1/(2 n_class).
CPU only; writes one JSON line.   python tools/experiments/reference_self_consistency.py [n]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "deep-fake-audio-classifier_b200")]
from dfs_b200 import synthetic as syn  # noqa: E402
from oracle import eer as oeer  # noqa: E402
from oracle import models_torch as ot  # noqa: E402


def run(x, sd, bs, threads):
    torch.set_num_threads(threads)
    out = []
    with torch.no_grad():
        for i in range(0, x.shape[0], bs):
            out.append(torch.sigmoid(ot.cnn2d_forward(sd, x[i:i + bs]).squeeze(-1)))
    return torch.cat(out).numpy()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    x = torch.from_numpy(syn.features(n, seed=1234))
    sd = syn.cnn2d_state(0)
    a = run(x, sd, 32, os.cpu_count() or 1)
    b = run(x, sd, 7, 1)
    ranks = np.argsort(np.argsort(a, kind="stable"), kind="stable")
    lab = (np.random.Generator(np.random.PCG64(7)).random(n) < 1 / (1 + np.exp(-6.0 * (ranks / n - 0.5)))).astype(np.uint8)
    ea, eb = oeer.calculate_eer(a, lab)[0], oeer.calculate_eer(b, lab)[0]
    rb = np.argsort(np.argsort(b, kind="stable"), kind="stable")
    srt = np.sort(a.astype(np.float64))
    print(json.dumps({"n": n, "max_rel_diff": float(np.max(np.abs(a - b) / np.abs(a))), "bit_identical": bool(np.array_equal(a, b)),
                      "utterances_with_changed_rank": int(np.sum(ranks != rb)), "max_rank_shift": int(np.max(np.abs(ranks - rb))),
                      "median_score_spacing": float(np.median(np.diff(srt))), "score_range": float(srt[-1] - srt[0]),
                      "fp32_ulp_at_score": float(np.spacing(np.float32(a.mean()))),
                      "distinct_scores": int(len(np.unique(a))),
                      "eer_a": ea, "eer_b": eb, "eer_delta_pp": 100 * abs(ea - eb)}))


if __name__ == "__main__":
    main()
