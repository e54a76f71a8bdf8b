"""One pass of the one-kernel 1D-CNN over 4,736 utterances, three times: target of the ncu --set full capture."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deep-fake-audio-classifier_b200"))
import torch  # noqa: E402

import dfs_b200 as D  # noqa: E402
from dfs_b200 import synthetic as syn  # noqa: E402

x = D.fill_features(4736)
c1 = D.Cnn1dScorer(syn.cnn1d_state(0))
for _ in range(3):
    s = c1.score(x, True)
torch.cuda.synchronize()
print("ok", float(s.sum()))
