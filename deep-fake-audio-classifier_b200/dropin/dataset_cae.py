"""Drop-in for the scoring-time part of ``src/dataset_cae.py``: ``FeatureNormalizer`` with the reference's
``fit / transform / save / load`` contract and ``.pt`` format ``{"mean": (180,), "std": (180,)}``
(/root/reference/src/dataset_cae.py:18-52).  On the scoring path the statistics are handed to the CAE scorer,
which applies them on load inside the first kernel and again in the fused MSE residual; ``transform`` is kept for
callers that normalise on the host (it is plain tensor arithmetic, no kernel of ours)."""
import torch


class FeatureNormalizer:
    def __init__(self):
        self.mean = None   # (F,)
        self.std = None    # (F,)

    def fit(self, features_list):
        all_feats = torch.cat(features_list, dim=0)           # (sum_T, F)
        self.mean = all_feats.mean(dim=0)
        self.std = all_feats.std(dim=0).clamp(min=1e-8)
        return self

    def transform(self, x):
        if self.mean is None:
            raise RuntimeError("Call .fit() first")
        return (x - self.mean.to(x.device)) / self.std.to(x.device)

    def save(self, path):
        torch.save({"mean": self.mean, "std": self.std}, path)

    @classmethod
    def load(cls, path):
        obj = cls()
        data = torch.load(path, map_location="cpu")
        obj.mean = data["mean"]
        obj.std = data["std"]
        return obj
