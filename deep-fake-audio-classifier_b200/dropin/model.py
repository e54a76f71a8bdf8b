"""Drop-in for the reference's ``src/model.py``: ``CNN2D(in_features=180, base_channels=32,
num_classes=1, dropout=0.2)`` with the same state-dict keys (conv.{0,1,5,6,10,11}, classifier) and
``forward(x, return_embedding=False)`` contract (/root/reference/src/model.py:12-42); eval-mode CUDA
forward = conv1 + two tcgen05 implicit-GEMM convs + head in libdfs_b200.so."""
import torch.nn as nn

from _base import NativeBackedModule


def _block(cin, cout, pool, dropout):
    layers = [nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout), nn.ReLU()]
    if pool:
        layers += [nn.AvgPool2d(kernel_size=(2, 1)), nn.Dropout(dropout)]
    return layers


class CNN2D(NativeBackedModule):
    HAS_SPLIT = True             # precision = "split": the accurate tensor-core mode (dfs_b200.Cnn2dScorer)

    def __init__(self, in_features=180, base_channels=32, num_classes=1, dropout=0.2):
        super().__init__()
        c = base_channels
        self.conv = nn.Sequential(*_block(1, c, True, dropout), *_block(c, 2 * c, True, dropout), *_block(2 * c, 4 * c, False, dropout))
        self.classifier = nn.Linear(4 * c * in_features, num_classes)

    def _make_scorer(self, sd, device_index):
        from dfs_b200 import Cnn2dScorer
        return Cnn2dScorer(sd, device=device_index, precision=self._precision())

    def forward(self, x, return_embedding=False):
        if self._use_native(x):
            out = self.native(x.device).score(x, apply_sigmoid=False, return_embedding=return_embedding)
            if return_embedding:
                return out[0].unsqueeze(-1), out[1]
            return out.unsqueeze(-1)                      # (B, 1) logits like the reference
        h = self.conv(x.unsqueeze(1)).mean(dim=2)         # training path: plain PyTorch layers
        emb = h.flatten(1)
        logits = self.classifier(emb)
        return (logits, emb) if return_embedding else logits
