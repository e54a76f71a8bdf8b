"""Drop-in for the reference's ``src/model_cnn1d.py``: ``CNN1D(in_features=180, base_channels=32,
num_classes=1, dropout=0.2)``; state-dict keys conv.{0,1,4,5,8,9}, classifier
(/root/reference/src/model_cnn1d.py:12-46)."""
import torch.nn as nn

from _base import NativeBackedModule


class CNN1D(NativeBackedModule):
    def __init__(self, in_features=180, base_channels=32, num_classes=1, dropout=0.2):
        super().__init__()
        c = base_channels
        layers = []
        for i, (cin, cout) in enumerate(((in_features, c), (c, 2 * c), (2 * c, 4 * c))):
            layers += [nn.Conv1d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm1d(cout), nn.ReLU()]
            if i < 2:
                layers.append(nn.Dropout(dropout))
        self.conv = nn.Sequential(*layers)
        self.pool = nn.AdaptiveAvgPool1d(1)
        self.classifier = nn.Linear(4 * c, num_classes)

    def _make_scorer(self, sd, device_index):
        from dfs_b200 import Cnn1dScorer
        return Cnn1dScorer(sd, device=device_index, precision=self._precision())

    def forward(self, x):
        if self._use_native(x):
            return self.native(x.device).score(x, apply_sigmoid=False).unsqueeze(-1)
        h = self.pool(self.conv(x.transpose(1, 2))).flatten(1)
        return self.classifier(h)
