"""Drop-in for the model part of ``src/dlqueen_model.py``: ``StatsPool``, ``ConvEncoder`` and
``DeepfakeDetector(in_ch, hidden=256, dropout=0.3)`` with the reference's state-dict keys (enc.net.{0,1,4,5,8,9},
head.{0,3}) and ``forward(x (B,C,T), lengths (B,)) -> logits (B,)`` (/root/reference/src/dlqueen_model.py:115-173).
Eval-mode CUDA forward runs in libdfs_b200.so (conv1d tensor-core template + masked stats pooling + head); train mode
uses the PyTorch layers below."""
import torch
import torch.nn as nn

from _base import NativeBackedModule


class StatsPool(nn.Module):
    """Mean+Std pooling over time (masked)."""
    def forward(self, x, lengths):
        B, C, T = x.shape
        mask = (torch.arange(T, device=x.device).unsqueeze(0) < lengths.unsqueeze(1)).unsqueeze(1).float()
        denom = mask.sum(dim=2).clamp(min=1.0)
        mean = (x * mask).sum(dim=2) / denom
        var = (mask * (x - mean.unsqueeze(-1)) ** 2).sum(dim=2) / denom
        return torch.cat([mean, torch.sqrt(var.clamp(min=1e-6))], dim=1)


class ConvEncoder(nn.Module):
    def __init__(self, in_ch, hidden=256, dropout=0.2):
        super().__init__()
        self.net = nn.Sequential(
            nn.Conv1d(in_ch, hidden, kernel_size=5, padding=2), nn.BatchNorm1d(hidden), nn.GELU(), nn.Dropout(dropout),
            nn.Conv1d(hidden, hidden, kernel_size=3, padding=1), nn.BatchNorm1d(hidden), nn.GELU(), nn.Dropout(dropout),
            nn.Conv1d(hidden, hidden, kernel_size=3, padding=1), nn.BatchNorm1d(hidden), nn.GELU(), nn.Dropout(dropout))

    def forward(self, x):
        return self.net(x)


class DeepfakeDetector(NativeBackedModule):
    def __init__(self, in_ch, hidden=256, dropout=0.3):
        super().__init__()
        self.enc = ConvEncoder(in_ch=in_ch, hidden=hidden, dropout=dropout)
        self.pool = StatsPool()
        self.head = nn.Sequential(nn.Linear(hidden * 2, hidden), nn.GELU(), nn.Dropout(dropout), nn.Linear(hidden, 1))

    def _make_scorer(self, sd, device_index):
        from dfs_b200 import DlqScorer
        return DlqScorer(sd, device=device_index)

    def forward(self, x, lengths):
        if self._use_native(x):
            if x.shape[2] != 321:
                raise ValueError(f"the native StatsPool detector is built for T = 321 frames (got {x.shape[2]}); pad the batch to 321")
            return self.native(x.device).score(x.transpose(1, 2), lengths)       # (B,C,T) storage read as the (B,T,C) view
        return self.head(self.pool(self.enc(x), lengths)).squeeze(1)
