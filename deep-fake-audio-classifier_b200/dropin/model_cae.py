"""Drop-in for the reference's ``src/model_cae.py``: ``ConvAutoencoder(base_channels=32)`` with
state-dict keys encoder.{0,1,4,5,8,9,12,13}, decoder.{0,1,3,4,6,7,9} and ``forward(x) -> (recon,
latent)`` (/root/reference/src/model_cae.py:23-125).

``forward`` must hand back a materialised reconstruction, so it is the slower compat path; the
fused path that never writes the reconstruction is ``score_mse(x)`` (what
``src/predict_hybrid.py::get_cae_scores`` computes, see scoring.py)."""
import torch.nn as nn
import torch.nn.functional as F

from _base import NativeBackedModule


class ConvAutoencoder(NativeBackedModule):
    def __init__(self, base_channels: int = 32):
        super().__init__()
        c = base_channels
        enc = []
        for cin, cout in ((1, c), (c, 2 * c), (2 * c, 4 * c), (4 * c, 8 * c)):
            enc += [nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True), nn.AvgPool2d(kernel_size=2)]
        self.encoder = nn.Sequential(*enc)
        dec = []
        for cin, cout, opad in ((8 * c, 4 * c, 0), (4 * c, 2 * c, (0, 1)), (2 * c, c, 0)):
            dec += [nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2, output_padding=opad), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]
        dec.append(nn.ConvTranspose2d(c, 1, kernel_size=2, stride=2))
        self.decoder = nn.Sequential(*dec)
        self._norm = None

    def set_normalizer(self, mean, std):
        """Optional FeatureNormalizer statistics (src/dataset_cae.py:18-52) applied inside score_mse."""
        self._norm = (mean, std)
        self._native_key = None

    def _make_scorer(self, sd, device_index):
        from dfs_b200 import CaeScorer
        mean, std = self._norm if self._norm is not None else (None, None)
        return CaeScorer(sd, mean, std, device=device_index, precision=self._precision())

    def score_mse(self, x, apply_normalizer=None):
        """Per-utterance reconstruction MSE, reconstruction never materialised (predict_hybrid.py:75-76)."""
        if not self._use_native(x):
            raise RuntimeError("score_mse is an eval-mode scoring call")
        return self.native(x.device).score(x, apply_normalizer)

    def forward(self, x):
        if self._use_native(x):
            return self.native(x.device).forward(x)
        latent = self.encoder(x.unsqueeze(1))
        recon = self.decoder(latent)
        t, tr = x.size(1), recon.size(2)
        if tr < t:
            recon = F.pad(recon, (0, 0, 0, t - tr))
        elif tr > t:
            recon = recon[:, :, :t, :]
        return recon.squeeze(1), latent
