"""dfs_b200 -- B200-native scoring engine for the Deep-Fake-Audio-Classifier hot path.

Host side only (ctypes over libdfs_b200.so); see DESIGN.md for the path and its boundary.
"""
from . import _native  # noqa: F401
from . import hostmem  # noqa: F401
from .engine import CaeScorer, Cnn1dScorer, Cnn2dScorer, DlqScorer, ScorerGroup, fill_features, pinned_empty  # noqa: F401
from .metrics import (alpha_sweep, bce_with_logits_mean, blend, calculate_eer, confusion_at_threshold,  # noqa: F401
                      eer_details, ensemble_mean, hybrid_blend, normalise_01)

__all__ = ["Cnn2dScorer", "Cnn1dScorer", "CaeScorer", "DlqScorer", "ScorerGroup", "fill_features", "pinned_empty", "calculate_eer", "confusion_at_threshold",
           "eer_details", "normalise_01", "hybrid_blend", "ensemble_mean", "blend", "alpha_sweep", "bce_with_logits_mean"]
