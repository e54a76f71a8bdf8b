"""Drop-in for ``scripts/evaluation.py`` / the metric half of ``src/evaluation.py``:
``calculate_eer(scores, labels) -> (eer, threshold)`` and ``confusion_at_threshold`` on the device
(radix sort + FAR/FRR sweep), plus ``evaluate(model, dataloader, ...)`` with the reference's return
contract (/root/reference/src/evaluation.py:51-104)."""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

import torch  # noqa: E402

from dfs_b200.metrics import calculate_eer, confusion_at_threshold, eer_details  # noqa: E402,F401


def evaluate(model, dataloader, criterion=None, device="cpu", apply_sigmoid=False, swap_tf: bool = False):
    """metrics dict (avg_loss, eer, threshold), scores, labels -- logits unless apply_sigmoid."""
    model.eval()
    chunks, label_chunks = [], []
    total_loss, total_count = 0.0, 0
    with torch.no_grad():
        for features, batch_labels in dataloader:
            features = features.to(device)
            batch_labels = batch_labels.to(device)
            if swap_tf:
                features = features.transpose(1, 2)
            logits = model(features).squeeze(-1)
            if criterion is not None:
                total_loss += criterion(logits, batch_labels).item() * batch_labels.size(0)
                total_count += batch_labels.size(0)
            chunks.append(torch.sigmoid(logits) if apply_sigmoid else logits)
            label_chunks.append(batch_labels)
    scores = torch.cat(chunks).tolist() if chunks else []
    labels = torch.cat(label_chunks).tolist() if label_chunks else []
    eer, threshold = (None, None)
    if scores and labels:
        eer, threshold = calculate_eer(scores, labels)
    return {"avg_loss": (total_loss / total_count) if total_count > 0 else None, "eer": eer, "threshold": threshold}, scores, labels
