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

from dfs_b200.metrics import bce_with_logits_mean, calculate_eer, confusion_at_threshold, eer_details  # noqa: E402,F401


def _is_plain_bce(criterion):
    """nn.BCEWithLogitsLoss() with default reduction and no weights: the loss the training loop evaluates with
    (src/train.py); it is then computed once over all logits by one fused device reduction."""
    return (isinstance(criterion, torch.nn.BCEWithLogitsLoss) and criterion.reduction == "mean"
            and criterion.weight is None and criterion.pos_weight is None)


def evaluate(model, dataloader, criterion=None, device="cpu", apply_sigmoid=False, swap_tf: bool = False):
    """metrics dict (avg_loss, eer, threshold), scores, labels -- logits unless apply_sigmoid
    (/root/reference/src/evaluation.py:51-104).  The model calls stay per batch (a DataLoader is the interface), but
    nothing is synchronised per batch: logits stay on the device, a plain nn.BCEWithLogitsLoss() is evaluated once over
    all of them (dfs_bce_with_logits) and the EER is the device radix select."""
    model.eval()
    logit_chunks, label_chunks = [], []
    total_loss, total_count = 0.0, 0
    fused_bce = criterion is not None and _is_plain_bce(criterion) and torch.device(device).type == "cuda"
    with torch.no_grad():
        for features, batch_labels in dataloader:
            features = features.to(device)
            batch_labels = batch_labels.to(device)
            if swap_tf:
                features = features.transpose(1, 2)
            logits = model(features).squeeze(-1)
            if criterion is not None and not fused_bce:     # any other criterion: the caller's own torch op, per batch
                total_loss += criterion(logits, batch_labels).item() * batch_labels.size(0)
                total_count += batch_labels.size(0)
            logit_chunks.append(logits)
            label_chunks.append(batch_labels)
    scores, labels = [], []
    if logit_chunks:
        all_logits, all_labels = torch.cat(logit_chunks), torch.cat(label_chunks)
        if fused_bce:
            total_count = all_logits.numel()
            total_loss = bce_with_logits_mean(all_logits, all_labels.float()) * total_count
        scores = (torch.sigmoid(all_logits) if apply_sigmoid else all_logits).tolist()
        labels = all_labels.tolist()
    eer, threshold = (None, None)
    if scores and labels:
        eer, threshold = calculate_eer(scores, labels)
    return {"avg_loss": (total_loss / total_count) if total_count > 0 else None, "eer": eer, "threshold": threshold}, scores, labels


def main(argv=None):
    """``python evaluation.py <prediction.pkl> <labels.pkl>`` -- the reference's scoring CLI (scripts/evaluation.py:59-95):
    same validation errors, same printout, EER and confusion counts on the device."""
    import pandas as pd
    argv = sys.argv[1:] if argv is None else list(argv)
    if len(argv) != 2:
        raise ValueError("Usage: python evaluation.py <prediction.pkl> <labels.pkl>")
    prediction_df = pd.read_pickle(argv[0])
    labels_df = pd.read_pickle(argv[1])
    if "uttid" not in prediction_df.columns or "predictions" not in prediction_df.columns:
        raise ValueError("prediction.pkl must have 'uttid' and 'predictions' columns")
    if "uttid" not in labels_df.columns or "label" not in labels_df.columns:
        raise ValueError("labels.pkl must have 'uttid' and 'label' columns")
    merged = pd.merge(prediction_df, labels_df, on="uttid", how="inner")
    if len(merged) != len(prediction_df) or len(merged) != len(labels_df):
        raise ValueError("uttid mismatch between prediction and labels")
    scores = merged["predictions"].values
    labels = merged["label"].values
    eer, threshold = calculate_eer(scores, labels)
    tp, fp, tn, fn, far, frr = confusion_at_threshold(scores, labels, threshold)
    print(f"EER: {eer:.6f}")
    print(f"Threshold: {threshold:.6f}")
    print(f"TP: {tp}  FP: {fp}  TN: {tn}  FN: {fn}")
    print(f"FAR: {far:.6f}  FRR: {frr:.6f}")
    return eer, threshold, (tp, fp, tn, fn, far, frr)


if __name__ == "__main__":
    main()
