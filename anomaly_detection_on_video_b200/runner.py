"""The reference's Lightning runner (src/runner.py) without Lightning: training and scoring with the native MGFN head.

    training_step              src/runner.py:29-39   cat(normal, abnormal) bags -> model(video, labels) -> loss
    configure_optimizers       src/runner.py:53-59   Adam(lr 1e-3, weight_decay 5e-4) -> the fused native Adam
    validation_step            src/runner.py:42-50   features (1, T, crops, C+1) -> permute -> model(video=...) -> scores
    on_validation_epoch_end    src/runner.py:62-79   np.repeat(preds, frames_per_clip); ROC-AUC and PR-AUC (sklearn)

Lightning, hydra and wandb are not dependencies of this module; ``fit`` is the plain loop Lightning would run.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Mapping, Optional, Sequence

import numpy as np
import torch


def training_step(model, batch, device: Optional[torch.device] = None) -> torch.Tensor:
    """src/runner.py:29-39: ``batch = (normal, abnormal)``, each a mapping with ``feature`` [b, crops, T, C + 1] and
    ``anomaly`` [b] labels; the bags are concatenated normal first.  Returns the loss (call ``.backward()`` on it, or let
    ``NativeAdam.step()`` consume the gradients the native step already produced)."""
    ninputs, ainputs = batch
    dev = device if device is not None else next(model.parameters()).device
    inputs = torch.cat((torch.as_tensor(ninputs["feature"]), torch.as_tensor(ainputs["feature"])), dim=0).float().to(dev)
    nlabels = torch.as_tensor(ninputs["anomaly"]).float().to(dev)
    alabels = torch.as_tensor(ainputs["anomaly"]).float().to(dev)
    outputs = model(video=inputs, abnormal_labels=alabels, normal_labels=nlabels)
    return outputs.loss


def configure_optimizers(model, learning_rate: float = 1e-3, weight_decay: float = 5e-4):
    """src/runner.py:53-59 with configs/runner/default.yaml:5-7: Adam over every parameter of the head, as the fused native
    optimizer (``mgfn.NativeAdam``: one kernel over the flat blobs, gradient all-reduce across ranks when distributed)."""
    from .mgfn import NativeAdam

    return [NativeAdam(model, lr=learning_rate, weight_decay=weight_decay)]


def fit(model, batches: Iterable, max_steps: Optional[int] = None, learning_rate: float = 1e-3, weight_decay: float = 5e-4,
        device: Optional[torch.device] = None) -> List[float]:
    """The loop Lightning runs around ``training_step``: zero_grad, forward + backward (one native call), optimizer step."""
    model.train()
    (opt,) = configure_optimizers(model, learning_rate, weight_decay)
    losses: List[float] = []
    for i, batch in enumerate(batches):
        if max_steps is not None and i >= max_steps:
            break
        opt.zero_grad()
        loss = training_step(model, batch, device)
        opt.step()
        losses.append(float(loss.detach()))
    return losses


@torch.no_grad()
def validation_step(model, batch: Mapping[str, np.ndarray], device: Optional[torch.device] = None) -> Dict[str, np.ndarray]:
    """One test video (src/runner.py:42-50).  ``batch["feature"]``: (T, crops, C + 1) or (1, T, crops, C + 1) as the
    reference's batch-size-1 loader delivers it; returns per-snippet ``preds`` (T,) and the per-frame ``labels``."""
    feat = torch.as_tensor(np.asarray(batch["feature"]), dtype=torch.float32)
    if feat.dim() == 3:
        feat = feat.unsqueeze(0)
    dev = device if device is not None else next(model.parameters()).device
    features = feat.permute(0, 2, 1, 3).contiguous().to(dev)          # (1, crops, T, C + 1)
    outputs = model(video=features)
    out = {"preds": outputs.scores.squeeze(0).squeeze(-1).cpu().numpy()}
    if "label" in batch:
        out["labels"] = np.asarray(batch["label"]).reshape(-1)
    return out


def frame_level_metrics(outputs: Sequence[Mapping[str, np.ndarray]], frames_per_clip: int = 16) -> Dict[str, float]:
    """src/runner.py:62-73: snippet scores repeated ``frames_per_clip`` times against the per-frame ground truth."""
    from sklearn.metrics import auc, precision_recall_curve, roc_curve

    preds = np.repeat(np.concatenate([o["preds"] for o in outputs]), frames_per_clip)
    labels = np.concatenate([np.asarray(o["labels"]) for o in outputs])
    if len(preds) != len(labels):
        raise ValueError(f"{len(preds)} frame predictions for {len(labels)} frame labels (frames_per_clip = {frames_per_clip})")
    fpr, tpr, _ = roc_curve(labels.tolist(), preds)
    precision, recall, _ = precision_recall_curve(labels.tolist(), preds)
    return {"valid/rec_auc": float(auc(fpr, tpr)), "valid/pr_auc": float(auc(recall, precision))}


def validate(model, dataset: Iterable[Mapping[str, np.ndarray]], frames_per_clip: int = 16,
             device: Optional[torch.device] = None) -> Dict[str, float]:
    """The validation loop: every item of a test ``FeatureDataset`` -> ``validation_step`` -> ``frame_level_metrics``."""
    model.eval()
    if getattr(model, "force_split", False):
        model.force_split = False
    outs: List[Dict[str, np.ndarray]] = [validation_step(model, dataset[i], device) for i in range(len(dataset))]
    return frame_level_metrics(outs, frames_per_clip)


__all__ = ["training_step", "configure_optimizers", "fit", "validation_step", "frame_level_metrics", "validate"]
