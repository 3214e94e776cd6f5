"""``get_accuracy`` and ``IOU2`` -- the two names of the reference's utils/metrics.py that its train / validate loops
import (main_ucf101.py:27,190,261).  Host-side bookkeeping outside the hot path; the reference's matplotlib plotting
helpers are not provided."""
import numpy as np
import torch


def get_accuracy(predicted_actor, actor):
    """Fraction of rows whose arg-max class equals the label.  predicted_actor (B, C), actor (B, 1) or (B,)."""
    pred = predicted_actor.detach().argmax(dim=1).reshape(-1)
    truth = actor.detach().reshape(-1).to(device=pred.device, dtype=pred.dtype)
    return float((pred == truth).sum().item()) / float(pred.numel())


def IOU2(gt, img):
    """Intersection over union of two {0,1} masks (numpy arrays or tensors); NaN when the ground truth is empty,
    which is what the validation loop tests for (main_ucf101.py:262)."""
    g = np.asarray(gt.detach().cpu() if torch.is_tensor(gt) else gt) > 0
    p = np.asarray(img.detach().cpu() if torch.is_tensor(img) else img) > 0
    if not g.any():
        return float("nan")
    return float(np.logical_and(g, p).sum()) / float(np.logical_or(g, p).sum())
