"""The two functions of the reference's utils/metrics.py that the training / validation loops call
(get_accuracy :7-13, IOU2 :171-193).  The plotting helpers (matplotlib) are outside the hot path."""
import numpy as np
import torch


def get_accuracy(predicted_actor, actor):
    maxm, prediction = torch.max(predicted_actor, 1)
    prediction = prediction.view(-1, 1)
    actor = actor.view(-1, 1).to(prediction.device)
    correct = torch.sum(actor == prediction.float()).item()
    return correct / float(prediction.shape[0])


def IOU2(gt, img):
    """IoU of two binary numpy masks; NaN when the ground truth is empty."""
    intersection = gt + img
    intersection[intersection < 2] = 0
    intersection[intersection > 0] = 1
    intersection_sum = intersection.sum()
    union = gt + img
    union[union > 1] = 1
    union_sum = union.sum()
    if gt.sum() > 0:
        return intersection_sum / union_sum
    return float('NaN')
