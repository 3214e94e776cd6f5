"""Drop-in for the reference's utils/helpers.py: the consistency attentive masks, computed on the device
(the reference copies every clip to the host and loops in numpy, helpers.py:29,87).
Both functions return CUDA float32 tensors; the callers immediately cast to cuda float (main_ucf101.py:117-118,131)."""
import torch

from b200caps import engine, ops


def measure_pixelwise_var_v2(pred, flip_pred, frames_cnt=5, use_sig_output=False):
    """Cyclic temporal variance mask (reference :8-67).  pred / flip_pred (B,1,8,H,W) -> (B,1,8,H,W)."""
    engine.require_cuda(pred, "pred")
    p = pred.detach().contiguous().float()
    f = flip_pred.detach().contiguous().float()
    B, H, W = p.shape[0], p.shape[-2], p.shape[-1]
    assert p.shape[-3] == 8, "the reference hard-codes 8-frame clips (helpers.py:14)"
    m = torch.empty((B, 1, 8, H, W), dtype=torch.float32, device=p.device)
    mm = torch.empty((B, 2), dtype=torch.float32, device=p.device)
    ops.bv_mask(p, f, m, mm, B, H, W, int(frames_cnt), bool(use_sig_output))
    return m


def measure_pixelwise_gradient(pred, conf_thresh_lower=None, conf_thresh_upper=None):
    """Second temporal derivative of sigmoid(pred), per-clip min-max (reference :70-95).
    Returns (B,8,H,W) -- no channel dimension, exactly like the reference (:76)."""
    engine.require_cuda(pred, "pred")
    p = pred.detach().contiguous().float()
    B, H, W = p.shape[0], p.shape[-2], p.shape[-1]
    m = torch.empty((B, 8, H, W), dtype=torch.float32, device=p.device)
    mm = torch.empty((B, 2), dtype=torch.float32, device=p.device)
    ops.gv_mask(p, m, mm, B, H, W, conf_thresh_lower, conf_thresh_upper)
    return m
