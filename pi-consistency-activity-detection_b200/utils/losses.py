"""Drop-in for the reference's utils/losses.py (SpreadLoss :6-37, DiceLoss :40-57, weighted_mse_loss :74-76).
Same names and call signatures; forward and backward run in the b200caps loss kernels (csrc/losses.cu).
`CapsuleLoss` (:61-72) is dead code in the reference and is not provided."""
import torch
import torch.nn as nn
from torch.nn.modules.loss import _Loss

from b200caps import engine, ops


class _SpreadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target, m_min):
        x = x.contiguous().float()
        b, E = x.shape
        dev = x.device
        idx = torch.arange(b, dtype=torch.int32, device=dev)
        tgt = target.to(dev).float().reshape(-1).contiguous()
        out = torch.empty(2, dtype=torch.float32, device=dev)
        ops.spread_loss(x, tgt, idx, b, E, m_min, out, 0.0, None)
        ctx.save_for_backward(x, tgt, idx)
        ctx.m_min = m_min
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_loss, g_abs):
        x, tgt, idx = ctx.saved_tensors
        b, E = x.shape
        dx = torch.zeros_like(x)
        ops.spread_loss(x, tgt, idx, b, E, ctx.m_min, None, 1.0, dx)
        return dx * g_loss, None, None


class SpreadLoss(_Loss):
    """utils/losses.py:6-37: fixed margin m_min (r = 0), divides by the batch size twice."""

    def __init__(self, m_min=0.2, m_max=0.9, num_class=24):
        super(SpreadLoss, self).__init__()
        self.m_min = m_min
        self.m_max = m_max
        self.num_class = num_class

    def forward(self, x, target):
        engine.require_cuda(x, "SpreadLoss input")
        b, E = x.shape
        assert E == self.num_class
        return _SpreadFn.apply(x, target, float(self.m_min))


class _SegFn(torch.autograd.Function):
    """BCE-with-logits (mean) and Dice over ALL rows of `logits`; returns both."""

    @staticmethod
    def forward(ctx, logits, targets):
        lg = logits.contiguous().float()
        tg = targets.to(lg.device).contiguous().float()
        P = lg.shape[0]
        V = lg.numel() // P
        idx = torch.arange(P, dtype=torch.int32, device=lg.device)
        sums = torch.empty(4, dtype=torch.float64, device=lg.device)
        out = torch.empty(2, dtype=torch.float32, device=lg.device)
        ops.seg_loss_fwd(lg, tg, idx, P, V, sums, out)
        ctx.save_for_backward(lg, tg, idx, sums)
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_bce, g_dice):
        lg, tg, idx, sums = ctx.saved_tensors
        P = lg.shape[0]
        V = lg.numel() // P
        d = torch.zeros_like(lg)
        # the upstream gradients are scalars (loss weights); read them once on the host side of the launch
        ops.seg_loss_bwd(lg, tg, idx, P, V, sums, float(g_bce), float(g_dice), d)
        return d, None


class DiceLoss(nn.Module):
    """utils/losses.py:40-57: 1 - (2*sum(p*t)+1)/(sum(p)+sum(t)+1) over the whole batch (one ratio)."""

    def __init__(self, weight=None, size_average=True):
        super(DiceLoss, self).__init__()

    def forward(self, inputs, targets, smooth=1):
        assert smooth == 1
        engine.require_cuda(inputs, "DiceLoss input")
        return _SegFn.apply(inputs, targets)[1]


class BCEWithLogitsLoss(nn.Module):
    """Kernel-backed stand-in for nn.BCEWithLogitsLoss(size_average=True) (main_ucf101.py:390)."""

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, inputs, targets):
        return _SegFn.apply(inputs, targets)[0]


class _WMSEFn(torch.autograd.Function):
    """mean(weight * (input - target)^2) for the two weight shapes the reference uses:
    same shape as the (B,1,8,H,W) maps, or (B,8,H,W) -> the (B,B,8,H,W) broadcast of main_ucf101.py:130-132."""

    @staticmethod
    def forward(ctx, inp, target, weight):
        a = inp.contiguous().float()
        t = target.contiguous().float()
        P, H, W = a.shape[0], a.shape[-2], a.shape[-1]
        assert a.shape == t.shape and a.numel() == P * 8 * H * W, "weighted_mse_loss expects (B,1,8,H,W) maps"
        w = weight.to(a.device).contiguous().float()
        ctx.gv = (w.dim() == 4)
        if not ctx.gv and w.numel() == 1:
            w = w.expand_as(a).contiguous()
        assert w.numel() == a.numel()
        acc = torch.empty(4, dtype=torch.float64, device=a.device)
        loss = torch.empty(4, dtype=torch.float32, device=a.device)
        if ctx.gv:
            ops.cons_reduce(t, a, None, None, w, acc, P, H, W, 0, 0)
            ops.cons_finish(acc, loss, P, H, W, 2, 0.0, 0.0, 0.0)
        else:
            ops.cons_reduce(t, a, w, None, None, acc, P, H, W, 0, 0)
            ops.cons_finish(acc, loss, P, H, W, 1, 1.0, 0.0, 0.0)
        ctx.save_for_backward(a, t, w)
        return loss[0].clone()

    @staticmethod
    def backward(ctx, g):
        a, t, w = ctx.saved_tensors
        P, H, W = a.shape[0], a.shape[-2], a.shape[-1]
        da = torch.zeros_like(a) if ctx.needs_input_grad[0] else None
        dt = torch.zeros_like(t) if ctx.needs_input_grad[1] else None
        gs = float(g)
        if ctx.gv:
            ops.cons_grad(t, a, None, None, w, dt, da, P, H, W, 0, 0, 0.0, 0.0, gs)
        else:
            ops.cons_grad(t, a, w, None, None, dt, da, P, H, W, 0, 0, 0.0, gs, 0.0)
        return da, dt, None


def weighted_mse_loss(input, target, weight):
    engine.require_cuda(input, "weighted_mse_loss input")
    return _WMSEFn.apply(input, target, weight)
