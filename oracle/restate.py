"""CPU restatement of the reference's semi-supervised training step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the shipped
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may use it, and only as the checker
or as the timed CPU baseline.

The reference (AKASH2907/pi-consistency-activity-detection) is pure Python on top
of torch / numpy library calls, so the restatement is a *functional* torch-CPU
program (dtype generic: fp64 is the parity yardstick, fp32 is the timed CPU
baseline) that takes the weights as a ``state_dict`` with the reference's key
names.  Every function cites the reference file:line it restates.

Pinning: the reference ships no tests / golden vectors, so
this file is pinned against the reference's own code executed in the build
container: ``oracle/make_golden.py`` imports ``/root/reference`` with CPU shims,
runs it on identical weights + inputs, asserts agreement with this restatement
(fp64, <=1e-9) and writes the small fixtures under ``tests/golden/`` that travel
to the GPU box (``/root/reference`` does not exist there).
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------
# optional emulation of the CUDA path's bf16 rounding points (operands of every tensor-core
# GEMM, stored activations) on top of the otherwise exact restatement.  The routing of this
# model is chaotically sensitive at random init: bf16 rounding anywhere upstream
# moves the mask logits by O(1).  Comparing the CUDA path with the *same roundings applied to
# the oracle* separates implementation fidelity from that conditioning.  Straight-through in
# autograd (rounding has identity gradient), like the CUDA backward.
# --------------------------------------------------------------------------------------
_EMULATE = None        # None | "bf16" | "tf32"


class emulate_bf16:
    kind = "bf16"

    def __enter__(self):
        global _EMULATE
        self.prev, _EMULATE = _EMULATE, self.kind

    def __exit__(self, *a):
        global _EMULATE
        _EMULATE = self.prev


class emulate_tf32(emulate_bf16):
    """The CUDA path's tf32 precision mode: the same rounding points, 10 mantissa bits (round to nearest, ties away)."""
    kind = "tf32"


def _round_tf32(x: Tensor) -> Tensor:
    i = x.to(torch.float32).contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32).to(x.dtype)


def _q(x: Tensor) -> Tensor:
    if _EMULATE is None:
        return x
    xd = x.detach()
    r = xd.to(torch.bfloat16).to(x.dtype) if _EMULATE == "bf16" else _round_tf32(xd)
    return x + (r - xd)

# --------------------------------------------------------------------------------------
# architecture tables (models/pytorch_i3d.py:221-281, truncated at Mixed_4f by
# models/capsules_ucf101.py:343)
# --------------------------------------------------------------------------------------
INCEPTION_CFG = [
    ("Mixed_3b", 192, [64, 96, 128, 16, 32, 32]),
    ("Mixed_3c", 256, [128, 128, 192, 32, 96, 64]),
    ("MaxPool3d_4a_3x3", None, None),
    ("Mixed_4b", 480, [192, 96, 208, 16, 48, 64]),
    ("Mixed_4c", 512, [160, 112, 224, 24, 64, 64]),
    ("Mixed_4d", 512, [128, 128, 256, 24, 64, 64]),
    ("Mixed_4e", 512, [112, 144, 288, 32, 64, 64]),
    ("Mixed_4f", 528, [256, 160, 320, 32, 128, 128]),
]


def state_dict_spec(num_classes: int = 24) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) for every entry of ``CapsNet().state_dict()`` in reference
    order (capsules_ucf101.py:337-384, pytorch_i3d.py:48-149,221-281)."""
    spec: List[Tuple[str, Tuple[int, ...], str]] = []

    def unit(prefix, cin, cout, k):
        spec.append((prefix + ".conv3d.weight", (cout, cin) + tuple(k), "conv"))
        spec.append((prefix + ".bn.weight", (cout,), "bn_w"))
        spec.append((prefix + ".bn.bias", (cout,), "bn_b"))
        spec.append((prefix + ".bn.running_mean", (cout,), "bn_rm"))
        spec.append((prefix + ".bn.running_var", (cout,), "bn_rv"))
        spec.append((prefix + ".bn.num_batches_tracked", (), "bn_n"))

    unit("conv1.Conv3d_1a_7x7", 3, 64, (7, 7, 7))
    unit("conv1.Conv3d_2b_1x1", 64, 64, (1, 1, 1))
    unit("conv1.Conv3d_2c_3x3", 64, 192, (3, 3, 3))
    for name, cin, oc in INCEPTION_CFG:
        if cin is None:
            continue
        p = "conv1." + name
        unit(p + ".b0", cin, oc[0], (1, 1, 1))
        unit(p + ".b1a", cin, oc[1], (1, 1, 1))
        unit(p + ".b1b", oc[1], oc[2], (3, 3, 3))
        unit(p + ".b2a", cin, oc[3], (1, 1, 1))
        unit(p + ".b2b", oc[3], oc[4], (3, 3, 3))
        unit(p + ".b3b", cin, oc[5], (1, 1, 1))
    C = num_classes
    spec += [
        ("primary_caps.pose.weight", (512, 832, 9, 9), "pc_w"),
        ("primary_caps.pose.bias", (512,), "bias:67392"),
        ("primary_caps.a.weight", (32, 832, 9, 9), "pc_w"),
        ("primary_caps.a.bias", (32,), "bias:67392"),
        ("conv_caps.beta_u", (C, 16), "randn"),
        ("conv_caps.beta_a", (C,), "randn"),
        ("conv_caps.weights", (1, 32, C, 4, 4), "randn"),
        ("upsample1.weight", (C * 16, 64, 9, 9), "up_w"),
        ("upsample1.bias", (64,), "bias:5184"),
        ("upsample2.weight", (128, 64, 3, 3, 3), "up_w"),
        ("upsample2.bias", (64,), "bias:1728"),
        ("upsample3.weight", (128, 64, 3, 3, 3), "up_w"),
        ("upsample3.bias", (64,), "bias:1728"),
        ("upsample4.weight", (128, 128, 3, 3, 3), "up_w"),
        ("upsample4.bias", (128,), "bias:3456"),
        ("smooth.weight", (128, 1, 3, 3, 3), "up_w"),
        ("smooth.bias", (1,), "bias:27"),
        ("conv28.weight", (64, 832, 3, 3), "conv"),
        ("conv28.bias", (64,), "bias:7488"),
        ("conv56.weight", (64, 192, 3, 3, 3), "conv"),
        ("conv56.bias", (64,), "bias:5184"),
        ("conv112.weight", (64, 64, 3, 3, 3), "conv"),
        ("conv112.bias", (64,), "bias:1728"),
    ]
    return spec


def _key_seed(key: str, seed: int) -> int:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF


def make_state_dict(num_classes: int = 24, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Name-keyed deterministic weights (so ~192 MB of weights never enter git).

    Distributions follow the reference's random init (torch default conv init;
    ``normal_(0, 0.1)`` capsules_ucf101.py:36,39; ``normal_(0, 0.02)`` :359-374;
    ``randn`` :95-101) except BatchNorm gamma/beta, which are perturbed away from
    (1, 0) so parity tests exercise them.  Generated in fp32 then cast.
    """
    sd: Dict[str, Tensor] = {}
    for key, shape, kind in state_dict_spec(num_classes):
        g = torch.Generator().manual_seed(_key_seed(key, seed))
        if kind == "conv":
            fan_in = int(np.prod(shape[1:]))
            b = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif kind == "bn_w":
            t = 0.8 + 0.4 * torch.rand(shape, generator=g)
        elif kind == "bn_b":
            t = 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
        elif kind == "bn_rm":
            t = torch.zeros(shape)
        elif kind == "bn_rv":
            t = torch.ones(shape)
        elif kind == "bn_n":
            sd[key] = torch.zeros((), dtype=torch.long)
            continue
        elif kind == "pc_w":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "up_w":
            t = torch.randn(shape, generator=g) * 0.02
        elif kind == "randn":
            t = torch.randn(shape, generator=g)
        elif kind.startswith("bias:"):
            b = 1.0 / math.sqrt(int(kind.split(":")[1]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        else:  # pragma: no cover
            raise ValueError(kind)
        sd[key] = t.to(dtype)
    return sd


# --------------------------------------------------------------------------------------
# I3D encoder
# --------------------------------------------------------------------------------------
def _same_pad(size: int, k: int, s: int) -> Tuple[int, int]:
    """TF-style 'same' padding; front = pad//2 (pytorch_i3d.py:15-19,35-42 / :82-109)."""
    pad = max(k - s, 0) if size % s == 0 else max(k - (size % s), 0)
    return pad // 2, pad - pad // 2


def _pad_same(x: Tensor, k: Sequence[int], s: Sequence[int]) -> Tensor:
    t, h, w = x.shape[2:]
    pt, ph, pw = _same_pad(t, k[0], s[0]), _same_pad(h, k[1], s[1]), _same_pad(w, k[2], s[2])
    return F.pad(x, (pw[0], pw[1], ph[0], ph[1], pt[0], pt[1]))


def maxpool_same(x: Tensor, k, s) -> Tensor:
    """MaxPool3dSamePadding.forward (pytorch_i3d.py:21-45): zero pad, then max_pool3d."""
    return F.max_pool3d(_pad_same(x, k, s), kernel_size=tuple(k), stride=tuple(s))


class BNState:
    """Collects the train-mode running-stat updates (pytorch_i3d.py:80: eps=1e-3,
    momentum=0.01; torch semantics: normalise with biased var, update with unbiased)."""

    def __init__(self, train: bool):
        self.train = train
        self.updates: Dict[str, Tuple[Tensor, Tensor]] = {}


def unit3d(x: Tensor, sd, prefix: str, k, s, bn: BNState) -> Tensor:
    """Unit3D.forward (pytorch_i3d.py:89-120): same-pad -> conv3d(no bias) -> BN -> ReLU."""
    x = _q(F.conv3d(_pad_same(_q(x), k, s), _q(sd[prefix + ".conv3d.weight"]), None, stride=tuple(s)))
    w, b = sd[prefix + ".bn.weight"], sd[prefix + ".bn.bias"]
    rm, rv = sd[prefix + ".bn.running_mean"].to(x.dtype), sd[prefix + ".bn.running_var"].to(x.dtype)
    if bn.train:
        mean = x.mean(dim=(0, 2, 3, 4))
        var = x.var(dim=(0, 2, 3, 4), unbiased=False)
        n = x.numel() // x.shape[1]
        with torch.no_grad():
            bn.updates[prefix] = (0.99 * rm + 0.01 * mean.detach(),
                                  0.99 * rv + 0.01 * var.detach() * n / max(n - 1, 1))
    else:
        mean, var = rm, rv
    sh = (1, -1, 1, 1, 1)
    x = (x - mean.view(sh)) / torch.sqrt(var.view(sh) + 1e-3) * w.view(sh) + b.view(sh)
    return _q(F.relu(x))


def inception(x: Tensor, sd, p: str, bn: BNState) -> Tensor:
    """InceptionModule.forward (pytorch_i3d.py:144-149)."""
    one, three = (1, 1, 1), (3, 3, 3)
    b0 = unit3d(x, sd, p + ".b0", one, one, bn)
    b1 = unit3d(unit3d(x, sd, p + ".b1a", one, one, bn), sd, p + ".b1b", three, one, bn)
    b2 = unit3d(unit3d(x, sd, p + ".b2a", one, one, bn), sd, p + ".b2b", three, one, bn)
    b3 = unit3d(maxpool_same(x, three, one), sd, p + ".b3b", one, one, bn)
    return torch.cat([b0, b1, b2, b3], dim=1)


def i3d_trunk(x: Tensor, sd, bn: BNState, prefix: str = "conv1.") -> Tuple[Tensor, Tensor, Tensor]:
    """InceptionI3d.forward up to Mixed_4f (pytorch_i3d.py:328-346)."""
    out112 = unit3d(x, sd, prefix + "Conv3d_1a_7x7", (7, 7, 7), (2, 2, 2), bn)
    x = maxpool_same(out112, (1, 3, 3), (1, 2, 2))
    x = unit3d(x, sd, prefix + "Conv3d_2b_1x1", (1, 1, 1), (1, 1, 1), bn)
    out56 = unit3d(x, sd, prefix + "Conv3d_2c_3x3", (3, 3, 3), (2, 1, 1), bn)
    x = maxpool_same(out56, (1, 3, 3), (1, 2, 2))
    for name, cin, _ in INCEPTION_CFG:
        if cin is None:
            x = maxpool_same(x, (3, 3, 3), (2, 1, 1))
        else:
            x = inception(x, sd, prefix + name, bn)
    return x, out56, out112


# --------------------------------------------------------------------------------------
# capsule head
# --------------------------------------------------------------------------------------
def primary_caps(x: Tensor, sd) -> Tensor:
    """PrimaryCaps.forward (capsules_ucf101.py:43-49) -> (B, 20, 20, 544)."""
    p = F.conv2d(_q(x), _q(sd["primary_caps.pose.weight"]), sd["primary_caps.pose.bias"])
    a = torch.sigmoid(F.conv2d(_q(x), _q(sd["primary_caps.a.weight"]), sd["primary_caps.a.bias"]))
    return torch.cat([p, a], dim=1).permute(0, 2, 3, 1)


def em_routing(poses: Tensor, a_in: Tensor, W: Tensor, beta_u: Tensor, beta_a: Tensor,
               iters: int = 3, eps: float = 1e-8, lam: float = 1e-6) -> Tuple[Tensor, Tensor]:
    """ConvCaps.forward with K=(1,1) (capsules_ucf101.py:290-309): votes
    (transform_view :247-268), caps_em_routing :184-211, m_step :108-156, e_step :158-182.

    poses (b, B, 16), a_in (b, B), W (B, C, 4, 4) -> mu (b, C, 16), a_out (b, C)."""
    b, B, _ = poses.shape
    C = W.shape[1]
    v = torch.einsum("nirk,ijkc->nijrc", poses.view(b, B, 4, 4), W).reshape(b, B, C, 16)
    a_in = a_in.view(b, B, 1)
    r = torch.full((b, B, C), 1.0 / C, dtype=v.dtype)
    ln_2pi = math.log(2 * math.pi)
    for it in range(iters):
        # m_step
        r = r * a_in
        r = r / (r.sum(dim=2, keepdim=True) + eps)
        r_sum = r.sum(dim=1, keepdim=True)                     # (b,1,C)
        coeff = (r / (r_sum + eps)).unsqueeze(-1)              # (b,B,C,1)
        mu = torch.sum(coeff * v, dim=1, keepdim=True)         # (b,1,C,16)
        sigma_sq = torch.sum(coeff * (v - mu) ** 2, dim=1, keepdim=True) + eps
        cost_h = (beta_u + torch.log(sigma_sq.view(b, C, 16).sqrt())) * r_sum.view(b, C, 1)
        cost_h = cost_h.sum(dim=2)                             # (b,C)
        cost_mean = cost_h.mean(dim=1, keepdim=True)
        # the reference squares the SUM of deviations (capsules_ucf101.py:144) - kept.
        cost_stdv = torch.sqrt(torch.sum(cost_h - cost_mean, dim=1, keepdim=True) ** 2 / C + eps)
        a_out = torch.sigmoid(lam * (beta_a - (cost_mean - cost_h) / (cost_stdv + eps)))
        if it < iters - 1:
            # e_step
            ln_p = -1.0 * (v - mu) ** 2 / (2 * sigma_sq) - torch.log(sigma_sq.sqrt()) - 0.5 * ln_2pi
            ln_ap = ln_p.sum(dim=3) + torch.log(eps + a_out.view(b, 1, C))
            r = torch.softmax(ln_ap, dim=2)
    return mu.view(b, C, 16), a_out


def encode(sd, img: Tensor, drop_mask1: Optional[Tensor], bn: BNState):
    """Segment 1 of CapsNet.forward (capsules_ucf101.py:425-431): trunk + Dropout3d."""
    x, cross56, cross112 = i3d_trunk(img, sd, bn)
    if drop_mask1 is not None:
        x = _q(x * drop_mask1.to(x.dtype))
    return x.view(-1, 832, 28, 28), cross56, cross112


def capsules(sd, x: Tensor):
    """Segment 2 (:432-436): PrimaryCaps + EM routing.  Returns caps (B,20,20,544) and rout (B,20,20,C*17)."""
    caps = primary_caps(x, sd)
    B_ = caps.shape[0]
    mu, a_out = em_routing(caps[..., :512].reshape(B_ * 400, 32, 16), caps[..., 512:].reshape(B_ * 400, 32),
                           sd["conv_caps.weights"][0], sd["conv_caps.beta_u"], sd["conv_caps.beta_a"])
    C = a_out.shape[-1]
    rout = torch.cat([mu.reshape(B_, 20, 20, C * 16), a_out.reshape(B_, 20, 20, C)], dim=-1)
    return caps, rout


def decode(sd, rout: Tensor, cross28: Tensor, cross56: Tensor, cross112: Tensor, classification: Tensor,
           concat_labels: Tensor, epoch: int, thresh_ep: int, train: bool, drop_mask2: Optional[Tensor],
           taps: Optional[dict] = None):
    """Segment 3 (:438-512): class activations, pose masking, localisation decoder.
    taps: optional dict that receives the intermediate activations (diagnostics: per-layer gradient bisection)."""
    B_ = rout.shape[0]
    C = rout.shape[-1] // 17
    poses = rout[..., :C * 16].reshape(B_, 20, 20, C, 16)
    activations = rout[..., C * 16:]
    feat = activations.reshape(B_, 400, C)
    class_act = activations.mean(dim=1).mean(dim=1)
    eye = torch.eye(C, dtype=poses.dtype)
    if train:
        lab = eye[classification.long().view(-1)]
        if epoch < thresh_ep:
            unl = torch.ones_like(lab)
        else:
            unl = eye[torch.argmax(class_act, dim=1)]
        sel = (concat_labels.view(-1, 1) == 0).to(poses.dtype)
        mask = sel * unl + (1 - sel) * lab
    else:
        mask = eye[torch.argmax(class_act, dim=1)]
    poses = poses * mask.view(B_, 1, 1, C, 1)
    x = _q(poses.reshape(B_, 20, 20, C * 16).permute(0, 3, 1, 2))
    W = {k: _q(sd[k + ".weight"]) for k in ("upsample1", "conv28", "upsample2", "conv56", "upsample3", "conv112",
                                             "upsample4", "smooth")}
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    tap("x0", x)
    x = tap("u1", _q(F.relu(F.conv_transpose2d(x, W["upsample1"], sd["upsample1.bias"]))))
    x = x.view(-1, 64, 1, 28, 28)
    c28 = _q(F.relu(F.conv2d(cross28, W["conv28"], sd["conv28.bias"], padding=1))).view(-1, 64, 1, 28, 28)
    x = tap("cat28", torch.cat((x, c28), dim=1))
    x = tap("u2", _q(F.relu(F.conv_transpose3d(x, W["upsample2"], sd["upsample2.bias"], stride=2, padding=1, output_padding=1))))
    c56 = _q(F.relu(F.conv3d(cross56, W["conv56"], sd["conv56.bias"], padding=1)))
    x = tap("cat56", torch.cat((x, c56), dim=1))
    x = tap("u3", _q(F.relu(F.conv_transpose3d(x, W["upsample3"], sd["upsample3.bias"], stride=2, padding=1, output_padding=1))))
    c112 = _q(F.relu(F.conv3d(cross112, W["conv112"], sd["conv112.bias"], padding=1)))
    x = tap("cat112", torch.cat((x, c112), dim=1))
    x = F.conv_transpose3d(x, W["upsample4"], sd["upsample4.bias"], stride=2, padding=1, output_padding=1)
    if drop_mask2 is not None:
        x = x * drop_mask2.to(x.dtype)
    x = tap("u4", _q(x))
    x = F.conv_transpose3d(x, W["smooth"], sd["smooth.bias"], padding=1)
    return x.view(-1, 1, 8, 224, 224), class_act, feat


def capsnet_forward(sd, img: Tensor, classification: Tensor, concat_labels: Tensor, epoch: int,
                    thresh_ep: int, train: bool, drop_masks: Optional[Sequence[Tensor]] = None,
                    bn: Optional[BNState] = None, num_classes: int = 24):
    """CapsNet.forward (capsules_ucf101.py:413-512) = encode -> capsules -> decode.

    drop_masks: the two Dropout3d keep-masks already scaled by 1/(1-p) -- shapes
    (B,832,1,1,1) and (B,128,1,1,1) -- or None for no dropout (eval / p=0).
    Returns (logits (B,1,8,224,224), class_act (B,C), feat (B,400,C))."""
    bn = bn or BNState(train)
    x, cross56, cross112 = encode(sd, img, None if drop_masks is None else drop_masks[0], bn)
    _, rout = capsules(sd, x)
    return decode(sd, rout, x, cross56, cross112, classification, concat_labels, epoch, thresh_ep, train,
                  None if drop_masks is None else drop_masks[1])


# --------------------------------------------------------------------------------------
# losses and consistency masks
# --------------------------------------------------------------------------------------
def spread_loss(x: Tensor, target: Tensor, m_min: float = 0.2, m_max: float = 0.9) -> Tuple[Tensor, Tensor]:
    """SpreadLoss.forward (utils/losses.py:14-37); r=0 so margin = m_min; double /b."""
    b, _ = x.shape
    at = x.gather(1, target.long().view(b, 1))
    loss = torch.clamp(m_min - (at - x), min=0) ** 2
    absloss = torch.clamp(0.9 - (at - x), min=0) ** 2
    absloss = absloss.sum() / b - 0.9 ** 2
    loss = (loss.sum() / b - m_min ** 2) / b
    return loss, absloss


def dice_loss(logits: Tensor, targets: Tensor, smooth: float = 1.0) -> Tensor:
    """DiceLoss.forward (utils/losses.py:44-57): ONE ratio over the whole sub-batch."""
    p = torch.sigmoid(logits).reshape(-1)
    t = targets.reshape(-1)
    inter = (p * t).sum()
    return 1 - (2.0 * inter + smooth) / (p.sum() + t.sum() + smooth)


def weighted_mse_loss(inp: Tensor, target: Tensor, weight: Tensor) -> Tensor:
    """utils/losses.py:74-76 (broadcasting semantics intentionally kept, see gv quirk)."""
    return (weight * (inp - target) ** 2).mean()


def pixelwise_var_mask(pred: Tensor, flip_pred: Tensor, frames_cnt: int = 5, use_sig: bool = False) -> Tensor:
    """measure_pixelwise_var_v2 (utils/helpers.py:8-67) as a closed form: 14-frame
    cycle pred[0..7] ++ flip_pred[1..6]; population variance over a cyclic window of
    ``frames_cnt`` centred at each frame (np.var on float32 data); fold; per-clip min-max.

    The reference computes in numpy float32 and stores into a float64 buffer; we follow
    that by computing in float32 unless the input is float64 (fp64 oracle)."""
    assert frames_cnt in (3, 5)
    pred, flip_pred = pred.detach(), flip_pred.detach()
    if use_sig:
        pred, flip_pred = torch.sigmoid(pred), torch.sigmoid(flip_pred)
    cyc = torch.cat([pred[:, 0], flip_pred[:, 0, 1:7]], dim=1)           # (B,14,H,W)
    h = frames_cnt // 2
    idx = (torch.arange(14).view(14, 1) + torch.arange(-h, h + 1).view(1, -1)) % 14
    win = cyc[:, idx]                                                       # (B,14,n,H,W)
    var = win.var(dim=2, unbiased=False)                                    # (B,14,H,W)
    out = torch.empty_like(var[:, :8])
    out[:, 0] = 2 * var[:, 0]
    out[:, 7] = 2 * var[:, 7]
    for k in range(1, 7):
        out[:, k] = var[:, k] + var[:, 14 - k]
    mn = out.amin(dim=(1, 2, 3), keepdim=True)
    out = out - mn
    mx = out.amax(dim=(1, 2, 3), keepdim=True)
    mn2 = out.amin(dim=(1, 2, 3), keepdim=True)    # == 0; the reference recomputes min after the shift
    out = out / (mx - mn2 + 1e-7)
    return out.unsqueeze(1)                                                 # (B,1,8,H,W)


def pixelwise_grad_mask(pred: Tensor, lower: Optional[float] = None, upper: Optional[float] = None) -> Tensor:
    """measure_pixelwise_gradient (utils/helpers.py:70-95): sigmoid, optional clamps,
    np.gradient twice along time (central inside, one-sided at the ends), per-clip
    min-max.  Returns (B, 8, H, W) -- NO channel dim (:76), which makes the caller's
    weighted MSE broadcast to (B, B, 8, H, W)."""
    p = torch.sigmoid(pred.detach())[:, 0]
    if lower is not None:
        p = torch.where(p < lower, torch.zeros_like(p), p)
    if upper is not None:
        p = torch.where(p > upper, torch.ones_like(p), p)

    def grad_t(a):
        g = torch.empty_like(a)
        g[:, 1:-1] = (a[:, 2:] - a[:, :-2]) / 2.0
        g[:, 0] = a[:, 1] - a[:, 0]
        g[:, -1] = a[:, -1] - a[:, -2]
        return g

    g = grad_t(grad_t(p))
    g = g - g.amin(dim=(1, 2, 3), keepdim=True)
    g = g / (g.amax(dim=(1, 2, 3), keepdim=True) - g.amin(dim=(1, 2, 3), keepdim=True) + 1e-7)
    return g


def exp_rampup(rampup_length: int):
    """utils/ramp_ups.py:15-24."""
    def f(epoch):
        if epoch < rampup_length:
            e = float(np.clip(epoch, 0.0, rampup_length))
            phase = 1.0 - e / rampup_length
            return float(np.exp(-5.0 * phase * phase))
        return 1.0
    return f


# --------------------------------------------------------------------------------------
# the step (main_ucf101.py:50-150) -- without the randperm shuffle, which the caller
# applies to the inputs (the permutation is an input of the step, not part of it).
# --------------------------------------------------------------------------------------
def train_step_losses(sd, data: Tensor, fl_data: Tensor, action: Tensor, seg: Tensor, labels: Tensor,
                      epoch: int = 1, thresh_epoch: int = 11, bv: bool = True, gv: bool = False,
                      n_frames: int = 5, wt_loc: float = 1.0, wt_cls: float = 1.0, wt_cons: float = 0.1,
                      wt_ramp: Optional[float] = None, bv_wt: float = 0.5, gv_wt: float = 0.5,
                      predict_maps: bool = False, lower=None, upper=None,
                      drop_masks: Optional[Sequence[Tensor]] = None, num_classes: int = 24,
                      rampup_epochs: int = 100, bn_states: Optional[list] = None):
    """train_model_interface (main_ucf101.py:50-150).  drop_masks: 4 masks in draw order
    (enc#1, dec#1, enc#2, dec#2) or None.  Returns dict of outputs and scalar losses."""
    dm1 = None if drop_masks is None else drop_masks[0:2]
    dm2 = None if drop_masks is None else drop_masks[2:4]
    bn1, bn2 = BNState(True), BNState(True)
    output, pred_action, feat = capsnet_forward(sd, data, action, labels, epoch, thresh_epoch, True, dm1, bn1,
                                                num_classes)
    flip_op, _, _ = capsnet_forward(sd, fl_data, action, labels, epoch, thresh_epoch, True, dm2, bn2, num_classes)
    if bn_states is not None:
        bn_states += [bn1, bn2]
    res = step_losses(output, flip_op, pred_action, action, seg, labels, epoch=epoch, bv=bv, gv=gv, n_frames=n_frames,
                      wt_loc=wt_loc, wt_cls=wt_cls, wt_cons=wt_cons, wt_ramp=wt_ramp, bv_wt=bv_wt, gv_wt=gv_wt,
                      predict_maps=predict_maps, lower=lower, upper=upper, rampup_epochs=rampup_epochs)
    res.update(output=output, flip_op=flip_op, pred_action=pred_action, feat=feat)
    return res


def step_losses(output: Tensor, flip_op: Tensor, pred_action: Tensor, action: Tensor, seg: Tensor, labels: Tensor,
                epoch: int = 1, bv: bool = True, gv: bool = False, n_frames: int = 5, wt_loc: float = 1.0,
                wt_cls: float = 1.0, wt_cons: float = 0.1, wt_ramp: Optional[float] = None, bv_wt: float = 0.5,
                gv_wt: float = 0.5, predict_maps: bool = False, lower=None, upper=None, rampup_epochs: int = 100):
    """The loss half of train_model_interface (main_ucf101.py:88-148) given the two forward passes' outputs."""
    if wt_ramp is None:
        wt_ramp = exp_rampup(rampup_epochs)(epoch)
    lab_idx = torch.where(labels.view(-1) == 1)[0]
    lab_op = output[lab_idx]
    lab_seg = seg[lab_idx].to(output.dtype)
    loc1 = F.binary_cross_entropy_with_logits(lab_op, lab_seg)
    loc2 = dice_loss(lab_op, lab_seg)
    cls_loss, _ = spread_loss(pred_action[lab_idx], action[lab_idx], 0.2, 0.9)
    flipped = torch.flip(flip_op, [4])
    l2 = weighted_mse_loss(flipped, output, torch.ones_like(output))
    cons1 = cons2 = None
    if bv:
        v_clk = pixelwise_var_mask(output, torch.flip(flipped, [2]), n_frames, predict_maps).to(output.dtype)
        v_anti = pixelwise_var_mask(torch.flip(output, [2]), flipped, n_frames, predict_maps).to(output.dtype)
        lv1 = weighted_mse_loss(flipped, output, v_clk)
        lv2 = weighted_mse_loss(flipped, output, torch.flip(v_anti, [2]))
        cons1 = wt_ramp * (lv1 + lv2) + (1 - wt_ramp) * l2
    if gv:
        g = pixelwise_grad_mask(output, lower, upper).to(output.dtype)
        cons2 = weighted_mse_loss(flipped, output, g)          # (B,B,8,H,W) broadcast, kept
    if bv and gv:
        cons = bv_wt * cons1 + gv_wt * cons2
    elif gv:
        cons = cons2
    elif bv:
        cons = cons1
    else:
        cons = l2
    loc = loc1 + loc2
    total = wt_loc * loc + wt_cls * cls_loss + wt_cons * cons
    return dict(total=total, loc=loc, bce=loc1, dice=loc2, cls=cls_loss, cons=cons, l2=l2)


def synthetic_batch(n_lab: int, n_unl: int, seed: int = 47, num_classes: int = 24, dtype=torch.float32,
                    labels_pattern: str = "ucf"):
    """SURVEY.md section 8(d) config-2 inputs: U[0,1) clips, flipped copies, random class,
    random axis-aligned box mask per clip.  Returned already concatenated (labeled first);
    the randperm shuffle of main_ucf101.py:73-79 is applied by the caller if wanted."""
    g = torch.Generator().manual_seed(seed)
    n = n_lab + n_unl
    data = torch.rand((n, 3, 8, 224, 224), generator=g, dtype=torch.float32)
    action = torch.randint(0, num_classes, (n, 1), generator=g).float()
    seg = torch.zeros((n, 1, 8, 224, 224))
    for i in range(n):
        y0, x0 = [int(v) for v in torch.randint(0, 150, (2,), generator=g)]
        hh, ww = [int(v) for v in torch.randint(30, 74, (2,), generator=g)]
        seg[i, 0, :, y0:y0 + hh, x0:x0 + ww] = 1.0
    labels = torch.cat([torch.ones(n_lab), torch.zeros(n_unl)])
    return dict(data=data.to(dtype), fl_data=torch.flip(data, [4]).to(dtype), action=action, seg=seg, labels=labels)


def make_drop_masks(n: int, seed: int, count: int = 4, dtype=torch.float32):
    """Dropout3d(0.5) keep-masks scaled by 2 (capsules_ucf101.py:371,428,507), shapes
    alternate (n,832,1,1,1), (n,128,1,1,1)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(count):
        c = 832 if i % 2 == 0 else 128
        out.append((torch.rand((n, c, 1, 1, 1), generator=g) < 0.5).to(dtype) * 2.0)
    return out
