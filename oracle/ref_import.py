"""Import the UNMODIFIED reference from /root/reference on CPU (build container only).

TEST INFRASTRUCTURE ONLY (see oracle/restate.py header).  The reference targets
torch 1.7 + CUDA; to execute its own code on CPU under torch 2.11 we install the
shims listed in SURVEY.md section 8(c):
  * stub modules: torchsummary, matplotlib(.pyplot), imageio, tensorboardX, skvideo.io,
    datasets.ucf_dataloader
  * torch.cuda.FloatTensor -> torch.FloatTensor (or DoubleTensor for the fp64 oracle)
  * Tensor.cuda -> identity ; torch.load -> {} (random init)
No reference source is copied; modules are loaded from where they lie.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("B200CAPS_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "capsules_ucf101.py"))


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_installed = False


def install_shims(double: bool = False):
    global _installed
    torch.cuda.FloatTensor = torch.DoubleTensor if double else torch.FloatTensor
    if _installed:
        return
    _installed = True
    _stub("torchsummary", summary=lambda *a, **k: None)
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    _stub("imageio")

    class _SW:  # tensorboardX.SummaryWriter stand-in
        def __init__(self, *a, **k):
            pass

        def add_scalars(self, *a, **k):
            pass

    _stub("tensorboardX", SummaryWriter=_SW)
    sk = _stub("skvideo")
    sk.io = _stub("skvideo.io", vread=None)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    _orig_load = torch.load
    torch.load = lambda *a, **k: {}
    torch._b200caps_orig_load = _orig_load


def import_reference(double: bool = False):
    """Returns a namespace with the reference modules (models.capsules_ucf101,
    utils.losses, utils.helpers, utils.ramp_ups)."""
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    install_shims(double)
    # the reference's models/ utils/ datasets/ are namespace packages; make sure OUR
    # drop-in packages (same names) are not the ones imported here.
    for k in [k for k in sys.modules if k.split(".")[0] in ("models", "utils", "datasets")]:
        del sys.modules[k]
    saved = list(sys.path)
    sys.path[:] = [REF_ROOT] + [p for p in sys.path if "pi-consistency-activity-detection_b200" not in p]
    try:
        ns = types.SimpleNamespace()
        ns.i3d = importlib.import_module("models.pytorch_i3d")
        ns.caps = importlib.import_module("models.capsules_ucf101")
        ns.losses = importlib.import_module("utils.losses")
        ns.helpers = importlib.import_module("utils.helpers")
        ns.ramp_ups = importlib.import_module("utils.ramp_ups")
    finally:
        sys.path[:] = saved
        for k in [k for k in sys.modules if k.split(".")[0] in ("models", "utils", "datasets")]:
            del sys.modules[k]
    return ns
