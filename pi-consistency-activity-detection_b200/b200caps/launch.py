"""Compat launcher: run the reference's OWN scripts (main_ucf101.py, main_jhmdb.py, evaluate_*.py), unmodified, on
the b200caps kernels under torch >= 2 / numpy >= 2 (hazards 1-6).

    cd /path/to/pi-consistency-activity-detection
    PYTHONPATH=/root/repo/pi-consistency-activity-detection_b200 python -m b200caps.launch main_ucf101.py --bv --n_frames 5 ...
    ... python -m b200caps.launch --check main_ucf101.py      # import only: print where models/utils/datasets resolve

What it does before handing control to the script with ``runpy`` (``__name__ == '__main__'``, ``sys.argv`` = the
script's own arguments):

  * module resolution: this package's parent directory is put on ``sys.path`` AHEAD of site-packages.  ``models``,
    ``utils`` and ``datasets`` here are regular packages, so they win over the script directory's namespace
    directories and over HuggingFace ``datasets`` -- verified after the run / by ``--check``, which fails loudly if any
    of the hot-path modules resolved elsewhere;
  * stub modules for imports the scripts make but never need on the hot path and that are absent from the image
    (torchsummary, imageio, tensorboardX, skvideo, matplotlib, wandb) -- only when the real module is missing;
  * ``np.int`` / ``np.float`` / ``np.bool`` aliases (removed in numpy >= 1.24; evaluate_ucf101.py:123);
  * ``ReduceLROnPlateau(verbose=...)`` accepted and dropped (removed kwarg; main_ucf101.py:417);
  * a CPU tensor indexed by a CUDA index tensor (main_ucf101.py:90 ``concat_seg[labeled_vid_index]``) moves the index
    to the host first, as torch 1.7 did implicitly;
  * ``torch.nn.BCEWithLogitsLoss`` -> the kernel-backed ``utils.losses.BCEWithLogitsLoss`` for CUDA inputs, so the
    supervised localisation loss of main_ucf101.py:390 runs in libb200caps.so as well (``B200CAPS_KEEP_TORCH_BCE=1``
    keeps torch's).

Host glue only; nothing here computes."""
from __future__ import annotations

import importlib
import importlib.util
import json
import os
import runpy
import sys
import types

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOT_MODULES = ("models.pytorch_i3d", "models.capsules_ucf101", "models.capsules_jhmdb_semi_sup_pa", "utils.losses",
               "utils.helpers", "utils.ramp_ups", "utils.metrics", "datasets.ucf_dataloader",
               "datasets.load_jhmdb_pytorch_multi")


def _missing(name: str) -> bool:
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__b200caps_stub__ = True
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _NullWriter:
    """tensorboardX.SummaryWriter / wandb stand-in: accepts every call, records nothing."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None


def install_stub_modules():
    installed = []
    if _missing("torchsummary"):
        _stub("torchsummary", summary=lambda *a, **k: None)
        installed.append("torchsummary")
    if _missing("imageio"):
        _stub("imageio")
        installed.append("imageio")
    if _missing("tensorboardX"):
        _stub("tensorboardX", SummaryWriter=_NullWriter)
        installed.append("tensorboardX")
    if _missing("skvideo"):
        sk = _stub("skvideo")
        sk.io = _stub("skvideo.io", vread=None)
        installed.append("skvideo")
    if _missing("matplotlib"):
        mpl = _stub("matplotlib", use=lambda *a, **k: None)
        mpl.pyplot = _stub("matplotlib.pyplot")
        installed.append("matplotlib")
    if _missing("wandb"):
        nw = _NullWriter()
        _stub("wandb", init=nw.init, log=nw.log, watch=nw.watch, finish=nw.finish, config=types.SimpleNamespace())
        installed.append("wandb")
    return installed


def install_numpy_aliases():
    import numpy as np
    for name, typ in (("int", int), ("float", float), ("bool", bool)):
        if name not in np.__dict__:
            setattr(np, name, typ)


def install_torch_shims():
    import torch
    from torch.optim import lr_scheduler

    # (1) ReduceLROnPlateau(verbose=...) -- keyword removed in torch 2.x
    base = lr_scheduler.ReduceLROnPlateau
    if not getattr(base, "__b200caps_shim__", False):
        class ReduceLROnPlateau(base):
            __b200caps_shim__ = True

            def __init__(self, *a, verbose=None, **k):
                super().__init__(*a, **k)

        ReduceLROnPlateau.__name__ = base.__name__
        ReduceLROnPlateau.__qualname__ = base.__qualname__
        lr_scheduler.ReduceLROnPlateau = ReduceLROnPlateau

    # (2) cpu_tensor[cuda_index]: torch 1.7 copied the index to the host; torch 2 raises
    orig_getitem = torch.Tensor.__getitem__
    if not getattr(orig_getitem, "__b200caps_shim__", False):
        def _host_index(self, idx):
            if torch.is_tensor(idx) and idx.is_cuda and not self.is_cuda:
                return idx.cpu()
            if isinstance(idx, tuple) and not self.is_cuda and any(torch.is_tensor(i) and i.is_cuda for i in idx):
                return tuple(i.cpu() if (torch.is_tensor(i) and i.is_cuda) else i for i in idx)
            return idx

        def __getitem__(self, idx):
            return orig_getitem(self, _host_index(self, idx))

        __getitem__.__b200caps_shim__ = True
        torch.Tensor.__getitem__ = __getitem__

    # (3) supervised BCE on the kernels
    if os.environ.get("B200CAPS_KEEP_TORCH_BCE", "0") != "1" and not getattr(torch.nn.BCEWithLogitsLoss, "__b200caps_shim__", False):
        torch_bce = torch.nn.BCEWithLogitsLoss

        class BCEWithLogitsLoss(torch_bce):
            __b200caps_shim__ = True

            def __init__(self, *a, **k):
                self._plain = not a and set(k) <= {"size_average", "reduce", "reduction"} and \
                    k.get("reduction", "mean") == "mean" and k.get("size_average", True) in (True, None) and \
                    k.get("reduce", True) in (True, None)
                k.pop("size_average", None)
                k.pop("reduce", None)
                super().__init__(*a, **k)

            def forward(self, input, target):
                if self._plain and input.is_cuda and input.dim() == 5:
                    from utils.losses import BCEWithLogitsLoss as KernelBCE
                    return KernelBCE()(input, target)
                return super().forward(input, target)

        torch.nn.BCEWithLogitsLoss = BCEWithLogitsLoss
        torch.nn.modules.loss.BCEWithLogitsLoss = BCEWithLogitsLoss


def ensure_path():
    """This package's parent ahead of site-packages (and of any other entry that could offer `datasets`)."""
    while PKG_ROOT in sys.path:
        sys.path.remove(PKG_ROOT)
    sys.path.insert(0, PKG_ROOT)
    # modules imported before the launcher ran (e.g. HuggingFace `datasets` pulled in by another import) must not linger
    for top in ("models", "utils", "datasets"):
        mod = sys.modules.get(top)
        f = getattr(mod, "__file__", None) if mod is not None else None
        if mod is not None and (f is None or not os.path.abspath(f).startswith(PKG_ROOT)):
            for k in [k for k in sys.modules if k == top or k.startswith(top + ".")]:
                del sys.modules[k]


def resolution_report(names=HOT_MODULES):
    rep = {}
    for n in names:
        try:
            spec = importlib.util.find_spec(n)
            rep[n] = getattr(spec, "origin", None) if spec else None
        except Exception as e:   # noqa: BLE001
            rep[n] = f"<{type(e).__name__}: {e}>"
    return rep


def check_resolution(report) -> list:
    return [n for n, origin in report.items() if not (origin and os.path.abspath(origin).startswith(PKG_ROOT))]


def prepare(script: str):
    script = os.path.abspath(script)
    if not os.path.isfile(script):
        raise SystemExit(f"b200caps.launch: script not found: {script}")
    stubs = install_stub_modules()
    install_numpy_aliases()
    install_torch_shims()
    ensure_path()
    # runpy puts the script directory at sys.path[0]; do the same for --check so both see the same search order
    sdir = os.path.dirname(script)
    sys.path.insert(0, sdir)
    return script, stubs


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    check = False
    while argv and argv[0].startswith("--") and not argv[0].endswith(".py"):
        flag = argv.pop(0)
        if flag == "--check":
            check = True
        else:
            raise SystemExit(f"b200caps.launch: unknown launcher flag {flag} (script arguments go AFTER the script path)")
    if not argv:
        raise SystemExit(__doc__)
    script, stubs = prepare(argv[0])
    sys.argv = [script] + argv[1:]
    report = resolution_report()
    bad = check_resolution(report)
    if bad:
        raise SystemExit("b200caps.launch: these hot-path modules would NOT resolve into b200caps: "
                         + json.dumps({k: report[k] for k in bad}, indent=1))
    if check:
        ns = runpy.run_path(script, run_name="__b200caps_check__")
        used = {}
        for name, obj in ns.items():
            mod = getattr(obj, "__module__", None)
            if isinstance(mod, str) and mod.split(".")[0] in ("models", "utils", "datasets"):
                used[name] = getattr(sys.modules.get(mod), "__file__", None)
        print(json.dumps({"script": script, "stubs": stubs, "resolves": report, "names_bound_by_script": used}, indent=1))
        wrong = [k for k, f in used.items() if not (f and os.path.abspath(f).startswith(PKG_ROOT))]
        if wrong:
            raise SystemExit(f"b200caps.launch --check: names bound from outside b200caps: {wrong}")
        return 0
    print(f"[b200caps.launch] running {script} on libb200caps.so; stubbed modules: {stubs or 'none'}", flush=True)
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
