"""Thin launch wrappers over the C ABI.  No math in Python: every function forwards device
pointers / sizes to libb200caps.so on torch's current stream.  CPU tensors are rejected."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _abi
from .plans import ConvPlan, View, fill_conv_desc, fill_wgrad_desc


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("b200caps: CPU tensor passed to a device kernel (no CPU fallback exists)")
    return t.data_ptr()


# ---- tcgen05 implicit GEMM ---------------------------------------------------------------
def conv_fprop(plan: ConvPlan, which: str, x: View, out: View, bias=None, scale_nc=None, relu=False,
               sigmoid_from=-1, accumulate=False, bn_tile=0):
    d = fill_conv_desc(plan, which, x, out, bias, scale_nc, relu, sigmoid_from, accumulate, bn_tile)
    _abi.call("b2c_conv_fprop", C.byref(d), stream())


def conv_wgrad(plan: ConvPlan, x: View, dy: View, dw: torch.Tensor, atomic=True, nsplit=0, bn_tile=0):
    d = fill_wgrad_desc(plan, x, dy, dw, atomic, nsplit, bn_tile)
    _abi.call("b2c_conv_wgrad", C.byref(d), stream())
