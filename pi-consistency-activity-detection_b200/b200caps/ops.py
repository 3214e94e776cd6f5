"""Thin launch wrappers over the C ABI.  No math in Python: every function forwards device
pointers / sizes to libb200caps.so on torch's current stream.  CPU tensors are rejected."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _abi
from .plans import PREC, ConvPlan, View, act_dtype, fill_conv_desc, fill_wgrad_desc


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("b200caps: CPU tensor passed to a device kernel (no CPU fallback exists)")
    return t.data_ptr()


# ---- tcgen05 implicit GEMM ---------------------------------------------------------------
TIMING = None   # bench.py: list collecting (kind, start_event, end_event, executed MACs) of every implicit-GEMM launch
BW_TIMING = None  # bench.py: list collecting (kernel family, start_event, end_event, algorithmic bytes) of the other kernels


def _bw(name: str, nbytes: int, *args):
    """Launch a bandwidth-class kernel; when bench.py instruments the step, bracket it with CUDA events and record the
    ALGORITHMIC bytes it has to move (every operand element once)."""
    if BW_TIMING is None:
        _abi.call(name, *args)
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _abi.call(name, *args)
    b.record()
    BW_TIMING.append((name[4:], a, b, int(nbytes)))


def _vb(v) -> int:
    """bytes of a channel window view"""
    return v.rows * v.C * v.t.element_size()


def _launch(kind: str, name: str, desc, macs: int = 0, nbytes: int = 0):
    if TIMING is None:
        _abi.call(name, C.byref(desc), stream())
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _abi.call(name, C.byref(desc), stream())
    b.record()
    TIMING.append((kind, a, b, macs, nbytes))


def _conv_bytes(plan: ConvPlan, which: str, x: View, out) -> int:
    """Compulsory bytes of a fprop / dgrad launch: the input window, the output, the packed operand -- each once."""
    classes = plan.fprop if which == "fprop" else plan.dgrad
    ob = (out.rows * out.C * out.t.element_size()) if isinstance(out, View) else out.numel() * out.element_size()
    wb = sum(c.packed.numel() * c.packed.element_size() for c in classes if c.packed is not None)
    return _vb(x) + ob + wb


def _macs(plan: ConvPlan, which: str, n: int) -> int:
    """MACs the launch executes (zero-padded channels excluded): positions x taps x Cin x Cout over the classes."""
    classes = plan.fprop if which in ("fprop", "wgrad") else plan.dgrad
    pk = plan.fprop_pack if which in ("fprop", "wgrad") else plan.dgrad_pack
    return n * sum(c.Q[0] * c.Q[1] * c.Q[2] * len(c.taps) for c in classes) * pk["C_real"] * pk["R"]


def conv_fprop(plan: ConvPlan, which: str, x: View, out, bias=None, scale_nc=None, relu=False,
               sigmoid_from=-1, accumulate=False, bn_tile=0, final=False):
    """out: View (bf16 / fp32 rows) or a 2-D fp32 tensor (Cout_pad, rows) for the planar epilogue."""
    d = fill_conv_desc(plan, which, x, out, bias, scale_nc, relu, sigmoid_from, accumulate, bn_tile, final)
    _launch(f"{which} {plan.spec.Cin}->{plan.spec.Cout} k{tuple(plan.spec.k)} s{tuple(plan.spec.stride)} in{plan.in_dims}"
            f"{' T' if plan.spec.transposed else ''}", "b2c_conv_fprop", d, _macs(plan, which, x.N) if TIMING is not None else 0,
            _conv_bytes(plan, which, x, out) if TIMING is not None else 0)


def split_bf16(v: View):
    """tf32 mode: fp32 view -> (hi, lo) compact bf16 Views with hi + lo = x to 2^-17."""
    shape = tuple(v.t.shape[:-1]) + (v.C,)
    hi = torch.empty(shape, dtype=torch.bfloat16, device=v.t.device)
    lo = torch.empty(shape, dtype=torch.bfloat16, device=v.t.device)
    _bw("b2c_split_bf16", _vb(v) * 2, v.ptr, v.row_stride, v.c_off, _p(hi), _p(lo), v.rows, v.C, stream())
    return View(hi), View(lo)


def conv_wgrad(plan: ConvPlan, x: View, dy: View, dw: torch.Tensor, atomic=True, nsplit=0, bn_tile=0, part=None, per_clip=False,
               pp=(0, 0, 0), presplit=None, segs=None):
    name = (f"wgrad {plan.spec.Cin}->{plan.spec.Cout} k{tuple(plan.spec.k)} s{tuple(plan.spec.stride)} in{plan.in_dims}"
            f"{' T' if plan.spec.transposed else ''}")
    wm = wb = 0
    if TIMING is not None:
        wb = _vb(x) + (_vb(dy) if part is None else dy.rows * part[1] * dy.t.element_size()) + \
            (sum(t.numel() for _, t in segs) if segs else dw.numel()) * 4
        cl, geo = plan.wgrad_cls, plan.wgrad_geom
        cp = part[2] if (part is not None and len(part) > 2) else (part[1] if part is not None else geo.get("Cp_real", geo["Cp"]))
        wm = x.N * geo["Q"][0] * geo["Q"][1] * geo["Q"][2] * len(cl.taps) * geo["Cg_real"] * cp
    if PREC.mode:
        # tf32 mode: 3 x bf16 split GEMMs accumulated into dw (see b2c_split_bf16); dw must be zero / an accumulator
        assert atomic, "tf32-mode wgrad accumulates"
        (xh, xl), (dh, dl) = presplit if presplit is not None else (split_bf16(x), split_bf16(dy))
        for a, b in ((xh, dh), (xh, dl), (xl, dh)):
            _launch(name, "b2c_conv_wgrad", fill_wgrad_desc(plan, a, b, dw, True, nsplit, bn_tile, part, per_clip, force_bf16=True, pp=pp,
                                                             segs=segs), wm, wb)
        return
    d = fill_wgrad_desc(plan, x, dy, dw, atomic, nsplit, bn_tile, part, per_clip, pp=pp, segs=segs)
    _launch(name, "b2c_conv_wgrad", d, wm, wb)


class PackRegistry:
    """Every weight-packing job issued so far (pointers are stable: parameters are views of the flat buffer, packed
    operands are persistent).  Once the fused step has seen all layers, `flush()` re-packs everything in one launch."""

    def __init__(self):
        self.jobs = {}           # (w_ptr, packed_ptr, r_off, col_off) -> PackJob fields
        self.table = None        # (jobs_dev, block_start_dev, njobs, nblocks)
        self.flushed_epoch = -1
        self.pre = {}            # key -> callable launched before the batched pack (derived fp32 sources: folded stem weights)

    def add_pre(self, key, fn):
        self.pre.setdefault(key, fn)

    def add(self, key, fields):
        if key not in self.jobs:
            self.jobs[key] = fields
            self.table = None

    def flush(self, epoch: int):
        if not self.jobs:
            return False
        if self.table is None:
            import numpy as np
            arr = (_abi.PackJob * len(self.jobs))()
            starts = [0]
            for i, f in enumerate(self.jobs.values()):
                for k, v in f.items():
                    setattr(arr[i], k, v)
                total = f["R"] * f["ntaps"] * f["C"]
                assert total < 2 ** 31
                starts.append(starts[-1] + max(1, (total + 4095) // 4096))       # 16 elements per thread, every job
            dev = torch.device("cuda", torch.cuda.current_device())
            raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
            bs = torch.tensor(starts, dtype=torch.int32, device=dev)
            self.table = (raw, bs, len(self.jobs), starts[-1])
        raw, bs, nj, nb = self.table
        for fn in self.pre.values():
            fn()
        _bw("b2c_pack_weights_batched", sum(j['R'] * j['ntaps'] * j['C'] for j in self.jobs.values()) * 6, _p(raw), _p(bs), nj, nb, stream())
        self.flushed_epoch = epoch
        return True


PACKS: Optional[PackRegistry] = None     # the registry of the TrainStep that is currently executing (else None)


def pack_part(weight, packed, wtap_dev, R, ntaps, C, C_real, s_r, s_c, tap_pitch, col_off, r_off, bn_tile, nkb):
    if PACKS is not None:
        key = (weight.data_ptr(), packed.data_ptr(), r_off, col_off)
        PACKS.add(key, dict(w=weight.data_ptr(), packed=packed.data_ptr(), wtap=wtap_dev.data_ptr(), s_r=s_r, s_c=s_c,
                            tap_pitch=tap_pitch, col_off=col_off, R=R, ntaps=ntaps, C=C, C_real=C_real, r_off=r_off,
                            bn_tile=bn_tile, nkb=nkb, dtype=PREC.mode))
    _abi.call("b2c_pack_weights", _p(weight), _p(packed), _p(wtap_dev), R, ntaps, C, C_real, s_r, s_c, tap_pitch, col_off,
              r_off, bn_tile, nkb, stream())


# ---- layout ---------------------------------------------------------------------------------
def ncdhw_to_cl(x: torch.Tensor, cpad: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(N,C,T,H,W) fp32 contiguous -> (N,T,H,W,cpad) bf16 (optionally into a preallocated slice)."""
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 5
    N, Cc, T, H, W = x.shape
    if out is None:
        out = torch.empty((N, T, H, W, cpad), dtype=act_dtype(), device=x.device)
    assert tuple(out.shape) == (N, T, H, W, cpad) and out.is_contiguous() and out.dtype == act_dtype()
    _bw("b2c_ncdhw_to_ndhwc", x.numel() * 4 + out.numel() * out.element_size(), _p(x), _p(out), N, Cc, T * H * W, cpad, stream())
    return out


def u8_clip_to_cl(u8: torch.Tensor, out: torch.Tensor):
    """(P,C,T,H,W) uint8 -> out (2P,T,H,W,8): [clips / 255 ; the clips mirrored in W] in the activation precision."""
    assert u8.dtype == torch.uint8 and u8.is_contiguous() and u8.dim() == 5
    P, Cc, T, H, W = u8.shape
    assert tuple(out.shape) == (2 * P, T, H, W, 8) and out.is_contiguous() and out.dtype == act_dtype()
    _bw("b2c_u8_clip_to_cl", u8.numel() + out.numel() * out.element_size(), _p(u8), _p(out), P, Cc, T * H, W, stream())
    return out


def u8_to_f32(u8: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    assert u8.dtype == torch.uint8 and u8.is_contiguous()
    out = torch.empty(u8.shape, dtype=torch.float32, device=u8.device)
    _bw("b2c_u8_to_f32", u8.numel() * 5, _p(u8), _p(out), u8.numel(), float(scale), stream())
    return out


def cl_to_ncdhw_f32(v: View) -> torch.Tensor:
    N = v.N
    T, H, W = v.dims
    out = torch.empty((N, v.C, T, H, W), dtype=torch.float32, device=v.t.device)
    _abi.call("b2c_ndhwc_to_ncdhw_f32", v.ptr, v.row_stride, v.c_off, _p(out), N, v.C, T * H * W, stream())
    return out


def im2col_small(x: View, out: torch.Tensor, C, out_dims, k, s, pf, Kpad):
    _bw("b2c_im2col_small", x.rows * C * x.t.element_size() + out.numel() * out.element_size(), x.ptr, _p(out), x.N, x.row_stride, C, *x.dims, *out_dims, *k, *s, *pf, Kpad, stream())


def stem_fold_input(x: View, xs: torch.Tensor, pt: int, Tp: int):
    T, H, W = x.dims
    _bw("b2c_stem_fold_input", _vb(x) // 8 * 3 + xs.numel() * xs.element_size(), x.ptr, _p(xs), x.N, T, H, W, x.row_stride, pt, Tp, stream())


def clips_to_folded(clips: torch.Tensor, xs: torch.Tensor, pt: int, Tp: int, mirror: bool):
    """clips (P,C,T,H,W) fp32 / uint8 -> folded stem input xs (N',1,H,W,Tp*4) (rows [0,P); with mirror also [P,2P))."""
    assert clips.is_contiguous() and clips.dim() == 5 and clips.dtype in (torch.float32, torch.uint8)
    P, Cc, T, H, W = clips.shape
    assert xs.is_contiguous() and xs.dtype == act_dtype() and tuple(xs.shape[1:]) == (1, H, W, Tp * 4) and xs.shape[0] >= (2 * P if mirror else P)
    _bw("b2c_clips_to_folded", clips.numel() * clips.element_size() + (2 if mirror else 1) * P * H * W * Tp * 4 * xs.element_size(),
        _p(clips), int(clips.dtype == torch.uint8), _p(xs), P, Cc, T, H, W, pt, Tp, int(mirror), stream())


def stem_fold_weights(w, w2, cout, cin, kt, khw, st, To, Kf):
    _abi.call("b2c_stem_fold_weights", _p(w), _p(w2), cout, cin, kt, khw, st, To, Kf, stream())


def stem_unfold_wgrad(dw2, dw, cout, cin, kt, khw, st, To, Kf):
    _abi.call("b2c_stem_unfold_wgrad", _p(dw2), _p(dw), cout, cin, kt, khw, st, To, Kf, stream())


# ---- batch norm -----------------------------------------------------------------------------
BN_FUSE_FINALIZE = os.environ.get("B2C_BN_FUSE_FINALIZE", "1") != "0"


def bn_relu_fwd(x: View, groups, ws, mean, rstd, rm, rv, momentum, eps, gamma, beta, y: View, relu=True):
    """Train-mode BatchNorm + ReLU of the view x into the view y: statistics + finalize (mean / rstd (groups, C) for the
    backward, running stats; done by the statistics launch's last block), apply.  Two launches on purpose: a fused single launch with a grid barrier between the
    statistics and the apply pass was built and measured slower in the captured step (cooperative launch 22.9 ms, ordinary
    launch + hand-rolled barrier 22.2 ms, separate launches 21.5 ms per step: the barrier wait exceeds what the next
    launch's overlap with this one's tail already hides)."""
    if BN_FUSE_FINALIZE:
        _bw("b2c_bn_sums_finalize", _vb(x), x.ptr, x.rows, x.C, x.row_stride, x.c_off, groups, _p(ws), _p(mean), _p(rstd), _p(rm), _p(rv),
            float(momentum), float(eps), stream())
    else:
        bn_sums(x, groups, ws)
        bn_finalize(ws, x.C, 0, x.C, groups, x.rows // groups, mean, rstd, rm, rv, momentum, eps)
    bn_relu_apply(x, groups, mean, rstd, gamma, beta, y, relu)


def bn_relu_bwd(dy: View, y, x: View, groups, mean, rstd, gamma, beta, ws, dx: View, dgamma, dbeta, relu=True):
    """Backward of bn_relu_fwd.  y = None: the kernels recompute the ReLU mask from x (forward's own arithmetic) instead of
    reading y -- 2 instead of 3 and 3 instead of 4 tensor passes."""
    bn_relu_bwd_reduce(dy, y, x, groups, mean, rstd, ws, relu, gamma, beta)
    bn_relu_bwd_apply(dy, y, x, groups, mean, rstd, gamma, ws, dx, dgamma, dbeta, relu, beta)


def bn_sums(x: View, groups: int, ws: torch.Tensor):
    _bw("b2c_bn_sums", _vb(x), x.ptr, x.rows, x.C, x.row_stride, x.c_off, groups, _p(ws), stream())


def bn_finalize(ws, ws_C, c_off, C, groups, rows_per_group, mean, rstd, rm, rv, momentum, eps):
    _abi.call("b2c_bn_finalize", _p(ws), ws_C, c_off, C, groups, rows_per_group, _p(mean), _p(rstd), _p(rm), _p(rv),
              momentum, eps, stream())


def bn_relu_apply(x: View, groups, mean, rstd, gamma, beta, y: View, relu=True):
    _bw("b2c_bn_relu_apply", 2 * _vb(x), x.ptr, x.rows, x.C, x.row_stride, x.c_off, groups, _p(mean), _p(rstd), _p(gamma),
              _p(beta), y.ptr, y.row_stride, y.c_off, int(relu), stream())


def bn_relu_bwd_reduce(dy: View, y, x: View, groups, mean, rstd, ws, relu=True, gamma=None, beta=None):
    _bw("b2c_bn_relu_bwd_reduce", (3 if y is not None else 2) * _vb(x), dy.ptr, dy.row_stride, dy.c_off, y.ptr if y is not None else None,
        y.row_stride if y is not None else 0, y.c_off if y is not None else 0, x.ptr, x.row_stride, x.c_off, x.rows, x.C, groups,
        _p(mean), _p(rstd), _p(gamma), _p(beta), _p(ws), int(relu), stream())


def bn_relu_bwd_apply(dy: View, y, x: View, groups, mean, rstd, gamma, ws, dx: View, dgamma, dbeta, relu=True, beta=None):
    _bw("b2c_bn_relu_bwd_apply", (4 if y is not None else 3) * _vb(x), dy.ptr, dy.row_stride, dy.c_off, y.ptr if y is not None else None,
        y.row_stride if y is not None else 0, y.c_off if y is not None else 0, x.ptr, x.row_stride, x.c_off, x.rows, x.C, groups,
        _p(mean), _p(rstd), _p(gamma), _p(beta), _p(ws), dx.ptr, dx.row_stride, dx.c_off, _p(dgamma), _p(dbeta), int(relu), stream())


# ---- pooling / elementwise ------------------------------------------------------------------
def maxpool_fwd(x: View, y: View, idx, k, s, p):
    _bw("b2c_maxpool_fwd", _vb(x) + _vb(y) + y.rows * y.C, x.ptr, x.row_stride, x.c_off, y.ptr, y.row_stride, y.c_off, _p(idx), x.N, x.C,
              *x.dims, *y.dims, *k, *s, *p, stream())


def maxpool_bwd(dy: View, idx, dx: View, k, s, p, accumulate=False):
    _bw("b2c_maxpool_bwd", _vb(dy) + dy.rows * dy.C + _vb(dx), dy.ptr, dy.row_stride, dy.c_off, _p(idx), dx.ptr, dx.row_stride, dx.c_off, dx.N, dx.C,
              *dx.dims, *dy.dims, *k, *s, *p, int(accumulate), stream())


def channel_scale(x: View, scale_nc, y: View):
    T, H, W = x.dims
    _bw("b2c_channel_scale", 2 * _vb(x), x.ptr, x.row_stride, x.c_off, _p(scale_nc), y.ptr, y.row_stride, y.c_off, x.N,
              T * H * W, x.C, stream())


def act_bwd(dy: View, y: Optional[View], scale_nc, dz: Optional[View], dbias, relu: bool):
    T, H, W = dy.dims
    _bw("b2c_act_bwd", _vb(dy) * (1 + (y is not None) + (dz is not None)), dy.ptr, dy.row_stride, dy.c_off, y.ptr if y is not None else None,
              y.row_stride if y is not None else 0, y.c_off if y is not None else 0, _p(scale_nc),
              dz.ptr if dz is not None else None, dz.row_stride if dz is not None else 0,
              dz.c_off if dz is not None else 0, _p(dbias), dy.N, T * H * W, dy.C, int(relu), stream())


def add(a: View, b: View, out: View):
    _bw("b2c_add", 3 * _vb(a), a.ptr, a.row_stride, a.c_off, b.ptr, b.row_stride, b.c_off, out.ptr, out.row_stride, out.c_off,
              a.rows, a.C, stream())


def stencil27_fwd(P, out, bias, N, T, H, W):
    _bw("b2c_stencil27_fwd", N * T * H * W * 4 * 28, _p(P), _p(out), _p(bias), N, T, H, W, stream())


def stencil27_bwd(dout, dP, dbias, N, T, H, W, cpad=32):
    _bw("b2c_stencil27_bwd", N * T * H * W * (4 + cpad * dP.element_size()), _p(dout), _p(dP), _p(dbias), N, T, H, W, cpad, stream())


# ---- collapsed decoder tail (upsample4 -> Dropout3d -> smooth as one per-clip transposed convolution) --------------
def tail_weff(w4, b4, ws, drop_nc, packed_f, f_stride, packed_d, d_stride, d_nkb, biasfield, N):
    _bw("b2c_tail_weff", 128 * 128 * 27 * 4 + N * (216 * 128 * 2) * packed_f.element_size(), _p(w4), _p(b4), _p(ws), _p(drop_nc), _p(packed_f), f_stride, _p(packed_d), d_stride, d_nkb,
              _p(biasfield), N, stream())


def tail_gather_fwd(y_planar, biasfield, bs, logits, N, It, Ih, Iw):
    _bw("b2c_tail_gather_fwd", N * It * Ih * Iw * (216 * 4 + 8 * 4), _p(y_planar), _p(biasfield), _p(bs), _p(logits), N, It, Ih, Iw, stream())


def tail_gather_bwd(dlogits, dy, class_sums, N, It, Ih, Iw):
    _bw("b2c_tail_gather_bwd", N * It * Ih * Iw * (224 * dy.element_size() + 8 * 4 * 2), _p(dlogits), _p(dy), _p(class_sums), N, It, Ih, Iw, stream())


def tail_chain_bwd(dweff, class_sums, w4, b4, ws, drop_nc, dw4, db4, dws, dbs, N):
    _bw("b2c_tail_chain_bwd", N * 128 * 216 * 4 + 128 * 128 * 27 * 8, _p(dweff), _p(class_sums), _p(w4), _p(b4), _p(ws), _p(drop_nc), _p(dw4), _p(db4), _p(dws),
              _p(dbs), N, stream())


# ---- capsule head ---------------------------------------------------------------------------
def em_routing_fwd(caps, W, beta_u, beta_a, out, b, C, state=None):
    """state: optional fp32 (b, routing_state_floats()) buffer -- the training forward saves the per-iteration routing
    state there for em_routing_bwd."""
    if state is None:
        _bw("b2c_em_routing_fwd", b * (544 + C * 17) * 4, _p(caps), _p(W), _p(beta_u), _p(beta_a), _p(out), b, C, stream())
    else:
        # the forward writes 8672 of the state's floats per location (assignments, normalisers, means, variances)
        _bw("b2c_em_routing_fwd_train", b * (544 + C * 17 + 8672) * 4, _p(caps), _p(W), _p(beta_u), _p(beta_a), _p(out),
            _p(state), b, C, stream())


def em_routing_bwd(caps, W, beta_u, beta_a, dout, dcaps, dW, dbu, dba, b, C, state=None):
    if state is None:
        _bw("b2c_em_routing_bwd", b * 2 * (544 + C * 17) * 4, _p(caps), _p(W), _p(beta_u), _p(beta_a), _p(dout), _p(dcaps), _p(dW),
            _p(dbu), _p(dba), b, C, stream())
    else:
        # coefficient kernel: reads what the forward saved, writes 2048 + 4128 floats; final kernel: reads the 11264-float prefix
        _bw("b2c_em_routing_bwd_state", b * (3 * 544 + 2 * C * 17 + 8672 + 6176 + 11264) * 4, _p(caps), _p(W), _p(beta_u), _p(beta_a), _p(dout),
            _p(state), _p(dcaps), _p(dW), _p(dbu), _p(dba), b, C, stream())


def routing_state_floats() -> int:
    return int(_abi.lib().b2c_em_routing_state_floats())


def primarycaps_finish(part, bias, out):
    """part (N, S, h, w, 544) fp32 partial sums of the S K-slices -> out (N, 1, h, w, 544) = sum + bias, sigmoid on the last 32"""
    N, S = part.shape[0], part.shape[1]
    L = part.shape[2] * part.shape[3]
    _bw("b2c_primarycaps_finish", (part.numel() + out.numel()) * 4, _p(part), S, _p(bias), _p(out), N, L, stream())


def primarycaps_bwd_prep2(g, out, dz, dz_rows, dbias, N, Hq, Wq, dz_pitch=544):
    """primarycaps_bwd_prep that also writes dz in (h, w, n) order for the rows-major dgrad"""
    _bw("b2c_primarycaps_bwd_prep2", N * Hq * Wq * (544 * 8 + 2 * dz_pitch * dz.element_size()), _p(g), _p(out), _p(dz), _p(dz_rows),
        _p(dbias), N, Hq, Wq, dz_pitch, stream())


def rows_to_clips(src, dst: View):
    """compact (1, H, W, N, C) -> the channel window `dst` of a clip-major (N, 1, H, W, Ctot) tensor"""
    _, H, W, N, C = src.shape
    assert dst.N == N and dst.dims == (1, H, W) and dst.C == C
    _bw("b2c_rows_to_clips", 2 * N * H * W * C * src.element_size(), _p(src), dst.ptr, dst.row_stride, dst.c_off, N, H, W, C, stream())


def clips_to_rows(src: View, dst):
    """channel window of a clip-major (N, 1, H, W, Ctot) tensor -> compact (1, H, W, N, C)"""
    _, H, W, N, C = dst.shape
    assert src.N == N and src.dims == (1, H, W) and src.C == C
    _bw("b2c_clips_to_rows", 2 * N * H * W * C * dst.element_size(), src.ptr, src.row_stride, src.c_off, _p(dst), N, H, W, C, stream())


def primarycaps_bwd_prep(g, out, dz, dbias, rows, dz_pitch=544):
    _bw("b2c_primarycaps_bwd_prep", rows * (544 * 8 + dz_pitch * dz.element_size()), _p(g), _p(out), _p(dz), _p(dbias), rows, dz_pitch, stream())


def class_mean_fwd(rout, act, N, L, C):
    _abi.call("b2c_class_mean_fwd", _p(rout), _p(act), N, L, C, stream())


def pose_mask_fwd(rout, mask, x, N, L, C):
    _abi.call("b2c_pose_mask_fwd", _p(rout), _p(mask), _p(x), N, L, C, stream())


def caps_head_bwd(dx, mask, dact, dfeat, drout, N, L, C):
    _abi.call("b2c_caps_head_bwd", _p(dx), _p(mask), _p(dact), _p(dfeat), _p(drout), N, L, C, stream())


# ---- losses ---------------------------------------------------------------------------------
def seg_loss_fwd(logits, targets, lab_idx, n_lab, V, sums, loss):
    _bw("b2c_seg_loss_fwd", n_lab * V * 8, _p(logits), _p(targets), _p(lab_idx), n_lab, V, _p(sums), _p(loss), stream())


def seg_loss_bwd(logits, targets, lab_idx, n_lab, V, sums, w_bce, w_dice, dlogits):
    _bw("b2c_seg_loss_bwd", n_lab * V * 12, _p(logits), _p(targets), _p(lab_idx), n_lab, V, _p(sums), float(w_bce), float(w_dice),
              _p(dlogits), stream())


def spread_loss(act, target, lab_idx, n_lab, C, m_min, loss, w, dact):
    _abi.call("b2c_spread_loss", _p(act), _p(target), _p(lab_idx), n_lab, C, float(m_min), _p(loss), float(w), _p(dact),
              stream())


def bv_mask(pred, flip_pred, m, mm, P, H, W, frames_cnt, use_sig, pred_tflip=0, fp_tflip=0, fp_wmirror=0):
    _bw("b2c_bv_mask", P * 8 * H * W * 4 * 3, _p(pred), _p(flip_pred), _p(m), _p(mm), P, H, W, frames_cnt, int(use_sig), int(pred_tflip),
              int(fp_tflip), int(fp_wmirror), stream())


def gv_mask(out, m, mm, P, H, W, lower, upper):
    _bw("b2c_gv_mask", P * 8 * H * W * 4 * 2, _p(out), _p(m), _p(mm), P, H, W, float(lower or 0.0), float(upper or 0.0),
              int(lower is not None), int(upper is not None), stream())


def cons_reduce(out, flp, w1, w2, wg, acc, P, H, W, mirror, w2_tflip):
    _bw("b2c_cons_reduce", P * 8 * H * W * 4 * (2 + (w1 is not None) + (w2 is not None) + (wg is not None)), _p(out), _p(flp), _p(w1), _p(w2), _p(wg), _p(acc), P, H, W, int(mirror), int(w2_tflip),
              stream())


def cons_finish(acc, loss, P, H, W, mode, wt_ramp, bv_wt, gv_wt, dev_scalars=None):
    """dev_scalars: optional device float[3] (wt_ramp, bv_wt, gv_wt) read by the kernel instead of the host values."""
    _abi.call("b2c_cons_finish", _p(acc), _p(loss), P, H, W, mode, float(wt_ramp), float(bv_wt), float(gv_wt), _p(dev_scalars),
              stream())


def cons_grad(out, flp, w1, w2, wg, dout, dflp, P, H, W, mirror, w2_tflip, a_l2, a_lv, a_lg, dev_scalars=None):
    _bw("b2c_cons_grad", P * 8 * H * W * 4 * (4 + (w1 is not None) + (w2 is not None) + (wg is not None)), _p(out), _p(flp), _p(w1), _p(w2), _p(wg), _p(dout), _p(dflp), P, H, W, int(mirror),
              int(w2_tflip), float(a_l2), float(a_lv), float(a_lg), _p(dev_scalars), stream())


def frame_iou_counts(logits, gt):
    """logits, gt: (..., H, W) fp32 CUDA with the same leading shape -> int32 (frames, 3): intersection, union, gt pixels."""
    assert logits.shape == gt.shape and logits.is_contiguous() and gt.is_contiguous()
    HW = logits.shape[-1] * logits.shape[-2]
    frames = logits.numel() // HW
    out = torch.empty((frames, 3), dtype=torch.int32, device=logits.device)
    _bw("b2c_frame_iou_counts", logits.numel() * 8, _p(logits), _p(gt), _p(out), frames, HW, stream())
    return out


def adam_step(p, g, m, v, n, lr, beta1, beta2, eps, step_dev, grad_scale=1.0, lr_dev=None):
    """step_dev: int32 device tensor holding the number of steps taken so far (incremented by the call);
    lr_dev: optional device float that overrides `lr` (graph replays with a scheduler-controlled learning rate)."""
    _bw("b2c_adam_step", n * 28, _p(p), _p(g), _p(m), _p(v), n, float(lr), float(beta1), float(beta2), float(eps), _p(step_dev),
              float(grad_scale), _p(lr_dev), stream())


def fill_f32(t, v):
    _bw("b2c_fill_f32", t.numel() * 4, _p(t), t.numel(), float(v), stream())


def set_deterministic(on: bool):
    """Tests only: bit-reproducible BatchNorm reductions (one block per statistic group)."""
    _abi.call("b2c_set_deterministic", int(bool(on)))


def set_pool_generic(on: bool):
    """Tests only: run the generic max-pool kernels where a specialised row kernel exists."""
    _abi.call("b2c_set_pool_generic", int(bool(on)))
