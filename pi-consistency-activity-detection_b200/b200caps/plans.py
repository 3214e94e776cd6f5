"""Host-side planning for the generalised implicit-GEMM convolution (csrc/igemm.cu).

A ``ConvPlan`` turns one torch-style convolution / transposed convolution into the
tap tables, output-parity classes and packed bf16 weight tiles the tcgen05 kernels consume:

  * conv fprop                         -> 1 class, taps d = k - pad_front, input stride = conv stride
  * conv dgrad / convT fprop           -> prod(stride) output-parity classes, only the taps that hit
                                          real inputs (no MACs on inserted zeros)
  * convT dgrad                        -> 1 class (a strided conv over the output gradient)
  * wgrad (conv and convT)             -> 1 launch, positions are the GEMM-K dimension

Pure host logic (unit-tested on CPU); device work happens only in ``ops``.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

from . import _abi


# Zero-padded channel tail: the packed weights give every tap ceil64(C) K-columns, so layers whose input channel count
# is not a multiple of 64 (Inception 3x3x3 inputs 16..160, the 480 / 528-channel 1x1 inputs, JHMDB's 336) run on the TMA
# im2col path -- TMA zero-fills the channels past C -- instead of the cp.async gather path.  B2C_TMA_TAIL=0 disables.
TMA_TAIL = os.environ.get("B2C_TMA_TAIL", "1") != "0"


class _Precision:
    """Process-wide precision mode of the activation tensors (include/b200caps.h: b2c_set_precision).
    0 = bf16 activations / bf16 operands; 1 = fp32 activations / tf32 operands (the reference's fp32 arithmetic,
    main_ucf101.py:52-55, within tensor-core tf32 rounding)."""
    mode = 0


PREC = _Precision()


def set_precision(name) -> None:
    """'bf16' (default) or 'tf32'.  Set it BEFORE building a TrainStep / running a model: packed operands and
    activation buffers created under one mode are not valid under the other (caches are keyed by the mode)."""
    mode = {"bf16": 0, "tf32": 1, "fp32": 1, 0: 0, 1: 1}[name]
    _abi.call("b2c_set_precision", mode)
    PREC.mode = mode


def precision() -> str:
    return "tf32" if PREC.mode else "bf16"


def act_dtype() -> torch.dtype:
    return torch.float32 if PREC.mode else torch.bfloat16


def kblock() -> int:
    """K elements per 128-byte swizzle row of a GEMM operand."""
    return 32 if PREC.mode else 64


# widest output-channel tile of layers with more than 256 output channels (PrimaryCaps 544, its dgrad 832)
BN_TILE_MAX = int(os.environ.get("B2C_BN_TILE_MAX", "256"))


def tap_pitch(C: int) -> int:
    """K columns per tap of a packed operand over C stored channels."""
    if PREC.mode:
        return (C + 31) // 32 * 32           # tf32 mode runs on the TMA path only
    return (C + 63) // 64 * 64 if (TMA_TAIL and C % 64) else C


def pick_bn_tile(cout: int) -> int:
    """Output-channel tile per CTA (UMMA N): the whole Cout when <= 256, else an even split (multiple of 16)."""
    if cout <= 256:
        return (cout + 15) // 16 * 16
    nt = (cout + BN_TILE_MAX - 1) // BN_TILE_MAX
    return ((cout + nt - 1) // nt + 15) // 16 * 16


def packed_geometry(rows_pad: int, K: int):
    """(bn_tile, n_tiles, nkb, elements) of the pre-swizzled weight operand."""
    bn = pick_bn_tile(rows_pad)
    nt = (rows_pad + bn - 1) // bn
    kb = kblock()
    nkb = (K + kb - 1) // kb
    return bn, nt, nkb, nt * nkb * bn * kb


def _tap_word(dt: int, dh: int, dw: int) -> int:
    for v in (dt, dh, dw):
        assert -128 <= v <= 127
    w = (dt & 0xFF) | ((dh & 0xFF) << 8) | ((dw & 0xFF) << 16)
    return w


@dataclass
class TapClass:
    taps: List[Tuple[int, int, int]]   # input offsets (dt, dh, dw)
    wtap: List[int]                    # linear tap offset inside the torch weight's kernel dims
    Q: Tuple[int, int, int]            # GEMM-M grid (per sample)
    po: Tuple[int, int, int]           # output offset
    taps_dev: Optional[torch.Tensor] = None
    wtap_dev: Optional[torch.Tensor] = None
    taps_host: Optional[object] = None       # ctypes int32 array (kept alive with the plan)
    packed: Optional[torch.Tensor] = None


def conv_out_size(i: int, k: int, s: int, pf: int, pb: int) -> int:
    return (i + pf + pb - k) // s + 1


def convT_out_size(i: int, k: int, s: int, p: int, op: int) -> int:
    return (i - 1) * s - 2 * p + k + op


def same_pad(size: int, k: int, s: int) -> Tuple[int, int]:
    """TF 'same' padding of the reference (pytorch_i3d.py:15-19, :82-86): front = pad // 2."""
    pad = max(k - s, 0) if size % s == 0 else max(k - (size % s), 0)
    return pad // 2, pad - pad // 2


def _dim_conv_like(k: int, p: int):
    """taps of a gather 'in[q*s + (kk - p)]' for one dimension: list of (kk, d)."""
    return [(kk, kk - p) for kk in range(k)]


def _dim_transposed_like(k: int, s: int, p: int, out_size: int):
    """Per output parity r (o = s*q + r): taps kk with (r + p - kk) % s == 0 reading in[q + d],
    d = (r + p - kk) // s.  Returns list over r of (taps [(kk, d)], Q_r)."""
    out = []
    for r in range(s):
        taps = [(kk, (r + p - kk) // s) for kk in range(k) if (r + p - kk) % s == 0]
        q = max(0, -(-(out_size - r) // s))
        out.append((taps, q))
    return out


def _prune_padding_taps(classes, si, in_dims):
    """Drop the taps that read nothing but padding for EVERY output position of their class (e.g. the dt = +-1 taps of the
    3x3x3 convolutions of Mixed_4b..4f, whose input has a single frame: 18 of their 27 taps).  Exact: such taps contribute
    zero to fprop / dgrad and receive a zero weight gradient (which is what the zero-initialised dw already holds)."""
    for cl in classes:
        keep = []
        for idx, tap in enumerate(cl.taps):
            ok = True
            for d, s, I, Q in zip(tap, si, in_dims, cl.Q):
                q_lo = max(0, -(d // s)) if d < 0 else 0          # smallest q with q*s + d >= 0
                q_hi = min(Q - 1, (I - 1 - d) // s)               # largest q with q*s + d <= I - 1
                ok = ok and q_lo <= q_hi
            if ok:
                keep.append(idx)
        if not keep:
            keep = [0]
        if len(keep) != len(cl.taps):
            cl.taps = [cl.taps[i] for i in keep]
            cl.wtap = [cl.wtap[i] for i in keep]


def _product_classes(per_dim, kdims):
    """per_dim: for each of 3 dims a list over parity of (taps, Q).  Build the class list."""
    classes = []
    kt, kh, kw = kdims
    for rt, (tt, qt) in enumerate(per_dim[0]):
        for rh, (th, qh) in enumerate(per_dim[1]):
            for rw, (tw, qw) in enumerate(per_dim[2]):
                taps, wtap = [], []
                for (a, da) in tt:
                    for (b, db) in th:
                        for (c, dc) in tw:
                            taps.append((da, db, dc))
                            wtap.append((a * kh + b) * kw + c)
                if qt * qh * qw == 0:
                    continue
                classes.append(TapClass(taps, wtap, (qt, qh, qw), (rt, rh, rw)))
    return classes


@dataclass
class ConvSpec:
    """A torch convolution (transposed=False: weight (Cout,Cin,k..)) or transposed convolution
    (transposed=True: weight (Cin,Cout,k..)).  2-D layers use kt=1."""
    Cin: int
    Cout: int
    k: Tuple[int, int, int]
    stride: Tuple[int, int, int] = (1, 1, 1)
    pad_front: Tuple[int, int, int] = (0, 0, 0)   # conv: explicit front pad ; convT: `padding`
    pad_back: Tuple[int, int, int] = (0, 0, 0)    # conv only
    out_pad: Tuple[int, int, int] = (0, 0, 0)     # convT only
    transposed: bool = False
    Cin_pad: int = 0                              # channels of the stored input tensor (>= Cin, % 8 == 0)
    Cout_pad: int = 0                             # channels of the stored output tensor

    def __post_init__(self):
        if not self.Cin_pad:
            self.Cin_pad = (self.Cin + 7) // 8 * 8
        if not self.Cout_pad:
            self.Cout_pad = (self.Cout + 7) // 8 * 8

    def out_dims(self, in_dims):
        if self.transposed:
            return tuple(convT_out_size(i, k, s, p, op) for i, k, s, p, op in
                         zip(in_dims, self.k, self.stride, self.pad_front, self.out_pad))
        return tuple(conv_out_size(i, k, s, pf, pb) for i, k, s, pf, pb in
                     zip(in_dims, self.k, self.stride, self.pad_front, self.pad_back))


class ConvPlan:
    """Geometry of one layer at one input size.  ``in_dims`` = (T, H, W) of the layer input."""

    def __init__(self, spec: ConvSpec, in_dims: Sequence[int]):
        self.spec = spec
        self.in_dims = tuple(int(v) for v in in_dims)
        self.out_dims = spec.out_dims(self.in_dims)
        k, s, p = spec.k, spec.stride, spec.pad_front
        T = k[0] * k[1] * k[2]
        self.ntaps_full = T
        if not spec.transposed:
            # weight (Cout, Cin, T)
            per = [[(_dim_conv_like(k[i], p[i]), self.out_dims[i])] for i in range(3)]
            self.fprop = _product_classes(per, k)
            self.fprop_si, self.fprop_so = s, (1, 1, 1)
            self.fprop_pack = dict(R=spec.Cout, R_pad=spec.Cout_pad, C=spec.Cin_pad, C_real=spec.Cin, s_r=spec.Cin * T, s_c=T)
            perd = [_dim_transposed_like(k[i], s[i], p[i], self.in_dims[i]) for i in range(3)]
            self.dgrad = _product_classes(perd, k)
            self.dgrad_si, self.dgrad_so = (1, 1, 1), s
            self.dgrad_pack = dict(R=spec.Cin, R_pad=spec.Cin_pad, C=spec.Cout_pad, C_real=spec.Cout, s_r=T, s_c=spec.Cin * T)
            # wgrad: g = x gathered with the fprop taps, p = dy at plain positions
            self.wgrad_cls = self.fprop[0]
            self.wgrad_geom = dict(g_is_input=True, sg=s, sp=(1, 1, 1), s_p=spec.Cin * T, s_g=T,
                                   Cg=spec.Cin_pad, Cg_real=spec.Cin, Cp=spec.Cout_pad, Cp_real=spec.Cout, Q=self.out_dims)
        else:
            # weight (Cin, Cout, T)
            perf = [_dim_transposed_like(k[i], s[i], p[i], self.out_dims[i]) for i in range(3)]
            self.fprop = _product_classes(perf, k)
            self.fprop_si, self.fprop_so = (1, 1, 1), s
            self.fprop_pack = dict(R=spec.Cout, R_pad=spec.Cout_pad, C=spec.Cin_pad, C_real=spec.Cin, s_r=T, s_c=spec.Cout * T)
            perd = [[(_dim_conv_like(k[i], p[i]), self.in_dims[i])] for i in range(3)]
            self.dgrad = _product_classes(perd, k)
            self.dgrad_si, self.dgrad_so = s, (1, 1, 1)
            self.dgrad_pack = dict(R=spec.Cin, R_pad=spec.Cin_pad, C=spec.Cout_pad, C_real=spec.Cout, s_r=spec.Cout * T, s_c=T)
            # wgrad: g = dOut gathered with the dgrad taps, p = x at plain positions
            self.wgrad_cls = self.dgrad[0]
            self.wgrad_geom = dict(g_is_input=False, sg=s, sp=(1, 1, 1), s_p=spec.Cout * T, s_g=T,
                                   Cg=spec.Cout_pad, Cg_real=spec.Cout, Cp=spec.Cin_pad, Cp_real=spec.Cin, Q=self.in_dims)
        _prune_padding_taps(self.fprop, self.fprop_si, self.in_dims)
        _prune_padding_taps(self.dgrad, self.dgrad_si, self.out_dims)
        self._device = None

    def split_fprop_k(self, nsplit: int) -> None:
        """Turn the single fprop class of a 2-D layer (one output frame) into `nsplit` classes over disjoint tap ranges,
        class s writing its partial sums to output frame s: the K dimension of the GEMM is split over scheduling units and a
        small kernel adds the frames (ops.primarycaps_finish) -- fixed order, bit-reproducible.  For a layer whose tile
        count is a poor multiple of the SM count (PrimaryCaps: 300 tiles on 148 SMs = 3 rounds for 2.03 rounds of work) the
        finer units fill the last round."""
        assert len(self.fprop) == 1 and not self.spec.transposed and self._device is None and self.out_dims[0] == 1
        cl = self.fprop[0]
        nt = len(cl.taps)
        nsplit = max(1, min(nsplit, nt, 8))
        if nsplit == 1:
            return
        cuts = [nt * i // nsplit for i in range(nsplit + 1)]
        self.fprop = [TapClass(cl.taps[a:b], cl.wtap[a:b], cl.Q, (s, 0, 0)) for s, (a, b) in enumerate(zip(cuts, cuts[1:]))]
        self.fprop_out_dims = (nsplit,) + tuple(self.out_dims[1:])     # the GEMM's output tensor: one frame per K slice

    @staticmethod
    def pointwise_from_strides(Cin: int, Cout: int, Cout_pad: int, s_co: int, s_ci: int, dims) -> "ConvPlan":
        """A 1x1x1 convolution whose (Cout, Cin) matrix lives inside another tensor with element strides
        (s_co, s_ci) -- used for `smooth`: W[c][0][k] viewed as the 128 -> 27 projection (s_co=1, s_ci=27)."""
        pl = ConvPlan(ConvSpec(Cin, Cout, (1, 1, 1), Cout_pad=Cout_pad), dims)
        pl.fprop_pack = dict(R=Cout, R_pad=Cout_pad, C=Cin, C_real=Cin, s_r=s_co, s_c=s_ci)
        pl.dgrad_pack = dict(R=Cin, R_pad=Cin, C=Cout_pad, C_real=Cout, s_r=s_ci, s_c=s_co)
        pl.wgrad_geom = dict(g_is_input=True, sg=(1, 1, 1), sp=(1, 1, 1), s_p=s_co, s_g=s_ci, Cg=Cin, Cg_real=Cin,
                             Cp=Cout_pad, Cp_real=Cout, Q=tuple(dims))
        return pl

    # ---- algorithmic work (for roofline accounting) ------------------------------------
    def macs_fprop(self, n: int) -> int:
        return n * sum(c.Q[0] * c.Q[1] * c.Q[2] * len(c.taps) for c in self.fprop) * self.spec.Cin * self.spec.Cout

    # ---- device state -------------------------------------------------------------------
    def to(self, device):
        if self._device == device:
            return self
        extra = [self.wgrad_cls] if not any(self.wgrad_cls is c for c in self.fprop + self.dgrad) else []
        for cl in self.fprop + self.dgrad + extra:
            words = [_tap_word(*t) for t in cl.taps]
            cl.taps_dev = torch.tensor(words, dtype=torch.int32, device=device)
            cl.taps_host = (C.c_int32 * len(words))(*words)
            cl.wtap_dev = torch.tensor(cl.wtap, dtype=torch.int32, device=device)
        self._device = device
        return self

    def pack(self, weight: torch.Tensor, which: str, stream_ptr: int):
        """(Re)build the packed bf16 weight tiles of the fprop or dgrad classes from the fp32 master."""
        assert weight.dtype == torch.float32 and weight.is_contiguous() and weight.is_cuda
        classes = self.fprop if which == "fprop" else self.dgrad
        pk = self.fprop_pack if which == "fprop" else self.dgrad_pack
        pitch = pk.get("pitch") or tap_pitch(pk["C"])
        for cl in classes:
            nt = len(cl.taps)
            bn, ntile, nkb, elems = packed_geometry(pk["R_pad"], nt * pitch)
            if cl.packed is None or cl.packed.dtype != act_dtype():
                cl.packed = torch.zeros(elems, dtype=act_dtype(), device=weight.device)
            from . import ops
            ops.pack_part(weight, cl.packed, cl.wtap_dev, pk["R"], nt, pk["C"], pk["C_real"], pk["s_r"], pk["s_c"], pitch, 0, 0,
                          bn, nkb)


def t_block_of(taps) -> int:
    """Taps per block of constant dt when the tap list is such blocks with strictly monotonic dt, every block walking the
    same strictly monotonic dh at one dw (a 2-D layer with its image rows on the T axis and its columns on the H axis,
    ConvPlan.rows_major), else 0."""
    if len(taps) < 2:
        return 0
    nw = 1
    while nw < len(taps) and taps[nw][0] == taps[0][0]:
        nw += 1
    if len(taps) % nw or nw == len(taps):
        return 0
    dts = []
    inner = [t[1] for t in taps[:nw]]
    for b in range(len(taps) // nw):
        blk = taps[b * nw:(b + 1) * nw]
        if len({t[0] for t in blk}) != 1 or [t[1] for t in blk] != inner or len({t[2] for t in blk}) != 1:
            return 0
        dts.append(blk[0][0])
    mono = lambda v: all(b > a for a, b in zip(v, v[1:])) or all(b < a for a, b in zip(v, v[1:])) or len(v) == 1
    return nw if (mono(dts) and mono(inner)) else 0


def h_block_of(taps) -> int:
    """Taps per block of constant (dt, dh) when the tap list is such blocks with one dt overall and strictly monotonic dh
    (2-D layers in (h, w) product order), else 0.  The kernel then skips blocks that only read padding (b2c_conv_class)."""
    # Off by default: measured no gain at the step's shapes (PrimaryCaps dgrad 1.51 -> 1.57 ms, upsample1 0.335 -> 0.336):
    # with the static round-robin tile order every CTA's tiles sit at the same phase of the 784-row clip period
    # (37 m-tiles between them = 6.04 periods), so the CTAs that own interior rows skip nothing and set the kernel time.
    if len(taps) < 2 or len({t[0] for t in taps}) != 1 or os.environ.get("B2C_TAP_SKIP", "0") != "1":
        return 0
    nw = 1
    while nw < len(taps) and taps[nw][1] == taps[0][1]:
        nw += 1
    if len(taps) % nw or nw == len(taps):
        return 0
    dhs = []
    for b in range(len(taps) // nw):
        blk = taps[b * nw:(b + 1) * nw]
        if len({t[1] for t in blk}) != 1:
            return 0
        dhs.append(blk[0][1])
    inc = all(b > a for a, b in zip(dhs, dhs[1:]))
    dec = all(b < a for a, b in zip(dhs, dhs[1:]))
    return nw if (inc or dec) else 0


@dataclass
class View:
    """Channels-last bf16/fp32 activation view: tensor (N,T,H,W,Ctot), channel window [c_off, c_off+C)."""
    t: torch.Tensor
    c_off: int = 0
    C: int = -1

    def __post_init__(self):
        assert self.t.dim() == 5 and self.t.is_contiguous(), (self.t.shape, self.t.stride())
        if self.C < 0:
            self.C = self.t.shape[-1] - self.c_off

    @property
    def N(self):
        return self.t.shape[0]

    @property
    def dims(self):
        return tuple(self.t.shape[1:4])

    @property
    def rows(self):
        return self.t.shape[0] * self.t.shape[1] * self.t.shape[2] * self.t.shape[3]

    @property
    def row_stride(self):
        return self.t.shape[-1]

    @property
    def ptr(self):
        return self.t.data_ptr()


def fill_conv_desc(plan: ConvPlan, which: str, x: View, out: View, bias=None, scale_nc=None, relu=False,
                   sigmoid_from=-1, accumulate=False, bn_tile=0, final=False) -> _abi.ConvDesc:
    """which = 'fprop' (x = layer input, out = layer output) or 'dgrad' (x = dY, out = dX).
    final (tf32 mode): the output is not a later GEMM's operand -- keep full fp32 instead of rounding to tf32."""
    classes = plan.fprop if which == "fprop" else plan.dgrad
    si, so = (plan.fprop_si, plan.fprop_so) if which == "fprop" else (plan.dgrad_si, plan.dgrad_so)
    pk = plan.fprop_pack if which == "fprop" else plan.dgrad_pack
    exp_in, exp_out = (plan.in_dims, getattr(plan, "fprop_out_dims", plan.out_dims)) if which == "fprop" else (plan.out_dims, plan.in_dims)
    assert x.dims == tuple(exp_in), (which, x.dims, exp_in)
    assert x.C == pk["C"], (which, x.C, pk["C"])
    assert x.t.dtype == act_dtype(), (x.t.dtype, precision())
    d = _abi.ConvDesc()
    d.dtype = PREC.mode
    d.inp = x.ptr
    d.bias = bias.data_ptr() if bias is not None else None
    d.scale_nc = scale_nc.data_ptr() if scale_nc is not None else None
    d.in_row_stride = x.row_stride
    d.in_c_off, d.Cin, d.Cout = x.c_off, pk["C"], pk["R_pad"]
    d.N = x.N
    d.Ti, d.Hi, d.Wi = x.dims
    d.To, d.Ho, d.Wo = exp_out
    d.si_t, d.si_h, d.si_w = si
    d.so_t, d.so_h, d.so_w = so
    if isinstance(out, View):
        assert out.dims == tuple(exp_out) and out.C == (pk.get("out_fold") or pk["R_pad"]), (which, out.dims, exp_out, out.C, pk["R_pad"])
        assert out.t.dtype in (act_dtype(), torch.float32)
        d.out, d.out_row_stride, d.out_c_off = out.ptr, out.row_stride, out.c_off
        d.out_fp32 = 1 if out.t.dtype == torch.float32 else 0
        d.round_out = 1 if (PREC.mode and not final) else 0
    else:   # planar fp32 (Cout_pad, rows)
        rows = x.N * exp_out[0] * exp_out[1] * exp_out[2]
        assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (pk["R_pad"], rows)
        d.out, d.out_row_stride, d.out_c_off = out.data_ptr(), rows, 0
        d.out_fp32 = 2
    assert bn_tile in (0, pick_bn_tile(pk["R_pad"])), "bn_tile is fixed by the packed weight layout"
    d.relu, d.sigmoid_from, d.accumulate, d.bn_tile = int(relu), int(sigmoid_from), int(accumulate), pick_bn_tile(pk["R_pad"])
    d.tap_pitch = pk.get("pitch") or tap_pitch(pk["C"])
    d.w_sample_stride = int(pk.get("sample_stride_bytes", 0))     # per-clip weight sets (collapsed decoder tail)
    d.out_fold = int(pk.get("out_fold", 0))                       # folded stem: column blocks -> output frames
    d.nclass = len(classes)
    assert 1 <= d.nclass <= 8
    for i, cl in enumerate(classes):
        assert cl.packed is not None, "weights not packed"
        c = d.cls[i]
        c.taps, c.w, c.ntaps = cl.taps_dev.data_ptr(), cl.packed.data_ptr(), len(cl.taps)
        c.Qt, c.Qh, c.Qw = cl.Q
        c.po_t, c.po_h, c.po_w = cl.po
        c.lo_t, c.lo_h, c.lo_w = (min(t[i] for t in cl.taps) for i in range(3))
        # rows-major plans (image rows on the T axis, clips on the H axis): skip padding-only tap rows per tile
        c.h_block = -t_block_of(cl.taps) if getattr(plan, "rows_major", False) else h_block_of(cl.taps)
    return d


def fill_wgrad_desc(plan: ConvPlan, x: View, dy: View, dw: torch.Tensor, atomic=True, nsplit=0, bn_tile=0,
                    part=None, per_clip=False, force_bf16=False, pp=(0, 0, 0), segs=None) -> _abi.WgradDesc:
    """x = layer input, dy = gradient of the layer output, dw = fp32 gradient in torch weight layout.
    part=(c_off, C): restrict a fused layer's wgrad to the output-channel window of one member weight."""
    geo = dict(plan.wgrad_geom)
    cl = plan.wgrad_cls
    if part is not None:
        assert geo["g_is_input"], "fused members are plain convolutions"
        dy = View(dy.t, dy.c_off + part[0], part[1])
        geo["Cp"] = part[1]
        geo["Cp_real"] = part[2] if len(part) > 2 else part[1]
    g, p = (x, dy) if geo["g_is_input"] else (dy, x)
    assert dw.dtype == torch.float32 and dw.is_contiguous()
    assert g.C == geo["Cg"] and p.C == (geo.get("p_fold") or geo["Cp"]), (g.C, geo["Cg"], p.C, geo["Cp"])
    want = torch.bfloat16 if force_bf16 else act_dtype()
    assert g.t.dtype == want and p.t.dtype == want, (g.t.dtype, p.t.dtype, precision())
    d = _abi.WgradDesc()
    d.dtype = 0          # tf32 mode: the caller passes bf16 hi / lo splits (ops.conv_wgrad)
    d.Cp_real = geo.get("Cp_real", geo["Cp"])
    d.g, d.p, d.dw = g.ptr, p.ptr, dw.data_ptr()
    d.taps, d.wtap = cl.taps_dev.data_ptr(), cl.wtap_dev.data_ptr()
    d.taps_host = C.cast(cl.taps_host, C.c_void_p)
    d.g_row_stride, d.p_row_stride, d.s_p, d.s_g = g.row_stride, p.row_stride, geo["s_p"], geo["s_g"]
    d.g_c_off, d.p_c_off, d.Cg, d.Cp, d.Cg_real = g.c_off, p.c_off, geo["Cg"], geo["Cp"], geo["Cg_real"]
    d.N = x.N
    d.Tg, d.Hg, d.Wg = g.dims
    d.Tp, d.Hp, d.Wp = p.dims
    d.Qt, d.Qh, d.Qw = geo["Q"]
    d.sg_t, d.sg_h, d.sg_w = geo["sg"]
    d.sp_t, d.sp_h, d.sp_w = geo["sp"]
    d.pp_t, d.pp_h, d.pp_w = pp        # position offset of the plain operand
    d.p_fold = int(geo.get("p_fold", 0))   # folded stem: p-channel blocks are the output frames
    d.ntaps, d.bn_tile, d.nsplit, d.atomic = len(cl.taps), int(bn_tile), int(nsplit), int(atomic)
    if per_clip:      # dw is (N, ...): one gradient per clip
        assert dw.shape[0] == x.N
        d.dw_sample_stride = dw.numel() // x.N
    if segs:          # fused layer: [(first p channel, member dw), ...] in channel order; dw is the first member's
        assert part is None and not per_clip and 1 <= len(segs) <= 4 and segs[0][0] == 0 and segs[0][1] is dw
        d.nseg = len(segs)
        for i, (begin, t) in enumerate(segs):
            assert t.dtype == torch.float32 and t.is_contiguous()
            d.seg_begin[i], d.seg_dw[i] = int(begin), t.data_ptr()
    return d
