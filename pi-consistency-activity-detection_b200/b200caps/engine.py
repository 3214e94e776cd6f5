"""Layer-level composition of the C-ABI kernels and the torch.autograd.Function wrappers the
drop-in nn.Modules call.  Host logic only: shapes, buffers (torch = allocator), launch order.
All arithmetic happens inside libb200caps.so.

Internal activation format: contiguous (N,T,H,W,C) bf16 tensors ("CL").  Module boundaries expose
them as logical (N,C,T,H,W) views (``from_cl``) so reference code such as ``x.view(-1,832,28,28)``
keeps working without copies.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .plans import PREC, ConvPlan, ConvSpec, View, act_dtype, same_pad

import os

# Decoder tail: "collapsed" (default) = upsample4 -> Dropout3d -> smooth as one per-clip transposed convolution
# (CollapsedTail below); B2C_TAIL=explicit keeps the layer-by-layer schedule (3.3 GB intermediate at 16+16 clips).
TAIL_COLLAPSED = os.environ.get("B2C_TAIL", "collapsed") != "explicit"
# EM routing: the training forward saves the per-iteration state for the backward kernel (B2C_ROUTING_STATE=0: recompute)
ROUTING_SAVE_STATE = os.environ.get("B2C_ROUTING_STATE", "1") != "0"
# Eval mode: BatchNorm folded into the convolutions (B2C_EVAL_FOLD_BN=0: conv -> running-statistics BatchNorm kernel)
EVAL_FOLD_BN = os.environ.get("B2C_EVAL_FOLD_BN", "1") != "0"

BN_EPS = 1e-3       # pytorch_i3d.py:80
BN_MOMENTUM = 0.01  # pytorch_i3d.py:80


class _State:
    weights_epoch = 0          # bumped by the fused optimiser (raw-pointer updates bypass tensor._version)
    dropout_source: Optional[Callable] = None   # tests inject Dropout3d masks: f(n, c, device) -> (n,c) fp32 keep*2
    bn_groups = 1              # >1: the batch holds several independent forward passes (own BN statistics each)
    direct_grads = False       # fused step: kernels accumulate straight into param.grad (flat buffer views)
    zero_arena = None          # fused step: ZeroArena (all small zero-initialised workspaces of a step, ONE fill launch)
    defer_bn_counters = False  # fused step: num_batches_tracked of all BatchNorms bumped with one foreach op


STATE = _State()


class ZeroArena:
    """Bump allocator over one fp32 buffer that the fused step zeroes with a single fill at step start: replaces the
    ~90 per-layer torch.zeros launches (BatchNorm sum workspaces) of a step.  The allocation sequence of a step is
    deterministic, so under CUDA-graph capture every slice keeps its address."""

    def __init__(self, numel: int, device):
        self.buf = torch.zeros(numel, dtype=torch.float32, device=device)
        self.off = 0

    def reset(self):
        ops.fill_f32(self.buf, 0.0)
        self.off = 0

    def take(self, shape):
        n = 1
        for d in shape:
            n *= int(d)
        n4 = (n + 3) // 4 * 4
        if self.off + n4 > self.buf.numel():
            return None
        out = self.buf[self.off:self.off + n].view(*shape)
        self.off += n4
        return out


def zeros_f32(shape, device):
    """Zero-initialised fp32 workspace: a slice of the step's arena when one is active, else torch.zeros."""
    if STATE.zero_arena is not None:
        t = STATE.zero_arena.take(shape)
        if t is not None:
            return t
    return torch.zeros(shape, dtype=torch.float32, device=device)


def bump_weights_epoch():
    STATE.weights_epoch += 1


def _packed_key(*weights):
    """Cache key of derived packed operands.  When the fused step has re-packed every registered operand in one batched
    launch for the current epoch, the epoch component is dropped so per-layer packing is skipped."""
    base = (PREC.mode,) + tuple((w.data_ptr(), w._version) for w in weights)
    if ops.PACKS is not None and ops.PACKS.flushed_epoch == STATE.weights_epoch:
        return base + ("batched",)
    return base + (STATE.weights_epoch,)


def _fresh(old, key) -> bool:
    """True when the packed operands recorded under `old` are valid for `key` (same tensors and versions, and either
    the same epoch or re-packed by this epoch's batched flush)."""
    if old is None or old[:-1] != key[:-1]:
        return False
    return old[-1] == key[-1] or key[-1] == "batched"


def grad_buf(p: torch.Tensor):
    """Where a parameter gradient is accumulated: the existing .grad (fused step, zeroed once per step) or a
    fresh zero tensor that is handed back to autograd."""
    if STATE.direct_grads and p.grad is not None:
        return p.grad, True
    return torch.zeros_like(p), False


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"b200caps: {what} must be a CUDA tensor -- the hot path has no CPU fallback")


# ---- layout helpers -------------------------------------------------------------------------------
class _ToCL(torch.autograd.Function):
    """(N,C,T,H,W) fp32 -> CL bf16 (channels zero-padded to a multiple of 8)."""

    @staticmethod
    def forward(ctx, x, cpad):
        ctx.C = x.shape[1]
        return ops.ncdhw_to_cl(x.contiguous(), cpad)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        return ops.cl_to_ncdhw_f32(View(g, 0, ctx.C)), None


class _FromCLf32(torch.autograd.Function):
    """CL bf16 -> (N,C,T,H,W) fp32 contiguous (user-facing boundary of a stand-alone module)."""

    @staticmethod
    def forward(ctx, x_cl):
        ctx.cpad = x_cl.shape[-1]
        return ops.cl_to_ncdhw_f32(View(x_cl))

    @staticmethod
    def backward(ctx, g):
        return ops.ncdhw_to_cl(g.float().contiguous(), ctx.cpad)


def to_cl(x: torch.Tensor, cpad: Optional[int] = None) -> torch.Tensor:
    """Accepts logical (N,C,T,H,W) / (N,C,H,W) tensors: bf16 channels-last views pass through without a copy,
    fp32 tensors are converted by the layout kernel."""
    require_cuda(x, "activation")
    if x.dim() == 4:
        x = x.unsqueeze(2)
    assert x.dim() == 5, x.shape
    C = x.shape[1]
    cpad = cpad or (C + 7) // 8 * 8
    if x.dtype == act_dtype() and cpad == C and (x.dtype != torch.float32 or x.permute(0, 2, 3, 4, 1).is_contiguous()):
        v = x.permute(0, 2, 3, 4, 1)
        return v if v.is_contiguous() else v.contiguous()
    return _ToCL.apply(x.float(), cpad)


def from_cl(x_cl: torch.Tensor) -> torch.Tensor:
    return x_cl.permute(0, 4, 1, 2, 3)


def grad_cl(g: torch.Tensor) -> torch.Tensor:
    """Incoming gradient for a CL-shaped output -> contiguous bf16."""
    if g.dtype != act_dtype():
        g = g.to(act_dtype())
    return g if g.is_contiguous() else g.contiguous()


# ---- conv layer state --------------------------------------------------------------------------------
class ConvLayer:
    """Plans + packed bf16 weights of one convolution parameter.  The fp32 nn.Parameter in the reference
    layout stays the parameter of record; packed tiles are derived caches, rebuilt when it changes."""

    def __init__(self, weight: torch.nn.Parameter, spec_fn: Callable[[Tuple[int, int, int]], ConvSpec]):
        self.weight = weight
        self.spec_fn = spec_fn
        self.plans: Dict[Tuple[int, int, int], ConvPlan] = {}
        self.keys: Dict[Tuple, Tuple] = {}

    def plan(self, in_dims) -> ConvPlan:
        in_dims = tuple(int(v) for v in in_dims)
        pl = self.plans.get(in_dims)
        if pl is None:
            pl = ConvPlan(self.spec_fn(in_dims), in_dims)
            self.plans[in_dims] = pl
        return pl.to(self.weight.device)

    def packed(self, in_dims, which: str) -> ConvPlan:
        in_dims = tuple(int(v) for v in in_dims)
        pl = self.plan(in_dims)
        w = self.weight
        key = _packed_key(w)
        if not _fresh(self.keys.get((in_dims, which)), key):
            pl.pack(w.detach(), which, ops.stream())
        self.keys[(in_dims, which)] = key
        return pl


def _rows_major_spec(sp: ConvSpec) -> ConvSpec:
    """The 2-D layer `sp` (kt = 1) with its image rows on the T axis, its columns on the H axis and nothing along W."""
    assert sp.k[0] == 1
    return ConvSpec(sp.Cin, sp.Cout, (sp.k[1], sp.k[2], 1), (sp.stride[1], sp.stride[2], 1), (sp.pad_front[1], sp.pad_front[2], 0),
                    (sp.pad_back[1], sp.pad_back[2], 0), (sp.out_pad[1], sp.out_pad[2], 0), sp.transposed, Cin_pad=sp.Cin_pad,
                    Cout_pad=sp.Cout_pad)


def _conv_rows_major_fprop(self, in_dims, N: int) -> ConvPlan:
    """fprop plan of a 2-D layer on (row, column, clip) positions (see FusedConvLayer.rows_major_dgrad): for the transposed
    9 x 9 convolution `upsample1` (20 x 20 -> 28 x 28) half of the (tap, position) pairs are padding, which the tiles of
    this order skip.  Shares the packed operand of the clip-major plan."""
    in_dims = tuple(int(v) for v in in_dims)
    base = self.packed(in_dims, "fprop")
    key = ("rows", in_dims, int(N))
    pl = self.plans.get(key)
    if pl is None:
        assert in_dims[0] == 1
        pl = ConvPlan(_rows_major_spec(base.spec), (in_dims[1], in_dims[2], int(N)))
        pl.rows_major = True
        assert [c.wtap for c in pl.fprop] == [c.wtap for c in base.fprop], "rows-major plan must share the packed operand"
        self.plans[key] = pl
    pl = pl.to(self.weight.device)
    for a, b in zip(pl.fprop, base.fprop):
        a.packed = b.packed
    return pl


ConvLayer.rows_major_fprop = _conv_rows_major_fprop


def _bn_fold(gamma, beta, rm, rv):
    """Eval-mode BatchNorm as a per-channel affine: y = conv(x) * s + b (pytorch_i3d.py:80,117 with running statistics)."""
    s = gamma.detach().float() * torch.rsqrt(rv.detach().float() + BN_EPS)
    return s, (beta.detach().float() - rm.detach().float() * s).contiguous()


def _conv_packed_eval(self, in_dims, gamma, beta, rm, rv):
    """Eval mode: BatchNorm folded into the convolution -- the per-channel scale goes into the packed weights, the shift
    becomes the epilogue bias, ReLU is applied in the epilogue: one kernel per Unit3D, no pre-activation tensor."""
    in_dims = tuple(int(v) for v in in_dims)
    if not hasattr(self, "eval_plans"):
        self.eval_plans, self.eval_keys, self.eval_bias = {}, {}, {}
    pl = self.eval_plans.get(in_dims)
    if pl is None:
        pl = ConvPlan(self.spec_fn(in_dims), in_dims).to(self.weight.device)
        self.eval_plans[in_dims] = pl
    key = _packed_key(self.weight, gamma, beta, rm, rv)
    if not _fresh(self.eval_keys.get(in_dims), key):
        s, b = _bn_fold(gamma, beta, rm, rv)
        wf = (self.weight.detach() * s.view(-1, 1, 1, 1, 1)).contiguous()
        pl.pack(wf, "fprop", ops.stream())
        self.eval_bias[in_dims] = b
    self.eval_keys[in_dims] = key
    return pl, self.eval_bias[in_dims]


ConvLayer.packed_eval = _conv_packed_eval


class StemLayer(ConvLayer):
    """Few-input-channel convolution (the RGB 7x7x7 stem, pytorch_i3d.py:224).

    "fold" (default): the time axis is folded into the channel axis -- xs[n][h][w][fp*4+c] -- so the layer is a 2-D
    convolution over 64 channels on the TMA im2col path, one output class (own shifted weight set) per output frame;
    nothing but the 205 MB folded input is materialised (include/b200caps.h, "Folded stem").
    "im2col" (B2C_STEM=im2col, or clips the fold does not cover): explicit im2col (bandwidth kernel) + a 1x1x1 GEMM over
    K = taps*Cin columns -- a 3.5 GB matrix at 16+16 clips, written once and read twice."""
    KF = 64            # folded channels: 16 padded frames x 4 channel slots

    def __init__(self, weight, cin, cout, k, stride):
        self.weight, self.cin, self.cout, self.k, self.stride = weight, cin, cout, tuple(k), tuple(stride)
        self.taps = k[0] * k[1] * k[2]
        self.K = self.taps * cin
        self.Kpad = (self.K + 63) // 64 * 64
        self.plans, self.keys = {}, {}
        self._wtap = None
        self._w2 = None
        self.fold_enabled = os.environ.get("B2C_STEM", "fold") != "im2col"

    def geometry(self, in_dims):
        pads = [same_pad(d, kk, ss) for d, kk, ss in zip(in_dims, self.k, self.stride)]
        od = tuple((d + p[0] + p[1] - kk) // ss + 1 for d, p, kk, ss in zip(in_dims, pads, self.k, self.stride))
        return od, tuple(p[0] for p in pads)

    # ---- folded path -------------------------------------------------------------------------------------
    def use_fold(self, in_dims) -> bool:
        od, pf = self.geometry(in_dims)
        last = self.stride[0] * (od[0] - 1) + self.k[0]          # padded frames touched
        return (self.fold_enabled and self.cin <= 4 and od[0] * self.cout <= 256 and self.cout % 64 == 0 and last <= 16
                and pf[0] + in_dims[0] <= 16 and (od[1] * od[2]) % 64 == 0)

    def fold_input(self, x: View) -> View:
        od, pf = self.geometry(x.dims)
        T, H, W = x.dims
        xs = torch.empty((x.N, 1, H, W, self.KF), dtype=act_dtype(), device=x.t.device)
        ops.stem_fold_input(x, xs, pf[0], self.KF // 4)
        return View(xs)

    def fold_plan(self, in_dims) -> ConvPlan:
        in_dims = tuple(int(v) for v in in_dims)
        key = ("fold", in_dims)
        pl = self.plans.get(key)
        if pl is None:
            od, pf = self.geometry(in_dims)
            T, H, W = in_dims
            To = od[0]
            pads = [same_pad(d, kk, ss) for d, kk, ss in zip(in_dims, self.k, self.stride)]
            # one GEMM row per output pixel (n, h, w); its To * Cout columns are the output frames (out_fold = Cout)
            spec = ConvSpec(self.KF, To * self.cout, (1, self.k[1], self.k[2]), (1, self.stride[1], self.stride[2]),
                            (0, pads[1][0], pads[2][0]), (0, pads[1][1], pads[2][1]))
            pl = ConvPlan(spec, (1, H, W))
            b = pl.fprop[0]
            b.wtap = [i * self.KF for i in range(len(b.taps))]        # tap offset inside W2 / dW2: (To * Cout, khw, KF)
            pl.dgrad = []                                             # the stem input needs no gradient
            pl.out_dims = tuple(od)
            khw = self.k[1] * self.k[2]
            pl.fprop_pack = dict(R=To * self.cout, R_pad=To * self.cout, C=self.KF, C_real=self.KF, s_r=khw * self.KF, s_c=1,
                                 out_fold=self.cout)
            pl.wgrad_cls = b
            pl.wgrad_geom = dict(pl.wgrad_geom, s_p=khw * self.KF, s_g=1, Q=(1, od[1], od[2]), p_fold=self.cout)
            self.plans[key] = pl
        return pl.to(self.weight.device)

    def _refresh_w2(self, od0: int):
        khw = self.k[1] * self.k[2]
        ops.stem_fold_weights(self.weight.detach(), self._w2, self.cout, self.cin, self.k[0], khw, self.stride[0], od0, self.KF)

    def packed_fold(self, in_dims) -> ConvPlan:
        in_dims = tuple(int(v) for v in in_dims)
        pl = self.fold_plan(in_dims)
        w = self.weight
        key = _packed_key(w)
        if not _fresh(self.keys.get(("fold", in_dims)), key):
            from .plans import packed_geometry
            To = pl.out_dims[0]
            khw = self.k[1] * self.k[2]
            if self._w2 is None or self._w2.shape[0] != To:
                self._w2 = torch.zeros((To, self.cout, khw, self.KF), dtype=torch.float32, device=w.device)
            self._refresh_w2(To)
            if ops.PACKS is not None:
                ops.PACKS.add_pre(("stem", id(self)), lambda To=To: self._refresh_w2(To))
            bn, _, nkb, elems = packed_geometry(To * self.cout, khw * self.KF)
            cl = pl.fprop[0]
            if cl.packed is None or cl.packed.dtype != act_dtype():
                cl.packed = torch.zeros(elems, dtype=act_dtype(), device=w.device)
            ops.pack_part(self._w2, cl.packed, cl.wtap_dev, To * self.cout, khw, self.KF, self.KF, khw * self.KF, 1, self.KF,
                          0, 0, bn, nkb)
        self.keys[("fold", in_dims)] = key
        return pl

    def packed_eval(self, in_dims, gamma, beta, rm, rv):
        """Eval mode, folded path: BatchNorm scale folded into the (time-folded) weights, shift as the epilogue bias."""
        in_dims = tuple(int(v) for v in in_dims)
        if not hasattr(self, "eval_plans"):
            self.eval_plans, self.eval_keys, self.eval_bias = {}, {}, {}
        pl = self.eval_plans.get(in_dims)
        if pl is None:
            train_plan = self.plans.pop(("fold", in_dims), None)
            pl = self.fold_plan(in_dims)                     # an independent plan object (own packed operand) for eval
            self.plans.pop(("fold", in_dims))
            if train_plan is not None:
                self.plans[("fold", in_dims)] = train_plan
            self.eval_plans[in_dims] = pl
        key = _packed_key(self.weight, gamma, beta, rm, rv)
        if not _fresh(self.eval_keys.get(in_dims), key):
            from .plans import packed_geometry
            s, b = _bn_fold(gamma, beta, rm, rv)
            To = pl.out_dims[0]
            khw = self.k[1] * self.k[2]
            wf = (self.weight.detach() * s.view(-1, 1, 1, 1, 1)).contiguous()
            w2 = torch.zeros((To, self.cout, khw, self.KF), dtype=torch.float32, device=wf.device)
            ops.stem_fold_weights(wf, w2, self.cout, self.cin, self.k[0], khw, self.stride[0], To, self.KF)
            bn, _, nkb, elems = packed_geometry(To * self.cout, khw * self.KF)
            cl = pl.fprop[0]
            if cl.packed is None or cl.packed.dtype != act_dtype():
                cl.packed = torch.zeros(elems, dtype=act_dtype(), device=wf.device)
            ops.pack_part(w2, cl.packed, cl.wtap_dev, To * self.cout, khw, self.KF, self.KF, khw * self.KF, 1, self.KF, 0, 0, bn, nkb)
            self.eval_bias[in_dims] = b.repeat(To).contiguous()
        self.eval_keys[in_dims] = key
        return pl, self.eval_bias[in_dims]

    def fold_wgrad(self, in_dims, xs: View, dy: View, dw: torch.Tensor):
        """dw (Cout, Cin, kt, kh, kw) += wgrad: ONE launch with N = To * Cout columns (dy's frames are the p-channel blocks)
        into dW2, then the adjoint of the weight fold."""
        pl = self.fold_plan(in_dims)
        To = pl.out_dims[0]
        khw = self.k[1] * self.k[2]
        dw2 = torch.zeros((To, self.cout, khw, self.KF), dtype=torch.float32, device=dw.device)
        ops.conv_wgrad(pl, xs, dy, dw2, atomic=True)
        ops.stem_unfold_wgrad(dw2, dw, self.cout, self.cin, self.k[0], khw, self.stride[0], To, self.KF)

    # ---- im2col path -------------------------------------------------------------------------------------
    def im2col(self, x: View) -> torch.Tensor:
        od, pf = self.geometry(x.dims)
        col = torch.empty((x.N,) + od + (self.Kpad,), dtype=act_dtype(), device=x.t.device)
        ops.im2col_small(x, col, self.cin, od, self.k, self.stride, pf, self.Kpad)
        return col

    def plan(self, col_dims) -> ConvPlan:
        col_dims = tuple(int(v) for v in col_dims)
        pl = self.plans.get(col_dims)
        if pl is None:
            pl = ConvPlan(ConvSpec(self.Kpad, self.cout, (1, 1, 1)), col_dims)
            pl.wgrad_geom.update(s_p=self.Kpad, s_g=1, Cg_real=self.K)    # wgrad lands in a (Cout, Kpad) scratch
            self.plans[col_dims] = pl
        return pl.to(self.weight.device)

    def packed(self, col_dims, which: str) -> ConvPlan:
        assert which == "fprop", "the stem input needs no gradient"
        col_dims = tuple(int(v) for v in col_dims)
        pl = self.plan(col_dims)
        w = self.weight
        key = _packed_key(w)
        if not _fresh(self.keys.get(col_dims), key):
            from .plans import packed_geometry
            cl = pl.fprop[0]
            bn, _, nkb, elems = packed_geometry(self.cout, self.Kpad)
            if cl.packed is None or cl.packed.dtype != act_dtype():
                cl.packed = torch.zeros(elems, dtype=act_dtype(), device=w.device)
                self._wtap = torch.arange(self.taps, dtype=torch.int32, device=w.device)
            # k = tap*cin + c  <-  w[co][c][tap]
            ops.pack_part(w.detach(), cl.packed, self._wtap, self.cout, self.taps, self.cin, self.cin, self.cin * self.taps,
                          self.taps, self.cin, 0, 0, bn, nkb)
        self.keys[col_dims] = key
        return pl

    def scatter_wgrad(self, scratch: torch.Tensor, dw: torch.Tensor):
        """(Cout, Kpad) scratch with column tap*cin + c  ->  += into dw (Cout, cin, kt, kh, kw)."""
        g = scratch[:, :self.K].view(self.cout, self.taps, self.cin).permute(0, 2, 1).reshape(dw.shape)
        dw.add_(g)


# ---- primitives (no autograd) -----------------------------------------------------------------------
class UnitSaved:
    __slots__ = ("raw", "mean", "rstd", "groups", "dims", "col")


def unit_fwd(layer: ConvLayer, gamma, beta, rm, rv, x: View, y: View, training: bool, groups: int,
             prefolded=None) -> UnitSaved:
    """Unit3D (pytorch_i3d.py:89-120): same-pad conv (tcgen05) -> BatchNorm3d -> ReLU, written into `y`.
    prefolded = (T, H, W): x is already the stem's folded input (ops.clips_to_folded) of clips with these dims."""
    col = None
    folded = prefolded is not None or (isinstance(layer, StemLayer) and layer.use_fold(x.dims))
    orig_dims = tuple(prefolded) if prefolded is not None else x.dims
    if not training and EVAL_FOLD_BN and (folded or not isinstance(layer, StemLayer)) and ops.PACKS is None:
        # inference: conv + folded BatchNorm + ReLU in ONE kernel, written straight into the concat slot
        pl, bias = layer.packed_eval(orig_dims, gamma, beta, rm, rv)
        if folded and prefolded is None:
            x = layer.fold_input(x)
        ops.conv_fprop(pl, "fprop", x, y, bias=bias, relu=True)
        sv = UnitSaved()
        sv.raw, sv.dims, sv.col, sv.mean, sv.rstd, sv.groups = None, orig_dims, None, None, None, 1
        return sv
    if folded:
        if prefolded is None:
            x = layer.fold_input(x)
        col = x.t
        pl = layer.packed_fold(orig_dims)
    else:
        if isinstance(layer, StemLayer):
            col = layer.im2col(x)
            x = View(col)
        pl = layer.packed(x.dims, "fprop")
    N = x.N
    Cout = layer.cout if folded else pl.spec.Cout_pad
    raw = torch.empty((N,) + tuple(pl.out_dims) + (Cout,), dtype=act_dtype(), device=x.t.device)
    ops.conv_fprop(pl, "fprop", x, View(raw))
    sv = UnitSaved()
    sv.raw, sv.dims, sv.col = raw, (orig_dims if folded else x.dims), col
    rv_ = View(raw)
    if training:
        g = groups
        assert N % g == 0
        ws = zeros_f32((g, 2, Cout), raw.device)
        sv.mean = torch.empty((g, Cout), dtype=torch.float32, device=raw.device)
        sv.rstd = torch.empty((g, Cout), dtype=torch.float32, device=raw.device)
        sv.groups = g
        ops.bn_relu_fwd(rv_, g, ws, sv.mean, sv.rstd, rm, rv, BN_MOMENTUM, BN_EPS, gamma.detach(), beta.detach(), y, relu=True)
        return sv
    sv.mean = rm.detach().float().view(1, -1).contiguous()
    sv.rstd = torch.rsqrt(rv.detach().float() + BN_EPS).view(1, -1).contiguous()
    sv.groups = 1
    ops.bn_relu_apply(rv_, sv.groups, sv.mean, sv.rstd, gamma.detach(), beta.detach(), y, relu=True)
    return sv


def unit_bwd(layer: ConvLayer, gamma, beta, sv: UnitSaved, x: View, y: View, gy: View, dx: Optional[View],
             accumulate: bool):
    """Returns (dweight, dgamma, dbeta) (None where accumulated directly into .grad); writes dx (+= when accumulate)."""
    dev = x.t.device
    Cout = sv.raw.shape[-1]
    raw = View(sv.raw)
    ws = zeros_f32((sv.groups, 2, Cout), dev)
    draw = torch.empty_like(sv.raw)
    dgamma, d1 = grad_buf(gamma)
    dbeta, d2 = grad_buf(beta)
    ops.bn_relu_bwd(gy, None if BN_REMASK else y, raw, sv.groups, sv.mean, sv.rstd, gamma.detach(), beta.detach(), ws, View(draw), dgamma,
                    dbeta, relu=True)
    dw, d0 = grad_buf(layer.weight)
    if isinstance(layer, StemLayer):
        assert dx is None, "the few-channel stem does not propagate a gradient to its input"
        if sv.col.shape[-1] == layer.KF and sv.col.shape[1] == 1 and layer.use_fold(sv.dims):
            layer.fold_wgrad(sv.dims, View(sv.col), View(draw), dw)
            return (None if d0 else dw), (None if d1 else dgamma), (None if d2 else dbeta)
        scratch = torch.zeros((layer.cout, layer.Kpad), dtype=torch.float32, device=dev)
        ops.conv_wgrad(layer.plan(sv.dims), View(sv.col), View(draw), scratch, atomic=True)
        layer.scatter_wgrad(scratch, dw)
        return (None if d0 else dw), (None if d1 else dgamma), (None if d2 else dbeta)
    if dx is not None:
        pl = layer.packed(sv.dims, "dgrad")
        ops.conv_fprop(pl, "dgrad", View(draw), dx, accumulate=accumulate)
    pl = layer.plan(sv.dims)
    ops.conv_wgrad(pl, x, View(draw), dw, atomic=True)
    return (None if d0 else dw), (None if d1 else dgamma), (None if d2 else dbeta)


def cba_fwd(layer: ConvLayer, bias, x: View, y: View, relu: bool, scale_nc=None, sigmoid_from=-1):
    """conv / transposed conv + bias (+ per-sample channel scale) (+ ReLU) fused in the GEMM epilogue."""
    pl = layer.packed(x.dims, "fprop")
    ops.conv_fprop(pl, "fprop", x, y, bias=bias.detach() if bias is not None else None, scale_nc=scale_nc, relu=relu,
                   sigmoid_from=sigmoid_from)


def cba_bwd(layer: ConvLayer, bias, x: View, y: Optional[View], gy: View, relu: bool, scale_nc, dx: Optional[View],
            accumulate: bool = False):
    """Returns (dweight, dbias) (None where accumulated directly into .grad)."""
    dev = x.t.device
    C = gy.C
    dbias, d1 = grad_buf(bias)
    if relu or scale_nc is not None:
        dz_t = torch.empty((gy.N,) + tuple(gy.dims) + (C,), dtype=act_dtype(), device=dev)
        dz = View(dz_t)
        ops.act_bwd(gy, y if relu else None, scale_nc, dz, dbias, relu)
    else:
        dz = gy
        ops.act_bwd(gy, None, None, None, dbias, False)
    if dx is not None:
        pl = layer.packed(x.dims, "dgrad")
        ops.conv_fprop(pl, "dgrad", dz, dx, accumulate=accumulate)
    pl = layer.plan(x.dims)
    dw, d0 = grad_buf(layer.weight)
    ops.conv_wgrad(pl, x, dz, dw, atomic=True)
    return (None if d0 else dw), (None if d1 else dbias)


def dropout_scale(n: int, c: int, device, p: float = 0.5) -> torch.Tensor:
    """Dropout3d keep mask scaled by 1/(1-p), one value per (sample, channel) (capsules_ucf101.py:371).
    RNG draw = torch's bernoulli on the device (plumbing); tests inject masks through STATE.dropout_source."""
    if STATE.dropout_source is not None:
        m = STATE.dropout_source(n, c, device)
        return m.to(device=device, dtype=torch.float32).reshape(n, c).contiguous()
    return torch.empty((n, c), dtype=torch.float32, device=device).bernoulli_(1.0 - p).div_(1.0 - p)


# ---- autograd functions -------------------------------------------------------------------------------
class Unit3DFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_cl, weight, gamma, beta, mod):
        layer = mod._layer
        prefolded = getattr(ctx, "prefolded", None)       # fused step: the input is already in the stem's folded layout
        if isinstance(layer, StemLayer):
            od, cpad = layer.geometry(tuple(prefolded or x_cl.shape[1:4]))[0], layer.cout
        else:
            pl = layer.plan(x_cl.shape[1:4])
            od, cpad = tuple(pl.out_dims), pl.spec.Cout_pad
        y = torch.empty((x_cl.shape[0],) + od + (cpad,), dtype=act_dtype(), device=x_cl.device)
        training = mod.training
        sv = unit_fwd(layer, gamma, beta, mod.bn.running_mean, mod.bn.running_var, View(x_cl), View(y), training,
                      STATE.bn_groups if training else 1, prefolded=prefolded)
        if training and not STATE.defer_bn_counters:
            mod.bn.num_batches_tracked += STATE.bn_groups      # one per forward pass held in the batch
        ctx.mod, ctx.sv, ctx.x, ctx.y, ctx.gamma = mod, sv, x_cl, y, gamma
        ctx.training = training
        return y

    @staticmethod
    def backward(ctx, gy):
        if not ctx.training:
            raise RuntimeError("b200caps: backward through eval-mode BatchNorm is not supported")
        gy = grad_cl(gy)
        need_dx = ctx.needs_input_grad[0] and not isinstance(ctx.mod._layer, StemLayer)
        dx = torch.empty_like(ctx.x) if need_dx else None
        dw, dg, db = unit_bwd(ctx.mod._layer, ctx.gamma, ctx.mod.bn.bias, ctx.sv, View(ctx.x), View(ctx.y), View(gy),
                              View(dx) if need_dx else None, False)
        return dx, dw, dg, db, None


class MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_cl, k, s):
        N, T, H, W, C = x_cl.shape
        pads = [same_pad(d, kk, ss) for d, kk, ss in zip((T, H, W), k, s)]
        od = tuple((d + p[0] + p[1] - kk) // ss + 1 for d, p, kk, ss in zip((T, H, W), pads, k, s))
        y = torch.empty((N,) + od + (C,), dtype=act_dtype(), device=x_cl.device)
        idx = torch.empty((N,) + od + (C,), dtype=torch.uint8, device=x_cl.device)
        pf = tuple(p[0] for p in pads)
        ops.maxpool_fwd(View(x_cl), View(y), idx, k, s, pf)
        ctx.geo = (tuple(k), tuple(s), pf, tuple(x_cl.shape))
        ctx.idx = idx
        return y

    @staticmethod
    def backward(ctx, gy):
        k, s, pf, xshape = ctx.geo
        gy = grad_cl(gy)
        dx = torch.empty(xshape, dtype=act_dtype(), device=gy.device)
        ops.maxpool_bwd(View(gy), ctx.idx, View(dx), k, s, pf, accumulate=False)
        return dx, None, None


def bn_fwd_part(raw: View, m, y: View, g: int):
    """Training-mode BatchNorm3d + ReLU of one channel window of a (possibly fused) convolution output -> y.
    Returns (mean, rstd) of the window (per statistics group)."""
    C = raw.C
    ws = zeros_f32((g, 2, C), raw.t.device)
    mean = torch.empty((g, C), dtype=torch.float32, device=raw.t.device)
    rstd = torch.empty((g, C), dtype=torch.float32, device=raw.t.device)
    ops.bn_relu_fwd(raw, g, ws, mean, rstd, m.bn.running_mean, m.bn.running_var, BN_MOMENTUM, BN_EPS, m.bn.weight.detach(),
                          m.bn.bias.detach(), y, relu=True)
    return mean, rstd


def bn_bwd_part(raw: View, y: View, gy: View, draw: View, m, mean, rstd, g: int):
    """Backward of bn_fwd_part: writes d(raw) into `draw`; returns (dgamma, dbeta) (None where accumulated into .grad)."""
    ws = zeros_f32((g, 2, raw.C), raw.t.device)
    dgamma, d1 = grad_buf(m.bn.weight)
    dbeta, d2 = grad_buf(m.bn.bias)
    ops.bn_relu_bwd(gy, None if BN_REMASK else y, raw, g, mean, rstd, m.bn.weight.detach(), m.bn.bias.detach(), ws, draw, dgamma, dbeta,
                    relu=True)
    return (None if d1 else dgamma), (None if d2 else dbeta)


# Inception blocks: the three 1x1x1 convolutions that read the block input (b0, b1a, b2a) run as ONE implicit GEMM with
# N = c0 + c1 + c3 output channels (176 .. 448), their three input gradients as ONE dgrad GEMM over the stacked output
# gradients: 2 launches instead of 6 per block and one read of x instead of three.  B2C_FUSE_SIBLINGS=0 disables.
FUSE_SIBLINGS = os.environ.get("B2C_FUSE_SIBLINGS", "1") != "0"
# PrimaryCaps forward: K (the 81 taps) split over this many scheduling classes (B2C_PC_KSPLIT=1: one class, fused epilogue)
PC_KSPLIT = int(os.environ.get("B2C_PC_KSPLIT", "8"))
# PrimaryCaps dgrad on (row, clip, column) positions with per-tile skipping of padding-only tap rows (B2C_PC_DGRAD_ROWS=0: clip-major)
PC_DGRAD_ROWS = os.environ.get("B2C_PC_DGRAD_ROWS", "1") != "0"
# upsample1 forward on (row, column, clip) positions with padding-tap skipping.  Off by default: bit-identical and measured
# neutral on the step (19.73 vs 19.74 ms) -- the layer's N = 64 tiles are bound by the tensor core's own shared-memory reads,
# and the two row permutations cost what the skipped taps save.  B2C_UP1_ROWS=1 enables it.
UP1_ROWS = os.environ.get("B2C_UP1_ROWS", "0") == "1"
# BatchNorm backward recomputes the ReLU mask from the raw convolution output instead of reading y (B2C_BN_REMASK=0: read y)
BN_REMASK = os.environ.get("B2C_BN_REMASK", "1") != "0"
# weight gradients of a fused layer's members in one wgrad launch (B2C_FUSED_WGRAD=0: one launch per member)
FUSED_WGRAD = os.environ.get("B2C_FUSED_WGRAD", "1") != "0"


def _sibling_layer(mod) -> "FusedConvLayer":
    fl = mod.__dict__.get("_fused_siblings")
    ws = [mod.b0.conv3d.weight, mod.b1a.conv3d.weight, mod.b2a.conv3d.weight]
    if fl is None or any(a is not b for a, b in zip(fl.weights, ws)):
        cin = int(ws[0].shape[1])
        cout = sum(int(w.shape[0]) for w in ws)
        fl = FusedConvLayer(ws, lambda d, cin=cin, cout=cout: ConvSpec(cin, cout, (1, 1, 1)))
        mod.__dict__["_fused_siblings"] = fl
    return fl


class InceptionFn(torch.autograd.Function):
    """InceptionModule (pytorch_i3d.py:124-149) hand-scheduled: six Unit3D + same-pad max-pool, every branch
    writes straight into its channel slot of the concatenated output (no torch.cat), gradients of the four
    consumers of x are accumulated in the dgrad epilogues."""
    UNITS = ("b0", "b1a", "b1b", "b2a", "b2b", "b3b")

    @staticmethod
    def forward(ctx, x_cl, mod, *params):
        N, T, H, W, Cin = x_cl.shape
        dev = x_cl.device
        u = {n: getattr(mod, n) for n in InceptionFn.UNITS}
        oc = [m.bn.weight.numel() for m in (u["b0"], u["b1a"], u["b1b"], u["b2a"], u["b2b"], u["b3b"])]
        c0, c1, c2, c3, c4, c5 = oc
        training = mod.training
        g = STATE.bn_groups if training else 1
        out = torch.empty((N, T, H, W, c0 + c2 + c4 + c5), dtype=act_dtype(), device=dev)
        mid1 = torch.empty((N, T, H, W, c1), dtype=act_dtype(), device=dev)
        mid2 = torch.empty((N, T, H, W, c3), dtype=act_dtype(), device=dev)
        pooled = torch.empty_like(x_cl)
        idx = torch.empty(x_cl.shape, dtype=torch.uint8, device=dev)
        xv = View(x_cl)

        def run(name, xin, yout):
            m = u[name]
            return unit_fwd(m._layer, m.bn.weight, m.bn.bias, m.bn.running_mean, m.bn.running_var, xin, yout, training, g)

        sv = {}
        fused = training and FUSE_SIBLINGS and c0 % 8 == 0 and c1 % 8 == 0 and c3 % 8 == 0
        raw3 = None
        if fused:
            fl = _sibling_layer(mod)
            raw3 = torch.empty((N, T, H, W, c0 + c1 + c3), dtype=act_dtype(), device=dev)
            ops.conv_fprop(fl.packed((T, H, W), "fprop"), "fprop", xv, View(raw3))
            sv["b0"] = bn_fwd_part(View(raw3, 0, c0), u["b0"], View(out, 0, c0), g)
            sv["b1a"] = bn_fwd_part(View(raw3, c0, c1), u["b1a"], View(mid1), g)
            sv["b2a"] = bn_fwd_part(View(raw3, c0 + c1, c3), u["b2a"], View(mid2), g)
        else:
            sv["b0"] = run("b0", xv, View(out, 0, c0))
            sv["b1a"] = run("b1a", xv, View(mid1))
            sv["b2a"] = run("b2a", xv, View(mid2))
        sv["b1b"] = run("b1b", View(mid1), View(out, c0, c2))
        sv["b2b"] = run("b2b", View(mid2), View(out, c0 + c2, c4))
        ops.maxpool_fwd(xv, View(pooled), idx, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        sv["b3b"] = run("b3b", View(pooled), View(out, c0 + c2 + c4, c5))
        if training and not STATE.defer_bn_counters:
            for m in u.values():
                m.bn.num_batches_tracked += STATE.bn_groups
        ctx.mod, ctx.sv, ctx.oc = mod, sv, oc
        ctx.x, ctx.out, ctx.mid1, ctx.mid2, ctx.pooled, ctx.idx = x_cl, out, mid1, mid2, pooled, idx
        ctx.training, ctx.raw3, ctx.groups = training, raw3, g
        return out

    @staticmethod
    def backward(ctx, gout):
        if not ctx.training:
            raise RuntimeError("b200caps: backward through eval-mode BatchNorm is not supported")
        gout = grad_cl(gout)
        mod, sv = ctx.mod, ctx.sv
        c0, c1, c2, c3, c4, c5 = ctx.oc
        u = {n: getattr(mod, n) for n in InceptionFn.UNITS}
        x, out = ctx.x, ctx.out
        xv = View(x)
        dx = torch.empty_like(x)
        dmid1 = torch.empty_like(ctx.mid1)
        dmid2 = torch.empty_like(ctx.mid2)
        dpool = torch.empty_like(x)
        grads = {}

        def run(name, xin, yv, gyv, dxv, acc):
            m = u[name]
            grads[name] = unit_bwd(m._layer, m.bn.weight, m.bn.bias, sv[name], xin, yv, gyv, dxv, acc)

        if ctx.raw3 is not None:
            run("b1b", View(ctx.mid1), View(out, c0, c2), View(gout, c0, c2), View(dmid1), False)
            run("b2b", View(ctx.mid2), View(out, c0 + c2, c4), View(gout, c0 + c2, c4), View(dmid2), False)
            raw3, gq = ctx.raw3, ctx.groups
            draw3 = torch.empty_like(raw3)
            bn = {}
            bn["b0"] = bn_bwd_part(View(raw3, 0, c0), View(out, 0, c0), View(gout, 0, c0), View(draw3, 0, c0), u["b0"], *sv["b0"], gq)
            bn["b1a"] = bn_bwd_part(View(raw3, c0, c1), View(ctx.mid1), View(dmid1), View(draw3, c0, c1), u["b1a"], *sv["b1a"], gq)
            bn["b2a"] = bn_bwd_part(View(raw3, c0 + c1, c3), View(ctx.mid2), View(dmid2), View(draw3, c0 + c1, c3), u["b2a"],
                                    *sv["b2a"], gq)
            fl = _sibling_layer(mod)
            dims = tuple(x.shape[1:4])
            ops.conv_fprop(fl.packed(dims, "dgrad"), "dgrad", View(draw3), View(dx))
            dws = fl.wgrad(dims, xv, View(draw3))
            for name, dwi in zip(("b0", "b1a", "b2a"), dws):
                grads[name] = (dwi,) + bn[name]
        else:
            run("b0", xv, View(out, 0, c0), View(gout, 0, c0), View(dx), False)
            run("b1b", View(ctx.mid1), View(out, c0, c2), View(gout, c0, c2), View(dmid1), False)
            run("b1a", xv, View(ctx.mid1), View(dmid1), View(dx), True)
            run("b2b", View(ctx.mid2), View(out, c0 + c2, c4), View(gout, c0 + c2, c4), View(dmid2), False)
            run("b2a", xv, View(ctx.mid2), View(dmid2), View(dx), True)
        run("b3b", View(ctx.pooled), View(out, c0 + c2 + c4, c5), View(gout, c0 + c2 + c4, c5), View(dpool), False)
        ops.maxpool_bwd(View(dpool), ctx.idx, View(dx), (3, 3, 3), (1, 1, 1), (1, 1, 1), accumulate=True)
        flat = []
        for n in InceptionFn.UNITS:
            flat += list(grads[n])
        return (dx, None) + tuple(flat)


class ChannelScaleFn(torch.autograd.Function):
    """Dropout3d as a per-(sample, channel) scale (capsules_ucf101.py:428)."""

    @staticmethod
    def forward(ctx, x_cl, scale_nc):
        y = torch.empty_like(x_cl)
        ops.channel_scale(View(x_cl), scale_nc, View(y))
        ctx.scale = scale_nc
        return y

    @staticmethod
    def backward(ctx, gy):
        gy = grad_cl(gy)
        dx = torch.empty_like(gy)
        ops.channel_scale(View(gy), ctx.scale, View(dx))
        return dx, None


class FusedConvLayer:
    """Several convolutions that read the same input (same kernel / stride / padding) executed as ONE implicit
    GEMM with N = sum(Cout_i): the packed fprop operand stacks the members' rows, the dgrad operand places them
    side by side along K.  Used for PrimaryCaps (pose | a, N = 544)."""

    def __init__(self, weights: Sequence[torch.nn.Parameter], spec_fn, grad_cpad: int = 0, fprop_ksplit: int = 1):
        self.weights = list(weights)
        self.fprop_ksplit = fprop_ksplit   # > 1: the forward GEMM's K (taps) is split over scheduling classes (one output frame each)
        self.grad_cpad = grad_cpad       # channel width of the output-gradient tensor (>= sum Cout, 64-aligned for TMA)
        self.couts = [int(w.shape[0]) for w in self.weights]
        self.offs = [sum(self.couts[:i]) for i in range(len(self.couts))]
        self.spec_fn = spec_fn
        self.plans: Dict[Tuple[int, int, int], ConvPlan] = {}
        self.keys: Dict[Tuple, Tuple] = {}

    def plan(self, in_dims) -> ConvPlan:
        in_dims = tuple(int(v) for v in in_dims)
        pl = self.plans.get(in_dims)
        if pl is None:
            pl = ConvPlan(self.spec_fn(in_dims), in_dims)
            assert not pl.spec.transposed and pl.spec.Cout == sum(self.couts)
            if self.fprop_ksplit > 1:
                pl.split_fprop_k(self.fprop_ksplit)
            if self.grad_cpad:
                # the gradient w.r.t. the fused output is stored grad_cpad channels wide (zero tail): dgrad's K and
                # wgrad's plain operand become 64-channel aligned -> TMA path
                pl.dgrad_pack = dict(pl.dgrad_pack, C=self.grad_cpad)
                pl.wgrad_geom = dict(pl.wgrad_geom, Cp=self.grad_cpad)
            self.plans[in_dims] = pl
        return pl.to(self.weights[0].device)

    def packed(self, in_dims, which: str) -> ConvPlan:
        pl = self.plan(in_dims)
        key = _packed_key(*self.weights)
        if _fresh(self.keys.get((tuple(in_dims), which)), key):
            self.keys[(tuple(in_dims), which)] = key
            return pl
        spec = pl.spec
        T = spec.k[0] * spec.k[1] * spec.k[2]
        dev = self.weights[0].device
        from .plans import packed_geometry
        from .plans import tap_pitch
        if which == "fprop":
            pitch = tap_pitch(spec.Cin_pad)
            for cl in pl.fprop:
                nt = len(cl.taps)
                bn, _, nkb, elems = packed_geometry(spec.Cout_pad, nt * pitch)
                if cl.packed is None or cl.packed.dtype != act_dtype():
                    cl.packed = torch.zeros(elems, dtype=act_dtype(), device=dev)
                for w, co, off in zip(self.weights, self.couts, self.offs):
                    ops.pack_part(w.detach(), cl.packed, cl.wtap_dev, co, nt, spec.Cin_pad, spec.Cin, spec.Cin * T, T,
                                  pitch, 0, off, bn, nkb)
        else:
            cg = self.grad_cpad or spec.Cout_pad
            pitch = tap_pitch(cg)
            for cl in pl.dgrad:
                nt = len(cl.taps)
                bn, _, nkb, elems = packed_geometry(spec.Cin_pad, nt * pitch)
                if cl.packed is None or cl.packed.dtype != act_dtype():
                    cl.packed = torch.zeros(elems, dtype=act_dtype(), device=dev)
                for w, co, off in zip(self.weights, self.couts, self.offs):
                    ops.pack_part(w.detach(), cl.packed, cl.wtap_dev, spec.Cin, nt, co, co, T, spec.Cin * T, pitch, off, 0, bn, nkb)
        self.keys[(tuple(in_dims), which)] = key
        return pl

    def rows_major_dgrad(self, in_dims, N: int) -> ConvPlan:
        """dgrad plan of a 2-D layer (T = 1) with the image rows on the kernel's T axis, the image columns on its H axis and
        the N clips on its W axis: GEMM positions then run (row, column, clip), a 128-position tile holds 128 / N columns of
        ONE image row of all clips, and the tiles skip the tap rows and tap columns that are padding for them
        (b2c_conv_class.h_block < 0).  PrimaryCaps' 9 x 9 'valid' convolution reads a 20 x 20 gradient from 28 x 28
        positions: 49 % of its (tap, position) pairs are such padding.  Same taps in the same order as the clip-major
        plan, so both share one packed operand."""
        in_dims = tuple(int(v) for v in in_dims)
        key = ("rows", in_dims, int(N))
        pl = self.plans.get(key)
        base = self.packed(in_dims, "dgrad")
        if pl is None:
            sp = base.spec
            assert in_dims[0] == 1 and sp.k[0] == 1 and not sp.transposed
            pl = ConvPlan(_rows_major_spec(sp), (in_dims[1], in_dims[2], int(N)))
            pl.rows_major = True
            if self.grad_cpad:
                pl.dgrad_pack = dict(pl.dgrad_pack, C=self.grad_cpad)
            assert [c.wtap for c in pl.dgrad] == [c.wtap for c in base.dgrad], "rows-major plan must share the packed operand"
            self.plans[key] = pl
        pl = pl.to(self.weights[0].device)
        for a, b in zip(pl.dgrad, base.dgrad):
            a.packed = b.packed
        return pl

    def wgrad(self, in_dims, x: View, dy: View) -> List[torch.Tensor]:
        pl = self.plan(in_dims)
        outs = []
        if FUSED_WGRAD and len(self.weights) <= 4:
            # ONE launch over all members: each output column block lands in its member's weight gradient
            # (b2c_wgrad_desc.seg_*).  The gathered operand is read once instead of once per member: PrimaryCaps'
            # 32-column activation member alone cost 0.41 ms next to 1.07 ms for the 512 pose columns.
            bufs = [grad_buf(w) for w in self.weights]
            cp = pl.wgrad_geom["Cp"]
            ops.conv_wgrad(pl, x, View(dy.t, dy.c_off, cp), bufs[0][0], atomic=True,
                           segs=[(off, b[0]) for off, b in zip(self.offs, bufs)])
            return [None if direct else dw for dw, direct in bufs]
        for w, co, off in zip(self.weights, self.couts, self.offs):
            dw, direct = grad_buf(w)
            width = co
            if co % 64 and self.grad_cpad and off % 64 == 0 and off + (co + 63) // 64 * 64 <= self.grad_cpad:
                width = (co + 63) // 64 * 64        # zero-padded window: keeps the plain operand on the TMA path
            ops.conv_wgrad(pl, x, dy, dw, atomic=True, part=(off, width, co))
            outs.append(None if direct else dw)
        return outs


class PrimaryCapsFn(torch.autograd.Function):
    """PrimaryCaps (capsules_ucf101.py:43-49): the pose (512) and activation (32) 9x9 convolutions run as ONE
    implicit GEMM with N = 544 whose epilogue adds the biases, applies the sigmoid to the last 32 columns and
    writes fp32 rows -- exactly the (B,20,20,544) permuted / concatenated layout the routing consumes."""

    @staticmethod
    def forward(ctx, x_cl, wp, bp, wa, ba, mod):
        layer: FusedConvLayer = mod._layer
        pl = layer.packed(x_cl.shape[1:4], "fprop")
        N = x_cl.shape[0]
        bias = torch.cat([bp.detach(), ba.detach()])      # 544 floats (plumbing)
        if len(pl.fprop) > 1:
            # K split (FusedConvLayer.fprop_ksplit): every slice writes its partial sums to its own frame; sum + bias + sigmoid
            part = torch.empty((N,) + tuple(pl.fprop_out_dims) + (544,), dtype=torch.float32, device=x_cl.device)
            out = torch.empty((N,) + tuple(pl.out_dims) + (544,), dtype=torch.float32, device=x_cl.device)
            ops.conv_fprop(pl, "fprop", View(x_cl), View(part), final=True)
            ops.primarycaps_finish(part, bias, out)
        else:
            out = torch.empty((N,) + tuple(pl.out_dims) + (544,), dtype=torch.float32, device=x_cl.device)
            ops.conv_fprop(pl, "fprop", View(x_cl), View(out), bias=bias, sigmoid_from=512, final=True)
        ctx.mod, ctx.x, ctx.out = mod, x_cl, out
        return out

    @staticmethod
    def backward(ctx, g):
        mod, x, out = ctx.mod, ctx.x, ctx.out
        layer: FusedConvLayer = mod._layer
        g = g.contiguous().float()
        rows = out.numel() // 544
        cg = layer.grad_cpad or 544
        alloc = torch.zeros if cg != 544 else torch.empty
        dzb = alloc(out.shape[:-1] + (cg,), dtype=act_dtype(), device=out.device)
        dbias = torch.zeros(544, dtype=torch.float32, device=out.device)
        dims = tuple(x.shape[1:4])
        N = x.shape[0]
        dx = None
        if ctx.needs_input_grad[0] and PC_DGRAD_ROWS and dims[0] == 1:
            # rows-major dgrad (FusedConvLayer.rows_major_dgrad): the prologue also writes dz with the clips innermost, the
            # GEMM runs on (row, column, clip) positions and skips padding-only tap rows / columns, the result is permuted back
            Hq, Wq = out.shape[2], out.shape[3]
            dzb_r = alloc((1, Hq, Wq, N, cg), dtype=act_dtype(), device=out.device)
            ops.primarycaps_bwd_prep2(g, out, dzb, dzb_r, dbias, N, Hq, Wq, cg)
            plr = layer.rows_major_dgrad(dims, N)
            dxr = torch.empty((1, dims[1], dims[2], N, x.shape[-1]), dtype=act_dtype(), device=x.device)
            ops.conv_fprop(plr, "dgrad", View(dzb_r), View(dxr))
            dx = torch.empty_like(x)
            ops.rows_to_clips(dxr, View(dx))
        else:
            ops.primarycaps_bwd_prep(g, out, dzb, dbias, rows, cg)
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x)
                ops.conv_fprop(layer.packed(dims, "dgrad"), "dgrad", View(dzb), View(dx))
        dwp, dwa = layer.wgrad(dims, View(x), View(dzb))
        dbp, dba = dbias[:512], dbias[512:]
        if STATE.direct_grads and mod.pose.bias.grad is not None:
            mod.pose.bias.grad.add_(dbp)
            mod.a.bias.grad.add_(dba)
            dbp = dba = None
        return dx, dwp, dbp, dwa, dba, None


class EMRoutingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, caps, W, beta_u, beta_a):
        """caps (N,h,w,544) fp32 -> (N,h,w,C*17) fp32 [mu | a]."""
        C = beta_a.numel()
        N, h, w, _ = caps.shape
        caps = caps.contiguous()
        out = torch.empty((N, h, w, C * 17), dtype=torch.float32, device=caps.device)
        Wc = W.detach().reshape(32, C, 4, 4).contiguous()
        # training: the forward saves its per-iteration state for the backward kernel (35 KB per location)
        need_bwd = ROUTING_SAVE_STATE and any(ctx.needs_input_grad)
        ctx.state = torch.empty((N * h * w, ops.routing_state_floats()), dtype=torch.float32, device=caps.device) if need_bwd else None
        ops.em_routing_fwd(caps, Wc, beta_u.detach().contiguous(), beta_a.detach().contiguous(), out, N * h * w, C, state=ctx.state)
        ctx.save_for_backward(caps, W, beta_u, beta_a)
        return out

    @staticmethod
    def backward(ctx, g):
        caps, W, beta_u, beta_a = ctx.saved_tensors
        C = beta_a.numel()
        N, h, w, _ = caps.shape
        g = g.contiguous().float()
        dcaps = torch.empty_like(caps)
        dW, d0 = grad_buf(W)
        dbu, d1 = grad_buf(beta_u)
        dba, d2 = grad_buf(beta_a)
        ops.em_routing_bwd(caps, W.detach().reshape(32, C, 4, 4).contiguous(), beta_u.detach().contiguous(),
                           beta_a.detach().contiguous(), g, dcaps, dW, dbu, dba, N * h * w, C, state=getattr(ctx, "state", None))
        ctx.state = None    # the state-based backward consumes the saved state; a second backward recomputes it
        return dcaps, (None if d0 else dW), (None if d1 else dbu), (None if d2 else dba)


class CapsHeadFn(torch.autograd.Function):
    """Class activation = spatial mean of a_out, feat = a_out, masked poses -> decoder input
    (capsules_ucf101.py:440-483).  mask (N,C) fp32 is a constant."""

    @staticmethod
    def forward(ctx, rout, mask):
        N, h, w, oc = rout.shape
        C = oc // 17
        L = h * w
        x0 = torch.empty((N, 1, h, w, C * 16), dtype=act_dtype(), device=rout.device)
        ops.pose_mask_fwd(rout, mask, x0, N, L, C)
        ctx.mask, ctx.shape = mask, (N, h, w, C)
        return x0

    @staticmethod
    def backward(ctx, gx):
        N, h, w, C = ctx.shape
        gx = grad_cl(gx)
        drout = torch.empty((N, h, w, C * 17), dtype=torch.float32, device=gx.device)
        ops.caps_head_bwd(gx, ctx.mask, None, None, drout, N, h * w, C)
        return drout, None


class ClassActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rout):
        N, h, w, oc = rout.shape
        C = oc // 17
        act = torch.empty((N, C), dtype=torch.float32, device=rout.device)
        ops.class_mean_fwd(rout, act, N, h * w, C)
        ctx.shape = (N, h, w, C)
        return act

    @staticmethod
    def backward(ctx, gact):
        N, h, w, C = ctx.shape
        drout = torch.empty((N, h, w, C * 17), dtype=torch.float32, device=gact.device)
        ops.caps_head_bwd(None, gact, gact.contiguous().float(), None, drout, N, h * w, C)
        return drout


class DecoderFn(torch.autograd.Function):
    """Localisation decoder (capsules_ucf101.py:486-510) hand-scheduled: transposed convs by output-parity
    class, skip convs, concatenations written in place, Dropout3d folded into the upsample4 epilogue,
    `smooth` = tensor-core projection onto its 27 taps + a 27-point stencil.  Output: fp32 logits."""
    ORDER = ("upsample1", "conv28", "upsample2", "conv56", "upsample3", "conv112", "upsample4", "smooth")

    @staticmethod
    def forward(ctx, x0, c28, c56, c112, drop_scale, mod, *params):
        dev = x0.device
        N = x0.shape[0]
        L = mod._layers
        bf = act_dtype()
        T1 = c56.shape[1]          # 2 for 8-frame clips
        H1 = c28.shape[2]          # 28
        cat28 = torch.empty((N, 1, H1, H1, 128), dtype=bf, device=dev)
        if UP1_ROWS and x0.shape[1] == 1 and L["upsample1"].plan(x0.shape[1:4]).spec.k[0] == 1:
            # upsample1 (transposed 9 x 9, 20 x 20 -> 28 x 28) on (row, column, clip) positions: tiles skip padding-only taps
            x0r = torch.empty((1, x0.shape[2], x0.shape[3], N, x0.shape[4]), dtype=bf, device=dev)
            ops.clips_to_rows(View(x0), x0r)
            plr = L["upsample1"].rows_major_fprop(x0.shape[1:4], N)
            y1r = torch.empty((1, H1, H1, N, 64), dtype=bf, device=dev)
            ops.conv_fprop(plr, "fprop", View(x0r), View(y1r), bias=mod.upsample1.bias.detach(), relu=True)
            ops.rows_to_clips(y1r, View(cat28, 0, 64))
        else:
            cba_fwd(L["upsample1"], mod.upsample1.bias, View(x0), View(cat28, 0, 64), relu=True)
        cba_fwd(L["conv28"], mod.conv28.bias, View(c28), View(cat28, 64, 64), relu=True)
        cat56 = torch.empty((N, T1, 2 * H1, 2 * H1, 128), dtype=bf, device=dev)
        cba_fwd(L["upsample2"], mod.upsample2.bias, View(cat28), View(cat56, 0, 64), relu=True)
        cba_fwd(L["conv56"], mod.conv56.bias, View(c56), View(cat56, 64, 64), relu=True)
        cat112 = torch.empty((N, 2 * T1, 4 * H1, 4 * H1, 128), dtype=bf, device=dev)
        cba_fwd(L["upsample3"], mod.upsample3.bias, View(cat56), View(cat112, 0, 64), relu=True)
        cba_fwd(L["conv112"], mod.conv112.bias, View(c112), View(cat112, 64, 64), relu=True)
        if TAIL_COLLAPSED:
            logits, tail_saved = L["tail"].forward(View(cat112), drop_scale)
            ctx.mod, ctx.drop, ctx.tail = mod, drop_scale, tail_saved
            ctx.t = (x0, c28, c56, c112, cat28, cat56, cat112, None)
            ctx.needs = ctx.needs_input_grad[:4]
            return logits
        To, Ho = 4 * T1, 8 * H1
        u4 = torch.empty((N, To, Ho, Ho, 128), dtype=bf, device=dev)
        cba_fwd(L["upsample4"], mod.upsample4.bias, View(cat112), View(u4), relu=False, scale_nc=drop_scale)
        rows = N * To * Ho * Ho
        P = torch.empty((32, rows), dtype=torch.float32, device=dev)
        ops.conv_fprop(L["smooth"].packed((To, Ho, Ho), "fprop"), "fprop", View(u4), P)
        logits = torch.empty((N, 1, To, Ho, Ho), dtype=torch.float32, device=dev)
        ops.stencil27_fwd(P, logits, mod.smooth.bias.detach(), N, To, Ho, Ho)
        ctx.mod, ctx.drop, ctx.tail = mod, drop_scale, None
        ctx.t = (x0, c28, c56, c112, cat28, cat56, cat112, u4)
        ctx.needs = ctx.needs_input_grad[:4]
        return logits

    @staticmethod
    def backward(ctx, glog):
        mod = ctx.mod
        L = mod._layers
        x0, c28, c56, c112, cat28, cat56, cat112, u4 = ctx.t
        dev = glog.device
        bf = act_dtype()
        glog = glog.contiguous().float()
        g = {}
        dcat112 = torch.empty_like(cat112)
        if ctx.tail is not None:
            g["upsample4"], g["smooth"] = L["tail"].backward(ctx.tail, View(cat112), glog, View(dcat112))
            return DecoderFn._backward_rest(ctx, g, dcat112)
        N, To, Ho = u4.shape[0], u4.shape[1], u4.shape[2]
        dP = torch.empty((N, To, Ho, Ho, SmoothLayer.BWD_CPAD), dtype=bf, device=dev)
        db_smooth, ds1 = grad_buf(mod.smooth.bias)
        ops.stencil27_bwd(glog, dP, db_smooth, N, To, Ho, Ho, SmoothLayer.BWD_CPAD)
        sm = L["smooth"]
        du4 = torch.empty_like(u4)      # = d(upsample4 pre-dropout output): the Dropout3d scale is applied in the epilogue
        ops.conv_fprop(sm.packed((To, Ho, Ho), "dgrad"), "dgrad", View(dP), View(du4), scale_nc=ctx.drop)
        dw_smooth, ds0 = grad_buf(mod.smooth.weight)
        ops.conv_wgrad(sm.plan((To, Ho, Ho), "dgrad"), View(u4), View(dP), dw_smooth, atomic=True)
        del dP
        g["upsample4"] = cba_bwd(L["upsample4"], mod.upsample4.bias, View(cat112), None, View(du4), False, None, View(dcat112))
        del du4
        g["smooth"] = (None if ds0 else dw_smooth, None if ds1 else db_smooth)
        return DecoderFn._backward_rest(ctx, g, dcat112)

    @staticmethod
    def _backward_rest(ctx, g, dcat112):
        mod = ctx.mod
        L = mod._layers
        x0, c28, c56, c112, cat28, cat56, cat112, _ = ctx.t
        dc112 = torch.empty_like(c112) if ctx.needs[3] else None
        g["conv112"] = cba_bwd(L["conv112"], mod.conv112.bias, View(c112), View(cat112, 64, 64), View(dcat112, 64, 64), True, None,
                               View(dc112) if dc112 is not None else None)
        dcat56 = torch.empty_like(cat56)
        g["upsample3"] = cba_bwd(L["upsample3"], mod.upsample3.bias, View(cat56), View(cat112, 0, 64), View(dcat112, 0, 64), True, None,
                                 View(dcat56))
        dc56 = torch.empty_like(c56) if ctx.needs[2] else None
        g["conv56"] = cba_bwd(L["conv56"], mod.conv56.bias, View(c56), View(cat56, 64, 64), View(dcat56, 64, 64), True, None,
                              View(dc56) if dc56 is not None else None)
        dcat28 = torch.empty_like(cat28)
        g["upsample2"] = cba_bwd(L["upsample2"], mod.upsample2.bias, View(cat28), View(cat56, 0, 64), View(dcat56, 0, 64), True, None,
                                 View(dcat28))
        dc28 = torch.empty_like(c28) if ctx.needs[1] else None
        g["conv28"] = cba_bwd(L["conv28"], mod.conv28.bias, View(c28), View(cat28, 64, 64), View(dcat28, 64, 64), True, None,
                              View(dc28) if dc28 is not None else None)
        dx0 = torch.empty_like(x0) if ctx.needs[0] else None
        g["upsample1"] = cba_bwd(L["upsample1"], mod.upsample1.bias, View(x0), View(cat28, 0, 64), View(dcat28, 0, 64), True, None,
                                 View(dx0) if dx0 is not None else None)
        flat = []
        for n in DecoderFn.ORDER:
            flat += [g[n][0], g[n][1]]
        return (dx0, dc28, dc56, dc112, None, None) + tuple(flat)


class CollapsedTail:
    """upsample4 -> Dropout3d -> smooth (capsules_ucf101.py:504-509) as ONE per-clip stride-2 transposed convolution
    128 -> 1 (include/b200caps.h "Collapsed decoder tail"): composite weights per clip (they contain the
    clip's Dropout3d mask), a 1x1x1 GEMM x -> 216 composite columns, and a stride-2 gather.  The (N,128,8,224,224)
    tensor is never formed; executed MACs drop from 23.6 G to ~1.4 G per clip-pass."""
    COLS, COLS_PAD = 216, 224

    def __init__(self, up4: torch.nn.Module, smooth: torch.nn.Module):
        self.up4, self.smooth = up4, smooth
        self.plans: Dict = {}
        self.bufs: Dict = {}

    def plan(self, dims) -> ConvPlan:
        dims = tuple(int(v) for v in dims)
        key = (dims, PREC.mode)
        pl = self.plans.get(key)
        if pl is None:
            from .plans import packed_geometry, tap_pitch
            pl = ConvPlan(ConvSpec(128, self.COLS, (1, 1, 1), Cout_pad=self.COLS_PAD), dims)
            # per-clip gradient dWeff[n][ci][224]
            pl.wgrad_geom = dict(pl.wgrad_geom, s_p=1, s_g=self.COLS_PAD)
            pl.geo_f = packed_geometry(self.COLS_PAD, 128)                       # (bn, nt, nkb, elems)
            pl.geo_d = packed_geometry(128, tap_pitch(self.COLS_PAD))
            self.plans[key] = pl
        return pl.to(self.up4.weight.device)

    def _weights(self, pl: ConvPlan, drop: Optional[torch.Tensor], N: int):
        """Composite weights (both operand images) + bias field for this step's Dropout3d masks."""
        dev = self.up4.weight.device
        nset = N if drop is not None else 1
        key = (nset, PREC.mode)
        b = self.bufs.get(key)
        if b is None:
            b = dict(f=torch.zeros(nset * pl.geo_f[3], dtype=act_dtype(), device=dev),
                     d=torch.zeros(nset * pl.geo_d[3], dtype=act_dtype(), device=dev),
                     bias=torch.zeros((nset, 27), dtype=torch.float32, device=dev),
                     ones=torch.ones((1, 128), dtype=torch.float32, device=dev))
            self.bufs[key] = b
        dr = drop if drop is not None else b["ones"]
        ops.tail_weff(self.up4.weight.detach(), self.up4.bias.detach(), self.smooth.weight.detach(), dr, b["f"], pl.geo_f[3],
                      b["d"], pl.geo_d[3], pl.geo_d[2], b["bias"], nset)
        esz = b["f"].element_size()
        per_clip = drop is not None
        pl.fprop[0].packed, pl.dgrad[0].packed = b["f"], b["d"]
        pl.fprop_pack = dict(pl.fprop_pack, sample_stride_bytes=pl.geo_f[3] * esz if per_clip else 0)
        pl.dgrad_pack = dict(pl.dgrad_pack, sample_stride_bytes=pl.geo_d[3] * esz if per_clip else 0)
        return b, dr

    def forward(self, x: View, drop: Optional[torch.Tensor]):
        """x: cat112 (N,It,Ih,Iw,128) -> fp32 logits (N,1,2It,2Ih,2Iw); returns (logits, saved state)."""
        N = x.N
        It, Ih, Iw = x.dims
        pl = self.plan(x.dims)
        b, dr = self._weights(pl, drop, N)
        dev = x.t.device
        rows = N * It * Ih * Iw
        Y = torch.empty((self.COLS_PAD, rows), dtype=torch.float32, device=dev)
        ops.conv_fprop(pl, "fprop", x, Y)
        logits = torch.empty((N, 1, 2 * It, 2 * Ih, 2 * Iw), dtype=torch.float32, device=dev)
        bf = b["bias"] if drop is not None else b["bias"].expand(N, 27).contiguous()
        ops.tail_gather_fwd(Y, bf, self.smooth.bias.detach(), logits, N, It, Ih, Iw)
        return logits, (pl, dr, drop is not None)

    def backward(self, saved, x: View, glog: torch.Tensor, dx: Optional[View]):
        """Returns ((dw4, db4), (dws, dbs)) -- None where accumulated straight into .grad; writes dx."""
        pl, dr, per_clip = saved
        assert per_clip, "backward through the eval-mode tail (no Dropout3d masks) is not supported"
        N = x.N
        It, Ih, Iw = x.dims
        dev = x.t.device
        dY = torch.empty((N, It, Ih, Iw, self.COLS_PAD), dtype=act_dtype(), device=dev)
        sums = zeros_f32((N, 27), dev)
        ops.tail_gather_bwd(glog, dY, sums, N, It, Ih, Iw)
        if dx is not None:
            ops.conv_fprop(pl, "dgrad", View(dY), dx)
        dweff = torch.zeros((N, 128, self.COLS_PAD), dtype=torch.float32, device=dev)
        ops.conv_wgrad(pl, x, View(dY), dweff, atomic=True, per_clip=True)
        dw4, d0 = grad_buf(self.up4.weight)
        db4, d1 = grad_buf(self.up4.bias)
        dws, d2 = grad_buf(self.smooth.weight)
        dbs, d3 = grad_buf(self.smooth.bias)
        ops.tail_chain_bwd(dweff, sums, self.up4.weight.detach(), self.up4.bias.detach(), self.smooth.weight.detach(), dr,
                           dw4, db4, dws, dbs, N)
        return (None if d0 else dw4, None if d1 else db4), (None if d2 else dws, None if d3 else dbs)


class SmoothLayer:
    """`smooth` weight (128,1,3,3,3) viewed as the 128 -> 27 projection matrix.  Forward writes 32 planar fp32 planes;
    the backward GEMMs see the tap gradients dP 64 channels wide (zero tail) so both run on the TMA path."""
    BWD_CPAD = 64

    def __init__(self, weight: torch.nn.Parameter):
        self.weight = weight
        self.plans: Dict = {}
        self.keys: Dict = {}

    def plan(self, dims, which="fprop") -> ConvPlan:
        dims = tuple(int(v) for v in dims)
        bwd = which != "fprop"
        pl = self.plans.get((dims, bwd))
        if pl is None:
            pl = ConvPlan.pointwise_from_strides(128, 27, self.BWD_CPAD if bwd else 32, 1, 27, dims)
            self.plans[(dims, bwd)] = pl
        return pl.to(self.weight.device)

    def packed(self, dims, which) -> ConvPlan:
        dims = tuple(int(v) for v in dims)
        pl = self.plan(dims, which)
        w = self.weight
        key = _packed_key(w)
        if not _fresh(self.keys.get((dims, which)), key):
            pl.pack(w.detach(), which, ops.stream())
        self.keys[(dims, which)] = key
        return pl
