"""Host logic: the conv plans (taps / parity classes / packing / dw index map) reproduce torch's
conv3d, conv_transpose3d and their gradients (fp64, CPU) for every layer type of the step."""
import pytest
import torch
import torch.nn.functional as F

import emu
from b200caps.plans import ConvPlan, ConvSpec, same_pad

torch.manual_seed(0)


def cl(x):   # NCDHW -> NDHWC
    return x.permute(0, 2, 3, 4, 1).contiguous()


def padc(x, c):
    if x.shape[-1] == c:
        return x
    return torch.cat([x, torch.zeros(x.shape[:-1] + (c - x.shape[-1],), dtype=x.dtype)], -1)


CASES = [
    # name, transposed, Cin, Cout, k, stride, in_dims, pad spec
    ("stem7x7 s2 same", False, 3, 8, (7, 7, 7), (2, 2, 2), (8, 10, 12), "same"),
    ("1x1", False, 16, 24, (1, 1, 1), (1, 1, 1), (2, 5, 6), "same"),
    ("3x3 s(2,1,1) same", False, 8, 16, (3, 3, 3), (2, 1, 1), (4, 6, 5), "same"),
    ("3x3 same", False, 16, 8, (3, 3, 3), (1, 1, 1), (2, 5, 5), "same"),
    ("3x3 T=1 same", False, 8, 8, (3, 3, 3), (1, 1, 1), (1, 6, 6), "same"),
    ("conv2d 9x9 valid", False, 8, 16, (1, 9, 9), (1, 1, 1), (1, 12, 12), 0),
    ("conv 3x3 p1", False, 8, 8, (3, 3, 3), (1, 1, 1), (2, 6, 6), 1),
    ("conv2d 3x3 p1", False, 8, 8, (1, 3, 3), (1, 1, 1), (1, 6, 6), (0, 1, 1)),
    ("convT2d 9x9", True, 16, 8, (1, 9, 9), (1, 1, 1), (1, 4, 4), 0),
    ("convT3d k3 s2 p1 op1", True, 8, 16, (3, 3, 3), (2, 2, 2), (2, 3, 4), 1),
    ("convT3d k3 s2 p1 op1 T=1", True, 16, 8, (3, 3, 3), (2, 2, 2), (1, 3, 3), 1),
    ("convT3d smooth k3 p1", True, 8, 8, (3, 3, 3), (1, 1, 1), (3, 4, 4), 1),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_plan_matches_torch(case):
    name, tr, Cin, Cout, k, s, dims, pad = case
    N = 2
    x = torch.randn((N, Cin) + dims, dtype=torch.float64, requires_grad=True)
    if tr:
        w = torch.randn((Cin, Cout) + k, dtype=torch.float64, requires_grad=True)
        p = (pad,) * 3 if isinstance(pad, int) else pad
        op = tuple(si - 1 for si in s)
        y = F.conv_transpose3d(x, w, None, stride=s, padding=p, output_padding=op)
        spec = ConvSpec(Cin, Cout, k, s, p, (0, 0, 0), op, True)
    else:
        w = torch.randn((Cout, Cin) + k, dtype=torch.float64, requires_grad=True)
        if pad == "same":
            pads = [same_pad(d, kk, ss) for d, kk, ss in zip(dims, k, s)]
        else:
            pp = (pad,) * 3 if isinstance(pad, int) else pad
            pads = [(v, v) for v in pp]
        xp = F.pad(x, (pads[2][0], pads[2][1], pads[1][0], pads[1][1], pads[0][0], pads[0][1]))
        y = F.conv3d(xp, w, None, stride=s)
        spec = ConvSpec(Cin, Cout, k, s, tuple(p[0] for p in pads), tuple(p[1] for p in pads))
    plan = ConvPlan(spec, dims)
    assert tuple(plan.out_dims) == tuple(y.shape[2:])
    gy = torch.randn_like(y)
    gx, gw = torch.autograd.grad(y, (x, w), gy)
    xc = padc(cl(x.detach()), spec.Cin_pad)
    yc = emu.conv(plan, "fprop", xc, w.detach(), plan.out_dims)
    assert torch.allclose(yc[..., :Cout], cl(y.detach()), atol=1e-10), name
    gyc = padc(cl(gy), spec.Cout_pad)
    gxc = emu.conv(plan, "dgrad", gyc, w.detach(), plan.in_dims)
    assert torch.allclose(gxc[..., :Cin], cl(gx), atol=1e-10), name
    gwc = emu.wgrad(plan, xc, gyc, tuple(w.shape))
    assert torch.allclose(gwc, gw, atol=1e-9), name


def test_transposed_classes_skip_zero_taps():
    spec = ConvSpec(128, 128, (3, 3, 3), (2, 2, 2), (1, 1, 1), (0, 0, 0), (1, 1, 1), True)
    plan = ConvPlan(spec, (4, 112, 112))
    assert sorted(len(c.taps) for c in plan.fprop) == [1, 2, 2, 2, 4, 4, 4, 8]
    # algorithmic MACs of upsample4 per clip (22.196 GMAC)
    assert abs(plan.macs_fprop(1) / 1e9 - 22.196) < 0.01


def test_fastdiv_formula_matches_integer_division():
    """Host restatement of csrc/igemm.cu make_fastdiv / fdiv (Granlund-Montgomery multiply-shift division used for the
    tile -> (clip, t, h, w) arithmetic of every warp role): exact for all 32-bit numerators the kernels can produce."""
    import random

    def make(d):
        l = 0
        while (1 << l) < d:
            l += 1
        m = ((((1 << l) - d) << 32) // d + 1) & 0xFFFFFFFF
        return m, min(l, 1), max(l - 1, 0)

    def fdiv(n, p):
        m, sh1, sh2 = p
        t = (m * n) >> 32
        return ((t + (((n - t) & 0xFFFFFFFF) >> sh1)) & 0xFFFFFFFF) >> sh2

    rng = random.Random(0)
    divisors = list(range(1, 300)) + [448, 544, 832, 1088, 4096, 50176, 401408, 12845056]
    for d in divisors:
        p = make(d)
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, (1 << 31) - 1, (1 << 31) - d] + [rng.randrange(0, 1 << 31) for _ in range(200)]
        for n in ns:
            assert fdiv(n, p) == n // d, (n, d)


def test_padding_only_taps_are_pruned_exactly():
    """plans._prune_padding_taps drops exactly the taps that read padding for EVERY output position (brute force)."""
    from b200caps.plans import ConvPlan, ConvSpec, same_pad
    cases = [
        (ConvSpec(160, 320, (3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 1, 1)), (1, 28, 28)),          # Mixed_4f.b1b: one frame
        (ConvSpec(96, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 1, 1)), (2, 28, 28)),           # two frames: nothing to prune
        (ConvSpec(128, 64, (3, 3, 3), (2, 2, 2), (1, 1, 1), (0, 0, 0), (1, 1, 1), True), (1, 28, 28)),   # upsample2
        (ConvSpec(832, 544, (1, 9, 9)), (1, 28, 28)),                                            # PrimaryCaps
    ]
    for spec, dims in cases:
        pl = ConvPlan(spec, dims)
        for classes, si, idims in ((pl.fprop, pl.fprop_si, pl.in_dims), (pl.dgrad, pl.dgrad_si, pl.out_dims)):
            for cl in classes:
                for tap in cl.taps:      # every kept tap touches real data somewhere
                    assert all(any(0 <= q * s + d < I for q in range(Q)) for d, s, I, Q in zip(tap, si, idims, cl.Q)), (spec, tap)
    one = ConvPlan(cases[0][0], cases[0][1])
    assert len(one.fprop[0].taps) == 9 and all(t[0] == 0 for t in one.fprop[0].taps)
    assert len(ConvPlan(cases[1][0], cases[1][1]).fprop[0].taps) == 27
    up2 = ConvPlan(cases[2][0], cases[2][1])
    assert [len(c.taps) for c in up2.fprop] == [1, 2, 2, 4, 1, 2, 2, 4] and len(up2.dgrad[0].taps) == 18
    # the weight-tap offsets stay aligned with the taps
    for cl in one.fprop:
        assert len(cl.wtap) == len(cl.taps) and cl.wtap == [9 + i for i in range(9)]


def test_h_block_detection():
    import os
    from b200caps.plans import ConvPlan, ConvSpec, h_block_of
    os.environ["B2C_TAP_SKIP"] = "1"
    try:
        pc = ConvPlan(ConvSpec(832, 544, (1, 9, 9)), (1, 28, 28))
        assert h_block_of(pc.dgrad[0].taps) == 9 and h_block_of(pc.fprop[0].taps) == 9
        c3 = ConvPlan(ConvSpec(96, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 1, 1)), (2, 28, 28))
        assert h_block_of(c3.fprop[0].taps) == 0          # several dt values: no single-frame block structure
    finally:
        os.environ.pop("B2C_TAP_SKIP")
    assert h_block_of(pc.dgrad[0].taps) == 0               # off by default


def test_folded_stem_plan_geometry():
    """engine.StemLayer.fold_plan: 2-D 7x7 stride-2 convolution over 64 folded channels, 4 x 64 output columns."""
    import torch
    from b200caps.engine import StemLayer
    w = torch.nn.Parameter(torch.zeros(64, 3, 7, 7, 7))
    st = StemLayer(w, 3, 64, (7, 7, 7), (2, 2, 2))
    assert st.use_fold((8, 224, 224)) and st.use_fold((8, 64, 64))
    assert not st.use_fold((32, 224, 224))                  # more padded frames than the 16 the fold holds
    od, pf = st.geometry((8, 224, 224))
    assert od == (4, 112, 112) and pf == (2, 2, 2)
    pl = st.fold_plan((8, 224, 224))
    assert pl.in_dims == (1, 224, 224) and pl.out_dims == (4, 112, 112) and len(pl.fprop) == 1 and len(pl.fprop[0].taps) == 49
    assert pl.fprop_pack["R"] == 256 and pl.fprop_pack["out_fold"] == 64 and pl.wgrad_geom["p_fold"] == 64
    assert pl.fprop[0].wtap == [i * 64 for i in range(49)]


def test_split_fprop_k_partitions_the_taps():
    """ConvPlan.split_fprop_k (PrimaryCaps forward): the K slices hold every tap exactly once in the original order, share
    the output grid, write to consecutive output frames, and leave the wgrad class (all taps) untouched."""
    from b200caps.plans import ConvPlan, ConvSpec
    pl = ConvPlan(ConvSpec(832, 544, (1, 9, 9)), (1, 28, 28))
    full = pl.fprop[0]
    taps, wtap = list(full.taps), list(full.wtap)
    assert len(taps) == 81 and pl.out_dims == (1, 20, 20)
    pl.split_fprop_k(8)
    assert len(pl.fprop) == 8 and pl.fprop_out_dims == (8, 20, 20)
    assert sum((c.taps for c in pl.fprop), []) == taps and sum((c.wtap for c in pl.fprop), []) == wtap
    assert [c.po for c in pl.fprop] == [(s, 0, 0) for s in range(8)] and all(c.Q == full.Q for c in pl.fprop)
    assert {len(c.taps) for c in pl.fprop} <= {10, 11}
    assert pl.wgrad_cls is full and len(pl.wgrad_cls.taps) == 81
    assert pl.macs_fprop(2) == 2 * 400 * 81 * 832 * 544
    # one slice = no split
    p1 = ConvPlan(ConvSpec(832, 544, (1, 9, 9)), (1, 28, 28))
    p1.split_fprop_k(1)
    assert len(p1.fprop) == 1 and not hasattr(p1, "fprop_out_dims")


def test_rows_major_plan_shares_the_packed_operand_and_marks_tap_blocks():
    """FusedConvLayer.rows_major_dgrad's geometry (PrimaryCaps dgrad): the layer with its image rows on T, columns on H, clips
    on W has the same taps in the same order as the clip-major plan (so both read one packed operand), its tap list is
    blocks of constant dt walking one monotonic dh sequence (what b2c_conv_class.h_block < 0 promises the kernel), and the
    fraction of (tap, position) pairs a perfect skipper would keep is 51 %."""
    from b200caps.plans import ConvPlan, ConvSpec, t_block_of, h_block_of
    base = ConvPlan(ConvSpec(832, 544, (1, 9, 9)), (1, 28, 28))
    rows = ConvPlan(ConvSpec(832, 544, (9, 9, 1)), (28, 28, 32))
    assert rows.out_dims == (20, 20, 32) and len(rows.dgrad) == len(base.dgrad) == 1
    assert rows.dgrad[0].wtap == base.dgrad[0].wtap
    assert [(t[0], t[1]) for t in rows.dgrad[0].taps] == [(t[1], t[2]) for t in base.dgrad[0].taps]
    assert t_block_of(rows.dgrad[0].taps) == 9 and rows.dgrad[0].Q == (28, 28, 32)
    # the clip-major tap list is no T-block list (one dt), and arbitrary lists are rejected
    assert t_block_of(base.dgrad[0].taps) == 0
    assert t_block_of([(0, 0, 0), (0, -1, 0), (-1, 0, 0), (-1, -2, 0)]) == 0
    # useful (tap, position) pairs of the 'valid' 9x9 gradient: per axis sum_x #{k: 0 <= x - k <= 19} = 180 of 28 * 9
    per_axis = sum(sum(1 for k in range(9) if 0 <= x - k <= 19) for x in range(28))
    assert per_axis == 180 and abs((per_axis / 252.0) ** 2 - 0.51) < 0.005
