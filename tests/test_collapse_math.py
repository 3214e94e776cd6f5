"""CPU: the algebra behind the collapsed decoder tail (csrc/collapse.cu; capsules_ucf101.py:504-509).
upsample4 (ConvTranspose3d k3 s2 p1 op1) -> Dropout3d -> smooth (ConvTranspose3d k3 p1) equals, in exact arithmetic, one
per-clip transposed convolution with 6 composite columns per dimension (e = k + m in 0..4, plus e = 2 without the
(k = 0, m = 2) term for input index 0) followed by the gather o = 2 i - 2 + e and a per-border-class bias field.
The CUDA kernels implement exactly the index conventions restated here."""
import itertools

import torch
import torch.nn.functional as F

PAIRS = {c: [(k, m) for k in range(3) for m in range(3) if k + m == c] for c in range(5)}
PAIRS[5] = [(1, 1), (2, 0)]


def cands(o, I):
    out = []
    for e in ((0, 2, 4) if o % 2 == 0 else (1, 3)):
        i = (o + 2 - e) // 2
        if 0 <= i < I:
            out.append((i, 5 if (e == 2 and i == 0) else e))
    return out


def test_collapsed_tail_is_exact_including_borders_and_bias():
    torch.manual_seed(0)
    dt = torch.float64
    N, Ci, C = 2, 6, 5
    I = (2, 3, 4)
    x = torch.randn(N, Ci, *I, dtype=dt)
    W4 = torch.randn(Ci, C, 3, 3, 3, dtype=dt) * 0.3
    b4 = torch.randn(C, dtype=dt)
    Ws = torch.randn(C, 1, 3, 3, 3, dtype=dt) * 0.3
    bs = torch.randn(1, dtype=dt)
    d = (torch.rand(N, C) < 0.5).to(dt) * 2
    u = F.conv_transpose3d(x, W4, b4, stride=2, padding=1, output_padding=1) * d.view(N, C, 1, 1, 1)
    ref = F.conv_transpose3d(u, Ws, bs, padding=1)[:, 0]
    # composite weights: T[n][ci][k][m] = sum_c d[n,c] W4[ci,c,k] Ws[c,m]; column = sum of its (k, m) triples
    T = torch.einsum("nc,ick,cm->nikm", d, W4.reshape(Ci, C, 27), Ws.reshape(C, 27))
    Weff = torch.zeros(N, Ci, 216, dtype=dt)
    for ct, ch, cw in itertools.product(range(6), repeat=3):
        for (kt, mt), (kh, mh), (kw, mw) in itertools.product(PAIRS[ct], PAIRS[ch], PAIRS[cw]):
            Weff[:, :, (ct * 6 + ch) * 6 + cw] += T[:, :, (kt * 3 + kh) * 3 + kw, (mt * 3 + mh) * 3 + mw]
    Y = torch.einsum("nithw,nic->nthwc", x, Weff)
    Bn = torch.einsum("nc,c,cm->nm", d, b4, Ws.reshape(C, 27))
    O = tuple(2 * i for i in I)
    ok = lambda o, m, Od: not ((o == 0 and m == 2) or (o == Od - 1 and m == 0))
    out = torch.zeros(N, *O, dtype=dt)
    for ot, oh, ow in itertools.product(*[range(v) for v in O]):
        acc = bs.expand(N).clone()
        for (it, ct), (ih, ch), (iw, cw) in itertools.product(cands(ot, I[0]), cands(oh, I[1]), cands(ow, I[2])):
            acc = acc + Y[:, it, ih, iw, (ct * 6 + ch) * 6 + cw]
        for mt, mh, mw in itertools.product(range(3), repeat=3):
            if ok(ot, mt, O[0]) and ok(oh, mh, O[1]) and ok(ow, mw, O[2]):
                acc = acc + Bn[:, (mt * 3 + mh) * 3 + mw]
        out[:, ot, oh, ow] = acc
    assert float((out - ref).abs().max()) < 1e-12 * float(ref.abs().max() + 1)
