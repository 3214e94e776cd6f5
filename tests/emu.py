"""CPU emulation of csrc/igemm.cu's data movement, driven by the SAME plan tables the kernels
consume (taps, parity classes, packed weights, dw index map).  Used by the CPU tests to prove the
host-side planning reproduces torch's conv / conv_transpose forward and gradients."""
import torch


def pack(plan, weight, which):
    classes = plan.fprop if which == "fprop" else plan.dgrad
    pk = plan.fprop_pack if which == "fprop" else plan.dgrad_pack
    wf = weight.reshape(-1)
    out = []
    for cl in classes:
        nt = len(cl.taps)
        p = torch.zeros((pk["R_pad"], nt, pk["C"]), dtype=weight.dtype)
        r = torch.arange(pk["R"]).view(-1, 1, 1)
        t = torch.tensor(cl.wtap).view(1, -1, 1)
        c = torch.arange(pk["C_real"]).view(1, 1, -1)
        p[:pk["R"], :, :pk["C_real"]] = wf[r * pk["s_r"] + c * pk["s_c"] + t]
        out.append(p.reshape(pk["R_pad"], nt * pk["C"]))
    return out


def _gather(x, q_dims, si, taps):
    """x (N,T,H,W,C) -> (N,Qt,Qh,Qw,ntaps,C) with zero fill out of bounds."""
    N, T, H, W, C = x.shape
    qt, qh, qw = q_dims
    out = torch.zeros((N, qt, qh, qw, len(taps), C), dtype=x.dtype)
    for ti, (dt, dh, dw) in enumerate(taps):
        it = torch.arange(qt) * si[0] + dt
        ih = torch.arange(qh) * si[1] + dh
        iw = torch.arange(qw) * si[2] + dw
        vt, vh, vw = (it >= 0) & (it < T), (ih >= 0) & (ih < H), (iw >= 0) & (iw < W)
        g = x[:, it.clamp(0, T - 1)][:, :, ih.clamp(0, H - 1)][:, :, :, iw.clamp(0, W - 1)]
        m = (vt.view(-1, 1, 1) & vh.view(1, -1, 1) & vw.view(1, 1, -1)).to(x.dtype)
        out[:, :, :, :, ti] = g * m.view(1, qt, qh, qw, 1)
    return out


def conv(plan, which, x, weight, out_dims):
    """Emulates b2c_conv_fprop for 'fprop' or 'dgrad'.  x channels-last (N,T,H,W,Cpad)."""
    classes = plan.fprop if which == "fprop" else plan.dgrad
    si, so = (plan.fprop_si, plan.fprop_so) if which == "fprop" else (plan.dgrad_si, plan.dgrad_so)
    pk = plan.fprop_pack if which == "fprop" else plan.dgrad_pack
    packed = pack(plan, weight, which)
    N = x.shape[0]
    out = torch.zeros((N,) + tuple(out_dims) + (pk["R_pad"],), dtype=x.dtype)
    for cl, wp in zip(classes, packed):
        a = _gather(x, cl.Q, si, cl.taps).reshape(N, cl.Q[0], cl.Q[1], cl.Q[2], -1)
        y = a @ wp.t()
        ot = torch.arange(cl.Q[0]) * so[0] + cl.po[0]
        oh = torch.arange(cl.Q[1]) * so[1] + cl.po[1]
        ow = torch.arange(cl.Q[2]) * so[2] + cl.po[2]
        out[:, ot.view(-1, 1, 1), oh.view(1, -1, 1), ow.view(1, 1, -1)] = y
    return out


def wgrad(plan, x, dy, dw_shape):
    """Emulates b2c_conv_wgrad: returns dw in the torch weight layout."""
    geo, cl = plan.wgrad_geom, plan.wgrad_cls
    g, p = (x, dy) if geo["g_is_input"] else (dy, x)
    a = _gather(g, geo["Q"], geo["sg"], cl.taps)            # (N,Q..,ntaps,Cg)
    N = a.shape[0]
    a = a.reshape(-1, len(cl.taps) * geo["Cg"])
    b = p.reshape(-1, geo["Cp"])
    assert a.shape[0] == b.shape[0]
    D = a.t() @ b                                            # (ntaps*Cg, Cp)
    dw = torch.zeros(int(torch.tensor(dw_shape).prod()), dtype=x.dtype)
    D = D.reshape(len(cl.taps), geo["Cg"], geo["Cp"])[:, :geo["Cg_real"], :]
    t = torch.tensor(cl.wtap).view(-1, 1, 1)
    gc = torch.arange(geo["Cg_real"]).view(1, -1, 1)
    pc = torch.arange(geo["Cp"]).view(1, 1, -1)
    dw.index_put_(((pc * geo["s_p"] + gc * geo["s_g"] + t).reshape(-1),), D.reshape(-1), accumulate=True)
    return dw.reshape(dw_shape)
