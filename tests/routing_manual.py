"""Hand-derived backward of the EM routing (the math csrc/routing.cu implements), written with
plain tensor ops so it can be checked against autograd of the oracle on CPU."""
import math

import torch


def forward_saved(poses, a_in, W, beta_u, beta_a, iters=3, eps=1e-8, lam=1e-6):
    b, B, _ = poses.shape
    C = W.shape[1]
    V = torch.einsum("nirk,ijkc->nijrc", poses.view(b, B, 4, 4), W).reshape(b, B, C, 16)
    r_prev = torch.full((b, B, C), 1.0 / C, dtype=V.dtype)
    saved = []
    for t in range(iters):
        rp = r_prev * a_in.view(b, B, 1)
        Z = rp.sum(2, keepdim=True) + eps
        rn = rp / Z
        R = rn.sum(1)                                   # (b,C)
        c = rn / (R.unsqueeze(1) + eps)                 # (b,B,C)
        mu = (c.unsqueeze(-1) * V).sum(1)               # (b,C,16)
        S = (c.unsqueeze(-1) * (V - mu.unsqueeze(1)) ** 2).sum(1) + eps
        T = (beta_u + 0.5 * torch.log(S)).sum(-1)       # (b,C)
        cost = T * R
        m = cost.mean(1, keepdim=True)
        s = torch.sqrt((cost - m).sum(1, keepdim=True) ** 2 / C + eps)
        a = torch.sigmoid(lam * (beta_a - (m - cost) / (s + eps)))
        saved.append(dict(r_prev=r_prev, Z=Z, rn=rn, R=R, c=c, mu=mu, S=S, T=T, a=a, s=s))
        if t < iters - 1:
            lnp = (-(V - mu.unsqueeze(1)) ** 2 / (2 * S.unsqueeze(1)) - 0.5 * torch.log(S.unsqueeze(1))
                   - 0.5 * math.log(2 * math.pi)).sum(-1)
            z = lnp + torch.log(eps + a).unsqueeze(1)
            r_prev = torch.softmax(z, dim=2)
    return V, saved


def backward_manual(poses, a_in, W, beta_u, beta_a, g_mu, g_a, iters=3, eps=1e-8, lam=1e-6):
    b, B, _ = poses.shape
    C = W.shape[1]
    V, saved = forward_saved(poses, a_in, W, beta_u, beta_a, iters, eps, lam)
    gV = torch.zeros_like(V)
    g_beta_u = torch.zeros_like(beta_u)
    g_beta_a = torch.zeros_like(beta_a)
    g_ain = torch.zeros_like(a_in)
    gmu, gS, ga = g_mu.clone(), torch.zeros_like(g_mu), g_a.clone()
    gr = None
    for t in reversed(range(iters)):
        sv = saved[t]
        mu, S, a, R, c, rn, Z, T = sv["mu"], sv["S"], sv["a"], sv["R"], sv["c"], sv["rn"], sv["Z"], sv["T"]
        dV = V - mu.unsqueeze(1)
        if t < iters - 1:
            r = saved[t + 1]["r_prev"]
            gz = r * (gr - (gr * r).sum(2, keepdim=True))
            ga = (gz / (eps + a).unsqueeze(1)).sum(1)
            gmu = (gz.unsqueeze(-1) * dV / S.unsqueeze(1)).sum(1)
            gS = (gz.unsqueeze(-1) * (dV ** 2 / (2 * S.unsqueeze(1) ** 2) - 0.5 / S.unsqueeze(1))).sum(1)
            gV = gV - gz.unsqueeze(-1) * dV / S.unsqueeze(1)
        gu = ga * a * (1 - a)
        g_beta_a = g_beta_a + lam * gu.sum(0)
        gcost = lam / (sv["s"] + eps) * (gu - gu.mean(1, keepdim=True))
        g_beta_u = g_beta_u + (gcost * R).unsqueeze(-1).expand(-1, -1, 16).sum(0)
        gS = gS + (gcost * R).unsqueeze(-1) * 0.5 / S
        gR = gcost * T
        csum = R / (R + eps)
        gmu = gmu - 2 * gS * mu * (1 - csum).unsqueeze(-1)
        gc = (gS.unsqueeze(1) * dV ** 2 + gmu.unsqueeze(1) * V).sum(-1)          # (b,B,C)
        gV = gV + c.unsqueeze(-1) * (gmu.unsqueeze(1) + 2 * gS.unsqueeze(1) * dV)
        D = (gc * c).sum(1)
        gR_tot = gR - D / (R + eps)
        grn = gc / (R.unsqueeze(1) + eps) + gR_tot.unsqueeze(1)
        grp = (grn - (grn * rn).sum(2, keepdim=True)) / Z
        g_ain = g_ain + (grp * sv["r_prev"]).sum(2)
        gr = grp * a_in.view(b, B, 1)
    gV4 = gV.view(b, B, C, 4, 4)
    g_poses = torch.einsum("nijrc,ijkc->nirk", gV4, W).reshape(b, B, 16)
    g_W = torch.einsum("nirk,nijrc->ijkc", poses.view(b, B, 4, 4), gV4)
    return g_poses, g_ain, g_W, g_beta_u, g_beta_a


def backward_split(poses, a_in, W, beta_u, beta_a, g_mu, g_a, iters=3, eps=1e-8, lam=1e-6):
    """The same backward in the form of the two-kernel split (csrc/routing.cu: em_routing_bwd_coef_kernel +
    em_routing_bwd_final_kernel).  Phase 1 walks t = 2, 1 and produces only per-(i,j) scalars gz^t and per-j vectors
    (X' = 2 gS' / (R+eps), G' = gmu' / (R+eps), U = 1/S); the sum D_j = sum_i gc_ij c_ij is closed-form
    (sum_i c (V-mu)^2 = S - eps, sum_i c V = mu).  Phase 2 assembles every vote gradient in ONE pass:
        gV_ij = sum_t rn^t (G'_t + X'_t (V - mu_t)) - sum_{t<2} gz^t (V - mu_t) U_t
    and does iteration 0's activation-gradient term."""
    assert iters == 3
    b, B, _ = poses.shape
    C = W.shape[1]
    V, saved = forward_saved(poses, a_in, W, beta_u, beta_a, iters, eps, lam)
    g_beta_u = torch.zeros_like(beta_u)
    g_beta_a = torch.zeros_like(beta_a)
    g_ain = torch.zeros_like(a_in)
    gmu, gS, ga = g_mu.clone(), torch.zeros_like(g_mu), g_a.clone()
    Xp, Gp, U, gz_saved = {}, {}, {}, {}
    gR_tot0 = None
    # ---- phase 1 ----
    for t in (2, 1, 0):
        sv = saved[t]
        mu, S, a, R, rn, Z, T = sv["mu"], sv["S"], sv["a"], sv["R"], sv["rn"], sv["Z"], sv["T"]
        invR = 1.0 / (R + eps)
        gu = ga * a * (1 - a)
        g_beta_a = g_beta_a + lam * gu.sum(0)
        gcost = lam / (sv["s"] + eps) * (gu - gu.mean(1, keepdim=True))
        g_beta_u = g_beta_u + (gcost * R).unsqueeze(-1).expand(-1, -1, 16).sum(0)
        gSh = gS + (gcost * R).unsqueeze(-1) * 0.5 / S
        gmh = gmu - 2 * gSh * mu * (1 - R * invR).unsqueeze(-1)
        D = (gSh * (S - eps) + gmh * mu).sum(-1)
        gR_tot = gcost * T - D * invR
        Xp[t] = 2 * gSh * invR.unsqueeze(-1)
        Gp[t] = gmh * invR.unsqueeze(-1)
        if t == 0:
            gR_tot0 = gR_tot
            break
        dV = V - mu.unsqueeze(1)
        gc = (gSh.unsqueeze(1) * dV * dV + gmh.unsqueeze(1) * V).sum(-1)
        grn = gc * invR.unsqueeze(1) + gR_tot.unsqueeze(1)
        grp = (grn - (grn * rn).sum(2, keepdim=True)) / Z
        g_ain = g_ain + (grp * sv["r_prev"]).sum(2)
        gr = grp * a_in.view(b, B, 1)
        r = sv["r_prev"]
        gz = r * (gr - (gr * r).sum(2, keepdim=True))
        gz_saved[t - 1] = gz
        p = saved[t - 1]
        iS = 1.0 / p["S"]
        U[t - 1] = iS
        d0 = V - p["mu"].unsqueeze(1)
        dv = d0 * iS.unsqueeze(1)
        ga = gz.sum(1) / (eps + p["a"])
        gmu = (gz.unsqueeze(-1) * dv).sum(1)
        gS = (gz.unsqueeze(-1) * 0.5 * iS.unsqueeze(1) * (dv * d0 - 1)).sum(1)
    # ---- phase 2 ----
    d = [V - saved[t]["mu"].unsqueeze(1) for t in range(3)]
    gV = torch.zeros_like(V)
    for t in range(3):
        gV = gV + saved[t]["rn"].unsqueeze(-1) * (Gp[t].unsqueeze(1) + Xp[t].unsqueeze(1) * d[t])
    for t in range(2):
        gV = gV - gz_saved[t].unsqueeze(-1) * d[t] * U[t].unsqueeze(1)
    gc0 = (0.5 * Xp[0].unsqueeze(1) * d[0] * d[0] + Gp[0].unsqueeze(1) * V).sum(-1)
    grn = gc0 + gR_tot0.unsqueeze(1)
    grp = (grn - (grn * saved[0]["rn"]).sum(2, keepdim=True)) / saved[0]["Z"]
    g_ain = g_ain + (grp * saved[0]["r_prev"]).sum(2)
    gV4 = gV.view(b, B, C, 4, 4)
    g_poses = torch.einsum("nijrc,ijkc->nirk", gV4, W).reshape(b, B, 16)
    g_W = torch.einsum("nirk,nijrc->ijkc", poses.view(b, B, 4, 4), gV4)
    return g_poses, g_ain, g_W, g_beta_u, g_beta_a
