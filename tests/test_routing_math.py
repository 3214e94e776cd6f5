"""The hand-derived EM-routing backward (tests/routing_manual.py == the math of csrc/routing.cu)
agrees with autograd of the oracle restatement, and with the reference's own gradient (golden)."""
import json
import os

import torch

import routing_manual
from oracle import restate

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_manual_backward_matches_autograd():
    g = torch.Generator().manual_seed(0)
    b, B, C = 7, 32, 24
    poses = (torch.randn((b, B, 16), generator=g, dtype=torch.float64) * 0.7).requires_grad_(True)
    a_in = torch.rand((b, B), generator=g, dtype=torch.float64).requires_grad_(True)
    W = torch.randn((B, C, 4, 4), generator=g, dtype=torch.float64).requires_grad_(True)
    bu = torch.randn((C, 16), generator=g, dtype=torch.float64).requires_grad_(True)
    ba = torch.randn((C,), generator=g, dtype=torch.float64).requires_grad_(True)
    mu, a = restate.em_routing(poses, a_in, W, bu, ba)
    gmu = torch.randn(mu.shape, generator=g, dtype=torch.float64)
    ga = torch.randn(a.shape, generator=g, dtype=torch.float64) * 1e3
    ref = torch.autograd.grad((mu * gmu).sum() + (a * ga).sum(), (poses, a_in, W, bu, ba))
    with torch.no_grad():
        man = routing_manual.backward_manual(poses, a_in, W, bu, ba, gmu, ga)
    for name, r, m in zip(("poses", "a_in", "W", "beta_u", "beta_a"), ref, man):
        err = float((r - m).abs().max() / (r.abs().max() + 1e-300))
        assert err < 1e-7, (name, err)


def test_manual_backward_matches_reference_golden():
    kat = json.load(open(os.path.join(GOLD, "kat_small.json")))["routing"]
    sd = restate.make_state_dict(24, seed=0, dtype=torch.float64)
    x = torch.tensor(kat["x"], dtype=torch.float64)
    gout = torch.tensor(kat["gout"], dtype=torch.float64)
    gin = torch.tensor(kat["gin"], dtype=torch.float64)
    poses, a_in = x[:, :512].reshape(6, 32, 16), x[:, 512:]
    mu, a = restate.em_routing(poses, a_in, sd["conv_caps.weights"][0], sd["conv_caps.beta_u"], sd["conv_caps.beta_a"])
    out = torch.cat([mu.reshape(6, 384), a], 1)
    assert float((out - torch.tensor(kat["out"], dtype=torch.float64)).abs().max()) < 1e-10
    gp, ga, _, _, _ = routing_manual.backward_manual(poses, a_in, sd["conv_caps.weights"][0], sd["conv_caps.beta_u"],
                                                     sd["conv_caps.beta_a"], gout[:, :384].reshape(6, 24, 16), gout[:, 384:])
    man = torch.cat([gp.reshape(6, 512), ga], 1)
    err = float((man - gin).abs().max() / gin.abs().max())
    assert err < 1e-7, err


def test_split_backward_matches_autograd():
    """The two-kernel formulation of the backward (csrc/routing.cu: em_routing_bwd_coef_kernel + em_routing_bwd_final_kernel;
    tensor-op form routing_manual.backward_split: closed-form D_j, per-(i,j) scalars gz^t, per-j vectors X' G' U, the vote
    gradient of all three iterations assembled in one pass) agrees with autograd of the oracle and with the
    iteration-by-iteration form."""
    g = torch.Generator().manual_seed(3)
    for C in (24, 21):
        b, B = 5, 32
        poses = (torch.randn((b, B, 16), generator=g, dtype=torch.float64) * 0.7).requires_grad_(True)
        a_in = torch.rand((b, B), generator=g, dtype=torch.float64).requires_grad_(True)
        W = torch.randn((B, C, 4, 4), generator=g, dtype=torch.float64).requires_grad_(True)
        bu = torch.randn((C, 16), generator=g, dtype=torch.float64).requires_grad_(True)
        ba = torch.randn((C,), generator=g, dtype=torch.float64).requires_grad_(True)
        mu, a = restate.em_routing(poses, a_in, W, bu, ba)
        gmu = torch.randn(mu.shape, generator=g, dtype=torch.float64)
        ga = torch.randn(a.shape, generator=g, dtype=torch.float64) * 1e3
        ref = torch.autograd.grad((mu * gmu).sum() + (a * ga).sum(), (poses, a_in, W, bu, ba))
        with torch.no_grad():
            spl = routing_manual.backward_split(poses, a_in, W, bu, ba, gmu, ga)
            man = routing_manual.backward_manual(poses, a_in, W, bu, ba, gmu, ga)
        for name, r, s_, m in zip(("poses", "a_in", "W", "beta_u", "beta_a"), ref, spl, man):
            err = float((r - s_).abs().max() / (r.abs().max() + 1e-300))
            assert err < 1e-7, (C, name, err)
            assert float((m - s_).abs().max() / (m.abs().max() + 1e-300)) < 1e-7, (C, name)
