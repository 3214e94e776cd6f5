"""Timing + parity probe for the EM-routing kernels (not a pytest file): CUDA-event time of forward / backward at the step's
size (12 800 locations, C = 24) for the kernel selected by B2C_ROUTING (cta | warp), and the deviation from the oracle on 64
locations.     B2C_ROUTING=cta python tests/gpu_routing_bench.py ; python tests/gpu_routing_bench.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "pi-consistency-activity-detection_b200")]


def main():
    from b200caps import ops
    from oracle import restate
    dev = torch.device("cuda")
    C, b = 24, 12800
    sd = restate.make_state_dict(C, seed=0)
    g = torch.Generator().manual_seed(0)
    caps = torch.cat([torch.randn((b, 512), generator=g) * 0.7, torch.rand((b, 32), generator=g)], 1).to(dev).contiguous()
    W = sd["conv_caps.weights"].to(dev).reshape(32, C, 4, 4).contiguous()
    bu, ba = sd["conv_caps.beta_u"].to(dev).contiguous(), sd["conv_caps.beta_a"].to(dev).contiguous()
    out = torch.empty((b, C * 17), device=dev)
    dout = torch.randn((b, C * 17), generator=g).to(dev)
    dcaps = torch.empty_like(caps)
    dW, dbu, dba = torch.zeros_like(W), torch.zeros_like(bu), torch.zeros_like(ba)

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return a.elapsed_time(e) / n

    use_state = os.environ.get("B2C_ROUTING_STATE", "1") != "0" and os.environ.get("B2C_ROUTING", "warp") != "cta"
    state = torch.empty((b, ops.routing_state_floats()), device=dev) if use_state else None
    t_f = timeit(lambda: ops.em_routing_fwd(caps, W, bu, ba, out, b, C, state=state))
    t_b = timeit(lambda: ops.em_routing_bwd(caps, W, bu, ba, dout, dcaps, dW, dbu, dba, b, C, state=state))
    # parity on 64 locations against the fp64 oracle (forward) and its autograd (backward)
    n = 64
    c64 = caps[:n].double().cpu().requires_grad_(True)
    W64 = sd["conv_caps.weights"][0].double().requires_grad_(True)
    bu64, ba64 = sd["conv_caps.beta_u"].double().requires_grad_(True), sd["conv_caps.beta_a"].double().requires_grad_(True)
    mu, a = restate.em_routing(c64[:, :512].reshape(n, 32, 16), c64[:, 512:], W64, bu64, ba64)
    ref = torch.cat([mu.reshape(n, C * 16), a], 1)
    e_f = float((out[:n].double().cpu() - ref.detach()).abs().max() / ref.detach().abs().max())
    gr = torch.autograd.grad(ref, [c64, W64, bu64, ba64], dout[:n].double().cpu())
    dW.zero_(); dbu.zero_(); dba.zero_()
    st_n = torch.empty((n, ops.routing_state_floats()), device=dev) if use_state else None
    out_n = torch.empty((n, C * 17), device=dev)
    ops.em_routing_fwd(caps[:n].contiguous(), W, bu, ba, out_n, n, C, state=st_n)
    ops.em_routing_bwd(caps[:n].contiguous(), W, bu, ba, dout[:n].contiguous(), dcaps[:n], dW, dbu, dba, n, C, state=st_n)
    torch.cuda.synchronize()
    rel = lambda x, y: float((x.double().cpu() - y).abs().max() / (y.abs().max() + 1e-30))
    print(f"B2C_ROUTING={os.environ.get('B2C_ROUTING', 'warp')} B2C_ROUTING_BWD={os.environ.get('B2C_ROUTING_BWD', 'split')} saved-state={use_state}: fwd {t_f:.3f} ms  bwd {t_b:.3f} ms | parity fwd {e_f:.2e} "
          f"dcaps {rel(dcaps[:n], gr[0]):.2e} dW {rel(dW.reshape(gr[1].shape), gr[1]):.2e} dbu {rel(dbu.reshape(gr[2].shape), gr[2]):.2e} "
          f"dba {rel(dba.reshape(gr[3].shape), gr[3]):.2e}")


if __name__ == "__main__":
    main()
