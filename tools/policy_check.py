"""Development aid: the fused tcgen05 policy kernel against the torch module it replaces (numerics + time).
usage: python tools/policy_check.py [N]"""
import sys

import torch

sys.path.insert(0, ".")
import ilrl_b200  # noqa: E402,F401
from ilrl_b200 import GaussianMLPPolicy  # noqa: E402
from ilrl_b200.rollout import FusedPolicy  # noqa: E402


def bf16_reference(policy, obs):
    """The same arithmetic contract in torch: bf16 operands and hidden activations, fp32 accumulation."""
    def run(seq):
        x = obs.to(torch.bfloat16).float()
        lins = [m for m in seq if isinstance(m, torch.nn.Linear)]
        for i, m in enumerate(lins):
            x = x @ m.weight.to(torch.bfloat16).float().t() + m.bias.float()
            if i < 2:
                x = torch.tanh(x).to(torch.bfloat16).float()
        return x
    return run(policy.pi), run(policy.vf).squeeze(-1)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    for obs_dim, act_dim in ((70, 17), (44, 2)):
        pol = GaussianMLPPolicy(obs_dim, act_dim).to(dev)
        with torch.no_grad():
            pol.log_std.copy_(torch.linspace(-1.0, 0.5, act_dim))
            for m in pol.modules():
                if isinstance(m, torch.nn.Linear):
                    m.bias.uniform_(-0.5, 0.5)
        obs = torch.randn(n, obs_dim, device=dev) * 2.0
        noise = torch.randn(n, act_dim, device=dev)
        with torch.no_grad():
            mean32, v32 = pol(obs)
            mean16, v16 = bf16_reference(pol, obs)
        if True:
            fp = FusedPolicy(pol)
            a = torch.full((n, act_dim), float("nan"), device=dev)
            ac, lp, v = torch.full_like(a, float("nan")), torch.full((n,), float("nan"), device=dev), torch.full((n,), float("nan"), device=dev)
            fp.step(obs, noise, a, ac, lp, v)
            torch.cuda.synchronize()
            std = pol.log_std.detach().exp()
            mean = a - std * noise
            lp_ref = (-0.5 * noise ** 2 - pol.log_std.detach() - 0.9189385332046727).sum(-1)
            print("dims %d->%d: |mean - bf16 ref| %.3e  |mean - fp32| %.3e  |v - bf16 ref| %.3e  |v - fp32| %.3e  "
                  "|logp| %.2e  clip ok %s  nan %d" % (
                      obs_dim, act_dim, (mean - mean16).abs().max().item(), (mean - mean32).abs().max().item(),
                      (v - v16).abs().max().item(), (v - v32).abs().max().item(), (lp - lp_ref).abs().max().item(),
                      bool(torch.equal(ac, a.clamp(-1, 1))), int(torch.isnan(a).sum() + torch.isnan(v).sum())))
    # value-only call and a ragged batch
    fp = FusedPolicy(pol)
    m = 1000
    v = torch.zeros(m, device=dev)
    fp.step(obs[:m].contiguous(), value=v)
    torch.cuda.synchronize()
    print("value-only ragged n=%d: |v - bf16 ref| %.3e" % (m, (v - v16[:m]).abs().max().item()))
    # time
    pol = GaussianMLPPolicy().to(dev)
    fp = FusedPolicy(pol)
    obs = torch.randn(n, 70, device=dev)
    noise = torch.randn(n, 17, device=dev)
    a, ac, lp, v = torch.empty(n, 17, device=dev), torch.empty(n, 17, device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev)
    for _ in range(10):
        fp.step(obs, noise, a, ac, lp, v)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        fp.step(obs, noise, a, ac, lp, v)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 200
    flop = 2 * n * ((70 * 256 + 256 * 256) * 2 + 256 * 18)
    print("fused policy step N=%d: %.1f us (%.1f TFLOP/s useful)" % (n, us, flop / us * 1e-6))
    import ctypes as C
    from ilrl_b200 import _lib
    tl = torch.zeros(32, dtype=torch.int64, device=dev)
    _lib.lib().ilrl_debug_policy_timeline(C.c_void_p(tl.data_ptr()))
    fp.step(obs, noise, a, ac, lp, v)
    torch.cuda.synchronize()
    _lib.lib().ilrl_debug_policy_timeline(None)
    t = tl.cpu().numpy()
    names = {0: "start", 1: "issued loads, alloc", 2: "inputs landed", 3: "A0 built", 4: "pi: W1 landed", 5: "pi: L1 mma done",
             6: "pi: L1 epilogue (+L2 issue)", 7: "pi: L2 mma done", 8: "pi: L2 epilogue (+L3 issue)", 9: "pi: L3 mma done",
             10: "pi: outputs", 12: "vf: W1 landed", 13: "vf: L1 mma done", 14: "vf: L1 epilogue", 15: "vf: L2 mma done",
             16: "vf: L2 epilogue", 17: "vf: L3 mma done", 18: "vf: outputs"}
    sub = {20: "next W2a requested", 21: "noise landed", 22: "D3 loaded", 23: "sampled + staged", 24: "cta barrier", 10: "tiles stored"}
    prev = t[9]
    for i in (20, 21, 22, 23, 24, 10):
        print("    pi outputs: %-24s +%6d clk" % (sub[i], t[i] - prev))
        prev = t[i]
    prev = t[0]
    for i in sorted(names):
        print("  %-32s +%6d clk  (t = %6d)" % (names[i], t[i] - prev, t[i] - t[0]))
        prev = t[i]
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for _ in range(10):
            pol(obs)
        e0.record()
        for _ in range(200):
            pol(obs)
        e1.record()
        torch.cuda.synchronize()
    print("torch bf16 autocast forward only: %.1f us" % (e0.elapsed_time(e1) * 1e3 / 200))


if __name__ == "__main__":
    main()
