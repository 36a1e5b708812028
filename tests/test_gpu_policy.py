"""The fused tcgen05 policy / value kernel (csrc/ilrl_policy.cu) against the torch module it replaces.

Arithmetic contract of the kernel: bf16 operands and hidden activations, fp32 accumulation, fp32 biases, tanh.approx.
The oracle here is plain torch: (a) an emulation of exactly that contract (tolerance 5e-3 absolute on O(1) outputs:
accumulation order and the 2^-11 tanh approximation), (b) the fp32 module (tolerance 3e-2: bf16 rounding)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

import ilrl_b200  # noqa: E402,F401
from ilrl_b200 import GaussianMLPPolicy, _lib  # noqa: E402
from ilrl_b200.rollout import FusedPolicy  # noqa: E402


def bf16_reference(policy, obs):
    def run(seq):
        x = obs.to(torch.bfloat16).float()
        lins = [m for m in seq if isinstance(m, torch.nn.Linear)]
        for i, m in enumerate(lins):
            x = x @ m.weight.to(torch.bfloat16).float().t() + m.bias.float()
            if i < 2:
                x = torch.tanh(x).to(torch.bfloat16).float()
        return x
    with torch.no_grad():
        return run(policy.pi), run(policy.vf).squeeze(-1)


def make_policy(obs_dim, act_dim, seed=0):
    torch.manual_seed(seed)
    pol = GaussianMLPPolicy(obs_dim, act_dim).cuda()
    with torch.no_grad():
        pol.log_std.copy_(torch.linspace(-1.0, 0.5, act_dim))
        for m in pol.modules():
            if isinstance(m, torch.nn.Linear):
                m.bias.uniform_(-0.5, 0.5)
    return pol


@pytest.mark.parametrize("dims", [(70, 17), (44, 2)])   # the reference's low-level and high-level models
@pytest.mark.parametrize("n", [1, 127, 128, 1000, 16384])
def test_fused_policy_matches_torch(dims, n):
    obs_dim, act_dim = dims
    pol = make_policy(obs_dim, act_dim)
    fp = FusedPolicy(pol)
    g = torch.Generator(device="cuda").manual_seed(n)
    obs = torch.randn(n, obs_dim, device="cuda", generator=g) * 2.0
    noise = torch.randn(n, act_dim, device="cuda", generator=g)
    nan = float("nan")
    a, ac = torch.full((n, act_dim), nan, device="cuda"), torch.full((n, act_dim), nan, device="cuda")
    lp, v = torch.full((n,), nan, device="cuda"), torch.full((n,), nan, device="cuda")
    fp.step(obs, noise, a, ac, lp, v)
    torch.cuda.synchronize()
    mean16, v16 = bf16_reference(pol, obs)
    with torch.no_grad():
        mean32, v32 = pol(obs)
    std = pol.log_std.detach().exp()
    assert bool(torch.isfinite(a).all()) and bool(torch.isfinite(v).all())
    mean = a - std * noise
    assert (mean - mean16).abs().max().item() < 5e-3 and (v - v16).abs().max().item() < 5e-3
    assert (mean - mean32).abs().max().item() < 3e-2 and (v - v32).abs().max().item() < 3e-2
    assert torch.equal(ac, a.clamp(-1.0, 1.0))
    lp_ref = (-0.5 * noise ** 2 - pol.log_std.detach() - 0.9189385332046727).sum(-1)
    torch.testing.assert_close(lp, lp_ref, rtol=1e-5, atol=1e-4)


def test_fused_policy_partial_outputs_and_repack():
    n = 300
    pol = make_policy(70, 17, seed=1)
    fp = FusedPolicy(pol)
    obs = torch.randn(n, 70, device="cuda")
    # value only (the bootstrap call of a rollout): policy outputs are not touched
    v = torch.zeros(n, device="cuda")
    fp.step(obs, value=v)
    _, v16 = bf16_reference(pol, obs)
    assert (v - v16).abs().max().item() < 5e-3
    # deterministic action, no value net
    a = torch.zeros(n, 17, device="cuda")
    fp.step(obs, action=a)
    mean16, _ = bf16_reference(pol, obs)
    assert (a - mean16).abs().max().item() < 5e-3
    # parameters change -> repack -> new outputs
    with torch.no_grad():
        for m in pol.modules():
            if isinstance(m, torch.nn.Linear):
                m.weight.mul_(0.5)
    a2 = torch.zeros(n, 17, device="cuda")
    fp.step(obs, action=a2)
    torch.cuda.synchronize()
    assert torch.equal(a2, a)   # still the packed copy
    fp.repack()
    fp.step(obs, action=a2)
    mean16b, _ = bf16_reference(pol, obs)
    assert (a2 - mean16b).abs().max().item() < 5e-3 and (a2 - a).abs().max().item() > 1e-2


def test_policy_argument_errors():
    L = _lib.lib()
    blob = torch.zeros(int(L.ilrl_policy_blob_bytes()), dtype=torch.uint8, device="cuda")
    obs, out = torch.zeros(8, 70, device="cuda"), torch.zeros(8, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    assert L.ilrl_policy_step(p(blob), p(obs), None, None, None, None, None, 70, 17, 8, None) == -1   # nothing to produce
    assert L.ilrl_policy_step(p(blob), p(obs), None, None, None, None, p(out), 81, 17, 8, None) == -1  # obs too wide
    assert L.ilrl_policy_step(p(blob), p(obs), None, None, None, None, p(out), 70, 33, 8, None) == -1
    assert L.ilrl_policy_step(p(blob), p(obs), None, None, None, None, p(out), 70, 17, 0, None) == -1
    assert L.ilrl_policy_step(None, p(obs), None, None, None, None, p(out), 70, 17, 8, None) == -1
    assert L.ilrl_policy_step(p(blob), p(obs), None, None, None, None, p(out), 70, 17, 8, None) == 0
    torch.cuda.synchronize()
