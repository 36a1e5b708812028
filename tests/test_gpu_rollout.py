"""GPU tests of the on-device rollout path: the GAE kernel against a NumPy restatement of
RLlib's discounted-cumsum formulation, and the graph-captured collector's SampleBatch invariants."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import ilrl_b200  # noqa: E402
from ilrl_b200 import BatchedHumanoidEnv, GaussianMLPPolicy, RolloutCollector, gae  # noqa: E402


def _gae_numpy(r, v, d, gamma, lam):
    T, n = r.shape
    adv = np.zeros((T, n))
    a = np.zeros(n)
    for t in range(T - 1, -1, -1):
        nd = 1.0 - d[t]
        delta = r[t] + gamma * v[t + 1] * nd - v[t]
        a = delta + gamma * lam * nd * a
        adv[t] = a
    return adv, adv + v[:-1]


@pytest.mark.parametrize("T,n", [(8, 16384), (1, 5), (37, 1000)])
def test_gae_kernel_matches_numpy(T, n):
    rng = np.random.default_rng(T * 1000 + n)
    r = rng.normal(size=(T, n)).astype(np.float32)
    v = rng.normal(size=(T + 1, n)).astype(np.float32)
    d = (rng.uniform(size=(T, n)) < 0.05).astype(np.uint8)
    adv, vt = gae(torch.from_numpy(r).cuda(), torch.from_numpy(v).cuda(), torch.from_numpy(d).cuda(), 0.99, 0.9)
    want_adv, want_vt = _gae_numpy(r.astype(np.float64), v.astype(np.float64), d.astype(np.float64), 0.99, 0.9)
    np.testing.assert_allclose(adv.cpu().numpy(), want_adv, rtol=1e-5, atol=1e-5)   # fp32 recurrence, stated tolerance
    np.testing.assert_allclose(vt.cpu().numpy(), want_vt, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("use_graph", [False, True])
def test_collector_sample_batch_invariants(use_graph, fused):
    n, T = 2048, 8
    torch.manual_seed(0)
    env = BatchedHumanoidEnv(n, "low", clips=["motion09_03"], seed=3, auto_reset=True)
    col = RolloutCollector(env, GaussianMLPPolicy(), horizon=T, gamma=0.99, lam=0.9, seed=1, use_graph=use_graph, fused=fused)
    env.stats()
    total_done = 0
    for it in range(3):
        b = col.collect()
        torch.cuda.synchronize()
        assert b["obs"].shape == (T, n, 70) and b["actions"].shape == (T, n, 17) and b["dones"].dtype == torch.uint8
        for k, t in b.items():
            assert bool(torch.isfinite(t.float()).all()), k
        # the state a step returns is the state the next step acts on (auto-reset included)
        assert torch.equal(b["new_obs"][:-1], b["obs"][1:])
        np.testing.assert_allclose((b["value_targets"] - b["advantages"]).cpu().numpy(), b["vf_preds"].cpu().numpy(),
                                   rtol=1e-5, atol=1e-5)
        # log-probability of the stored action under the policy that produced it
        with torch.no_grad():
            mean, _ = col.policy(b["obs"][0])
            std = col.policy.log_std.exp()
            lp = (-0.5 * ((b["actions"][0] - mean) / std) ** 2 - col.policy.log_std - 0.9189385332046727).sum(-1)
        # fused: the kernel's mean is the bf16 model's (|delta mean| <= 3e-2, tests/test_gpu_policy.py), and the fp32
        # module's log-density of the same action differs by sum_j z_j delta_j
        np.testing.assert_allclose(b["action_logp"][0].cpu().numpy(), lp.cpu().numpy(), rtol=1e-4,
                                   atol=0.3 if fused else 1e-3)
        total_done += int(b["dones"].sum())
    st = env.stats().cpu().numpy()
    # the statistics kernel saw the same episodes and steps (+ the graph warm-up pass, which is rolled back in state
    # but not in the counters)
    extra = T * n if use_graph else 0
    assert st[3] in (3 * T * n, 3 * T * n + extra)
    assert st[0] >= total_done
    env.close()
