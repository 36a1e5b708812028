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
    t_chk, ep_chk = np.zeros(n, np.int64), np.zeros(n, np.int64)     # NumPy restatement of RLlib's t / eps_id columns
    for it in range(3):
        b = col.collect()
        torch.cuda.synchronize()
        assert b["obs"].shape == (T, n, 70) and b["actions"].shape == (T, n, 17) and b["dones"].dtype == torch.uint8
        d, tc, ec = b["dones"].cpu().numpy(), b["t"].cpu().numpy(), b["eps_id"].cpu().numpy()
        for k in range(T):
            np.testing.assert_array_equal(tc[k], t_chk)
            np.testing.assert_array_equal(ec[k], (np.arange(n, dtype=np.int64) << 32) | ep_chk)
            t_chk = np.where(d[k] != 0, 0, t_chk + 1)
            ep_chk = ep_chk + (d[k] != 0)
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
    # the statistics kernel saw exactly the same episodes and steps (no hidden warm-up pass: the first fragment of a
    # graphed collector runs eagerly and counts)
    assert st[3] == 3 * T * n
    assert st[0] == total_done
    env.close()


def test_graphed_and_eager_collectors_give_the_same_trajectories():
    """use_graph=True must not change what is collected: same seed -> bit-identical batches, fragment by fragment
    (round-1 advisor finding: the graph warm-up used to roll back the state but not the envs' draw counters)."""
    n, T = 1024, 8
    outs = []
    for use_graph in (False, True):
        torch.manual_seed(0)
        env = BatchedHumanoidEnv(n, "low", clips=["motion09_03"], seed=5, auto_reset=True)
        col = RolloutCollector(env, GaussianMLPPolicy(), horizon=T, gamma=0.99, lam=0.9, seed=1, use_graph=use_graph)
        frag = []
        for it in range(4):   # eager, capture + replay, replay, replay
            b = col.collect()
            torch.cuda.synchronize()
            frag.append({k: v.clone() for k, v in b.items()})
        outs.append(frag)
        env.close()
    for fa, fb in zip(*outs):
        for k in fa:
            assert torch.equal(fa[k], fb[k]), k


def _gae_decisions_numpy(r, f, v, gamma, lam):
    """Forward-looking restatement: for each decision (bit2) find the next readout that reports an outcome (bit1),
    then A = delta + gamma lam A(next decision), resolved recursively from the end of the fragment."""
    T, n = r.shape[0] - 1, r.shape[1]
    adv, ret, valid = np.zeros((T, n)), np.zeros((T, n)), np.zeros((T, n), np.uint8)
    for i in range(n):
        memo = {}

        def A(t):   # advantage of the decision taken at tick t (None: not a decision with a known outcome)
            if t >= T or not (f[t, i] & 4):
                return None
            if t in memo:
                return memo[t]
            tau = next((k for k in range(t + 1, T + 1) if f[k, i] & 2), None)
            if tau is None:
                memo[t] = None
                return None
            ended = bool(f[tau, i] & 1)
            delta = r[tau, i] + (0.0 if ended else gamma * v[tau, i]) - v[t, i]
            nxt = None if ended else A(tau)
            memo[t] = delta + (gamma * lam * nxt if nxt is not None else 0.0)
            return memo[t]
        for t in range(T):
            if f[t, i] & 4:
                a = A(t)
                if a is None:
                    ret[t, i] = v[t, i]
                else:
                    adv[t, i], ret[t, i], valid[t, i] = a, a + v[t, i], 1
    return adv, ret, valid


@pytest.mark.parametrize("T,n", [(10, 4096), (1, 7), (23, 300)])
def test_gae_decisions_kernel_matches_numpy(T, n):
    rng = np.random.default_rng(T * 1000 + n)
    r = rng.normal(size=(T + 1, n)).astype(np.float32)
    v = rng.normal(size=(T + 1, n)).astype(np.float32)
    # arbitrary flag patterns, including outcome-without-decision, decision-without-outcome and endings
    f = (rng.integers(0, 2, (T + 1, n)) * 4 * (rng.random((T + 1, n)) < 0.4)
         + 2 * (rng.random((T + 1, n)) < 0.35) + 1 * (rng.random((T + 1, n)) < 0.15)).astype(np.uint8)
    adv, ret, valid = ilrl_b200.gae_decisions(torch.from_numpy(r).cuda(), torch.from_numpy(f).cuda(),
                                              torch.from_numpy(v).cuda(), 0.99, 0.9)
    want_adv, want_ret, want_valid = _gae_decisions_numpy(r.astype(np.float64), f, v.astype(np.float64), 0.99, 0.9)
    assert np.array_equal(valid.cpu().numpy(), want_valid)
    np.testing.assert_allclose(adv.cpu().numpy(), want_adv, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ret.cpu().numpy(), want_ret, rtol=1e-5, atol=1e-5)


def test_hier_collector_fragments():
    from ilrl_b200 import HierRolloutCollector
    n, T = 2048, 10

    def make(use_graph):
        torch.manual_seed(0)
        env = BatchedHumanoidEnv(n, "hier", clips=["motion08_03", "motion09_03"], clip_of_env=np.ones(n, np.int32),
                                 seed=5, auto_reset=True)
        hp, lp = GaussianMLPPolicy(44, 2), GaussianMLPPolicy(70, 17)
        return env, HierRolloutCollector(env, hp, lp, horizon=T, gamma=0.99, lam=0.9, seed=2, use_graph=use_graph)

    env_e, eager = make(False)
    env_g, graph = make(True)
    total_done = 0
    for it in range(3):
        be, bg = eager.collect(), graph.collect()
        torch.cuda.synchronize()
        for agent in ("low", "high"):      # the CUDA graph replays exactly what the eager loop does
            for k in be[agent]:
                assert torch.equal(be[agent][k], bg[agent][k]), (it, agent, k)
        lo, hi = bg["low"], bg["high"]
        assert lo["obs"].shape == (T, n, 70) and hi["obs"].shape == (T, n, 44) and hi["actions"].shape == (T, n, 2)
        for agent in ("low", "high"):
            for k, t in bg[agent].items():
                assert bool(torch.isfinite(t.float()).all()), (agent, k)
        flags = hi["flags"].cpu().numpy()
        decided, valid = hi["decided"].cpu().numpy(), hi["valid"].cpu().numpy().astype(bool)
        dones = lo["dones"].cpu().numpy().astype(bool)
        assert np.array_equal(decided, (flags[:T] & 4) != 0)
        assert not (valid & ~decided).any()
        if it == 0:
            assert decided[0].all()        # after reset every env waits for its first heading
        # the protocol of REF hier_env.py:583-642: an outcome is reported step_per_level = 5 low-level steps after the
        # decision, or at the end of the episode if that comes first; an episode end is always an outcome; after an
        # outcome the env waits for the next heading (auto-reset included)
        assert ((flags & 1) <= ((flags >> 1) & 1)).all()
        assert (((flags >> 1) & 1) <= ((flags >> 2) & 1))[1:].all()
        assert np.array_equal((flags[1:] & 1) != 0, dones)
        for i in range(0, n, 37):
            last = None
            for t in range(T + 1):
                if last is not None and flags[t, i] & 2:
                    assert t - last == 5 or (flags[t, i] & 1 and t - last <= 5), (i, t, last)
                if t < T and decided[t, i]:
                    last = t
        # low level: the obs a step returns is the obs the next step acts on, except where a new heading was taken in
        # between (high_level_step returns a refreshed low-level obs)
        same = torch.isclose(lo["new_obs"][:-1], lo["obs"][1:], rtol=0, atol=0).all(-1).cpu().numpy()
        assert same[~decided[1:]].all()
        np.testing.assert_allclose((lo["value_targets"] - lo["advantages"]).cpu().numpy(), lo["vf_preds"].cpu().numpy(),
                                   rtol=1e-5, atol=1e-5)
        v = torch.from_numpy(valid).cuda()
        np.testing.assert_allclose((hi["value_targets"] - hi["advantages"])[v].cpu().numpy(), hi["vf_preds"][v].cpu().numpy(),
                                   rtol=1e-5, atol=1e-5)
        # decisions whose outcome is inside the fragment: all but the last one of each env, at most
        assert valid.sum() >= decided.sum() - n and valid.sum() > 0
        # log-density of the stored high-level action under the fp32 module (bf16 kernel mean: |delta| <= 3e-2)
        with torch.no_grad():
            mean, _ = graph.high_policy(hi["obs"][0])
            ls = graph.high_policy.log_std
            lp = (-0.5 * ((hi["actions"][0] - mean) / ls.exp()) ** 2 - ls - 0.9189385332046727).sum(-1)
        np.testing.assert_allclose(hi["action_logp"][0].cpu().numpy(), lp.cpu().numpy(), rtol=1e-4, atol=0.3)
        total_done += int(dones.sum())
    st = env_g.stats().cpu().numpy()
    assert st[3] == 3 * T * n and st[0] == total_done
    env_e.close()
    env_g.close()
