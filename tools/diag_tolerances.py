"""Observed worst cases behind the tolerance asserts of tests/test_gpu_parity.py and tests/test_gpu_api.py
(development aid: the asserts are set to ~2x these numbers)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import ilrl_b200
from ilrl_b200 import BatchedHumanoidEnv
from oracle import oracle as O
import test_gpu_parity as P

for seed in (6, 16, 26):
    rng = np.random.default_rng(seed)
    n = 512
    p0 = P._random_states(rng, n, False)
    tau = rng.uniform(-40, 40, (n, 17)) * (rng.uniform(size=(n, 1)) < 0.7)
    tau = tau.astype(np.float32).astype(np.float64)
    want = np.stack([O.physics_step(p0[i], tau[i]) for i in range(n)])
    env = BatchedHumanoidEnv(n, "low", auto_reset=False)
    env.set_state(p0.astype(np.float32), None)
    env.physics_only(tau.astype(np.float32))
    got = env.get_state()[0].cpu().numpy().astype(np.float64)
    t = P._phys_tol(True)
    eq, eqd, ep = np.abs(got[:, 13:30] - want[:, 13:30]).max(1), np.abs(got[:, 30:47] - want[:, 30:47]).max(1), np.abs(got[:, 0:3] - want[:, 0:3]).max(1)
    bad = (eq > t["q"]) | (eqd > t["qd"]) | (ep > t["pos"])
    print("seed %d: tol %s  outside: %d of %d; worst q %.2e qd %.2e pos %.2e; outliers:" % (seed, t, bad.sum(), n, eq.max(), eqd.max(), ep.max()),
          [(int(i), "%.1e" % eq[i], "%.1e" % eqd[i], "%.1e" % ep[i]) for i in np.nonzero(bad)[0]])
    print("   percentiles qd err: 50%% %.1e 90%% %.1e 99%% %.1e" % tuple(np.percentile(eqd, [50, 90, 99])))
    env.close()

z = P._load("low_traj_motion09_03.npz")
n = len(z["reward"])
env = BatchedHumanoidEnv(n, "low", clips=["motion09_03"], auto_reset=False)
e = P._pad_env(z["env_before"]); e[:, 1] = 0
env.set_state(z["phys_before"].astype(np.float32), e)
env.set_forced_target_deg(np.where(z["rand_deg"] == -999, P.INT32_MIN, z["rand_deg"]).astype(np.int64))
obs, rew, done, terms = env.step(z["action"].astype(np.float32))
phys = env.get_state()[0].cpu().numpy()
rew, done = rew.cpu().numpy(), done.cpu().numpy()
mism = np.nonzero(done.astype(bool) != z["done"].astype(bool))[0]
print("cfg1: done mismatches %d of %d; torso z of those (oracle): %s ; reward max err %.2e (excluding mismatches %.2e)" % (
    len(mism), n, [round(float(z["phys_after"][i, 2]), 5) for i in mism], np.abs(rew - z["reward"]).max(),
    np.abs(np.delete(rew - z["reward"], mism)).max()))
d = np.abs(phys.astype(np.float64) - z["phys_after"])
print("cfg1 phys err: q %.2e qd %.2e pos %.2e" % (d[:, 13:30].max(), d[:, 30:47].max(), d[:, 0:3].max()))
env.close()
