"""Diagnostic: find env steps in which the torso leaves the ground fast (|vz| > 8 m/s) on the CUDA path and replay
exactly those steps (same state, same action) in the fp64 oracle: is the event the model's, or the kernel's?"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import ilrl_b200
from ilrl_b200.batched_env import BatchedHumanoidEnv
from oracle import oracle as O

n = 32768
env = BatchedHumanoidEnv(n, "low", clips=["motion09_03"], seed=3, auto_reset=True)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(5)
events, zmax, count_fast, total = [], 0.0, 0, 0
for t in range(400):
    a = torch.rand(n, 17, device="cuda", generator=g) * 2 - 1
    ph0, ef0 = [x.clone() for x in env.get_state()]
    o, r, d, tm = env.step(a)
    ph1, ef1 = env.get_state()
    total += n
    fast = (ph1[:, 9].abs() > 8.0) & (ph0[:, 9].abs() < 4.0) & (d == 0)
    count_fast += int(fast.sum())
    zmax = max(zmax, float(ph1[:, 2].max()))
    if len(events) < 12 and bool(fast.any()):
        i = int(torch.nonzero(fast)[0])
        events.append((ph0[i].cpu().numpy(), ef0[i].cpu().numpy(), a[i].cpu().numpy(), ph1[i].cpu().numpy()))
print("env-steps %d, steps where |vz| jumps from < 4 to > 8 m/s: %d (%.2e), max torso z %.2f" % (total, count_fast, count_fast / total, zmax))
for k, (p0, e0, act, p1) in enumerate(events):
    v = O.OracleEnv("motion09_03", 0)
    v.reset(0, 0.0, 0)
    v.set(p0.astype(np.float64), e0.astype(np.float64))
    v.low_step(act.astype(np.float64))
    q1, _, _ = v.get()
    print("event %d: z0 %.3f vz0 %+.2f -> CUDA z %.3f vz %+.2f | oracle z %.3f vz %+.2f | max |dq| %.2e" % (
        k, p0[2], p0[9], p1[2], p1[9], q1[2], q1[9], np.abs(q1[13:30] - p1[13:30]).max()))
env.close()
