"""Diagnostic: joint-limit overshoot statistics of a random-action rollout (development aid)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import ilrl_b200
from ilrl_b200 import BatchedHumanoidEnv
from oracle import oracle as O
n = 4096
env = BatchedHumanoidEnv(n, "low", clips=["motion09_03"], seed=12, auto_reset=True)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(9)
m = O.load_model(); lo, hi = np.array(m["joint_lo"]), np.array(m["joint_hi"]); names = m["joint_name"]
worst = np.zeros(17); hist = []
for t in range(200):
    a = torch.rand(n, 17, device="cuda", generator=g) * 2 - 1
    p0 = env.get_state()[0].cpu().numpy().astype(np.float64)
    env.step(a)
    p = env.get_state()[0].cpu().numpy().astype(np.float64)
    v = np.maximum(lo - p[:, 13:30], p[:, 13:30] - hi)
    worst = np.maximum(worst, v.max(0))
    i, j = np.unravel_index(np.argmax(v), v.shape)
    if v[i, j] > 0.8 and len(hist) < 3:
        hist.append((t, i, j, v[i, j]))
        print("step %d env %d joint %s overshoot %.3f: q %.3f -> %.3f  qd %.2f -> %.2f  lo %.3f hi %.3f  action(motor order) max %.2f torso z %.2f" % (
            t, i, names[j], v[i, j], p0[i, 13 + j], p[i, 13 + j], p0[i, 30 + j], p[i, 30 + j], lo[j], hi[j], float(a[i].abs().max()), p[i, 2]))
        # replay this env on the oracle from p0 with the same action
        tau = O.action_to_torque(a[i].cpu().numpy().astype(np.float64))
        po = O.physics_step(p0[i], tau)
        print("   oracle from the same state: q -> %.3f qd -> %.2f" % (po[13 + j], po[30 + j]))
print("worst overshoot per joint (rad):")
for j in range(17):
    print("  %-18s %.3f" % (names[j], worst[j]))
