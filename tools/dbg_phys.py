import sys, os
import numpy as np, torch
sys.path.insert(0, ".")
import ilrl_b200
from ilrl_b200.batched_env import BatchedHumanoidEnv
from oracle import oracle as O
m = O.load_model()
lo, hi = np.array(m["joint_lo"]), np.array(m["joint_hi"])
sc = np.array(m["sphere_c"]).reshape(-1, 3); sr = np.array(m["sphere_r"]); sb = m["sphere_body"]
z = dict(np.load("tests/golden/low_traj_motion09_03.npz"))
p0 = z["phys_before"].astype(np.float32).astype(np.float64)
n = len(p0)
tau = np.stack([O.action_to_torque(a) for a in z["action"]]).astype(np.float32).astype(np.float64)
nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 4
want = p0.copy()
for i in range(n):
    q = p0[i]
    for s in range(nsub): q = O.substep(q, tau[i], 0.0165 / 4)
    want[i] = q
env = BatchedHumanoidEnv(n, "low", auto_reset=False)
env.L.ilrl_debug_substeps(env.h, nsub)
env.set_state(p0.astype(np.float32), None)
env.physics_only(tau.astype(np.float32))
got = env.get_state()[0].cpu().numpy().astype(np.float64)
err = np.abs(got - want)
eq = err[:, 13:30].max(1); eqd = err[:, 30:47].max(1); ep = err[:, :3].max(1)
print("substeps", nsub, "q err pct [50,90,99,100]:", np.percentile(eq, [50, 90, 99, 100]))
print("qd err pct:", np.percentile(eqd, [50, 90, 99, 100]))
print("pos err pct:", np.percentile(ep, [50, 90, 99, 100]))
def info(p):
    bo, ao, br = O.fk(p)
    h = np.array([(bo[b] + br[b] @ c)[2] - r for b, c, r in zip(sb, sc, sr)])
    viol = np.where((p[13:30] <= lo) | (p[13:30] >= hi))[0]
    return "contacts(<.02): %s  viol joints: %s" % ([(int(i), round(float(h[i]), 4)) for i in np.where(h < 0.02)[0]], viol.tolist())
for i in np.argsort(-eqd)[:6]:
    print("rec", i, "eq %.2e eqd %.2e epos %.2e" % (eq[i], eqd[i], ep[i]), info(p0[i]))
    j = np.argmax(err[i, 30:47]); print("   worst qd joint", j, "got", got[i, 30 + j], "want", want[i, 30 + j], " base vel got", got[i,7:10], "want", want[i,7:10])
# classify error by presence of contacts / limits
nc = np.array([ (np.array([(O.fk(p)[0][b] + O.fk(p)[2][b] @ c)[2] - r for b, c, r in zip(sb, sc, sr)]) < 0.02).sum() for p in p0[:300]])
nl = np.array([((p[13:30] <= lo) | (p[13:30] >= hi)).sum() for p in p0[:300]])
for c in range(0, 4):
    for l in range(0, 4):
        sel = (nc == c) & (nl == l)
        if sel.sum(): print("ncontact", c, "nlimit", l, "count", sel.sum(), "median eqd %.2e max eqd %.2e" % (np.median(eqd[:300][sel]), eqd[:300][sel].max()))
