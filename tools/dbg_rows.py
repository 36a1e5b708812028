import sys
import numpy as np, torch, ctypes as C
sys.path.insert(0, ".")
import ilrl_b200
from ilrl_b200.batched_env import BatchedHumanoidEnv
from oracle import oracle as O
z = dict(np.load("tests/golden/low_traj_motion09_03.npz"))
recs = [714, 900, 100]
p0 = z["phys_before"][recs].astype(np.float32)
tau = np.stack([O.action_to_torque(a) for a in z["action"][recs]]).astype(np.float32)
for i, rec in enumerate(recs):
    env = BatchedHumanoidEnv(len(recs), "low", auto_reset=False)
    env.L.ilrl_debug_substeps(env.h, 1)
    dbg = torch.zeros(400, device="cuda")
    env.L.ilrl_debug_dump(env.h, C.c_void_p(dbg.data_ptr()), i)
    env.set_state(p0, None)
    env.physics_only(tau)
    d = dbg.cpu().numpy()
    nl, nc = int(d[0]), int(d[1])
    print("rec", rec, "nlim", nl, "ncon", nc)
    for r in range(nl + 3 * nc):
        print("   row", r, "link %d dir %g rhs %.5g dinv %.5g lam %.5g" % tuple(d[2 + 5 * r: 7 + 5 * r]))
    print("   contacts idx", d[230:230+nc], "dist", d[240:240+nc], "recomputed", d[250:250+nc], "act", int(d[259]))
    print("   sd", d[260:289].round(4))
    print("   nu*", d[300:323].round(3)); print("   dv ", d[330:353].round(3)); print("   resp0", d[360:383].round(4))
    env.close()
