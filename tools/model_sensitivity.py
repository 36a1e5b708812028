#!/usr/bin/env python3
"""Sensitivity of the rollout statistics to the UNPINNED Bullet-side constants of the restated dynamics.

No reference test, fixture or notebook pins a post-step state and PyBullet is not installable here, so the dynamics of
oracle/ilrl_oracle.c (and of the CUDA product, which agrees with it) are "restated from the published algorithm".  This
tool varies each uncertain constant ONE AT A TIME in the oracle and reports what it does to the statistics the
reference does expose:
  * mean episode length of an untrained policy: ~18-20 env steps at t = 0 of the reference's own PPO / PG / A2C runs
    (REF img/Komparasi algoritma eps len.png; RLlib's initial Gaussian policy = N(0,1) actions clipped to [-1,1]);
  * plausibility indicators the soak tests bound: largest joint-limit overshoot, largest change of the torso's
    vertical velocity in one substep ("launch events"), largest torso height.
Runs the ORACLE only (test infrastructure), single-threaded: python tools/model_sensitivity.py [envs] [steps]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402

NJ = 17
# MJCF joint attributes of REF humanoid_symmetric_2.xml:4, 23-80 in pybullet `ordered_joints` order (default class:
# armature 1 / damping 1 where a joint does not override it; MuJoCo semantics - Bullet's MJCF importer is believed to
# ignore all three)
DAMPING = [5, 5, 5, 5, 5, 5, 1, 5, 5, 5, 1, 1, 1, 1, 1, 1, 1]
ARMATURE = [0.02, 0.02, 0.02, 0.01, 0.01, 0.01, 0.006, 0.01, 0.01, 0.01, 0.006, 0.0068, 0.0051, 0.0028, 0.0068, 0.0051, 0.0028]
STIFFNESS = [20, 10, 10, 10, 10, 20, 1, 10, 10, 20, 1, 1, 1, 0, 1, 1, 0]
BASE = dict(limit_erp=0.2, contact_erp=0.9, friction=1.6, max_coord_vel=100.0, link_damp_scale=1.0, solver_iters=5,
            limit_rows_always=0, damping=0.0, armature=0.0, stiffness=0.0)
VARIANTS = [
    ("baseline (= product constants)", {}),
    ("limit ERP 0.2 -> 0.9", dict(limit_erp=0.9)),
    ("contact ERP 0.9 -> 0.2", dict(contact_erp=0.2)),
    ("limit rows always (speculative when not violated)", dict(limit_rows_always=1)),
    ("MJCF joint damping on", dict(damping=1.0)),
    ("MJCF joint armature on", dict(armature=1.0)),
    ("MJCF damping + armature on", dict(damping=1.0, armature=1.0)),
    ("MJCF damping + armature + stiffness on (MuJoCo semantics)", dict(damping=1.0, armature=1.0, stiffness=1.0)),
    ("Bullet link velocity damping off", dict(link_damp_scale=0.0)),
    ("solver iterations 5 -> 10", dict(solver_iters=10)),
    ("solver iterations 5 -> 50", dict(solver_iters=50)),
    ("max coordinate velocity 100 -> 1000 rad/s", dict(max_coord_vel=1000.0)),
    ("friction 1.6 -> 0.8 (plane only)", dict(friction=0.8)),
    # the reference loads the robot with self_collision = True (REF humanoid.py:13); the CUDA product does not model it
    ("Bullet self-collision on (all geom pairs but ancestors)", dict(selfcol=1)),
]


BOTH_ACTION_MODES = [VARIANTS[k][0] for k in (0, 3, 4, 6, 7, -1)]


def run(params, n, steps, action_mode):
    L = O.lib()
    L.ilrl_oracle_rollout.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_long), C.POINTER(C.c_double)]
    L.ilrl_oracle_rollout.restype = C.c_long
    L.ilrl_oracle_set_params.argtypes = [C.POINTER(C.c_double)]
    L.ilrl_oracle_diag.argtypes = [C.POINTER(C.c_double), C.c_int]
    L.ilrl_oracle_rollout_action_mode.argtypes = [C.c_int]
    p = dict(BASE)
    p.update(params)
    L.ilrl_oracle_set_self_collision(int(p.pop("selfcol", 0)))
    flat = np.array([p["limit_erp"], p["contact_erp"], p["friction"], p["max_coord_vel"], p["link_damp_scale"], p["solver_iters"],
                     p["limit_rows_always"]] + [p["damping"] * d for d in DAMPING] + [p["armature"] * a for a in ARMATURE] +
                    [p["stiffness"] * k for k in STIFFNESS], dtype=np.float64)
    L.ilrl_oracle_set_params(flat.ctypes.data_as(C.POINTER(C.c_double)))
    L.ilrl_oracle_rollout_action_mode(action_mode)
    envs = [O.OracleEnv("motion09_03", 0) for _ in range(n)]
    for k, e in enumerate(envs):
        e.reset((7 * k) % 80, 0.0, (37 * k) % 360 - 180)
    arr = (C.c_void_p * n)(*[e.h for e in envs])
    L.ilrl_oracle_rollout(arr, n, 40, 1, None, None)   # reach the steady state of the reset-on-done process
    L.ilrl_oracle_diag(None, 1)
    eps, rs = C.c_long(0), C.c_double(0)
    done = L.ilrl_oracle_rollout(arr, n, steps, 2, C.byref(eps), C.byref(rs))
    d = np.zeros(8)
    L.ilrl_oracle_diag(d.ctypes.data_as(C.POINTER(C.c_double)), 1)
    L.ilrl_oracle_set_params(None)
    L.ilrl_oracle_rollout_action_mode(0)
    sc, ss = C.c_long(0), C.c_long(0)
    L.ilrl_oracle_self_collision_stats(C.byref(sc), C.byref(ss))
    L.ilrl_oracle_set_self_collision(0)
    if ss.value:
        print("  (self-contacts per substep %.3f)" % (sc.value / ss.value), file=sys.stderr)
    return dict(len=done / max(eps.value, 1), ret=rs.value / max(eps.value, 1), over_max=d[1], over_rate=d[2] / d[0],
                dvz_max=d[3], dvz_rate=d[4] / d[0], zmax=d[5])


def main(n=192, steps=250):
    print("| variant | actions | mean episode length | mean return | max limit overshoot (rad) | substeps > 0.5 rad over | "
          "max torso dv_z per substep (m/s) | substeps dv_z > 2 m/s | max torso z (m) |")
    print("|---|---|---|---|---|---|---|---|---|")
    for name, prm in VARIANTS:
        for am, an in ((0, "U(-1,1)"), (1, "clip N(0,1)")):
            if am == 1 and name not in BOTH_ACTION_MODES:
                continue
            r = run(prm, n, steps, am)
            print("| %s | %s | %.1f | %.2f | %.2f | %.1e | %.1f | %.1e | %.2f |" % (
                name, an, r["len"], r["ret"], r["over_max"], r["over_rate"], r["dvz_max"], r["dvz_rate"], r["zmax"]), flush=True)


if __name__ == "__main__":
    if os.environ.get("ILRL_SENS_ONLY"):   # e.g. ILRL_SENS_ONLY=0,13: a subset of the variants
        VARIANTS = [VARIANTS[int(k)] for k in os.environ["ILRL_SENS_ONLY"].split(",")]
    main(*[int(x) for x in sys.argv[1:]])
