"""CPU twin of tools/policy_transfer.py on the fp64 ORACLE (test infrastructure), with the oracle's model knobs: which
of the UNPINNED Bullet-side constants lets the policies the reference trained in PyBullet walk?  
usage: python tools/policy_transfer_oracle.py [--run 6d114] [--clip motion09_03] [--skip 1] [--trials 2] [--steps 1000]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tools.model_sensitivity import ARMATURE, BASE, DAMPING, STIFFNESS  # noqa: E402


def set_params(**kw):
    L = O.lib()
    L.ilrl_oracle_set_params.argtypes = [C.POINTER(C.c_double)]
    if not kw:
        L.ilrl_oracle_set_params(None)
        return
    p = dict(BASE)
    p.update(kw)
    flat = np.array([p["limit_erp"], p["contact_erp"], p["friction"], p["max_coord_vel"], p["link_damp_scale"], p["solver_iters"],
                     p["limit_rows_always"]] + [p["damping"] * d for d in DAMPING] + [p["armature"] * a for a in ARMATURE] +
                    [p["stiffness"] * k for k in STIFFNESS], dtype=np.float64)
    L.ilrl_oracle_set_params(flat.ctypes.data_as(C.POINTER(C.c_double)))


def target_chain(deg, n=100, length=5.0):
    t, out = np.zeros(2), []
    for i in range(n):
        a = np.deg2rad(deg * i)
        t = t + length * np.array([-np.sin(a), np.cos(a)])
        out.append(t.copy())
    return np.array(out)


def run(run="6d114", clip="motion09_03", skip=1, degs=(0, 30, 90), trials=2, max_steps=1000, seed=0, stochastic=True,
        obs_fix=None):
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_policies.npz"))
    W = {k: z["%s/%s" % (run, k)].astype(np.float64) for k in ("fc_1/kernel", "fc_1/bias", "fc_2/kernel", "fc_2/bias",
                                                              "fc_out/kernel", "fc_out/bias", "log_std")}
    L = O.lib()
    L.ilrl_oracle_env_set_skip.argtypes = [C.c_void_p, C.c_int]
    rng = np.random.default_rng(seed)
    envs, chains = [], []
    for d in degs:
        for _ in range(trials):
            e = O.OracleEnv(clip, 0)
            L.ilrl_oracle_env_set_skip(e.h, skip)
            envs.append(e)
            chains.append(target_chain(d))
    n = len(envs)
    obs = np.stack([e.reset(0, 0.0, 90) for e in envs])
    idx = np.zeros(n, int)
    alive = np.ones(n, bool)
    steps = np.zeros(n)
    drift = np.zeros(n)
    od = W["fc_1/kernel"].shape[0]
    std = np.exp(W["log_std"])
    for t in range(max_steps):
        x = obs[:, :od] if obs_fix is None else obs_fix(obs, envs)
        h = np.tanh(x @ W["fc_1/kernel"] + W["fc_1/bias"])
        h = np.tanh(h @ W["fc_2/kernel"] + W["fc_2/bias"])
        a = h @ W["fc_out/kernel"] + W["fc_out/bias"]
        if stochastic:
            a = a + std * rng.standard_normal((n, 17))
        for i, e in enumerate(envs):
            if not alive[i]:
                continue
            o, r, d = e.low_step(a[i], 0)
            obs[i] = o
            p, w, terms = e.get()
            cur = chains[i][idx[i]]
            if abs(w[O.E_TARGET_X] - cur[0]) + abs(w[O.E_TARGET_Y] - cur[1]) > 1e-9:
                idx[i] = (idx[i] + 1) % len(chains[i])
                new = chains[i][idx[i]]
                rx, ry = w[O.E_ROBOT_X], w[O.E_ROBOT_Y]
                hd = np.arctan2(new[1] - ry, new[0] - rx)
                w[O.E_TARGET_X], w[O.E_TARGET_Y] = new
                w[O.E_LOW_TARGET_SCORE] = -np.hypot(new[0] - w[O.E_START_X], new[1] - w[O.E_START_Y])
                w[O.E_HLDEG] = hd
                w[O.E_WALK_X], w[O.E_WALK_Y] = rx + 10 * np.cos(hd), ry + 10 * np.sin(hd)
                e.set(p, w)
            ax, ay = w[O.E_START_X], w[O.E_START_Y]
            bx, by = w[O.E_TARGET_X] - ax, w[O.E_TARGET_Y] - ay
            px, py = w[O.E_ROBOT_X] - ax, w[O.E_ROBOT_Y] - ay
            tt = min(max((px * bx + py * by) / (bx * bx + by * by), 0.0), 1.0)
            drift[i] += np.hypot(px - tt * bx, py - tt * by)
            steps[i] += 1
            if terms[5] <= 0:
                alive[i] = False
        if not alive.any():
            break
    return steps.reshape(len(degs), trials), (drift / np.maximum(steps, 1)).reshape(len(degs), trials)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--run", default="6d114")
    ap.add_argument("--clip", default="motion09_03")
    ap.add_argument("--skip", type=int, default=1)
    ap.add_argument("--trials", type=int, default=2)
    ap.add_argument("--steps", type=int, default=1000)
    a = ap.parse_args()
    variants = [("baseline (product constants)", {}), ("MJCF damping on", dict(damping=1.0)), ("MJCF armature on", dict(armature=1.0)),
                ("damping + armature", dict(damping=1.0, armature=1.0)),
                ("damping + armature + stiffness", dict(damping=1.0, armature=1.0, stiffness=1.0)),
                ("limit rows always", dict(limit_rows_always=1)), ("friction 0.8", dict(friction=0.8)),
                ("contact ERP 0.2", dict(contact_erp=0.2)), ("solver iters 50", dict(solver_iters=50)),
                ("link damping off", dict(link_damp_scale=0.0))]
    for name, kw in variants:
        set_params(**kw)
        st, dr = run(a.run, a.clip, a.skip, trials=a.trials, max_steps=a.steps)
        print("%-36s survival per angle (0, 30, 90 deg): %s   mean %.0f   drift %.3f" % (name, np.round(st.mean(1)), st.mean(), dr.mean()), flush=True)
    set_params()
