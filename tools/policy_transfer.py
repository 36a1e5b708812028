"""Policy-transfer check of the dynamics: run the low-level policies the REFERENCE trained in
PyBullet (weights from its Log/ checkpoints, tests/golden/ref_policies.npz) in this repo's CUDA env under the protocol of
the reference's own evaluation script (REF env_check.py: startFrame 0, a chain of targets 5 m apart turning by a fixed
angle, stochastic policy, episode over when the robot is no longer alive or after 3000 steps) and print survival time
and drift next to what the reference logged for the same checkpoint in PyBullet (Log/data_*.json).
usage: python tools/policy_transfer.py [--run 6d114] [--clip motion09_03] [--skip 1] [--trials 10] [--json out.json]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ilrl_b200  # noqa: F401,E402
from ilrl_b200 import batched_env as B  # noqa: E402
from ilrl_b200.batched_env import BatchedHumanoidEnv  # noqa: E402


def load_policy(run, dev):
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_policies.npz"))
    g = lambda k: torch.tensor(z["%s/%s" % (run, k)], device=dev)  # noqa: E731
    P = dict(w1=g("fc_1/kernel"), b1=g("fc_1/bias"), w2=g("fc_2/kernel"), b2=g("fc_2/bias"), w3=g("fc_out/kernel"),
             b3=g("fc_out/bias"), log_std=g("log_std"))
    ref = dict(deg=z["%s/eval_deg" % run], timestep=z["%s/eval_timestep" % run], drift=z["%s/eval_drift" % run])
    return P, ref


def target_chain(deg, n=100, length=5.0):
    """REF env_check.py:100-108"""
    t, out = np.zeros(2), []
    for i in range(n):
        a = np.deg2rad(deg * i)
        t = t + length * np.array([-np.sin(a), np.cos(a)])   # Rz(a) applied to (0, 1, 0)
        out.append(t.copy())
    return np.array(out)


def run_transfer(run="6d114", clip="motion09_03", skip=1, trials=10, degs=None, max_steps=3000, seed=0, stochastic=True,
                 device=0, obs_dim=None):
    dev = torch.device("cuda", device)
    P, ref = load_policy(run, dev)
    degs = list(ref["deg"]) if degs is None else list(degs)
    n = len(degs) * trials
    chains = torch.tensor(np.stack([target_chain(d) for d in degs for _ in range(trials)]), device=dev, dtype=torch.float32)
    env = BatchedHumanoidEnv(n, "low", clips=[clip], device=device, seed=seed, auto_reset=False, max_timestep=10 ** 8,
                             skip_frame=skip)
    obs = env.reset(start_frame=np.zeros(n, np.int32), reset_yaw_deg=np.zeros(n, np.float32),
                    target_xy=chains[:, 0].contiguous()).clone()
    od = P["w1"].shape[0] if obs_dim is None else obs_dim
    g = torch.Generator(device=dev).manual_seed(seed)
    idx = torch.zeros(n, dtype=torch.long, device=dev)       # predefinedTargetIndex
    ar = torch.arange(n, device=dev)
    alive = torch.ones(n, dtype=torch.bool, device=dev)
    steps = torch.zeros(n, device=dev)
    drift_sum = torch.zeros(n, device=dev)
    std = P["log_std"].exp()
    for t in range(max_steps):
        h = torch.tanh(obs[:, :od] @ P["w1"] + P["b1"])
        h = torch.tanh(h @ P["w2"] + P["b2"])
        a = h @ P["w3"] + P["b3"]
        if stochastic:
            a = a + std * torch.randn(n, 17, device=dev, generator=g)
        obs, rew, done, terms = env.step(a.contiguous())
        phys, envf = env.get_state()
        cur = chains[ar, idx]
        switched = ((envf[:, B.E_TARGET_X] - cur[:, 0]).abs() + (envf[:, B.E_TARGET_Y] - cur[:, 1]).abs()) > 1e-6
        if bool(switched.any()):                              # REF low_level_env.py:419-429 with usePredefinedTarget
            idx = torch.where(switched, (idx + 1) % chains.shape[1], idx)
            new = chains[ar, idx]
            e = envf.clone()
            rx, ry = e[:, B.E_ROBOT_X], e[:, B.E_ROBOT_Y]
            hd = torch.atan2(new[:, 1] - ry, new[:, 0] - rx)
            score = -torch.hypot(new[:, 0] - e[:, B.E_START_X], new[:, 1] - e[:, B.E_START_Y])
            for col, val in ((B.E_TARGET_X, new[:, 0]), (B.E_TARGET_Y, new[:, 1]), (B.E_LOW_TARGET_SCORE, score),
                             (B.E_HLDEG, hd), (B.E_WALK_X, rx + 10 * torch.cos(hd)), (B.E_WALK_Y, ry + 10 * torch.sin(hd))):
                e[:, col] = torch.where(switched, val, e[:, col])
            env.set_state(None, e)
            envf = e
        # drift: distance of robot_pos to the segment starting_robot_pos -> target (REF env_check.py:44-46)
        ax, ay = envf[:, B.E_START_X], envf[:, B.E_START_Y]
        bx, by = envf[:, B.E_TARGET_X] - ax, envf[:, B.E_TARGET_Y] - ay
        px, py = envf[:, B.E_ROBOT_X] - ax, envf[:, B.E_ROBOT_Y] - ay
        tt = ((px * bx + py * by) / (bx * bx + by * by)).clamp(0, 1)
        d = torch.hypot(px - tt * bx, py - tt * by)
        drift_sum += torch.where(alive, d, torch.zeros_like(d))
        steps += alive.float()
        alive = alive & (terms[:, 5] > 0)                     # `debug=True`: only falling ends the episode
        if not bool(alive.any()):
            break
    env.close()
    steps = steps.cpu().numpy().reshape(len(degs), trials)
    drift = (drift_sum / steps.new_tensor(1).clamp(min=1) if False else drift_sum).cpu().numpy().reshape(len(degs), trials) / np.maximum(steps, 1)
    return dict(deg=degs, timestep=steps, drift=drift, ref=ref)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--run", default="6d114")
    ap.add_argument("--clip", default="motion09_03")
    ap.add_argument("--skip", type=int, default=1)
    ap.add_argument("--trials", type=int, default=10)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--deterministic", action="store_true")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    r = run_transfer(a.run, a.clip, a.skip, a.trials, max_steps=a.steps, stochastic=not a.deterministic)
    ref = r["ref"]
    print("run %s clip %s skipFrame %d %s policy, %d trials per angle" % (a.run, a.clip, a.skip,
                                                                         "deterministic" if a.deterministic else "stochastic", a.trials))
    print(" deg | survival steps: this env (mean, n>=%d) | PyBullet (mean, n>=3000 of 10) | drift: this env | PyBullet" % a.steps)
    for k, g in enumerate(r["deg"]):
        j = list(ref["deg"]).index(g)
        print("%4d | %7.0f  %2d/%d | %7.0f  %2d/10 | %.3f | %.3f" % (g, r["timestep"][k].mean(), (r["timestep"][k] >= a.steps).sum(),
                                                                   a.trials, ref["timestep"][j].mean(), (ref["timestep"][j] >= 3000).sum(),
                                                                   r["drift"][k].mean(), ref["drift"][j].mean()))
    print("all angles: survival %.0f (PyBullet %.0f)   drift %.3f (PyBullet %.3f)" % (
        r["timestep"].mean(), ref["timestep"].mean(), r["drift"].mean(), ref["drift"].mean()))
    if a.json:
        json.dump(dict(run=a.run, clip=a.clip, skip=a.skip, deg=[int(x) for x in r["deg"]], timestep=r["timestep"].tolist(),
                       drift=r["drift"].tolist(), ref_timestep=ref["timestep"].tolist(), ref_drift=ref["drift"].tolist()),
                  open(a.json, "w"))


if __name__ == "__main__":
    main()
