"""Soak of the f4 instantiations of the step kernel (heightfield terrain, self-collision, both) under bang-bang and
missing-action inputs with auto-reset: non-finite counts, episode statistics, height range, worst limit overshoot."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import ilrl_b200
from ilrl_b200.batched_env import BatchedHumanoidEnv
tr = np.random.default_rng(99)
H = np.repeat(np.repeat(tr.uniform(0, 0.5, (128, 128)), 2, axis=0), 2, axis=1)
H[126:130, 126:130] = 0.0
for mode, n, steps, terrain, selfc in (("low", 4099, 3000, True, False), ("low", 4099, 3000, False, True), ("low", 16389, 1500, True, True),
                                       ("hier", 4099, 2000, False, True), ("low", 65537, 400, True, True)):
    env = BatchedHumanoidEnv(n, mode, clips=["motion08_03", "motion09_03"], clip_of_env=np.arange(n, dtype=np.int32) % 2, seed=11, auto_reset=True,
                             self_collision=selfc)
    if terrain:
        env.set_heightfield(H.reshape(-1))
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    bad = 0
    for t in range(steps):
        a = (torch.rand(n, 17, device="cuda", generator=g) * 2 - 1) * (4.0 if t % 7 == 0 else 1.0)
        if t % 11 == 0:
            a[::5, 0] = float("nan")          # rows without an action in this call
        if mode == "hier":
            env.high_step(torch.rand(n, 2, device="cuda", generator=g) * 2 - 1)
        o, r, d, tm = env.step(a)
        if t % 50 == 49:
            bad += int((~torch.isfinite(o)).sum()) + int((~torch.isfinite(r)).sum())
    ph, ef = env.get_state()
    st = env.stats().cpu().numpy()
    print(mode, n, "terrain" if terrain else "flat", "selfcol" if selfc else "-", "steps", steps, "nonfinite", bad, int((~torch.isfinite(ph)).sum()),
          "episodes", int(st[0]), "mean_len %.1f" % (st[2] / max(st[0], 1)), "z range %.2f..%.2f" % (float(ph[:, 2].min()), float(ph[:, 2].max())), flush=True)
    env.close()
print("soak done")
