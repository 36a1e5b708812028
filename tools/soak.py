import sys, numpy as np, torch
sys.path.insert(0, ".")
import ilrl_b200
from ilrl_b200.batched_env import BatchedHumanoidEnv
for mode, n, steps in (("low", 4099, 3000), ("hier", 16389, 1500), ("low", 65537, 600), ("hier", 777, 3000)):
    env = BatchedHumanoidEnv(n, mode, clips=["motion08_03", "motion09_03"], clip_of_env=np.arange(n, dtype=np.int32) % 2, seed=11, auto_reset=True)
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
        if t % 100 == 99:
            bad += int((~torch.isfinite(o)).sum()) + int((~torch.isfinite(r)).sum())
    ph, ef = env.get_state()
    st = env.stats().cpu().numpy()
    print(mode, n, "steps", steps, "nonfinite", bad, int((~torch.isfinite(ph)).sum()), "episodes", int(st[0]), "mean_len %.1f" % (st[2] / max(st[0], 1)),
          "z range %.2f..%.2f" % (float(ph[:, 2].min()) if ph.shape[1] > 2 else 0, float(ph[:, 2].max()) if ph.shape[1] > 2 else 0))
    env.close()
print("soak done")
