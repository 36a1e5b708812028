"""Do the two instantiations of the substep (compile-time full shuffle masks / run-time masks) give bit-identical
states?  (development aid)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
import ilrl_b200
from ilrl_b200.batched_env import BatchedHumanoidEnv

n = 256
g = torch.Generator(device="cuda"); g.manual_seed(7)
acts = [torch.rand(n, 17, device="cuda", generator=g) * 2 - 1 for _ in range(30)]
res = {}
for flag in ("0", "1"):
    os.environ["ILRL_NO_FULLWARP"] = flag
    env = BatchedHumanoidEnv(n, "low", clips=["motion09_03"], seed=21, auto_reset=True)
    env.reset()
    out = []
    for a in acts:
        env.step(a)
        p, e = env.get_state()
        out.append((p.cpu().numpy().copy(), e.cpu().numpy().copy()))
    res[flag] = out
    env.close()
for t in range(len(acts)):
    dp = np.abs(res["0"][t][0] - res["1"][t][0])
    if dp.max() > 0:
        i, w = np.unravel_index(dp.argmax(), dp.shape)
        print("step", t, "first difference: max", dp.max(), "env", i, "word", w, "envs differing", int((dp.max(1) > 0).sum()),
              "words", sorted(set(np.nonzero(dp)[1].tolist()))[:20])
        # physics-only replay of that env from the previous state, one substep at a time
        break
else:
    print("bit-identical over", len(acts), "steps")
