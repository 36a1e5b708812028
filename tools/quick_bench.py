"""Quick device-side timing of the fused step kernel (development aid; bench.py is the contract)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import ilrl_b200
from ilrl_b200.batched_env import BatchedHumanoidEnv

for n in [int(x) for x in (sys.argv[1:] or ["4096", "32768", "131072"])]:
    env = BatchedHumanoidEnv(n, "low", clips=["motion09_03"], auto_reset=True, seed=1234)
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    acts = [torch.rand(n, 17, device="cuda", generator=g) * 2 - 1 for _ in range(8)]
    for i in range(20): env.step(acts[i % 8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 100
    e0.record()
    for i in range(K): env.step(acts[i % 8])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    st = env.stats().cpu().numpy()
    print("N=%d  %.3f ms/step  %.3e env-steps/s  episodes=%d mean_len=%.1f mean_ret=%.2f mean_rew=%.3f" % (
        n, ms, n / ms * 1e3, st[0], st[2] / max(st[0], 1), st[1] / max(st[0], 1), st[4] / max(st[3], 1)))
    env.close()
