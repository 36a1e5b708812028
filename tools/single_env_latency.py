"""Latency of the reference-shaped N = 1 env classes (the path the reference's unchanged drivers take):
microseconds per LowLevelHumanoidEnv.step / HierarchicalHumanoidEnv.step, next to the reference's own Python floor of
595 us per step (its reward / obs Python alone, PyBullet excluded)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import ilrl_b200
from ilrl_b200 import LowLevelHumanoidEnv, HierarchicalHumanoidEnv

env = LowLevelHumanoidEnv(reference_name="motion09_03", seed=1)
env.reset()
rng = np.random.default_rng(0)
acts = rng.uniform(-1, 1, (2000, 17))
for i in range(100):
    _, _, d, _ = env.step(acts[i])
    if d:
        env.reset()
t0 = time.perf_counter(); n = 0; resets = 0
for i in range(2000):
    _, _, d, _ = env.step(acts[i]); n += 1
    if d:
        env.reset(); resets += 1
dt = time.perf_counter() - t0
print("LowLevelHumanoidEnv: %.1f us per step (%d steps, %d resets included)" % (dt / n * 1e6, n, resets))
env.close()
h = HierarchicalHumanoidEnv(seed=1)
h.reset()
t0 = time.perf_counter(); n = 0
obs = h.reset()
for i in range(2000):
    if "high_level_agent" in obs and len(obs) == 1:
        obs, r, d, _ = h.step({"high_level_agent": rng.uniform(-1, 1, 2)})
    else:
        obs, r, d, _ = h.step({"low_level_agent": acts[i]}); n += 1
    if d["__all__"]:
        obs = h.reset()
dt = time.perf_counter() - t0
print("HierarchicalHumanoidEnv: %.1f us per low-level step (high-level steps and resets included)" % (dt / max(n, 1) * 1e6))
h.close()
