"""End-to-end env-steps/s through ilrl_step_host_async / ilrl_wait for several part counts (development aid)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import ilrl_b200
from ilrl_b200.batched_env import BatchedHumanoidEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = BatchedHumanoidEnv(n, "low", clips=["motion09_03"], auto_reset=True, seed=1234)
env.reset()
rng = np.random.default_rng(0)
NH = 16
host_act = [torch.from_numpy(rng.uniform(-1, 1, (n, 17)).astype(np.float32)).pin_memory().numpy() for _ in range(NH)]
obs_h = torch.zeros(n, 70).pin_memory().numpy(); rew_h = torch.zeros(n).pin_memory().numpy()
done_h = torch.zeros(n, dtype=torch.uint8).pin_memory().numpy()
K = 400
for _ in range(50):
    env.step_host(host_act[0], obs_h, rew_h, done_h)
t0 = time.perf_counter()
for k in range(K):
    env.step_host(host_act[k % NH], obs_h, rew_h, done_h)
dt = time.perf_counter() - t0
print("sync        %.1f us/step  %.1f M env-steps/s" % (dt / K * 1e6, n * K / dt / 1e6))
for P in (2, 3, 4, 6, 8):
    for rep in range(2):
        for p in range(P):
            env.step_host_async(p, P, host_act[0], obs_h, rew_h, done_h)
        t0 = time.perf_counter()
        for k in range(1, K):
            a = host_act[k % NH]
            for p in range(P):
                env.wait(p)
                env.step_host_async(p, P, a, obs_h, rew_h, done_h)
        for p in range(P):
            env.wait(p)
        dt = time.perf_counter() - t0
    print("async P=%d   %.1f us/step  %.1f M env-steps/s" % (P, dt / (K - 1) * 1e6, n * (K - 1) / dt / 1e6))
    for p in range(P):
        env.step_host_async(p, P, host_act[0], obs_h, rew_h, done_h)
    t0 = time.perf_counter()
    for k in range(1, K):
        a = host_act[k % NH]
        for p in range(P):
            env.step_host_async(p, P, a, obs_h, rew_h, done_h, wait_first=True)
    for p in range(P):
        env.wait(p)
    dt = time.perf_counter() - t0
    print("   combined wait+submit: %.1f us/step  %.1f M env-steps/s" % (dt / (K - 1) * 1e6, n * (K - 1) / dt / 1e6))
env.close()
