"""Development aid: where the end-to-end step time of `ilrl_step_host` goes (kernel with zero-copy I/O vs launch + sync)."""
import sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, ".")
import ilrl_b200
from ilrl_b200.batched_env import BatchedHumanoidEnv
n = 4096
env = BatchedHumanoidEnv(n, "low", clips=["motion09_03"], seed=1234, auto_reset=True)
env.reset()
pin = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt).pin_memory().numpy()
acts = [np.random.default_rng(i).uniform(-1, 1, (n, 17)).astype(np.float32) for i in range(8)]
pacts = []
for a in acts:
    p = pin(n, 17); p[:] = a; pacts.append(p)
obs, rew, done = pin(n, 70), pin(n), pin(n, dt=torch.uint8)
for i in range(50): env.step_host(pacts[i % 8], obs, rew, done)
ms, cnt = C.c_float(), C.c_int64()
env.L.ilrl_kernel_timing(env.h, 1, C.byref(ms), C.byref(cnt))
K = 500
t0 = time.perf_counter()
for i in range(K): env.step_host(pacts[i % 8], obs, rew, done)
dt = time.perf_counter() - t0
env.L.ilrl_kernel_timing(env.h, 0, C.byref(ms), C.byref(cnt))
print("zero-copy e2e: wall %.1f us/step (timing events add overhead), kernel %.1f us/launch over %d launches" % (dt / K * 1e6, ms.value / max(cnt.value, 1) * 1e3, cnt.value))
t0 = time.perf_counter()
for i in range(K): env.step_host(pacts[i % 8], obs, rew, done)
dt = time.perf_counter() - t0
print("zero-copy e2e without timing events: wall %.1f us/step" % (dt / K * 1e6))
# device path for comparison
a_dev = [torch.from_numpy(a).cuda() for a in acts]
for i in range(50): env.step(a_dev[i % 8])
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(K): env.step(a_dev[i % 8])
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("device path: wall %.1f us/step (launches queued, one sync)" % (dt / K * 1e6))
t0 = time.perf_counter()
for i in range(K):
    env.step(a_dev[i % 8]); torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("device path with a sync per step: wall %.1f us/step" % (dt / K * 1e6))
env.close()
