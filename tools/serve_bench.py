"""End-to-end step latency of the three host paths at one batch size (development aid; bench.py is the contract):
blocking ilrl_step_host, pipelined ilrl_step_host_async in 4 parts, persistent ilrl_serve_step."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import ilrl_b200  # noqa: F401,E402
from ilrl_b200.batched_env import BatchedHumanoidEnv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = 2000
pin = lambda shape, dt=torch.float32: torch.zeros(*shape, dtype=dt).pin_memory().numpy()  # noqa: E731
rng = np.random.default_rng(0)
acts = [pin((n, 17)) for _ in range(16)]
for a in acts:
    a[:] = rng.uniform(-1, 1, (n, 17))
o_h, r_h, d_h = pin((n, 70)), pin((n,)), pin((n,), torch.uint8)
env = BatchedHumanoidEnv(n, "low", clips=["motion09_03"], auto_reset=True, seed=1)
env.reset()
torch.cuda.synchronize()
for _ in range(100):
    env.step_host(acts[0], o_h, r_h, d_h)
t0 = time.perf_counter()
for k in range(K):
    env.step_host(acts[k % 16], o_h, r_h, d_h)
t = time.perf_counter() - t0
print("step_host        %.2f us/step  %.2f M env-steps/s" % (t / K * 1e6, n * K / t / 1e6))
P = 4
for p in range(P):
    env.step_host_async(p, P, acts[0], o_h, r_h, d_h)
t0 = time.perf_counter()
for k in range(K):
    for p in range(P):
        env.step_host_async(p, P, acts[k % 16], o_h, r_h, d_h, wait_first=True)
for p in range(P):
    env.wait(p)
t = time.perf_counter() - t0
print("async x4 parts   %.2f us/step  %.2f M env-steps/s" % (t / K * 1e6, n * K / t / 1e6))
env.serve_begin(o_h, r_h, d_h)
for k in range(100):
    env.serve_step(acts[k % 16])
t0 = time.perf_counter()
s = 0.0
for k in range(K):
    env.serve_step(acts[k % 16])
    s += r_h[0]
t = time.perf_counter() - t0
env.serve_end()
print("serve_step       %.2f us/step  %.2f M env-steps/s" % (t / K * 1e6, n * K / t / 1e6))
for P in (2, 4, 8):
    sl = [env.part_slice(p, P) for p in range(P)]
    env.serve_begin(o_h, r_h, d_h, nparts=P)
    for p in range(P):
        env.serve_post(acts[0], p)
    for k in range(100):
        for p in range(P):
            env.serve_wait(p); env.serve_post(acts[k % 16], p)
    t0 = time.perf_counter()
    for k in range(K):
        a = acts[k % 16]
        for p in range(P):
            env.serve_wait(p)
            s += r_h[sl[p].start]
            env.serve_post(a, p)
    for p in range(P):
        env.serve_wait(p)
    t = time.perf_counter() - t0
    env.serve_end()
    print("serve x%d parts   %.2f us/step  %.2f M env-steps/s" % (P, t / K * 1e6, n * K / t / 1e6))
env.close()
