"""Throughput of on-device rollout collection (BASELINE cfg 5 shape: 16384 envs x 8 steps per GPU per iteration).
Development / reporting aid; bench.py remains the contract.  usage: python tools/rollout_bench.py [N] [T] [iters]"""
import sys, time
import torch
sys.path.insert(0, ".")
import ilrl_b200
from ilrl_b200 import BatchedHumanoidEnv, GaussianMLPPolicy, RolloutCollector

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
for use_graph, amp, fused in ((False, None, False), (True, None, False), (True, torch.bfloat16, False), (False, None, True), (True, None, True)):
    env = BatchedHumanoidEnv(n, "low", clips=["motion09_03"], seed=1, auto_reset=True)
    col = RolloutCollector(env, GaussianMLPPolicy(), horizon=T, use_graph=use_graph, autocast_dtype=amp, fused=fused)
    for _ in range(5):
        col.collect()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        b = col.collect()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("rollout N=%d T=%d graph=%s policy=%s: %.3f ms/iteration, %.3e env-steps/s (policy + env + GAE on device)" % (
        n, T, use_graph, "fused tcgen05 kernel" if fused else "torch fp32" if amp is None else "torch bf16 autocast", ms, n * T / ms * 1e3))
    env.close()
