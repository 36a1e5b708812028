"""Cost of the per-step grid barrier: K env steps as K launches (graph replay) against ONE ilrl_step_sequence launch.
Usage: python tools/seq_bench.py [num_envs] [K]"""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("imitation-learning-rl_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
env = pkg.BatchedHumanoidEnv(n, "low", clips=["motion09_03"], seed=3, auto_reset=True)
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = torch.rand(K, n, 17, device="cuda", generator=g) * 2 - 1
obs = torch.empty(K, n, 70, device="cuda"); rew = torch.empty(K, n, device="cuda")
done = torch.empty(K, n, dtype=torch.uint8, device="cuda")
def loop():
    for t in range(K): env.step_into(acts[t], obs[t], rew[t], done[t])
def seq():
    env.step_sequence(acts, obs, rew, done)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): loop(); seq()
    s.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s): loop()
    for name, f in (("graph of K launches", gr.replay), ("one sequence launch", seq), ("graph of K launches", gr.replay), ("one sequence launch", seq)):
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        R = 50
        e0.record(s)
        for _ in range(R): f()
        e1.record(s); s.synchronize()
        ms = e0.elapsed_time(e1) / (R * K)
        print(f"{name}: {ms*1e3:.2f} us per env step, {n/ms/1e3:.2f} M env-steps/s", flush=True)
