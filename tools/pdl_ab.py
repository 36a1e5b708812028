"""Programmatic dependent launch on / off for one configuration (graph replay of 20 steps).
Usage: ILRL_PDL=0|1 python tools/pdl_ab.py num_envs [low|hier] [terrain|selfcol]"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("imitation-learning-rl_b200")
n = int(sys.argv[1]); mode = sys.argv[2] if len(sys.argv) > 2 else "low"; opt = sys.argv[3] if len(sys.argv) > 3 else ""
env = pkg.BatchedHumanoidEnv(n, mode, clips=["motion08_03", "motion09_03"], clip_of_env=np.ones(n, np.int32), seed=3, auto_reset=True,
                             self_collision=(opt == "selfcol"))
if opt == "terrain":
    tr = np.random.default_rng(99)
    h = np.repeat(np.repeat(tr.uniform(0, 0.5, (128, 128)), 2, axis=0), 2, axis=1); h[126:130, 126:130] = 0.0
    env.set_heightfield(h.reshape(-1))
env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
K = 20
acts = torch.rand(K, n, 17, device="cuda", generator=g) * 2 - 1
hacts = torch.rand(K, n, 2, device="cuda", generator=g) * 2 - 1
def loop():
    for t in range(K):
        if mode == "hier": env.high_step(hacts[t])
        env.step(acts[t])
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): loop()
    s.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s): loop()
    res = []
    for rep in range(3):
        for _ in range(3): gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        R = 30
        e0.record(s)
        for _ in range(R): gr.replay()
        e1.record(s); s.synchronize()
        res.append(e0.elapsed_time(e1) / (R * K) * 1e3)
print("PDL=%s n=%d %s %s: %.2f us per step (%.2f M env-steps/s)" % (os.environ.get("ILRL_PDL", "1"), n, mode, opt, min(res), n / min(res)))
