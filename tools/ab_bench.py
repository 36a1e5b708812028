"""A/B timing of step-kernel builds (development aid): for every exp/var_*.so (or the paths given) run the device-timed
step loop in a fresh process (ILRL_SO selects the library) and print best / median block time at each batch size.
usage: python tools/ab_bench.py [--sizes 4096,65536] [so ...]"""
import glob
import os
import subprocess
import sys

CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, ".")
import ilrl_b200
from ilrl_b200.batched_env import BatchedHumanoidEnv
for n in [int(x) for x in sys.argv[1].split(",")]:
    env = BatchedHumanoidEnv(n, "low", clips=["motion09_03"], auto_reset=True, seed=1234)
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    acts = [torch.rand(n, 17, device="cuda", generator=g) * 2 - 1 for _ in range(8)]
    for i in range(100): env.step(acts[i % 8])
    torch.cuda.synchronize()
    ts = []
    for rep in range(9):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 100
        e0.record()
        for i in range(K): env.step(acts[i % 8])
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / K * 1e3)
    st = env.stats().cpu().numpy()
    ts.sort()
    print("  N=%6d best %.2f us median %.2f us -> %.2f M/s   (mean_len %.2f mean_rew %.4f)" % (n, ts[0], ts[4], n / ts[4], st[2] / max(st[0], 1), st[4] / max(st[3], 1)), flush=True)
    env.close()
'''
args = sys.argv[1:]
sizes = "4096,65536"
if args and args[0] == "--sizes":
    sizes, args = args[1], args[2:]
sos = args or sorted(glob.glob("exp/var_*.so"))
for so in sos:
    print(os.path.basename(so), flush=True)
    env = dict(os.environ, ILRL_SO=os.path.abspath(so))
    subprocess.call([sys.executable, "-c", CHILD, sizes], env=env)
