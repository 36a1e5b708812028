"""Would grouping envs of similar cost into the same warps pay?  Run a batch to a steady-state mix, then time ONE step
from that state with the envs PHYSICALLY permuted (coalescing kept): as is, sorted by torso height, sorted by the
kernel's own cost key, random; then the same through the kernel's indirection.
Usage: python tools/sort_experiment.py [num_envs]"""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("imitation-learning-rl_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
g = torch.Generator(device="cuda"); g.manual_seed(1)
os.environ["ILRL_GROUP_EVERY"] = "1000000"     # cost recorded, perm stays the identity
env = pkg.BatchedHumanoidEnv(n, "low", clips=["motion09_03"], seed=3, auto_reset=True)
env.reset()
for t in range(45):
    env.step(torch.rand(n, 17, device="cuda", generator=g) * 2 - 1)
phys, envf = env.get_state()
cost = torch.empty(n, dtype=torch.uint8, device="cuda")
env.L.ilrl_get_grouping(env.h, cost.data_ptr(), None, None)
torch.cuda.synchronize()
a = torch.rand(n, 17, device="cuda", generator=g) * 2 - 1
z = phys[:, 2]
print("cost histogram", torch.bincount(cost.long(), minlength=8).tolist())
def timed(e, perm):
    p, ev, aa = phys[perm].contiguous(), envf[perm].contiguous(), a[perm].contiguous()
    ts = []
    for r in range(12):
        e.set_state(p, ev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); e.step(aa); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))
idn = torch.arange(n, device="cuda")
print("as is                : %.1f us" % (timed(env, idn) * 1e3))
print("sorted by z          : %.1f us" % (timed(env, torch.argsort(z)) * 1e3))
print("sorted by kernel key : %.1f us" % (timed(env, torch.argsort(cost.long(), stable=True)) * 1e3))
c = torch.zeros_like(z)
for name, key in (("z + 0.02 rows", z - 0.02 * c), ("z + 0.05 rows", z - 0.05 * c), ("z + 0.1 rows", z - 0.1 * c),
                  ("z bucket 0.1, rows", torch.floor(z * 10) - 0.01 * c), ("z bucket 0.05, rows", torch.floor(z * 20) - 0.01 * c),
                  ("z desc", -z), ("|z - 1.25|", -(z - 1.25).abs()), ("min(z, 1.3)", torch.clamp(z, max=1.3)),
                  ("min(z, 1.1)", torch.clamp(z, max=1.1)), ("min(z, 1.0)", torch.clamp(z, max=1.0))):
    print("%-21s: %.1f us" % (name, timed(env, torch.argsort(key, stable=True)) * 1e3))
print("random perm          : %.1f us" % (timed(env, torch.randperm(n, device="cuda")) * 1e3))
env.close()
os.environ["ILRL_GROUP_EVERY"] = "1"
env2 = pkg.BatchedHumanoidEnv(n, "low", clips=["motion09_03"], seed=3, auto_reset=True)
env2.reset()
g.manual_seed(1)
for t in range(45):
    env2.step(torch.rand(n, 17, device="cuda", generator=g) * 2 - 1)
print("indirect, sort + step: %.1f us" % (timed(env2, idn) * 1e3))
