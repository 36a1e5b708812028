"""Per-warp phase cycles of the step kernel from the measurement build (-DILRL_PROF): where the SLOWEST warp of a
launch spends its time against the average warp (at one wave of CTAs the step ends with the slowest warp).
usage (GPU box): python tools/warp_profile.py [N]      builds build_prof.so next to the repo root"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "build_prof.so")


def build():
    src = os.path.join(ROOT, "imitation-learning-rl_b200", "csrc")
    subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-ftz=true",
                           "-prec-div=false", "-prec-sqrt=false", "-DILRL_PROF", "-Xcompiler", "-fPIC", "-shared", "-o", SO,
                           os.path.join(src, "ilrl_capi.cu"), os.path.join(src, "ilrl_policy.cu")])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
        sys.exit(0)
    os.environ["ILRL_SO"] = SO
    sys.path.insert(0, ROOT)
    import torch
    import ilrl_b200  # noqa: F401
    from ilrl_b200.batched_env import BatchedHumanoidEnv
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    env = BatchedHumanoidEnv(n, "low", clips=["motion09_03"], auto_reset=True, seed=1234,
                             self_collision=bool(os.environ.get("ILRL_SELFCOL")))   # ILRL_SELFCOL=1: the SELFC kernel
    env.reset()
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    acts = [torch.rand(n, 17, device="cuda", generator=g) * 2 - 1 for _ in range(8)]
    for i in range(100):
        env.step(acts[i % 8])
    nw = (n + 7) // 8
    buf = np.zeros((nw, 16), np.int64)
    env.L.ilrl_debug_profile.argtypes = [C.c_void_p, C.c_void_p]
    env.L.ilrl_debug_profile(env.h, buf.ctypes.data)   # clear
    names = ["head", "fk", "inward", "outward", "rows", "pgs", "integ", "tail", "barrier", "total", "maxrows",
             "t_pose", "t_reward", "t_obs", "t_store"]   # parts of the tail ("tail" = its rest: obs rows out + statistics)
    K = 50
    acc_mean = np.zeros(16)
    acc_slow = np.zeros(16)
    tot_max = []
    for i in range(K):
        env.step(acts[i % 8])
        env.L.ilrl_debug_profile(env.h, buf.ctypes.data)
        t = buf[:, 9]
        w = int(t.argmax())
        acc_mean += buf.mean(0)
        acc_slow += buf[w]
        tot_max.append(t.max())
    print("N=%d, %d launches; cycles per launch: mean warp / slowest warp of the launch" % (n, K))
    for k, nm in enumerate(names):
        print("  %-8s %9.0f %9.0f" % (nm, acc_mean[k] / K, acc_slow[k] / K))
    print("  slowest-warp total: mean %.0f  min %.0f  max %.0f cycles" % (np.mean(tot_max), np.min(tot_max), np.max(tot_max)))
    env.close()
