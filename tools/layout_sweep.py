"""Device-side timing of the fused step kernel for every shared-memory layout over a range of batch sizes
(development aid for the layout model in ilrl_create; bench.py is the contract).
usage: python tools/layout_sweep.py [N ...]    env ILRL_SWEEP_LAYOUTS=small,large,dense4"""
import os
import sys

import torch

sys.path.insert(0, ".")
import ilrl_b200  # noqa: F401,E402
from ilrl_b200.batched_env import BatchedHumanoidEnv  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [4096, 8192, 16384, 65536]
layouts = os.environ.get("ILRL_SWEEP_LAYOUTS", "small,large,dense4").split(",")
for n in sizes:
    out = []
    for lay in layouts:
        os.environ["ILRL_LAYOUT"] = lay
        env = BatchedHumanoidEnv(n, "low", clips=["motion09_03"], auto_reset=True, seed=1234)
        env.reset()
        g = torch.Generator(device="cuda")
        g.manual_seed(0)
        acts = [torch.rand(n, 17, device="cuda", generator=g) * 2 - 1 for _ in range(8)]
        for i in range(60):
            env.step(acts[i % 8])
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            K = 100
            e0.record()
            for i in range(K):
                env.step(acts[i % 8])
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / K)
        out.append("%s %.1f us %.1f M" % (lay, best * 1e3, n / best / 1e3))
        env.close()
    print("N=%6d  " % n + "   ".join(out), flush=True)
os.environ.pop("ILRL_LAYOUT", None)
