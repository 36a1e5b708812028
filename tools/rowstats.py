#!/usr/bin/env python3
"""Histogram of (violated joint limits, kept ground contacts) per physics substep in the steady state of the bench
workload (random actions, reset on done), from the CPU oracle.  Used to size the constraint phases of the step kernel
(DESIGN.md section 5): rows = limits + 3 x contacts; the Gauss-Seidel phase of a warp lasts as long as its env with
the most rows.  Test/diagnostic infrastructure: runs the oracle, never the product."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402


def main(n=256, warm=60, steps=60):
    L = O.lib()
    L.ilrl_oracle_rollout.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_long),
                                      C.POINTER(C.c_double)]
    L.ilrl_oracle_rollout.restype = C.c_long
    L.ilrl_oracle_set_rowstats.argtypes = [C.POINTER(C.c_long)]
    envs = [O.OracleEnv("motion09_03", 0) for _ in range(n)]
    for k, e in enumerate(envs):
        e.reset((7 * k) % 80, 0.0, (37 * k) % 360 - 180)
    arr = (C.c_void_p * n)(*[e.h for e in envs])
    L.ilrl_oracle_rollout(arr, n, warm, 1, None, None)
    hist = np.zeros(18 * 9, dtype=np.int64)
    L.ilrl_oracle_set_rowstats(hist.ctypes.data_as(C.POINTER(C.c_long)))
    eps = C.c_long(0)
    done = L.ilrl_oracle_rollout(arr, n, steps, 2, C.byref(eps), None)
    L.ilrl_oracle_set_rowstats(None)
    h = hist.reshape(18, 9).astype(float)
    tot = h.sum()
    print("substeps %d, env steps %d, episodes %d (mean length %.1f)" % (tot, done, eps.value, done / max(eps.value, 1)))
    nl = np.arange(18)[:, None] + 0 * np.arange(9)[None, :]
    nc = 0 * np.arange(18)[:, None] + np.arange(9)[None, :]
    rows = nl + 3 * nc
    items = (nl + 2) // 3 + nc
    print("mean limits %.2f contacts %.2f rows %.2f items %.2f" % ((h * nl).sum() / tot, (h * nc).sum() / tot,
                                                                   (h * rows).sum() / tot, (h * items).sum() / tot))
    print("contacts histogram:", np.round(h.sum(0) / tot, 3))
    print("limits histogram  :", np.round(h.sum(1) / tot, 3))
    rh = np.zeros(45)
    for a in range(18):
        for b in range(9):
            rh[a + 3 * b] += h[a, b]
    cdf = np.cumsum(rh) / tot
    print("rows cdf:", " ".join("%d:%.3f" % (r, cdf[r]) for r in range(0, 45, 2)))
    # expected max of rows over 8 envs (one warp) if envs were independent
    pm = np.diff(np.concatenate([[0], cdf ** 8]))
    print("E[max rows over 8 envs] %.1f   E[rows] %.1f   E[rows^2] %.1f" % ((pm * np.arange(45)).sum(), (rh * np.arange(45)).sum() / tot,
                                                                        (rh * np.arange(45) ** 2).sum() / tot))
    pm512 = np.diff(np.concatenate([[0], cdf ** 4096]))
    print("E[max rows over 4096 envs] %.1f" % (pm512 * np.arange(45)).sum())


if __name__ == "__main__":
    main(*[int(x) for x in sys.argv[1:]])
