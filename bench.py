#!/usr/bin/env python3
"""bench.py — imitation env-steps/s of the B200-native batched humanoid env (the driver's measurement contract).

Workload (BASELINE.json configs[1], SURVEY.md section 8d cfg 2): low-level imitation env, 4096 batched envs PER GPU,
clip motion09_03, random reference start frames, uniform random actions, auto-reset on.  One "step" = one fused env
step (apply_action -> 4 physics substeps -> calc_state -> reward -> frame advance -> target bookkeeping -> obs ->
done -> reset) of all 4096 envs of a rank = ONE kernel launch.  Weak scaling: every rank owns its own 4096 envs; the
only collective is the NCCL all-reduce of the 16-float statistics vector at the end of the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W]            -> this framework
  python bench.py --impl reference [--steps K] [--warmup W]      -> CPU arm: the oracle port of the reference path
                                                                    (PyBullet is not installable in this image)
  python bench.py --workload hier16384|multiclip65536|rollout16384x8|hier_rollout16384x10   -> extra measurement modes (BASELINE cfg 3/4/5)
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "imitation env-steps/sec (physics+reward)"
UNIT = "env-steps/s"
ENVS_PER_GPU = 4096
CLIP = "motion09_03"
# BASELINE.json configs: the contract line is cfg 2 ("low4096"); the others are extra measurement modes
WORKLOADS = {
    "low4096": dict(mode="low", envs=4096, total=None, clips=[CLIP],
                    name="low-level imitation env, 4096 batched envs per GPU, motion09_03, random start frames, random actions, auto-reset"),
    "hier16384": dict(mode="hier", envs=16384, total=None, clips=["motion08_03", "motion09_03"],
                      name="hierarchical env (HumanoidBulletEnvHier-v0), 16384 envs per GPU, selected_motion=1, step_per_level=5, "
                           "random high (heading) and low (torque) actions, auto-reset; low-level steps counted"),
    "multiclip65536": dict(mode="low", envs=None, total=65536, clips=["motion02_04", "motion08_03", "motion09_03", "motion13_13"],
                           name="multi-clip imitation (02_04, 08_03, 09_03, 13_13; clip = env id mod 4), 65536 envs sharded "
                                "over the GPUs, random actions, auto-reset"),
}
BYTES_PER_ENV_STEP = 929  # 2 x 288 B state + 68 B action + 280 B obs + 4 B reward + 1 B done
WORKLOAD = "low-level imitation env, %d batched envs per GPU, %s, random start frames, random actions, auto-reset" % (
    ENVS_PER_GPU, CLIP)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_rollout(cores, envs_per_core, warmup, steps, seed=7):
    """`cores` host threads, each advancing its own `envs_per_core` oracle envs (C loop, GIL released) by
    warmup + steps env steps.  Returns (total env steps in the timed part, seconds)."""
    from oracle import oracle as O
    L = O.lib()
    L.ilrl_oracle_rollout.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_long),
                                      C.POINTER(C.c_double)]
    L.ilrl_oracle_rollout.restype = C.c_long
    groups = []
    for c in range(cores):
        envs = [O.OracleEnv(CLIP, 0) for _ in range(envs_per_core)]
        for k, e in enumerate(envs):
            e.reset((7 * k + c) % 80, 0.0, (37 * k + 11 * c) % 360 - 180)
        groups.append((envs, (C.c_void_p * envs_per_core)(*[e.h for e in envs])))
    bar = threading.Barrier(cores + 1)
    counts = [0] * cores

    def work(c):
        arr = groups[c][1]
        if warmup:
            L.ilrl_oracle_rollout(arr, envs_per_core, warmup, seed + c, None, None)
        bar.wait()
        counts[c] = L.ilrl_oracle_rollout(arr, envs_per_core, steps, seed + 1000 + c, None, None)
        bar.wait()

    th = [threading.Thread(target=work, args=(c,)) for c in range(cores)]
    for t in th:
        t.start()
    bar.wait()
    t0 = time.perf_counter()
    bar.wait()
    dt = time.perf_counter() - t0
    for t in th:
        t.join()
    return sum(counts), dt


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    envs_per_core = 64  # bounded sample of the 4096-env workload: `cores` x 64 envs
    # one "step" of this arm advances every sampled env by `inner` env steps, sized so that the timed part lasts about
    # 12 s whatever K is (a 20-step run of single env steps would time 20 ms of cold threads: measured 0.12 M env-steps/s
    # against 0.21 M in steady state)
    inner = max(1, min(200, int(12.0 * 2.0e5 / (cores * envs_per_core * max(args.steps, 1)))))
    total, dt = cpu_rollout(cores, envs_per_core, min(args.warmup * inner, 200), args.steps * inner)
    value = total / dt
    sample = "%d host threads x %d oracle envs (C, fp64, dense 23x23 dynamics) x %d steps of %d env steps of the same workload" % (
        cores, envs_per_core, args.steps, inner)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_step": cores * envs_per_core, "env_steps_per_env_per_step": inner,
                   "note": "reference arm = oracle/ilrl_oracle.c (port of the reference path; the reference's own "
                           "implementation needs PyBullet, which is not installable here)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    EMIT(json.dumps(line))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = "/tmp/ilrl_clocks_%d_%d.csv" % (os.getpid(), gpu_index)
        self.p = None
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = args.steps, max(args.warmup, 3)

    # CPU baseline first (before CUDA is initialised in this process), rank 0 at N=1 only
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        steps = 2400  # ~12 s of CPU work: cores x 64 envs x 2400 steps at ~13k env-steps/s/core
        total, dt = cpu_rollout(cores, 64, 10, steps)
        cpu_base = {"value": total / dt, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": "%d host threads x 64 oracle envs (oracle/ilrl_oracle.c, fp64) x %d env steps of the "
                              "same workload (%d env-steps, %.1f s)" % (cores, steps, total, dt)}

    import torch
    import torch.distributed as dist
    import ilrl_b200
    from ilrl_b200.batched_env import BatchedHumanoidEnv

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = WORKLOADS[args.workload]
    hier = wl["mode"] == "hier"
    if wl["total"]:  # strong scaling over a fixed total (cfg 4): contiguous env-id blocks, clip = global id mod 4
        first, n = ilrl_b200.stats.shard_envs(wl["total"], world, rank)
    else:
        first, n = rank * wl["envs"], wl["envs"]
    cid = ilrl_b200.stats.clip_of_env(first, n, len(wl["clips"])) if len(wl["clips"]) > 2 else (
        np.ones(n, np.int32) if hier else None)
    env = BatchedHumanoidEnv(n, wl["mode"], clips=wl["clips"], clip_of_env=cid, device=local_rank, seed=1234,
                             auto_reset=True, env_id_base=first)  # same seed, global env ids: sharding-invariant
    env.reset()
    # action pool larger than L2 (126 MB), e.g. 512 batches x 4096 x 17 x 4 B = 142 MB, rotated through -> inputs are
    # never L2-resident from the previous use.  (The persistent env state IS on-chip between steps: that is the
    # workload — state never leaves the GPU.)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    POOL = max(8, min(512, (160 << 20) // (n * 17 * 4)))  # keep the pool just above L2 size at every N
    pool = torch.rand(POOL, n, 17, device=dev, generator=g) * 2 - 1
    hpool = torch.rand(64, n, 2, device=dev, generator=g) * 2 - 1 if hier else None
    calls = [0]

    def one_step(i):
        """one low-level env step of every env; in hier mode the waiting envs first get their heading action
        (the kernel ignores the others), so that every call advances every env by one low-level step"""
        if hier:
            env.high_step(hpool[calls[0] % 64])
        calls[0] += 1
        env.step(pool[i % POOL])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        one_step(i)
    env.stats()
    barrier()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    l0 = env.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        one_step(W + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = env.launch_count() - l0
    st = env.stats()  # {episodes, sum return, sum length, steps, sum reward, ...}: the only cross-rank exchange
    t_ms = torch.tensor([ms], device=dev)
    ilrl_b200.stats.allreduce_stats(st)  # NCCL sum of 16 floats over the ranks (no-op at N=1)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    clk = clocks.stop() if clocks else None
    n_all = torch.tensor([float(n)], device=dev)
    if world > 1:
        dist.all_reduce(n_all, op=dist.ReduceOp.SUM)
    value = float(n_all.item()) * K / (ms * 1e-3)

    # per-launch kernel time: events around every launch, on the launching stream, in a separate pass
    KT = min(K, 200)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(KT)]
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(evs):
        if hier:
            env.high_step(hpool[i % 64])
        a.record()
        env.step(pool[(W + K + i) % POOL])
        b.record()
    torch.cuda.synchronize()
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))

    # end-to-end through the C ABI with HOST buffers (H2D actions, kernel, D2H obs/reward/done inside the timing)
    # pinned host memory on both sides (the contract's e2e definition): DMA endpoints, no staging copies
    NH = min(64, POOL)
    host_act = pool[:NH].cpu().pin_memory().numpy()
    obs_h = torch.zeros(n, 70).pin_memory().numpy(); rew_h = torch.zeros(n).pin_memory().numpy()
    done_h = torch.zeros(n, dtype=torch.uint8).pin_memory().numpy()
    KE = min(K, 300)
    def host_step(i):
        if hier:
            env.high_step(hpool[i % 64])
        env.step_host(host_act[i % NH], obs_h, rew_h, done_h)

    for i in range(5):
        host_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(KE):
        host_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_val = float(n_all.item()) * KE / float(t_e.item())

    if rank == 0:
        peak, which = measured_peak()
        achieved = BYTES_PER_ENV_STEP * n / (kern_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        stn = st.cpu().numpy()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if wl["total"] else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "envs_per_gpu": n, "clips": wl["clips"], "auto_reset": True,
                       "l2": "action batches rotate through a %d MB pool (> 126 MB L2); the env state is "
                             "persistent on-device state by design" % (POOL * n * 17 * 4 >> 20),
                       "episodes": float(stn[0]), "mean_episode_len": float(stn[2] / max(stn[0], 1)),
                       "mean_step_reward": float(stn[4] / max(stn[3], 1))},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": which, "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_launch": BYTES_PER_ENV_STEP * n,
                         "note": "latency/issue-bound path (SURVEY 8d): HBM fraction is reported as the tier asks; "
                                 "see profiles/ for SM issue utilisation",
                         # informative second view: what actually bounds the kernel.  6018 warp-instructions per
                         # env-step (smsp__inst_executed.sum / 4096, profiles/r1_v7_step_kernel_ncu.md) against
                         # 4 issue slots per SM per cycle at the sampled SM clock
                         "issue": {"achieved_gwarp_inst_s": 6018.0 * n / (kern_ms * 1e-3) * 1e-9,
                                   "peak_gwarp_inst_s": 148 * 4 * float((clk or {}).get("sm_mhz") or 1965.0) * 1e-3,
                                   "warp_inst_per_env_step": 6018}},
            "cpu_baseline": cpu_base,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": n * 17 * 4,
                    "d2h_bytes_per_step": n * (70 * 4 + 4 + 1), "steps": KE,
                    "api": "ilrl_step_host (C ABI, pinned + mapped host buffers: actions read and obs/reward/done written in place by the step kernel)"},
            "gpu_launches": int(launches),
            "clocks": clk,
        }
        EMIT(json.dumps(line))
    env.close()
    if world > 1:
        dist.destroy_process_group()


def _quiet_stdout():
    """Libraries (NCCL's version banner, torchrun notices) may write to fd 1; the contract is ONE JSON line on
    stdout.  Point fd 1 at stderr for the whole run and return a writer bound to the real stdout for the final line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real, "w")

    def emit(line):
        out.write(line + "\n")
        out.flush()
    return emit


EMIT = print


def run_rollout(args, hier=False):
    """hier=True: the same for the hierarchical env (both policies on device, 5 launches per tick, horizon 10 =
    rollout_fragment_length of REF train_config.py:257; counts low-level env steps).
    Extra measurement mode (BASELINE cfg 5): on-device rollout collection, 16384 envs x 8 steps per GPU per
    iteration (= 1 M env-steps per iteration on 8 GPUs): fused tcgen05 policy / value / sampling kernel -> fused env step,
    2 launches per step captured in one CUDA graph, + GAE (ilrl_gae).  A "step" is one iteration; value = env-steps/s of the whole job."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    import torch.distributed as dist
    import ilrl_b200
    from ilrl_b200 import BatchedHumanoidEnv, GaussianMLPPolicy, HierRolloutCollector, RolloutCollector
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, T = 16384, (10 if hier else 8)
    K, W = min(args.steps, 200), max(3, min(args.warmup, 20))
    torch.manual_seed(0)
    if hier:
        env = BatchedHumanoidEnv(n, "hier", clips=["motion08_03", "motion09_03"], clip_of_env=np.ones(n, np.int32),
                                 device=local_rank, seed=1234, auto_reset=True, env_id_base=rank * n)
        col = HierRolloutCollector(env, horizon=T, gamma=0.99, lam=0.9, seed=rank)
    else:
        env = BatchedHumanoidEnv(n, "low", clips=[CLIP], device=local_rank, seed=1234, auto_reset=True, env_id_base=rank * n)
        col = RolloutCollector(env, GaussianMLPPolicy(), horizon=T, gamma=0.99, lam=0.9, seed=rank)
    for _ in range(W):
        col.collect()
    env.stats()
    l0 = env.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K):
        batch = col.collect()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    st = ilrl_b200.stats.allreduce_stats(env.stats())
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    if rank == 0:
        summ = ilrl_b200.stats.summarize(st)
        EMIT(json.dumps({
            "metric": "rollout env-steps/sec (%spolicy + physics + reward + GAE on device)" % ("both policies: " if hier else ""), "value": world * n * T * K / (ms * 1e-3),
            "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("on-device PPO rollout collection, hierarchical env: %d envs x %d low-level steps per GPU per "
                                    "iteration, selected_motion=1, step_per_level=5, 44-256-256-2 and 70-256-256-17 tanh Gaussian "
                                    "policies + value nets (fused tcgen05 kernel), gamma 0.99 lambda 0.9" % (n, T)) if hier else
                                   ("on-device PPO rollout collection: %d envs x %d steps per GPU per iteration, %s, "
                                    "70-256-256-17 tanh Gaussian policy + value net (fused tcgen05 kernel, bf16 operands / fp32 accumulation), "
                                    "gamma 0.99 lambda 0.9" % (n, T, CLIP)), "envs_per_gpu": n, "horizon": T,
                       "episode_len_mean": summ["episode_len_mean"],
                       "sample_batch_columns": sorted(batch["low"].keys()) if hier else sorted(batch.keys())},
            # per iteration, replayed from the graph: low: T env steps + (T + 1) policy steps, + 1 GAE launch;
            # hier: 5 launches per tick + readout and two value calls after the last + 2 GAE launches
            "gpu_launches": K * ((5 * T + 5) if hier else (2 * T + 2))}))
    env.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    global EMIT
    EMIT = _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="low4096", choices=sorted(WORKLOADS) + ["rollout16384x8", "hier_rollout16384x10"],
                    help="low4096 = the contract line (BASELINE cfg 2); the others are extra measurement modes")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("rollout16384x8", "hier_rollout16384x10"):
        run_rollout(args, hier=args.workload.startswith("hier"))
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
