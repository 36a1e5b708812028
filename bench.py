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
    "hier2_16384": dict(mode="hier2", envs=16384, total=None, clips=["motion08_03", "motion09_03"],
                        name="hier_env_2.py variant (SURVEY row a18), 16384 envs per GPU, selected_motion=1, step_per_level=20, "
                             "skipFrame=5, random 36-d high (joint-target) and 17-d low (torque) actions, auto-reset; low-level "
                             "steps counted"),
    "terrain4096": dict(mode="low", envs=4096, total=None, clips=[CLIP], terrain=True,
                        name="low-level imitation env on the CustomScene heightfield terrain (SURVEY row f4: 256 x 256 samples, "
                             "2 x 2 plateaus of U(0, 0.5) m), 4096 envs per GPU, motion09_03, random actions, auto-reset"),
    "selfcol4096": dict(mode="low", envs=4096, total=None, clips=[CLIP], self_collision=True,
                        name="low-level imitation env with Bullet-style self-collision on (SURVEY row f4: 66 capsule pairs, "
                             "two-body contact rows), 4096 envs per GPU, motion09_03, random actions, auto-reset"),
    "multiclip65536": dict(mode="low", envs=None, total=65536, clips=["motion02_04", "motion08_03", "motion09_03", "motion13_13"],
                           name="multi-clip imitation (02_04, 08_03, 09_03, 13_13; clip = env id mod 4), 65536 envs sharded "
                                "over the GPUs, random actions, auto-reset"),
}
E2E_PARTS = 4   # parts of the pipelined end-to-end leg (measured 2 / 3 / 4 / 8: 42.9 / 44.5 / 44.8 / 45.0 M env-steps/s)
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


def reference_envs_per_core(cores):
    """The CPU arm steps the SAME number of envs as one GPU of the contract workload (4096), split over the host threads."""
    return max(1, ENVS_PER_GPU // max(cores, 1))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    envs_per_core = reference_envs_per_core(cores)
    n_envs = cores * envs_per_core
    # one "step" of this arm advances every env of the 4096-env workload by `inner` env steps, sized so that the timed
    # part lasts about 12 s whatever K is (a 20-step run of single env steps would time 0.4 s of cold threads)
    inner = max(1, min(200, int(12.0 * 2.0e5 / (n_envs * max(args.steps, 1)))))
    total, dt = cpu_rollout(cores, envs_per_core, min(args.warmup * inner, 20), args.steps * inner)
    value = total / dt
    sample = "%d host threads x %d oracle envs = %d envs (C, fp64, dense 23x23 dynamics) x %d steps of %d env steps of the same workload" % (
        cores, envs_per_core, n_envs, args.steps, inner)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs_per_gpu": n_envs, "env_steps_per_env_per_step": inner,
                   "note": "reference arm = oracle/ilrl_oracle.c, a CPU PORT of the reference path written for this repo "
                           "(the reference's own implementation needs PyBullet, which is not installable here): ratios "
                           "against this line are ratios against that port, not against PyBullet"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    EMIT(json.dumps(line))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock, power and throttle reasons of one GPU, sampled IN PROCESS through NVML every 2 ms by a thread, from
    before the warm-up to the end of the last timed region (nvidia-smi -lms cannot sample a 2 ms region).  `mark()`
    brackets the regions that count as "under load"."""

    def __init__(self, torch_device_index):
        self.samples = []          # (t, sm_mhz, power_w, reasons_bitmask)
        self.marks = []            # [t0, t1] load intervals
        self.ok = False
        self._stop = threading.Event()
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            h = None
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(torch_device_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(torch_device_index)
            self.nv, self.h = pynvml, h
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()
        except Exception as e:  # no NVML: the driver's own sampler is the only evidence
            self.err = repr(e)

    def _run(self):
        nv, h = self.nv, self.h
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = float("nan")
                self.samples.append((time.perf_counter(), sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def mark(self, t0, t1):
        self.marks.append((t0, t1))

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "samples_under_load": 0, "source": "nvml in-process, 2 ms period"}
        if not self.ok:
            out["source"] = "unavailable: " + getattr(self, "err", "?")
            return out
        self._stop.set()
        self.th.join(timeout=1.0)
        nv = self.nv
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        load = [s for s in self.samples if any(a <= s[0] <= b for a, b in self.marks)]
        use = load or self.samples
        reasons = set()
        for s in use:
            for nm, bit in bits.items():
                if s[3] & bit:
                    reasons.add(nm)
        if use:
            pw = [s[2] for s in use if s[2] == s[2]]
            out.update(sm_mhz=float(np.median([s[1] for s in use])), sm_max_mhz=self.sm_max, reasons=sorted(reasons),
                       samples=len(self.samples), samples_under_load=len(load),
                       power_w_max=float(max(pw)) if pw else None)
        return out


# ------------------------------------------------------------------------------------------------ GPU arm
def make_env(wl, rank, world, local_rank):
    import ilrl_b200
    from ilrl_b200.batched_env import BatchedHumanoidEnv
    hier = wl["mode"] in ("hier", "hier2")
    if wl["total"]:  # strong scaling over a fixed total (cfg 4): contiguous env-id blocks, clip = global id mod 4
        first, n = ilrl_b200.stats.shard_envs(wl["total"], world, rank)
    else:
        first, n = rank * wl["envs"], wl["envs"]
    cid = ilrl_b200.stats.clip_of_env(first, n, len(wl["clips"])) if len(wl["clips"]) > 2 else (
        np.ones(n, np.int32) if hier else None)
    env = BatchedHumanoidEnv(n, wl["mode"], clips=wl["clips"], clip_of_env=cid, device=local_rank, seed=1234,
                             auto_reset=True, env_id_base=first,  # same seed, global env ids: sharding-invariant
                             self_collision=bool(wl.get("self_collision")))
    if wl.get("terrain"):   # CustomScene.episode_restart (REF humanoid.py:88-124), one terrain per handle
        tr = np.random.default_rng(99)
        h = np.repeat(np.repeat(tr.uniform(0, 0.5, (128, 128)), 2, axis=0), 2, axis=1)
        h[126:130, 126:130] = 0.0
        env.set_heightfield(h.reshape(-1))
    env.reset()
    return env, n, hier


def measure_device(wl, K, W, rank, world, local_rank, clocks=None, min_region_s=0.5, max_blocks=200, kernel_events=True,
                   use_graph=True):
    """Device-timed throughput of one workload: W warm-up steps, then R blocks of EXACTLY K steps, each bracketed by a
    barrier + synchronize on both sides and timed with CUDA events on the launching stream; per block the MAX over
    ranks; reported: the MEDIAN block (R is chosen so that the blocks together last >= min_region_s: a single block of
    20 steps is 2 ms).  Returns (result dict, env, pools) - the env is kept for the end-to-end legs."""
    import torch
    import torch.distributed as dist
    import ilrl_b200
    dev = torch.device("cuda", local_rank)
    env, n, hier = make_env(wl, rank, world, local_rank)
    # action pool larger than L2 (126 MB), e.g. 512 batches x 4096 x 17 x 4 B = 142 MB, rotated through -> inputs are
    # never L2-resident from the previous use.  (The persistent env state IS on-chip between steps: that is the
    # workload — state never leaves the GPU.)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    POOL = max(8, min(512, (160 << 20) // (n * 17 * 4)))  # keep the pool just above L2 size at every N
    pool = torch.rand(POOL, n, 17, device=dev, generator=g) * 2 - 1
    hpool = torch.rand(64, n, env.hact_w, device=dev, generator=g) * 2 - 1 if hier else None
    calls = [0]

    def one_step():
        """one low-level env step of every env; in hier mode the waiting envs first get their heading action
        (the kernel ignores the others), so that every call advances every env by one low-level step"""
        i = calls[0]
        if hier:
            env.high_step(hpool[i % 64])
        calls[0] += 1
        env.step(pool[i % POOL])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def block(k, graph=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        if graph is not None:
            graph.replay()
        else:
            for _ in range(k):
                one_step()
        e1.record()
        barrier()
        if clocks:
            clocks.mark(t0, time.perf_counter())
        return e0.elapsed_time(e1)

    for _ in range(W):
        one_step()
    env.stats()
    est = block(K)  # untimed extra block: sizes R
    R = int(min(max_blocks, max(3, np.ceil(min_region_s * 1e3 / max(est, 1e-3)))))
    if world > 1:  # every rank must run the same number of blocks
        r_t = torch.tensor([R], device=dev)
        dist.all_reduce(r_t, op=dist.ReduceOp.MAX)
        R = int(r_t.item())
    # The K launches of a block are replayed from a CUDA graph (as rollout.py replays its fragments): the timed region
    # then holds no Python and no per-launch driver call, so host jitter (8 ranks share the box's cores) stays out of it.
    # G graphs over consecutive slices of the action pool are used in turn, so the actions still rotate through > L2.
    l0 = env.launch_count()
    block(K)
    launches_per_block = env.launch_count() - l0
    graphs = []
    if use_graph:
        G = int(max(1, min(R, POOL // max(K, 1))))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        for _ in range(G):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                for _ in range(K):
                    one_step()
            graphs.append(gr)
        torch.cuda.current_stream(dev).wait_stream(side)
        block(K, graphs[0])   # first replay untimed
    ms_blocks = torch.tensor([block(K, graphs[r % len(graphs)] if graphs else None) for r in range(R)], device=dev,
                             dtype=torch.float64)
    st = env.stats()  # {episodes, sum return, sum length, steps, sum reward, ...}: the only cross-rank exchange
    ilrl_b200.stats.allreduce_stats(st)  # NCCL sum of 16 floats over the ranks (no-op at N=1)
    n_all = torch.tensor([float(n)], device=dev)
    if world > 1:
        dist.all_reduce(ms_blocks, op=dist.ReduceOp.MAX)   # per block: the slowest rank
        dist.all_reduce(n_all, op=dist.ReduceOp.SUM)
    mb = np.sort(ms_blocks.cpu().numpy())
    ms = float(mb[len(mb) // 2])
    res = {"n": n, "n_all": float(n_all.item()), "hier": hier, "ms_block": ms, "ms_per_step": ms / K,
           "value": float(n_all.item()) * K / (ms * 1e-3), "blocks": R, "block_ms_min": float(mb[0]), "block_ms_max": float(mb[-1]),
           "launches_per_block": int(launches_per_block), "stats": st.cpu().numpy(), "pool_mb": POOL * n * 17 * 4 >> 20,
           "graphs": len(graphs)}
    if kernel_events:
        # time of the step kernel alone, measured ON THE DEVICE inside one more block of exactly K back-to-back steps:
        # every step kernel stamps %globaltimer at its first CTA's start and its last CTA's end (ilrl_kernel_timing; no
        # extra launch, no synchronisation), so the sum of the kernel times cannot exceed the block they ran in
        env.kernel_timing(True)
        kb = block(K)   # (launched one by one: the slot a launch stamps is a launch argument)
        kms, kcnt = env.kernel_timing(False)
        res["kernel_ms"] = kms / max(kcnt, 1)
        res["kernel_launches_timed"] = kcnt
        res["kernel_block_ms_per_step"] = kb / K
    return res, env, (pool, hpool, POOL)


def measure_sequence(env, n, pools, K, world, local_rank, blocks=60):
    """The same K steps per block through ilrl_step_sequence: ONE launch per block, a CTA takes its tile through all K
    steps (possible only because the K action batches exist before the block starts, as they do in this benchmark and
    in open-loop playback; a policy in the loop cannot use it).  Same timing rules as measure_device."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    pool, _, POOL = pools
    obs = torch.empty(K, n, 70, device=dev)
    rew = torch.empty(K, n, device=dev)
    done = torch.empty(K, n, dtype=torch.uint8, device=dev)
    slices = max(1, POOL // K)

    def block(r):
        a = pool[(r % slices) * K:(r % slices) * K + K]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        env.step_sequence(a, obs, rew, done)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for r in range(3):
        block(r)
    ms = torch.tensor([block(r) for r in range(blocks)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    mb = np.sort(ms.cpu().numpy())
    return {"ms_block": float(mb[len(mb) // 2]), "blocks": blocks}


def measure_e2e(env, n, hier, pools, K, rank, world, local_rank, clocks=None, min_region_s=0.5):
    """The same metric end to end through the C ABI with HOST buffers: every step's actions come from pinned host memory
    and its observations / rewards / dones land in pinned host memory inside the timed region.
    (1) `async`: ilrl_step_host_async / ilrl_wait with the batch cut in E2E_PARTS parts, each on its own stream, stepped
        as a software pipeline (the way a rollout worker overlaps its own work with the env; the host reads the part's
        results between its wait and its next submit): the headline end-to-end number;
    (2) `sync`: one blocking ilrl_step_host per step (launch + PCIe + synchronise exposed every step);
    (3) `serve`: one blocking ilrl_serve_step per step against the resident serving kernel (no launch, no stream
        synchronise; begin / end of the session inside every timed block).  Batches that fit one wave only."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    pool, hpool, POOL = pools
    NH = min(64, POOL)
    host_act = [pool[j].cpu().pin_memory().numpy() for j in range(NH)]
    obs_h = torch.zeros(n, env.obs_w).pin_memory().numpy()
    rew_h = torch.zeros(n).pin_memory().numpy()
    done_h = torch.zeros(n, dtype=torch.uint8).pin_memory().numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = {}
    steps = [0]
    P = E2E_PARTS
    sl = [env.part_slice(p, P) for p in range(P)]
    sink = [0.0]
    for kind in ("async", "sync", "serve"):
        if kind == "async" and hier:
            continue   # the hierarchical env needs its high-level call between the parts: only the blocking path is timed
        if kind == "serve" and (hier or n > 4096 or not os.environ.get("ILRL_BENCH_SERVE")):
            continue   # opt-in (ILRL_BENCH_SERVE=1): tools/serve_bench.py is where the serving path is measured
        def run(k):
            if kind == "serve":
                env.serve_begin(obs_h, rew_h, done_h)
                for _ in range(k):
                    i = steps[0]
                    steps[0] += 1
                    env.serve_step(host_act[i % NH])
                env.serve_end()
            elif kind == "sync":
                for _ in range(k):
                    i = steps[0]
                    if hier:
                        env.high_step(hpool[i % 64])
                    steps[0] += 1
                    env.step_host(host_act[i % NH], obs_h, rew_h, done_h)
            else:
                # software pipeline over two parts: while part p steps on the GPU, the host consumes the outputs of the
                # other part and submits its next step
                i = steps[0]
                for p in range(P):
                    env.step_host_async(p, P, host_act[i % NH], obs_h, rew_h, done_h)
                for j in range(1, k):
                    a = host_act[(i + j) % NH]
                    for p in range(P):
                        env.wait(p)
                        sink[0] += rew_h[sl[p].start]            # the part's results are in host memory: read one
                        env.step_host_async(p, P, a, obs_h, rew_h, done_h)
                for p in range(P):
                    env.wait(p)
                steps[0] += k
        run(5)
        barrier()
        t0 = time.perf_counter()
        run(K)
        torch.cuda.synchronize()
        est = time.perf_counter() - t0
        R = int(min(200, max(3, np.ceil(min_region_s / max(est, 1e-6)))))
        if world > 1:
            r_t = torch.tensor([R], device=dev)
            dist.all_reduce(r_t, op=dist.ReduceOp.MAX)
            R = int(r_t.item())
        ts = []
        for _ in range(R):
            barrier()
            t0 = time.perf_counter()
            run(K)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            ts.append(t1 - t0)
            if clocks:
                clocks.mark(t0, t1)
        t_e = torch.tensor(ts, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        tb = np.sort(t_e.cpu().numpy())
        out[kind] = {"s_block": float(tb[len(tb) // 2]), "blocks": R}
    return out


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = args.steps, max(args.warmup, 3)
    if world > 1:   # every rank on its own share of the host cores (the ranks of one box otherwise migrate over all of them)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]) or set(cores))
        except Exception:
            pass

    # CPU baseline first (before CUDA is initialised in this process), rank 0 at N=1 only
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        epc = reference_envs_per_core(cores)
        steps = max(10, int(12.0 * 2.0e5 / (cores * epc)))  # ~12 s of CPU work at ~13k env-steps/s/core
        total, dt = cpu_rollout(cores, epc, 5, steps)
        cpu_base = {"value": total / dt, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": "%d host threads x %d oracle envs = %d envs (oracle/ilrl_oracle.c, fp64: a CPU port of the "
                              "reference path, not PyBullet) x %d env steps of the same workload (%d env-steps, %.1f s)" % (
                                  cores, epc, cores * epc, steps, total, dt)}

    import torch
    import torch.distributed as dist
    import ilrl_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    clocks = ClockSampler(local_rank) if rank == 0 else None

    wl = WORKLOADS[args.workload]
    res, env, pools = measure_device(wl, K, W, rank, world, local_rank, clocks, use_graph=not args.no_graph)
    n, hier = res["n"], res["hier"]
    seq = None
    if not hier and not wl.get("terrain") and not wl.get("self_collision") and K <= pools[2]:
        seq = measure_sequence(env, n, pools, K, world, local_rank)
    e2e = measure_e2e(env, n, hier, pools, K, rank, world, local_rank, clocks)
    env.close()
    del pools
    torch.cuda.empty_cache()

    # the other BASELINE configs, short runs in the same process (the contract line stays cfg 2): cfg 3 hierarchical
    # 16384 envs per GPU, cfg 4 multi-clip 65536 envs sharded over the ranks, cfg 5 on-device rollout collection
    extra = {}
    if args.workload == "low4096" and not args.no_extra_configs:
        for key, name in (("cfg3_hier16384", "hier16384"), ("cfg4_multiclip65536", "multiclip65536"),
                          ("a18_hier2_16384", "hier2_16384"), ("f4_terrain4096", "terrain4096"),
                          ("f4_selfcol4096", "selfcol4096")):
            r2, env2, p2 = measure_device(WORKLOADS[name], K, max(W, 10), rank, world, local_rank, clocks, min_region_s=0.3,
                                          kernel_events=False, use_graph=not args.no_graph)
            env2.close()
            del p2
            torch.cuda.empty_cache()
            extra[key] = {"workload": WORKLOADS[name]["name"], "value": r2["value"], "unit": UNIT, "ms_per_step": r2["ms_per_step"],
                          "envs_per_gpu": r2["n"], "steps": K, "blocks": r2["blocks"],
                          "scaling": "strong" if WORKLOADS[name]["total"] else "weak",
                          "mean_episode_len": float(r2["stats"][2] / max(r2["stats"][0], 1))}
        try:
            extra["cfg5_rollout16384x8"] = rollout_measure(rank, world, local_rank, hier=False, K=max(3, min(K, 50)), W=5, blocks=5)
        except Exception as e:  # the widened row must never take the contract line down with it
            extra["cfg5_rollout16384x8"] = {"error": repr(e)}
    clk = clocks.stop() if clocks else None

    if rank == 0:
        peak, which = measured_peak()
        kern_ms = res["kernel_ms"]
        achieved = BYTES_PER_ENV_STEP * n / (kern_ms * 1e-3) / 1e9
        # DRAM traffic and executed warp-instructions are ncu counters: they cannot be measured inside this process.
        # For the contract workload they are read from the committed ncu summary of THIS kernel generation
        # (profiles/step_kernel_traffic.json names the capture); for every other workload they are null.
        traffic, wipes, prof_src = None, None, None
        tp = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
        if args.workload == "low4096" and os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic, wipes, prof_src = tj.get("dram_bytes_per_launch"), tj.get("warp_inst_per_env_step"), tj.get("source")
            except Exception:
                pass
        stn = res["stats"]
        e2e_head = e2e.get("async") or e2e["sync"]
        line = {
            "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong" if wl["total"] else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "envs_per_gpu": n, "clips": wl["clips"], "auto_reset": True,
                       "l2": "action batches rotate through a %d MB pool (> 126 MB L2); the env state is "
                             "persistent on-device state by design" % res["pool_mb"],
                       "timing": "median of %d blocks of exactly %d steps, each block bracketed by barrier + synchronize and "
                                 "timed with CUDA events (max over ranks per block); %s" % (
                                     res["blocks"], K, ("the %d launches of a block replayed from one of %d CUDA graphs" % (
                                         res["launches_per_block"], res["graphs"])) if res["graphs"] else "launched one by one"),
                       "blocks": res["blocks"], "block_ms_min": res["block_ms_min"], "block_ms_max": res["block_ms_max"],
                       "episodes": float(stn[0]), "mean_episode_len": float(stn[2] / max(stn[0], 1)),
                       "mean_step_reward": float(stn[4] / max(stn[3], 1))},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": prof_src, "peak_source": which, "kernel_ms": kern_ms,
                         "kernel_block_ms_per_step": res["kernel_block_ms_per_step"],
                         "kernel_timing": "%globaltimer stamps written by the step kernel itself (first CTA past its grid-"
                                          "dependency wait, last CTA end) in a block of the same K back-to-back EAGER launches "
                                          "(kernel_block_ms_per_step); `ms_per_step` is the graph-replayed block",
                         "algorithmic_bytes_per_launch": BYTES_PER_ENV_STEP * n,
                         "note": "latency/issue-bound path (SURVEY 8d): HBM fraction is reported as the tier asks; "
                                 "see profiles/ for SM issue utilisation",
                         # informative second view: what actually bounds the kernel (issue slots), from the same capture
                         "issue": None if wipes is None else {
                             "achieved_gwarp_inst_s": wipes * n / (kern_ms * 1e-3) * 1e-9,
                             "peak_gwarp_inst_s": 148 * 4 * float((clk or {}).get("sm_mhz") or 1965.0) * 1e-3,
                             "warp_inst_per_env_step": wipes}},
            "cpu_baseline": cpu_base,
            "vs_cpu_port": None if not cpu_base else res["value"] / cpu_base["value"],
            "e2e": {"value": res["n_all"] * K / e2e_head["s_block"], "unit": UNIT, "h2d_bytes_per_step": n * 17 * 4,
                    "d2h_bytes_per_step": n * (70 * 4 + 4 + 1), "steps": K, "blocks": e2e_head["blocks"],
                    "api": ("ilrl_step_host_async + ilrl_wait (C ABI): the batch in %d pipelined parts, each on its own "
                            "stream; pinned + mapped host buffers, actions read and obs/reward/done written in place by the "
                            "step kernel" % E2E_PARTS) if "async" in e2e else
                           "ilrl_step_host (C ABI, pinned + mapped host buffers, blocking)",
                    "sync_value": res["n_all"] * K / e2e["sync"]["s_block"],
                    "sync_api": "ilrl_step_host: one blocking call per step (launch + PCIe + synchronise exposed every step)",
                    "serve_value": None if "serve" not in e2e else res["n_all"] * K / e2e["serve"]["s_block"],
                    "serve_api": "ilrl_serve_step: one blocking call per step against the resident serving kernel (doorbell in "
                                 "mapped host memory: no launch, no stream synchronise; session begin / end inside each block)"},
            "gpu_launches": res["launches_per_block"],
            # NOT the headline: the same K steps as one ilrl_step_sequence launch (no grid-wide barrier between steps)
            "sequence": None if seq is None else {
                "value": res["n_all"] * K / (seq["ms_block"] * 1e-3), "unit": UNIT, "ms_per_step": seq["ms_block"] / K,
                "launches_per_block": 1, "blocks": seq["blocks"],
                "api": "ilrl_step_sequence: K steps of open-loop actions in one launch; bit-identical to K ilrl_step calls "
                       "(tests/test_gpu_api.py); `value` above stays the launch-per-step path every RL caller uses"},
            "clocks": clk,
        }
        if extra:
            line["configs"] = extra
        EMIT(json.dumps(line))
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


def rollout_measure(rank, world, local_rank, hier, K, W, blocks=1):
    """On-device rollout collection (BASELINE cfg 5): 16384 envs x 8 steps per GPU per iteration (= 1 M env-steps per
    iteration on 8 GPUs): fused tcgen05 policy / value / sampling kernel -> fused env step, 2 launches per step captured
    in one CUDA graph, + GAE (ilrl_gae).  hier=True: the same for the hierarchical env (both policies on device, 5 launches
    per tick, horizon 10 = rollout_fragment_length of REF train_config.py:257; counts low-level env steps).
    A "step" is one iteration.  The process group (world > 1) is the caller's."""
    import torch
    import torch.distributed as dist
    import ilrl_b200
    from ilrl_b200 import BatchedHumanoidEnv, GaussianMLPPolicy, HierRolloutCollector, RolloutCollector
    dev = torch.device("cuda", local_rank)
    n, T = 16384, (10 if hier else 8)
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
    block_ms = []
    for _ in range(blocks):   # median of `blocks` timed blocks of exactly K iterations (max over ranks per block)
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
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        block_ms.append(float(t_ms.item()))
    st = ilrl_b200.stats.allreduce_stats(env.stats())
    ms = sorted(block_ms)[len(block_ms) // 2]
    summ = ilrl_b200.stats.summarize(st)
    out = {"metric": "rollout env-steps/sec (%spolicy + physics + reward + GAE on device)" % ("both policies: " if hier else ""),
           "value": world * n * T * K / (ms * 1e-3), "unit": UNIT, "steps": K, "warmup": W, "ms_per_step": ms / K,
           "blocks": blocks, "block_ms": [round(b, 3) for b in block_ms],
           "workload": ("on-device PPO rollout collection, hierarchical env: %d envs x %d low-level steps per GPU per "
                        "iteration, selected_motion=1, step_per_level=5, 44-256-256-2 and 70-256-256-17 tanh Gaussian "
                        "policies + value nets (fused tcgen05 kernel), gamma 0.99 lambda 0.9" % (n, T)) if hier else
                       ("on-device PPO rollout collection: %d envs x %d steps per GPU per iteration, %s, "
                        "70-256-256-17 tanh Gaussian policy + value net (fused tcgen05 kernel, bf16 operands / fp32 accumulation), "
                        "gamma 0.99 lambda 0.9" % (n, T, CLIP)),
           "envs_per_gpu": n, "horizon": T, "episode_len_mean": summ["episode_len_mean"],
           "sample_batch_columns": sorted(batch["low"].keys()) if hier else sorted(batch.keys()),
           # per iteration, replayed from the graph: low: T env steps + (T + 1) policy steps, + 1 GAE launch;
           # hier: 4 launches per tick + readout and two value calls after the last + 2 GAE launches
           "gpu_launches": K * ((5 * T + 5) if hier else (2 * T + 2))}
    env.close()
    del col
    torch.cuda.empty_cache()
    return out


def run_rollout(args, hier=False):
    """Extra measurement mode (BASELINE cfg 5), see rollout_measure."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    r = rollout_measure(rank, world, local_rank, hier, min(args.steps, 200), max(3, min(args.warmup, 20)))
    if rank == 0:
        EMIT(json.dumps({
            "metric": r["metric"], "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": r["steps"], "warmup": r["warmup"],
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": r["workload"], "envs_per_gpu": r["envs_per_gpu"], "horizon": r["horizon"],
                       "episode_len_mean": r["episode_len_mean"], "sample_batch_columns": r["sample_batch_columns"]},
            "gpu_launches": r["gpu_launches"]}))
    if world > 1:
        dist.destroy_process_group()


def main():
    global EMIT
    EMIT = _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the steps of a timed block one by one instead of replaying a CUDA graph")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the short cfg 3 / 4 / 5 runs of the default line")
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
