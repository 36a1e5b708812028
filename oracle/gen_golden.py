#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference Python
(/root/reference, through oracle/ref_shim.py).  Runs only in the build container (the GPU box has no
/root/reference); the fixtures it writes are committed and travel.

  low_traj_motion09_03.npz  BASELINE cfg 1: one LowLevelHumanoidEnv, motion09_03, default_rng(0), 1000 random-action
                            steps, reset on done.  Physics = the C oracle (NOT PyBullet: "reference logic + restated
                            physics"); everything else is the reference's own code.
  low_injected.npz          reference `step` on randomly injected states (physics skipped), 4 clips x 400 states:
                            the reward / observation / frame-index / target / done parity vectors.
  hier_traj.npz             HierarchicalHumanoidEnv protocol trace with physics, 1000 low-level steps.
  hier_injected.npz         hier low/high steps on injected states, 2 x 400.
  reset_vectors.npz         resetFromFrame / reset outputs for fixed (start_frame, yaw, target) triples, low + hier.
  hier2_injected.npz        REF hier_env_2.py (row a18) reset / high_level_step / low_level_step on injected states, under
                            the two declared substitutions of ref_shim.make_hier2_env (data directory, robot stand-in).
  hier2_traj.npz            hier_env_2 protocol trace with physics (reset -> high -> 20 low -> high ...), 600 low steps.
  ref_policies.npz          weights of the low-level policies the reference trained in PyBullet (Log/Best Model) and its own
                            evaluation logs of them (Log/data_*.json): oracle/extract_ref_policies.py, DESIGN.md section 4.
  notebook_vectors.json     the recorded cell outputs of "Eksplor Ray RLLib.ipynb" that pin layout facts.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle import ref_shim as S  # noqa: E402

OUT = os.environ.get("ILRL_GOLDEN_OUT") or os.path.join(ROOT, "tests", "golden")  # (the reproducibility test writes elsewhere)
os.makedirs(OUT, exist_ok=True)
CLIPS = O.CLIPS
LO, HI = np.array(S.MODEL["joint_lo"]), np.array(S.MODEL["joint_hi"])


def f32(x):
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def random_phys(rng, upright=True):
    p = np.zeros(47)
    p[0:2] = rng.uniform(-6, 6, 2)
    p[2] = rng.uniform(0.55, 1.45)
    yaw = rng.uniform(-np.pi, np.pi)
    tilt = rng.uniform(-0.6, 0.6, 2) if upright else rng.uniform(-1.5, 1.5, 2)
    from scipy.spatial.transform import Rotation as R
    p[3:7] = (R.from_euler("z", yaw) * R.from_euler("xy", tilt)).as_quat()
    p[7:10] = rng.uniform(-2, 2, 3)
    p[10:13] = rng.uniform(-3, 3, 3)
    span = HI - LO
    p[13:30] = LO + span * rng.uniform(-0.03, 1.03, 17)
    p[30:47] = rng.uniform(-12, 12, 17) * rng.choice([0.1, 1.0, 5.0], 17)
    return f32(p)


def random_env_words(rng, phys, max_frame, hier):
    e = np.zeros(26)
    e[0] = rng.integers(0, max_frame - 1)
    e[2] = rng.choice([0, 1, 17, 500, 2998, 2999])
    bo, ao, _ = O.fk(phys)
    mean = (bo[:, :2].sum(0) + ao[:, :2].sum(0)) / 33.0
    mode = rng.uniform()
    if mode < 0.2:   # target about to be reached
        e[3:5] = mean + rng.uniform(-0.45, 0.45, 2)
    elif mode < 0.3:  # far from target: out-of-range termination
        e[3:5] = mean + rng.uniform(6, 12) * np.array([np.cos(mode * 50), np.sin(mode * 50)])
    else:
        th = rng.uniform(-np.pi, np.pi)
        e[3:5] = mean + rng.uniform(0.6, 6.5) * np.array([np.cos(th), np.sin(th)])
    e[5:7] = mean + rng.uniform(-3, 3, 2)
    e[7:10] = [mean[0] + rng.uniform(-1, 1), mean[1] + rng.uniform(-1, 1), 0.0]
    e[10:12] = mean + rng.uniform(-0.05, 0.05, 2)    # robot_pos of the previous step (used as-is only by hier)
    e[12] = rng.uniform(-np.pi, np.pi)
    e[13:15] = mean + 10 * np.array([np.cos(e[12]), np.sin(e[12])])
    e[15] = -np.linalg.norm(e[3:5] - mean) + rng.uniform(-0.05, 0.05)
    e[16:19] = rng.uniform(0, 1, 3)
    if hier:
        e[15] = 0.0
        e[21] = rng.integers(1, 6)
        e[22] = rng.uniform(0, 4)
        e[23] = -np.linalg.norm(e[3:5] - mean) + rng.uniform(-0.05, 0.05)
        e[24] = rng.choice([0, 2, 4, 6])
    return f32(e)


def gen_low_traj(clip="motion09_03", steps=1000, seed=0):
    env = S.make_low_env(clip, seed=seed, physics="oracle")
    arng = np.random.default_rng(seed + 1)
    rec = {k: [] for k in ["phys_before", "env_before", "action", "rand_deg", "phys_after", "env_after", "obs",
                           "reward", "done", "terms", "reset_before", "reset_start_frame", "reset_target_deg",
                           "reset_obs"]}
    need_reset = True
    for _ in range(steps):
        if need_reset:
            n0 = len(env.rng.log)
            robs = env.reset()
            sf, tdeg = env.rng.log[n0], env.rng.log[n0 + 1]
            rec["reset_before"].append(1); rec["reset_start_frame"].append(sf); rec["reset_target_deg"].append(tdeg)
            rec["reset_obs"].append(np.asarray(robs, dtype=np.float64))
        else:
            rec["reset_before"].append(0); rec["reset_start_frame"].append(-1); rec["reset_target_deg"].append(0)
            rec["reset_obs"].append(np.zeros(70))
        a = arng.uniform(-1, 1, 17)
        rec["phys_before"].append(env.flat_env.phys.copy())
        rec["env_before"].append(S.env_words(env))
        n0 = len(env.rng.log)
        obs, rew, done, _ = env.step(a)
        rec["rand_deg"].append(env.rng.log[n0] if len(env.rng.log) > n0 else -999)
        rec["action"].append(a)
        rec["phys_after"].append(env.flat_env.phys.copy())
        rec["env_after"].append(S.env_words(env))
        rec["obs"].append(np.asarray(obs, dtype=np.float64)); rec["reward"].append(rew); rec["done"].append(done)
        rec["terms"].append(S.terms_of(env))
        need_reset = done
    out = {k: np.array(v) for k, v in rec.items()}
    np.savez_compressed(os.path.join(OUT, "low_traj_%s.npz" % clip), **out)
    print("low_traj", clip, "episodes", int(out["reset_before"].sum()), "mean reward %.3f" % out["reward"].mean(),
          "target switches", int((out["rand_deg"] != -999).sum()))


def gen_low_injected(per_clip=400, seed=10):
    rng = np.random.default_rng(seed)
    rec = {k: [] for k in ["clip", "phys", "env_before", "action", "rand_deg", "env_after", "obs", "reward", "done",
                           "terms", "endpoint_score"]}
    for ci, clip in enumerate(CLIPS):
        env = S.make_low_env(clip, seed=seed, physics="none")
        env.reset()
        mf = O.load_clip(clip)["max_frame"]
        for _ in range(per_clip):
            phys = random_phys(rng)
            e = random_env_words(rng, phys, mf, hier=False)
            a = f32(rng.uniform(-1.3, 1.3, 17))
            deg = int(rng.integers(-180, 180))
            env.flat_env.phys[:] = phys
            S.set_env_words(env, e)
            env.cur_obs = env.flat_env.robot.calc_state()
            rec["endpoint_score"].append(env.calcEndPointScore(useExp=True))
            env.rng.forced = [deg]
            obs, rew, done, _ = env.step(a)
            used = len(env.rng.forced) == 0
            env.rng.forced = []
            rec["clip"].append(ci); rec["phys"].append(phys); rec["env_before"].append(e); rec["action"].append(a)
            rec["rand_deg"].append(deg if used else -999)
            rec["env_after"].append(S.env_words(env)); rec["obs"].append(np.asarray(obs, dtype=np.float64))
            rec["reward"].append(rew); rec["done"].append(done); rec["terms"].append(S.terms_of(env))
    out = {k: np.array(v) for k, v in rec.items()}
    np.savez_compressed(os.path.join(OUT, "low_injected.npz"), **out)
    print("low_injected", out["phys"].shape, "done frac %.2f" % out["done"].mean(), "switches",
          int((out["rand_deg"] != -999).sum()))


def gen_hier_traj(low_steps=1000, seed=3):
    env = S.make_hier_env(seed=seed, physics="oracle")
    arng = np.random.default_rng(seed + 1)
    K = ["kind", "phys_before", "env_before", "action", "rand_deg", "phys_after", "env_after", "low_obs", "low_reward",
         "high_obs", "high_reward", "flags", "terms", "draws"]
    rec = {k: [] for k in K}

    def push(kind, pb, eb, a, deg, lo, lr, ho, hr, flags, draws=(0, 0, 0)):
        aa = np.zeros(17); aa[:len(a)] = a
        rec["kind"].append(kind); rec["phys_before"].append(pb); rec["env_before"].append(eb); rec["action"].append(aa)
        rec["rand_deg"].append(deg); rec["phys_after"].append(env.flat_env.phys.copy())
        rec["env_after"].append(S.env_words(env, hier=True))
        rec["low_obs"].append(np.zeros(70) if lo is None else np.asarray(lo, dtype=np.float64))
        rec["low_reward"].append(lr)
        rec["high_obs"].append(np.zeros(44) if ho is None else np.asarray(ho, dtype=np.float64))
        rec["high_reward"].append(hr); rec["flags"].append(flags); rec["terms"].append(S.terms_of(env))
        rec["draws"].append(list(draws))

    n_low = 0
    need_reset, need_high = True, False
    while n_low < low_steps:
        pb, eb = env.flat_env.phys.copy(), S.env_words(env, hier=True)
        if need_reset:
            n0 = len(env.rng.log)
            obs = env.reset()
            push(0, pb, eb, [], -999, None, 0.0, obs["high_level_agent"], 0.0, 0, env.rng.log[n0:n0 + 3])
            need_reset, need_high = False, True
        elif need_high:
            a = arng.uniform(-1, 1, 2)
            obs, rew, done, _ = env.step({"high_level_agent": a})
            assert list(obs) == ["low_level_agent"] and rew["low_level_agent"] == 0 and not done["__all__"]
            push(1, pb, eb, a, -999, obs["low_level_agent"], 0.0, None, 0.0, 0)
            need_high = False
        else:
            a = arng.uniform(-1, 1, 17)
            n0 = len(env.rng.log)
            obs, rew, done, _ = env.step({"low_level_agent": a})
            deg = env.rng.log[n0] if len(env.rng.log) > n0 else -999
            has_high = "high_level_agent" in obs
            has_low = "low_level_agent" in obs
            flags = int(done["__all__"]) | (int(has_high) << 1) | (int(has_low) << 2)
            push(2, pb, eb, a, deg, obs.get("low_level_agent"), rew.get("low_level_agent", 0.0),
                 obs.get("high_level_agent"), rew.get("high_level_agent", 0.0), flags)
            n_low += 1
            need_reset = done["__all__"]
            need_high = has_high and not done["__all__"]
    out = {k: np.array(v) for k, v in rec.items()}
    np.savez_compressed(os.path.join(OUT, "hier_traj.npz"), **out)
    print("hier_traj records", len(out["kind"]), "resets", int((out["kind"] == 0).sum()), "high", int((out["kind"] == 1).sum()))


def gen_hier_injected(n=400, seed=20):
    rng = np.random.default_rng(seed)
    env = S.make_hier_env(seed=seed, physics="none")
    env.reset()
    mf = O.load_clip("motion09_03")["max_frame"]
    rec = {k: [] for k in ["kind", "phys", "env_before", "action", "rand_deg", "env_after", "low_obs", "low_reward",
                           "high_obs", "high_reward", "flags", "terms", "obs_sincos"]}
    for kind in (1, 2):
        for _ in range(n):
            phys = random_phys(rng)
            e = random_env_words(rng, phys, mf, hier=True)
            env.flat_env.phys[:] = phys
            S.set_env_words(env, e, hier=True)
            env.aliveReward = 0
            env.cur_obs = env.flat_env.robot.calc_state()   # also refreshes robot.body_xyz
            sc = np.array(env.cur_obs[1:3], dtype=np.float64)
            aa = np.zeros(17)
            if kind == 1:
                a = f32(rng.uniform(-1, 1, 2)); aa[:2] = a
                obs, rew, done, _ = env.step({"high_level_agent": a})
                lo, lr, ho, hr, flags, deg = obs["low_level_agent"], 0.0, np.zeros(44), 0.0, 0, -999
            else:
                a = f32(rng.uniform(-1.3, 1.3, 17)); aa[:] = a
                deg = int(rng.integers(-180, 180))
                env.rng.forced = [deg]
                obs, rew, done, _ = env.step({"low_level_agent": a})
                used = len(env.rng.forced) == 0
                env.rng.forced = []
                deg = deg if used else -999
                has_high, has_low = "high_level_agent" in obs, "low_level_agent" in obs
                flags = int(done["__all__"]) | (int(has_high) << 1) | (int(has_low) << 2)
                lo = obs.get("low_level_agent", np.zeros(70)); lr = rew.get("low_level_agent", 0.0)
                ho = obs.get("high_level_agent", np.zeros(44)); hr = rew.get("high_level_agent", 0.0)
            rec["kind"].append(kind); rec["phys"].append(phys); rec["env_before"].append(e); rec["action"].append(aa)
            rec["rand_deg"].append(deg); rec["env_after"].append(S.env_words(env, hier=True))
            rec["low_obs"].append(np.asarray(lo, dtype=np.float64)); rec["low_reward"].append(lr)
            rec["high_obs"].append(np.asarray(ho, dtype=np.float64)); rec["high_reward"].append(hr)
            rec["flags"].append(flags); rec["terms"].append(S.terms_of(env)); rec["obs_sincos"].append(sc)
    out = {k: np.array(v) for k, v in rec.items()}
    np.savez_compressed(os.path.join(OUT, "hier_injected.npz"), **out)
    print("hier_injected", out["phys"].shape, "flags hist", np.bincount(out["flags"]))


def gen_reset_vectors(seed=30):
    rng = np.random.default_rng(seed)
    rec = {k: [] for k in ["mode", "clip", "start_frame", "yaw", "target_deg", "obs", "phys_after", "env_after",
                           "sep_before"]}
    for ci, clip in enumerate(CLIPS):
        env = S.make_low_env(clip, seed=seed, physics="none")
        mf = O.load_clip(clip)["max_frame"]
        for _ in range(24):
            sf, yaw, deg = int(rng.integers(0, mf - 5)), float(rng.choice([0, 0, 30, -77])), int(rng.integers(-180, 180))
            env.rng.forced = [deg]
            obs = env.resetFromFrame(startFrame=sf, resetYaw=yaw, startFromRef=True, initVel=True)
            pad = np.zeros(70); pad[:] = obs
            rec["mode"].append(0); rec["clip"].append(ci); rec["start_frame"].append(sf); rec["yaw"].append(yaw)
            rec["target_deg"].append(deg); rec["obs"].append(pad); rec["phys_after"].append(env.flat_env.phys.copy())
            rec["env_after"].append(S.env_words(env)); rec["sep_before"].append(np.zeros(3))
    env = S.make_hier_env(seed=seed, physics="none")
    mf = O.load_clip("motion09_03")["max_frame"]
    for _ in range(24):
        sf, yaw, deg = int(rng.integers(0, mf - 5)), int(rng.integers(-180, 180)), int(rng.integers(-180, 180))
        sep = f32(rng.uniform(-2, 2, 3)); sep[2] = 0
        env.starting_ep_pos = sep.copy()
        env.rng.forced = [sf, yaw, deg]
        obs = env.reset()["high_level_agent"]
        pad = np.zeros(70); pad[:44] = obs
        rec["mode"].append(1); rec["clip"].append(2); rec["start_frame"].append(sf); rec["yaw"].append(yaw)
        rec["target_deg"].append(deg); rec["obs"].append(pad); rec["phys_after"].append(env.flat_env.phys.copy())
        rec["env_after"].append(S.env_words(env, hier=True)); rec["sep_before"].append(sep)
    out = {k: np.array(v) for k, v in rec.items()}
    np.savez_compressed(os.path.join(OUT, "reset_vectors.npz"), **out)
    print("reset_vectors", out["obs"].shape)


def _pad(v, n):
    out = np.zeros(n)
    if v is not None:
        v = np.asarray(v, dtype=np.float64).ravel()
        out[:len(v)] = v
    return out


def gen_hier2_injected(n=400, n_reset=48, seed=40):
    rng = np.random.default_rng(seed)
    env = S.make_hier2_env(seed=seed, physics="none")
    env.reset()
    mf = O.load_clip("motion09_03")["max_frame"]
    K = ["kind", "phys", "env_before", "jt_before", "action", "rand_deg", "draws", "phys_after", "env_after", "jt_after",
         "low_obs", "low_reward", "high_obs", "high_reward", "flags", "terms", "obs_sincos"]
    rec = {k: [] for k in K}
    for kind in (0, 1, 2):
        for _ in range(n_reset if kind == 0 else n):
            phys = random_phys(rng)
            e = random_env_words(rng, phys, mf, hier=True)
            e[21] = 1 if rng.uniform() < 0.3 else rng.integers(1, 21)
            e[15] = rng.uniform(0, 10); e[24] = rng.uniform(0, 10)
            e = f32(e)
            jt = f32(rng.uniform(-1.5, 1.5, 34))
            env.flat_env.phys[:] = phys
            S.set_env_words2(env, e)
            env.jointTarget = jt.copy()
            env.aliveReward = 0
            env.cur_obs = env.flat_env.robot.calc_state()   # also refreshes robot.body_xyz
            sc = np.array(env.cur_obs[1:3], dtype=np.float64)
            aa, deg, draws = np.zeros(36), -999, [0, 0, 0]
            lo = lr = ho = hr = None
            flags = 0
            if kind == 0:
                draws = [int(rng.integers(0, mf - 5)), int(rng.integers(-180, 180)), int(rng.integers(-180, 180))]
                env.rng.forced = list(draws)
                ho = env.reset()["high_level_agent"]
            elif kind == 1:
                aa = f32(rng.uniform(-1, 1, 36))
                obs, rew, done, _ = env.step({"high_level_agent": aa})
                assert list(obs) == ["low_level_agent"] and rew["low_level_agent"] == 0 and not done["__all__"]
                lo = obs["low_level_agent"]
            else:
                aa[:17] = f32(rng.uniform(-1.3, 1.3, 17))
                deg = int(rng.integers(-180, 180))
                env.rng.forced = [deg]
                obs, rew, done, _ = env.step({"low_level_agent": aa[:17].copy()})
                deg = deg if len(env.rng.forced) == 0 else -999
                env.rng.forced = []
                has_high, has_low = "high_level_agent" in obs, "low_level_agent" in obs
                flags = int(done["__all__"]) | (int(has_high) << 1) | (int(has_low) << 2)
                lo, lr = obs.get("low_level_agent"), rew.get("low_level_agent")
                ho, hr = obs.get("high_level_agent"), rew.get("high_level_agent")
            rec["kind"].append(kind); rec["phys"].append(phys); rec["env_before"].append(e); rec["jt_before"].append(jt)
            rec["action"].append(aa); rec["rand_deg"].append(deg); rec["draws"].append(draws)
            rec["phys_after"].append(env.flat_env.phys.copy()); rec["env_after"].append(S.env_words2(env))
            rec["jt_after"].append(np.asarray(env.jointTarget, dtype=np.float64))
            rec["low_obs"].append(_pad(lo, 72)); rec["low_reward"].append(0.0 if lr is None else lr)
            rec["high_obs"].append(_pad(ho, 60)); rec["high_reward"].append(0.0 if hr is None else hr)
            rec["flags"].append(flags); rec["terms"].append(S.terms_of2(env)); rec["obs_sincos"].append(sc)
    out = {k: np.array(v) for k, v in rec.items()}
    np.savez_compressed(os.path.join(OUT, "hier2_injected.npz"), **out)
    lowk = out["kind"] == 2
    print("hier2_injected", out["phys"].shape, "low-step flags hist", np.bincount(out["flags"][lowk]), "switches",
          int((out["rand_deg"] != -999).sum()))


def gen_hier2_traj(low_steps=600, seed=41):
    env = S.make_hier2_env(seed=seed, physics="oracle")
    arng = np.random.default_rng(seed + 1)
    K = ["kind", "phys_before", "env_before", "jt_before", "action", "rand_deg", "draws", "phys_after", "env_after",
         "low_obs", "low_reward", "high_obs", "high_reward", "flags", "terms"]
    rec = {k: [] for k in K}
    n_low = 0
    need_reset, need_high = True, False
    while n_low < low_steps:
        pb, eb, jb = env.flat_env.phys.copy(), S.env_words2(env), _pad(env.jointTarget, 34)
        aa, deg, draws, flags = np.zeros(36), -999, [0, 0, 0], 0
        lo = lr = ho = hr = None
        if need_reset:
            n0 = len(env.rng.log)
            ho = env.reset()["high_level_agent"]
            draws = env.rng.log[n0:n0 + 3]
            kind, need_reset, need_high = 0, False, True
        elif need_high:
            aa = arng.uniform(-1, 1, 36)
            obs, rew, done, _ = env.step({"high_level_agent": aa})
            lo = obs["low_level_agent"]
            kind, need_high = 1, False
        else:
            aa[:17] = arng.uniform(-1, 1, 17)
            n0 = len(env.rng.log)
            obs, rew, done, _ = env.step({"low_level_agent": aa[:17].copy()})
            deg = env.rng.log[n0] if len(env.rng.log) > n0 else -999
            has_high, has_low = "high_level_agent" in obs, "low_level_agent" in obs
            flags = int(done["__all__"]) | (int(has_high) << 1) | (int(has_low) << 2)
            lo, lr = obs.get("low_level_agent"), rew.get("low_level_agent")
            ho, hr = obs.get("high_level_agent"), rew.get("high_level_agent")
            kind = 2
            n_low += 1
            need_reset = done["__all__"]
            need_high = has_high and not done["__all__"]
        rec["kind"].append(kind); rec["phys_before"].append(pb); rec["env_before"].append(eb); rec["jt_before"].append(jb)
        rec["action"].append(aa); rec["rand_deg"].append(deg); rec["draws"].append(list(draws))
        rec["phys_after"].append(env.flat_env.phys.copy()); rec["env_after"].append(S.env_words2(env))
        rec["low_obs"].append(_pad(lo, 72)); rec["low_reward"].append(0.0 if lr is None else lr)
        rec["high_obs"].append(_pad(ho, 60)); rec["high_reward"].append(0.0 if hr is None else hr)
        rec["flags"].append(flags); rec["terms"].append(S.terms_of2(env))
    out = {k: np.array(v) for k, v in rec.items()}
    np.savez_compressed(os.path.join(OUT, "hier2_traj.npz"), **out)
    print("hier2_traj records", len(out["kind"]), "resets", int((out["kind"] == 0).sum()), "high", int((out["kind"] == 1).sum()),
          "high outcomes", int(((out["flags"] & 2) != 0).sum()))


def gen_notebook_vectors():
    nb = json.load(open(os.path.join(S.REF, "Eksplor Ray RLLib.ipynb")))
    found = {}
    for cell in nb["cells"]:
        for o in cell.get("outputs", []):
            txt = "".join(o.get("data", {}).get("text/plain", [])) + "".join(o.get("text", []))
            if txt.startswith("array([ 0.37      ,  0.12967981"):
                found["reset_obs_frame0_78"] = [float(x) for x in
                                                txt.replace("array(", "").replace(")", "").strip("[] \n").replace("\n", " ").split(",")]
            if "-0.38268343" in txt and "1.17" in txt and "reset_pose" not in found:
                found["reset_pose"] = [float(x) for x in txt.strip("[] \n").split()]
            if "'floor'" in txt and "parts_keys_has_floor" not in found:
                found["parts_keys_has_floor"] = True
            if "right_shoulder_y" in txt and "jdict_order" not in found and "abdomen_z" in txt:
                import re
                names = re.findall(r"'([a-z_]+)':", txt)
                if len(names) >= 17:
                    found["jdict_order"] = names
    found["source"] = "recorded cell outputs of /root/reference/Eksplor Ray RLLib.ipynb (SURVEY.md section 4)"
    json.dump(found, open(os.path.join(OUT, "notebook_vectors.json"), "w"), indent=1)
    print("notebook vectors:", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in found.items()})


if __name__ == "__main__":
    which = sys.argv[1:] or ["nb", "reset", "low_inj", "hier_inj", "low_traj", "hier_traj", "hier2_inj", "hier2_traj", "ref_policies"]
    if "nb" in which: gen_notebook_vectors()
    if "reset" in which: gen_reset_vectors()
    if "low_inj" in which: gen_low_injected()
    if "hier_inj" in which: gen_hier_injected()
    if "low_traj" in which: gen_low_traj()
    if "hier_traj" in which: gen_hier_traj()
    if "hier2_inj" in which: gen_hier2_injected()
    if "hier2_traj" in which: gen_hier2_traj()
    if "ref_policies" in which:   # weights + PyBullet evaluation logs of the reference's shipped checkpoints (data files)
        from oracle import extract_ref_policies
        extract_ref_policies.OUT = OUT
        extract_ref_policies.main()
