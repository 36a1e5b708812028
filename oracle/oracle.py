"""ctypes binding of oracle/ilrl_oracle.c — TEST INFRASTRUCTURE ONLY.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs (as the
checker or the CPU baseline).  The product package never imports this module.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, "_build", "libilrl_oracle.so")
_DATA = os.path.join(_ROOT, "imitation-learning-rl_b200", "data")

PHYS_WORDS, ENV_WORDS, TERM_WORDS = 47, 28, 12
CLIPS = ["motion02_04", "motion08_03", "motion09_03", "motion13_13"]
# envf word indices (ilrl_constants.h)
(E_FRAME, E_CLIP, E_T, E_TARGET_X, E_TARGET_Y, E_START_X, E_START_Y, E_SEP_X, E_SEP_Y, E_SEP_Z, E_ROBOT_X, E_ROBOT_Y,
 E_HLDEG, E_WALK_X, E_WALK_Y, E_LOW_TARGET_SCORE, E_JOINT_SCORE, E_JVEL_SCORE, E_POSTURE_SCORE, E_OBS_SIN, E_OBS_COS,
 E_STEPS_REMAINING, E_CUM_DRIFT, E_HIGH_TARGET_SCORE, E_CUM_ALIVE, E_HIGH_PENDING, E_EP_RETURN,
 E_EP_LEN) = range(28)


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "ilrl_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.ilrl_oracle_fk.argtypes = [_dp, _dp, _dp, _dp]
        L.ilrl_oracle_energy.argtypes = [_dp]
        L.ilrl_oracle_energy.restype = C.c_double
        L.ilrl_oracle_physics_step.argtypes = [_dp, _dp, C.c_int]
        L.ilrl_oracle_substep.argtypes = [_dp, _dp, C.c_double, C.c_int]
        L.ilrl_oracle_action_to_torque.argtypes = [_dp, _dp]
        L.ilrl_oracle_calc_state.argtypes = [_dp, C.c_double, C.c_double, _fp, _dp, _fp, _ip, _dp]
        L.ilrl_oracle_env_create.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int]
        L.ilrl_oracle_env_create.restype = C.c_void_p
        L.ilrl_oracle_env_destroy.argtypes = [C.c_void_p]
        L.ilrl_oracle_env_get.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.ilrl_oracle_env_set.argtypes = [C.c_void_p, _dp, _dp]
        L.ilrl_oracle_env_reset.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, _dp]
        L.ilrl_oracle_low_obs.argtypes = [C.c_void_p, _dp]
        L.ilrl_oracle_high_obs.argtypes = [C.c_void_p, _dp]
        L.ilrl_oracle_low_step.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, _dp, _dp]
        L.ilrl_oracle_low_step.restype = C.c_int
        L.ilrl_oracle_high_step.argtypes = [C.c_void_p, _dp, _dp]
        L.ilrl_oracle_hier_low_step.argtypes = [C.c_void_p, _dp, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
        L.ilrl_oracle_hier_low_step.restype = C.c_int
        L.ilrl_oracle_endpoint_score.argtypes = [C.c_void_p]
        L.ilrl_oracle_endpoint_score.restype = C.c_double
        L.ilrl_oracle_env_refresh.argtypes = [C.c_void_p]
        L.ilrl_oracle_set_heightfield.argtypes = [_dp, C.c_int, C.c_int, C.c_double]
        L.ilrl_oracle_set_self_collision.argtypes = [C.c_int]
        L.ilrl_oracle_self_collision_stats.argtypes = [C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.ilrl_oracle_self_collision_min_sep.argtypes = [C.c_int]
        L.ilrl_oracle_self_collision_min_sep.restype = C.c_double
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def load_model():
    with open(os.path.join(_DATA, "model.json")) as f:
        return json.load(f)


_clip_cache = {}


def load_clip(name):
    """-> dict(pos, rel, vel, ep float64 C-contiguous, max_frame).  max_frame = len(pos)-1 (REF low_level_env.py:80-82)
    clamped to the velocity table length for motion13_13 (declared divergence: the reference raises IndexError)."""
    if name not in _clip_cache:
        z = np.load(os.path.join(_DATA, "clips.npz"))
        d = {k: np.ascontiguousarray(z["%s_%s" % (name, k)], dtype=np.float64) for k in ("pos", "rel", "vel", "ep")}
        d["max_frame"] = int(min(len(d["pos"]) - 1, len(d["vel"])))
        _clip_cache[name] = d
    return _clip_cache[name]


def default_phys():
    p = np.zeros(PHYS_WORDS)
    p[2] = 1.4
    p[6] = 1.0
    return p


def fk(phys):
    phys = np.ascontiguousarray(phys, dtype=np.float64)
    bo, ao, br = np.zeros((15, 3)), np.zeros((17, 3)), np.zeros((15, 9))
    lib().ilrl_oracle_fk(_d(phys), _d(bo), _d(ao), _d(br))
    return bo, ao, br.reshape(15, 3, 3)


_hf_keep = None


def heightfield_zoff(data, body_z=0.25):
    """world z of a sample = data + zoff: Bullet centres the shape on (min + max) / 2 and REF humanoid.py:130 puts the
    terrain body at z = 0.25"""
    d = np.asarray(data, dtype=np.float64)
    return body_z - 0.5 * (float(d.min()) + float(d.max()))


def set_heightfield(data, rows=256, cols=256, body_z=0.25):
    """Heightfield terrain for every physics call of this process (None = flat ground).  data[i + j * rows], as the
    reference's CustomScene.heightfieldData (REF humanoid.py:74)."""
    global _hf_keep
    if data is None:
        _hf_keep = None
        lib().ilrl_oracle_set_heightfield(None, 0, 0, 0.0)
        return
    _hf_keep = np.ascontiguousarray(data, dtype=np.float64).reshape(-1)
    assert _hf_keep.size == rows * cols
    lib().ilrl_oracle_set_heightfield(_d(_hf_keep), rows, cols, heightfield_zoff(_hf_keep, body_z))


def set_self_collision(on):
    """Bullet-style self-collision for every physics call of this process (off by default, as in the CUDA product)."""
    lib().ilrl_oracle_set_self_collision(int(bool(on)))


def self_collision_stats():
    """(self-contacts, substeps) counted since set_self_collision"""
    c, s = C.c_long(0), C.c_long(0)
    lib().ilrl_oracle_self_collision_stats(C.byref(c), C.byref(s))
    return c.value, s.value


def self_collision_min_sep(reset=True):
    """smallest distance between the axes of a colliding capsule pair seen since the last reset (1e30: none)"""
    return lib().ilrl_oracle_self_collision_min_sep(int(reset))


def energy(phys):
    phys = np.ascontiguousarray(phys, dtype=np.float64)
    return lib().ilrl_oracle_energy(_d(phys))


def physics_step(phys, tau, flags=0):
    phys = np.array(phys, dtype=np.float64)
    tau = np.ascontiguousarray(tau, dtype=np.float64)
    lib().ilrl_oracle_physics_step(_d(phys), _d(tau), flags)
    return phys


def substep(phys, tau, dt, flags=0):
    phys = np.array(phys, dtype=np.float64)
    tau = np.ascontiguousarray(tau, dtype=np.float64)
    lib().ilrl_oracle_substep(_d(phys), _d(tau), dt, flags)
    return phys


def action_to_torque(action):
    a = np.ascontiguousarray(action, dtype=np.float64)
    t = np.zeros(17)
    lib().ilrl_oracle_action_to_torque(_d(a), _d(t))
    return t


def calc_state(phys, wtx, wty):
    phys = np.ascontiguousarray(phys, dtype=np.float64)
    obs = np.zeros(42, np.float32)
    xyz = np.zeros(3)
    js = np.zeros(17, np.float32)
    lim = C.c_int(0)
    rpy = np.zeros(3)
    lib().ilrl_oracle_calc_state(_d(phys), wtx, wty, obs.ctypes.data_as(_fp), _d(xyz), js.ctypes.data_as(_fp),
                                 C.byref(lim), _d(rpy))
    return obs, xyz, js, lim.value, rpy


class OracleEnv:
    """One env of the restated hot path.  mode 0 = LowLevelHumanoidEnv, 1 = HierarchicalHumanoidEnv."""

    def __init__(self, clip="motion09_03", mode=0):
        self.L = lib()
        self.clip = load_clip(clip)
        self.mode = mode
        c = self.clip
        self.h = self.L.ilrl_oracle_env_create(mode, _d(c["pos"]), _d(c["rel"]), _d(c["vel"]), _d(c["ep"]),
                                               len(c["pos"]), len(c["vel"]), c["max_frame"])

    def __del__(self):
        try:
            self.L.ilrl_oracle_env_destroy(self.h)
        except Exception:
            pass

    def reset(self, start_frame, reset_yaw_deg, target_deg):
        obs = np.zeros(44 if self.mode == 1 else 70)
        self.L.ilrl_oracle_env_reset(self.h, int(start_frame), float(reset_yaw_deg), int(target_deg), _d(obs))
        return obs

    def get(self):
        p, e, t = np.zeros(PHYS_WORDS), np.zeros(ENV_WORDS), np.zeros(TERM_WORDS)
        self.L.ilrl_oracle_env_get(self.h, _d(p), _d(e), _d(t))
        return p, e, t

    def set(self, phys, envf):
        p = np.ascontiguousarray(phys, dtype=np.float64)
        e = np.zeros(ENV_WORDS)
        e[:len(envf)] = envf  # golden fixtures carry the first 26 words
        self.L.ilrl_oracle_env_set(self.h, _d(p), _d(e))

    def low_step(self, action, rand_deg=0, skip_physics=False):
        a = np.ascontiguousarray(action, dtype=np.float64)
        obs = np.zeros(70)
        rew = C.c_double(0)
        done = self.L.ilrl_oracle_low_step(self.h, _d(a), int(rand_deg), int(skip_physics), _d(obs), C.byref(rew))
        return obs, rew.value, bool(done)

    def high_step(self, action2):
        a = np.ascontiguousarray(action2, dtype=np.float64)
        obs = np.zeros(70)
        self.L.ilrl_oracle_high_step(self.h, _d(a), _d(obs))
        return obs

    def hier_low_step(self, action, rand_deg=0, skip_physics=False):
        a = np.ascontiguousarray(action, dtype=np.float64)
        lo, hi = np.zeros(70), np.zeros(44)
        lr, hr = C.c_double(0), C.c_double(0)
        ret = self.L.ilrl_oracle_hier_low_step(self.h, _d(a), int(rand_deg), int(skip_physics), _d(lo), C.byref(lr),
                                               _d(hi), C.byref(hr))
        return lo, lr.value, hi, hr.value, bool(ret & 1), bool(ret & 2)

    def endpoint_score(self):
        return self.L.ilrl_oracle_endpoint_score(self.h)
