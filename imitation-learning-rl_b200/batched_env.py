"""BatchedHumanoidEnv — N humanoid imitation envs on one B200 behind the C ABI (include/ilrl.h).

Host-side plumbing only: torch owns the device buffers and the stream, every computation happens in
libilrl_b200.so's CUDA kernels.  There is no CPU path: constructing this class without a CUDA device raises.
The single-env, reference-shaped views (`LowLevelHumanoidEnv`, `HierarchicalHumanoidEnv`) and the RLlib-shaped
adapters are thin wrappers around this class.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .clips import CLIP_NAMES, load_clip

OBS_LOW, OBS_HIGH, ACT_LOW, ACT_HIGH = 70, 44, 17, 2
OBS_LOW2, OBS_HIGH2, ACT_HIGH2, JT_WORDS = 72, 60, 36, 34   # mode "hier2" (REF hier_env_2.py:48-60)
PHYS_WORDS, ENV_WORDS, TERM_WORDS, STATS_WORDS = 47, 28, 12, 16
INT32_MIN = -2 ** 31
TERM_NAMES = ["deltaJoints", "deltaVelJoints", "delta_lowTargetScore", "electricityScore", "jointLimitScore",
              "aliveReward", "bodyPostureScore", "lowTargetScore", "deltaEndPoints", "highTargetScore", "driftScore",
              "delta_highTargetScore"]
# envf word indices (csrc/ilrl_constants.h)
(E_FRAME, E_CLIP, E_T, E_TARGET_X, E_TARGET_Y, E_START_X, E_START_Y, E_SEP_X, E_SEP_Y, E_SEP_Z, E_ROBOT_X, E_ROBOT_Y,
 E_HLDEG, E_WALK_X, E_WALK_Y, E_LOW_TARGET_SCORE, E_JOINT_SCORE, E_JVEL_SCORE, E_POSTURE_SCORE, E_OBS_SIN, E_OBS_COS,
 E_STEPS_REMAINING, E_CUM_DRIFT, E_HIGH_TARGET_SCORE, E_CUM_ALIVE, E_HIGH_PENDING, E_EP_RETURN,
 E_EP_LEN) = range(28)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchedHumanoidEnv:
    """mode "low": LowLevelHumanoidEnv semantics (REF low_level_env.py); "hier": HierarchicalHumanoidEnv
    (REF hier_env.py); "hier2": the hier_env_2.py variant (joint-target tracking low level: low obs 72, high obs 60,
    high action 36, skipFrame 5, step_per_level 20 unless given).  clips: list of clip names staged in HBM; clip_of_env: per-env index into that list.
    self_collision: Bullet-style self-collision of the robot (REF humanoid.py:13; off by default, `set_self_collision`).
    env_id_base: global id of env 0 when a batch is sharded over several handles / GPUs (with the same seed the
    shards then reproduce exactly what one handle holding the whole batch would do)."""

    def __init__(self, num_envs, mode="low", clips=("motion09_03",), clip_of_env=None, device=0, seed=0,
                 auto_reset=True, max_timestep=3000, step_per_level=None, env_id_base=0, skip_frame=None,
                 self_collision=False):
        if not torch.cuda.is_available():
            raise _lib.IlrlError("BatchedHumanoidEnv needs a CUDA device (sm_100a); there is no CPU fallback")
        self.L = _lib.lib()
        self.num_envs = int(num_envs)
        self.mode = {"low": 0, "hier": 1, "hier2": 2}[mode]
        self.obs_w, self.hobs_w, self.hact_w = (OBS_LOW2, OBS_HIGH2, ACT_HIGH2) if self.mode == 2 else (OBS_LOW, OBS_HIGH, ACT_HIGH)
        self.device = torch.device("cuda", device)
        self.clip_names = list(clips)
        cfg = _lib.Config(device=device, num_envs=self.num_envs, mode=self.mode, auto_reset=int(bool(auto_reset)),
                          seed=seed, skip_frame=int(skip_frame or 0), max_timestep=max_timestep,
                          step_per_level=int(step_per_level or 0),
                          env_id_base=int(env_id_base))
        h = C.c_void_p()
        rc = self.L.ilrl_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise _lib.IlrlError("ilrl_create failed (%d): %s" % (rc, self.L.ilrl_last_error(None).decode()))
        self.h = h
        self.max_frame = []
        for ci, name in enumerate(self.clip_names):
            c = load_clip(name)
            self._ck(self.L.ilrl_load_clip(self.h, ci, c["pos"].ctypes.data, len(c["pos"]), c["rel"].ctypes.data,
                                           len(c["rel"]), c["vel"].ctypes.data, len(c["vel"]), c["ep"].ctypes.data,
                                           len(c["ep"]), c["max_frame"]))
            self.max_frame.append(c["max_frame"])
        ids = None
        if clip_of_env is not None:
            ids = np.ascontiguousarray(clip_of_env, dtype=np.int32)
            assert ids.shape == (self.num_envs,)
        self._ck(self.L.ilrl_set_clip_ids(self.h, None if ids is None else ids.ctypes.data))
        n, dev = self.num_envs, self.device
        self.obs = torch.zeros(n, self.obs_w, device=dev)
        self.reward = torch.zeros(n, device=dev)
        self.done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.terms = torch.zeros(n, TERM_WORDS, device=dev)
        if self.mode >= 1:
            self.high_obs = torch.zeros(n, self.hobs_w, device=dev)
            self.high_reward = torch.zeros(n, device=dev)
            self.high_flags = torch.zeros(n, dtype=torch.uint8, device=dev)
        if self_collision:
            self._ck(self.L.ilrl_set_self_collision(self.h, 1))
        self._forced = None
        self._noise = None
        self._host_key = None
        self._async_key = None
        self._pull_buf = None

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc != 0:
            raise _lib.IlrlError("ilrl call failed (%d): %s" % (rc, self.L.ilrl_last_error(self.h).decode()))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "h", None) is not None:
            self.L.ilrl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _f32(self, x, shape):
        t = torch.as_tensor(x, dtype=torch.float32, device=self.device).contiguous()
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), shape)
        return t

    # ------------------------------------------------------------------ env API (device tensors)
    def reset(self, mask=None, start_frame=None, target_deg=None, reset_yaw_deg=None, target_xy=None):
        """reset() / resetFromFrame() of the masked envs (all if mask is None).  Returns the obs tensor
        ([N,70] low-level obs in "low" mode, [N,44] high-level obs in "hier" mode); unmasked rows keep old values."""
        n = self.num_envs
        m = None if mask is None else torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        sf = None if start_frame is None else torch.as_tensor(start_frame, device=self.device).to(torch.int32).contiguous()
        td = None if target_deg is None else torch.as_tensor(target_deg, device=self.device).to(torch.int32).contiguous()
        yw = None if reset_yaw_deg is None else self._f32(reset_yaw_deg, (n,))
        xy = None if target_xy is None else self._f32(target_xy, (n, 2))
        out = self.high_obs if self.mode >= 1 else self.obs
        self._ck(self.L.ilrl_reset(self.h, _ptr(m), _ptr(sf), _ptr(td), _ptr(yw), _ptr(xy), _ptr(out), self._stream()))
        if self.mode >= 1:
            self._ck(self.L.ilrl_high_readout(self.h, None, None, _ptr(self.high_flags), self._stream()))
        return out

    def step(self, action, physics=True):
        """One low-level step of every env.  action [N,17] (device tensor).  -> (obs [N,70], reward [N], done [N] u8,
        terms [N,12]) — views of buffers owned by this object, overwritten by the next call."""
        a = self._f32(action, (self.num_envs, ACT_LOW))
        fn = self.L.ilrl_step if physics else self.L.ilrl_step_no_physics
        self._ck(fn(self.h, _ptr(a), _ptr(self.obs), _ptr(self.reward), _ptr(self.done), _ptr(self.terms), self._stream()))
        return self.obs, self.reward, self.done, self.terms

    def step_into(self, action, obs, reward, done, terms=None):
        """`step` writing straight into caller-owned device tensors (rollout buffers): obs [N,70] f32, reward [N] f32,
        done [N] u8, optional terms [N,12]; all contiguous, on this env's device."""
        n = self.num_envs
        for t, shape, dt in ((action, (n, ACT_LOW), torch.float32), (obs, (n, self.obs_w), torch.float32),
                             (reward, (n,), torch.float32), (done, (n,), torch.uint8)) + (
                                 () if terms is None else ((terms, (n, TERM_WORDS), torch.float32),)):
            assert t.is_cuda and t.device == self.device and t.dtype == dt and t.is_contiguous() and tuple(t.shape) == shape
        self._ck(self.L.ilrl_step(self.h, _ptr(action), _ptr(obs), _ptr(reward), _ptr(done), _ptr(terms), self._stream()))

    def step_sequence(self, actions, obs, reward, done, terms=None):
        """K low-level steps in one launch (mode "low"): actions [K,N,17] -> obs [K,N,70], reward [K,N], done [K,N] u8,
        optional terms [K,N,12], all caller-owned contiguous device tensors.  Same results as K `step_into` calls."""
        k, n = int(actions.shape[0]), self.num_envs
        for t, shape, dt in ((actions, (k, n, ACT_LOW), torch.float32), (obs, (k, n, self.obs_w), torch.float32),
                             (reward, (k, n), torch.float32), (done, (k, n), torch.uint8)) + (
                                 () if terms is None else ((terms, (k, n, TERM_WORDS), torch.float32),)):
            assert t.is_cuda and t.device == self.device and t.dtype == dt and t.is_contiguous() and tuple(t.shape) == shape
        self._ck(self.L.ilrl_step_sequence(self.h, k, _ptr(actions), _ptr(obs), _ptr(reward), _ptr(done), _ptr(terms),
                                           self._stream()))

    def step_host(self, action_np, obs_np, reward_np, done_np, terms_np=None):
        """Same step through host (numpy) buffers: H2D of the actions, kernel, D2H of obs/reward/done, synchronous."""
        # the argument tuple of the previous call is reused while the same buffers come back (the usual loop): dtype /
        # layout checks and pointer extraction cost more Python time than the ctypes call itself
        key = (id(action_np), id(obs_np), id(reward_np), id(done_np), id(terms_np))
        if self._host_key != key:
            assert action_np.dtype == np.float32 and action_np.flags.c_contiguous and action_np.shape == (self.num_envs, ACT_LOW)
            assert obs_np.dtype == np.float32 and obs_np.flags.c_contiguous and obs_np.shape == (self.num_envs, self.obs_w)
            assert reward_np.dtype == np.float32 and done_np.dtype == np.uint8
            assert reward_np.shape == (self.num_envs,) and done_np.shape == (self.num_envs,)
            assert terms_np is None or (terms_np.dtype == np.float32 and terms_np.shape == (self.num_envs, TERM_WORDS))
            ptr = lambda a: a.__array_interface__["data"][0]  # noqa: E731  (cheaper than ndarray.ctypes)
            self._host_args = (ptr(action_np), ptr(obs_np), ptr(reward_np), ptr(done_np),
                               None if terms_np is None else ptr(terms_np))
            self._host_refs = (action_np, obs_np, reward_np, done_np, terms_np)  # keep the ids valid
            self._host_key = key
        rc = self.L.ilrl_step_host(self.h, *self._host_args, self._stream())
        if rc != 0:
            self._ck(rc)

    def step_host_async(self, part, nparts, action_np, obs_np, reward_np, done_np, terms_np=None, wait_first=False):
        """Enqueue one step of part `part` of `nparts` (contiguous blocks of envs, see `part_slice`) on the part's own
        stream and return at once.  The arrays are the FULL [N, ...] pinned buffers (torch `pin_memory()`); the part's
        rows of obs / reward / done are valid after `wait(part)`.  Lets a worker prepare the actions of one part while
        the other steps (double-buffered rollout).  wait_first: first wait for the part's previous step (one call
        instead of `wait` + `step_host_async`; the caller must not need the previous outputs any more... they stay
        valid until the kernel of the new step overwrites them, i.e. read them BEFORE this call)."""
        key = (id(obs_np), id(reward_np), id(done_np), id(terms_np))
        if self._async_key != key:   # output buffers of the loop: pointers extracted once
            ptr = lambda a: None if a is None else a.__array_interface__["data"][0]  # noqa: E731
            self._async_out = (ptr(obs_np), ptr(reward_np), ptr(done_np), ptr(terms_np))
            self._async_refs = (obs_np, reward_np, done_np, terms_np)
            self._async_key = key
        fn = self.L.ilrl_wait_step_host_async if wait_first else self.L.ilrl_step_host_async
        rc = fn(self.h, part, nparts, action_np.__array_interface__["data"][0], *self._async_out)
        if rc != 0:
            self._ck(rc)

    def wait(self, part):
        rc = self.L.ilrl_wait(self.h, int(part))
        if rc != 0:
            self._ck(rc)

    # ---- persistent serving (ilrl_serve_*): one resident kernel, no launch / synchronise per step
    def serve_begin(self, obs_np, reward_np, done_np, terms_np=None, nparts=1):
        """Start the resident step kernels (one per part, `part_slice`) writing into these pinned [N,...] numpy buffers.
        Until `serve_end` only `serve_step` / `serve_post` / `serve_wait` may be called, and nothing may synchronise the
        whole device."""
        ptr = lambda a: None if a is None else a.__array_interface__["data"][0]  # noqa: E731
        assert obs_np.dtype == np.float32 and obs_np.shape == (self.num_envs, self.obs_w) and obs_np.flags.c_contiguous
        assert reward_np.dtype == np.float32 and reward_np.shape == (self.num_envs,) and done_np.dtype == np.uint8
        self._serve_refs = (obs_np, reward_np, done_np, terms_np)
        self._ck(self.L.ilrl_serve_begin(self.h, int(nparts), ptr(obs_np), ptr(reward_np), ptr(done_np), ptr(terms_np)))

    def serve_step(self, action_np):
        rc = self.L.ilrl_serve_step(self.h, action_np.__array_interface__["data"][0])
        if rc != 0:
            self._ck(rc)

    def serve_post(self, action_np, part=0):
        rc = self.L.ilrl_serve_post(self.h, part, action_np.__array_interface__["data"][0])
        if rc != 0:
            self._ck(rc)

    def serve_wait(self, part=0):
        rc = self.L.ilrl_serve_wait(self.h, part)
        if rc != 0:
            self._ck(rc)

    def serve_end(self):
        self._ck(self.L.ilrl_serve_end(self.h))

    def part_slice(self, part, nparts):
        """env index range of a part: ceil(N / nparts) rounded up to whole 16-env tiles"""
        per = -(-self.num_envs // nparts)
        per = -(-per // 16) * 16
        first = min(part * per, self.num_envs)
        return slice(first, min(first + per, self.num_envs))

    # packed row of ilrl_step_pull / ilrl_pull (include/ilrl.h)
    PULL_WORDS = 257
    PULL_OBS, PULL_REWARD, PULL_DONE, PULL_TERMS, PULL_ENVF, PULL_PHYS = slice(0, 70), 72, 73, slice(74, 86), slice(86, 114), slice(114, 161)
    PULL_HIGH_OBS, PULL_HIGH_REWARD, PULL_HIGH_FLAGS = slice(161, 205), 221, 222
    PULL_OBS2, PULL_HIGH_OBS2, PULL_JT = slice(0, 72), slice(161, 221), slice(223, 257)   # mode "hier2" widths

    def step_pull(self, action_np, forced_target_deg=None):
        """One blocking step from host actions [N,17] returning the packed mirror rows [N,257] (numpy, owned by this
        object): obs | reward | done | terms | envf | phys | high obs | high reward | high flags | jointTarget - one
        transfer (PULL_* slices)."""
        if self._pull_buf is None:
            self._pull_buf = np.zeros((self.num_envs, self.PULL_WORDS), np.float32)
        a = np.ascontiguousarray(action_np, dtype=np.float32).reshape(self.num_envs, ACT_LOW)
        rc = self.L.ilrl_step_pull(self.h, a.ctypes.data, INT32_MIN if forced_target_deg is None else int(forced_target_deg),
                                   self._pull_buf.ctypes.data, self._stream())
        if rc != 0:
            self._ck(rc)
        return self._pull_buf

    def pull(self, obs=None):
        """The packed mirror rows without stepping (after a reset / high-level step).  obs: device tensor [N,70] to take
        the observation columns from (default: the observation of the last step_pull)."""
        if self._pull_buf is None:
            self._pull_buf = np.zeros((self.num_envs, self.PULL_WORDS), np.float32)
        self._ck(self.L.ilrl_pull(self.h, _ptr(obs), self._pull_buf.ctypes.data, self._stream()))
        return self._pull_buf

    def set_config(self, max_timestep=0, step_per_level=0, skip_frame=0):
        """Change `max_timestep` / `step_per_level` / `skipFrame` of the live handle (0 = keep)."""
        self._ck(self.L.ilrl_set_config(self.h, int(max_timestep), int(step_per_level), int(skip_frame)))

    def high_step(self, action2):
        """hier mode: heading action [N,2] for the envs waiting for one; returns the low-level obs tensor [N,70]
        (rows of envs that were not waiting are untouched)."""
        a = self._f32(action2, (self.num_envs, self.hact_w))
        self._ck(self.L.ilrl_high_step(self.h, _ptr(a), _ptr(self.obs), self._stream()))
        return self.obs

    def high_readout(self):
        """hier mode: (high_obs [N,44], high_reward [N], flags [N] u8: bit0 episode ended, bit1 high agent present,
        bit2 waiting for a high-level action)."""
        self._ck(self.L.ilrl_high_readout(self.h, _ptr(self.high_obs), _ptr(self.high_reward), _ptr(self.high_flags),
                                          self._stream()))
        return self.high_obs, self.high_reward, self.high_flags

    def high_step_into(self, action2, low_obs):
        """`high_step` writing the low-level obs rows of the envs that were waiting into a caller-owned [N,70] tensor."""
        n = self.num_envs
        for t, shape in ((action2, (n, self.hact_w)), (low_obs, (n, self.obs_w))):
            assert t.is_cuda and t.device == self.device and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shape
        self._ck(self.L.ilrl_high_step(self.h, _ptr(action2), _ptr(low_obs), self._stream()))

    def high_readout_into(self, high_obs, high_reward, high_flags):
        """`high_readout` into caller-owned tensors: [N,44] f32, [N] f32, [N] u8."""
        n = self.num_envs
        for t, shape, dt in ((high_obs, (n, self.hobs_w), torch.float32), (high_reward, (n,), torch.float32),
                             (high_flags, (n,), torch.uint8)):
            assert t.is_cuda and t.device == self.device and t.dtype == dt and t.is_contiguous() and tuple(t.shape) == shape
        self._ck(self.L.ilrl_high_readout(self.h, _ptr(high_obs), _ptr(high_reward), _ptr(high_flags), self._stream()))

    def set_clip_of_env(self, clip_of_env):
        """Re-assign which staged clip each env imitates (`selected_motion`).  Takes effect at once, as in the reference;
        change it between episodes (a frame index valid in one clip may not exist in a shorter one)."""
        ids = np.ascontiguousarray(clip_of_env, dtype=np.int32)
        assert ids.shape == (self.num_envs,)
        self._ck(self.L.ilrl_set_clip_ids(self.h, ids.ctypes.data))

    # ------------------------------------------------------------------ parity-harness / introspection entry points
    def get_state(self):
        phys = torch.empty(self.num_envs, PHYS_WORDS, device=self.device)
        envf = torch.empty(self.num_envs, ENV_WORDS, device=self.device)
        self._ck(self.L.ilrl_get_state(self.h, _ptr(phys), _ptr(envf), self._stream()))
        return phys, envf

    def set_state(self, phys=None, envf=None):
        p = None if phys is None else self._f32(phys, (self.num_envs, PHYS_WORDS))
        e = None if envf is None else self._f32(envf, (self.num_envs, ENV_WORDS))
        self._ck(self.L.ilrl_set_state(self.h, _ptr(p), _ptr(e), self._stream()))

    def set_forced_target_deg(self, deg):
        """Per-env heading (int degrees) used by the next target re-sampling instead of the env's own draw;
        INT32_MIN entries (or deg=None) = draw normally."""
        if deg is None:
            self._forced = None
            self._ck(self.L.ilrl_set_forced_target_deg(self.h, None))
        else:
            self._forced = torch.as_tensor(deg, device=self.device).to(torch.int32).contiguous()
            self._ck(self.L.ilrl_set_forced_target_deg(self.h, _ptr(self._forced)))

    def set_heightfield(self, data, rows=256, cols=256, body_z=0.25):
        """Heightfield terrain for every env of the handle ("low" mode; REF humanoid.py:68-144 CustomScene /
        replaceHeightfieldData).  data[i + j * rows] as the reference's heightfieldData; None = flat ground.  The world
        height of a sample is data - (min + max) / 2 + body_z, as Bullet places a heightfield shape."""
        if data is None:
            self._ck(self.L.ilrl_set_heightfield(self.h, None, 0, 0, 0.0))
            return
        d = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
        assert d.size == rows * cols
        zoff = float(body_z) - 0.5 * (float(d.min()) + float(d.max()))
        self._ck(self.L.ilrl_set_heightfield(self.h, d.ctypes.data, int(rows), int(cols), zoff))

    def set_self_collision(self, on=True):
        """Bullet-style self-collision of the robot ("low" mode; REF humanoid.py:13 `self_collision = True`): 66 capsule
        pairs, two-body contact rows.  Off by default."""
        self._ck(self.L.ilrl_set_self_collision(self.h, int(bool(on))))

    def set_forced_reset_noise(self, noise17):
        """mode "hier2" harness: the joint noise [N,17] resets use instead of their own uniform(-0.1, 0.1) draws (only the
        six arm joints keep it, REF hier_env_2.py:214-252); None = draw."""
        self._noise = None if noise17 is None else self._f32(noise17, (self.num_envs, ACT_LOW))
        self._ck(self.L.ilrl_set_forced_reset_noise(self.h, _ptr(self._noise)))

    def get_joint_target(self):
        out = torch.empty(self.num_envs, JT_WORDS, device=self.device)
        self._ck(self.L.ilrl_get_joint_target(self.h, _ptr(out), self._stream()))
        return out

    def set_joint_target(self, jt):
        self._ck(self.L.ilrl_set_joint_target(self.h, _ptr(self._f32(jt, (self.num_envs, JT_WORDS))), self._stream()))

    def physics_only(self, torque):
        t = self._f32(torque, (self.num_envs, ACT_LOW))
        self._ck(self.L.ilrl_physics_only(self.h, _ptr(t), self._stream()))

    def endpoint_score(self):
        out = torch.empty(self.num_envs, device=self.device)
        self._ck(self.L.ilrl_endpoint_score(self.h, _ptr(out), self._stream()))
        return out

    def stats(self):
        """Device tensor [16]: {episodes, sum return, sum length, steps, sum reward, sums of the first 11 terms}
        accumulated since the previous call.  All-reduce it across ranks with torch.distributed (NCCL) if needed."""
        out = torch.empty(STATS_WORDS, device=self.device)
        self._ck(self.L.ilrl_stats(self.h, _ptr(out), self._stream()))
        return out

    def kernel_timing(self, on):
        """Device-side timing of the step kernels (ilrl_kernel_timing): returns (milliseconds, launches) accumulated
        since the previous call and switches the stamping on / off."""
        ms, cnt = C.c_float(0), C.c_int64(0)
        self._ck(self.L.ilrl_kernel_timing(self.h, int(bool(on)), C.byref(ms), C.byref(cnt)))
        return float(ms.value), int(cnt.value)

    def launch_count(self):
        return int(self.L.ilrl_launch_count(self.h))
