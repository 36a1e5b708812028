"""Reference-shaped single-env classes: the drop-in faces of the hot path.

`LowLevelHumanoidEnv` mirrors REF low_level_env.py:36-526 (a `gym.Env`), `HierarchicalHumanoidEnv` mirrors
REF hier_env.py:39-642 (an RLlib `MultiAgentEnv`): same class names, constructor arguments, spaces, methods, return
shapes/dtypes, agent ids, public attributes (the ones `custom_callback.py`, `env_check*.py` and `env_vis*.py` read or
write) and error behaviour (`assert np.isfinite(action).all()` REF humanoid.py:55, `assert len(action_dict) == 1`
REF hier_env.py:356).  Each instance is an N = 1 view of `BatchedHumanoidEnv`: every number it returns is computed by
the CUDA kernels behind the C ABI (include/ilrl.h); the host only moves 1 action up and 1 obs row + the env words
down per call and mirrors the reference's Python-side bookkeeping that is not per-step arithmetic (the
`usePredefinedTarget` list, the `debug` done rule, the `rng` the reference draws its integers from).

For throughput use `BatchedHumanoidEnv` or the adapters in rllib_adapters.py; these classes exist so the reference's
drivers (`train_config.make_env_low/make_env_hier`, `env_check.py`, ...) run unchanged on top of the new path.
"""
import copy

import numpy as np

from . import batched_env as B
from .batched_env import BatchedHumanoidEnv
from .clips import load_clip

try:  # gym is not part of this image; the reference needs it only for the Box spaces and the Env base class
    from gym import Env as _GymEnv
    from gym.spaces import Box
except Exception:  # pragma: no cover - exercised in this image
    class _GymEnv(object):
        pass

    class Box(object):
        """Minimal stand-in for gym.spaces.Box (low/high/shape/dtype/sample/contains)."""

        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.dtype = np.dtype(dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)
            self._rng = np.random.default_rng()

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)

try:
    from ray.rllib.env import MultiAgentEnv as _MultiAgentEnv
except Exception:  # pragma: no cover - ray is not part of this image
    class _MultiAgentEnv(object):
        pass

# the reference's dictionaries, kept as data because drivers read them (REF low_level_env.py:86-152)
JOINT_MAP = {
    "right_knee": "rightKnee", "right_hip_x": "rightHipX", "right_hip_y": "rightHipY", "right_hip_z": "rightHipZ",
    "left_knee": "leftKnee", "left_hip_x": "leftHipX", "left_hip_y": "leftHipY", "left_hip_z": "leftHipZ",
    "right_shoulder_x": "rightShoulderX", "right_shoulder_y": "rightShoulderY", "right_elbow": "rightElbow",
    "left_shoulder_x": "leftShoulderX", "left_shoulder_y": "leftShoulderY", "left_elbow": "leftElbow",
}
JOINT_COLS = ["rightHipX", "rightHipY", "rightHipZ", "rightKnee", "leftHipX", "leftHipY", "leftHipZ", "leftKnee",
              "rightShoulderX", "rightShoulderY", "rightElbow", "leftShoulderX", "leftShoulderY", "leftElbow"]
EP_COLS = ["%s_%sposition" % (b, a) for b in ("LeftLeg", "LeftFoot", "RightLeg", "RightFoot", "Head", "LeftForeArm",
                                                "LeftHand", "RightForeArm", "RightHand") for a in "XYZ"]
JOINT_WEIGHT = {"right_knee": 3, "right_hip_x": 1, "right_hip_y": 3, "right_hip_z": 1, "left_knee": 3, "left_hip_x": 1,
                "left_hip_y": 3, "left_hip_z": 1, "right_shoulder_x": 0.1, "right_shoulder_y": 0.3, "right_elbow": 0.3,
                "left_shoulder_x": 0.1, "left_shoulder_y": 0.3, "left_elbow": 0.3}        # REF low_level_env.py:103-119
JOINT_VEL_WEIGHT = {k: (1 if ("knee" in k or "hip" in k) else 0.1) for k in JOINT_WEIGHT}  # REF low_level_env.py:121-137
END_POINT_MAP = {"link0_11": "RightLeg", "right_foot": "RightFoot", "link0_18": "LeftLeg", "left_foot": "LeftFoot"}
END_POINT_WEIGHT = {"link0_11": 1, "right_foot": 3, "link0_18": 1, "left_foot": 3}         # REF low_level_env.py:139-152
_REWARD_ATTRS = dict(deltaJoints=0, deltaVelJoints=1, delta_lowTargetScore=2, electricityScore=3, jointLimitScore=4,
                     aliveReward=5, bodyPostureScore=6, lowTargetScore=7, deltaEndPoints=8, highTargetScore=9,
                     driftScore=10, delta_highTargetScore=11)


class _SingleEnv(object):
    """Shared plumbing of the two N = 1 views."""

    metadata = {"render.modes": ["human", "rgb_array"], "video.frames_per_second": 60}

    def _make(self, mode, clips, selected, device, seed, self_collision=False):
        # self_collision: the reference's robot class has `self_collision = True` (REF humanoid.py:13); the B200 path models
        # it as an option of the handle (ilrl_set_self_collision), off unless asked for
        self._env = BatchedHumanoidEnv(1, mode, clips=clips, clip_of_env=np.array([selected], np.int32), device=device,
                                       seed=0 if seed is None else int(seed), auto_reset=False,
                                       self_collision=self_collision)
        self.rng = np.random.default_rng(seed)
        self.cur_timestep = 0
        self._max_timestep, self._skip_frame = 3000, 2
        self.frame = 0
        self.joint_map = dict(JOINT_MAP)
        self.joint_weight, self.joint_vel_weight = dict(JOINT_WEIGHT), dict(JOINT_VEL_WEIGHT)
        self.joint_weight_sum = sum(self.joint_weight.values())
        self.joint_vel_weight_sum = sum(self.joint_vel_weight.values())
        self.end_point_map, self.end_point_weight = dict(END_POINT_MAP), dict(END_POINT_WEIGHT)
        self.end_point_weight_sum = sum(self.end_point_weight.values())
        self.target = np.array([1.0, 0.0, 0.0])
        self.highLevelDegTarget = 0.0
        self.predefinedTarget = np.array([[]])
        self.predefinedTargetIndex = 0
        self.usePredefinedTarget = False
        self.starting_ep_pos = np.zeros(3)
        self.starting_robot_pos = np.zeros(3)
        self.robot_pos = np.zeros(3)
        self.last_robotPos = np.zeros(3)
        self.frame_update_cnt = 0
        self.cur_obs = np.zeros(42, np.float32)
        self.initReward()

    # REF low_level_env.py:174-197 / hier_env.py:174-199
    def initReward(self):
        for k in _REWARD_ATTRS:
            setattr(self, k, 0)
        self.baseReward = 0
        self.last_lowTargetScore = 0
        self.bodySpeedScore = 0
        self.cumulative_driftScore = 0
        self.cumulative_aliveReward = 0
        self.delta_deltaJoints = self.delta_deltaVelJoints = self.delta_deltaEndPoints = 0
        self.delta_bodyPostureScore = 0

    def close(self):
        self._env.close()

    # Attributes the reference's drivers ASSIGN on a live env (REF env_vis_hier.py:52 `env.max_timestep = 100000`): they
    # are kernel parameters here, so assignment is forwarded to the handle instead of staying a dead Python field.
    max_timestep = property(lambda self: self._max_timestep)
    skipFrame = property(lambda self: self._skip_frame)

    @max_timestep.setter
    def max_timestep(self, v):
        self._max_timestep = int(v)
        self._env.set_config(max_timestep=int(v))

    @skipFrame.setter
    def skipFrame(self, v):
        self._skip_frame = int(v)
        self._env.set_config(skip_frame=int(v))

    @property
    def targetLen(self):
        return 5

    @targetLen.setter
    def targetLen(self, v):
        if float(v) != 5.0:   # REF low_level_env.py:155 / hier_env.py:160: a compile-time constant of the kernels
            raise NotImplementedError("targetLen is fixed at 5 m in the B200 path (ILRL_TARGET_LEN)")

    # the four reference tables as pandas DataFrames with the CSV's own column names (REF low_level_env.py:58-71 reads
    # them with pd.read_csv; drivers index them with .iloc[frame][column]).  Built on first use.
    def _frames(self, name):
        import pandas as pd
        c = load_clip(name)
        return (pd.DataFrame(c["pos"].astype(np.float64), columns=JOINT_COLS),
                pd.DataFrame(c["vel"].astype(np.float64), columns=JOINT_COLS),
                pd.DataFrame(c["rel"].astype(np.float64), columns=JOINT_COLS),
                pd.DataFrame(c["ep"].astype(np.float64), columns=EP_COLS))

    def _table(self, k):
        if getattr(self, "_df", None) is None:
            names = [self.reference_name] if hasattr(self, "reference_name") else self.motion_list
            self._df = [self._frames(n) for n in names]
        t = [d[k] for d in self._df]
        return t[0] if hasattr(self, "reference_name") else t   # a list per motion in the hierarchical env

    joints_df = property(lambda self: self._table(0))
    joints_vel_df = property(lambda self: self._table(1))
    joints_rel_df = property(lambda self: self._table(2))
    end_point_df = property(lambda self: self._table(3))

    def render(self, mode="human"):
        """The reference forwards to PyBullet's renderer (REF low_level_env.py:202-203); there is no renderer on
        this path.  "rgb_array" returns a black frame of the reference's render size so gym's Monitor keeps working."""
        if mode == "rgb_array":
            return np.zeros((240, 320, 3), np.uint8)
        return None

    def seed(self, seed=None):
        self.rng = np.random.default_rng(seed)
        return [seed]

    # ---- state mirror: ONE packed row per call (ilrl_step_pull / ilrl_pull) holds everything the reference keeps as
    # Python attributes; it is exposed under the reference's names
    def _mirror(self, row, with_terms=False):
        E = BatchedHumanoidEnv
        e = row[E.PULL_ENVF].astype(np.float64)
        self._row = row
        self._phys = row[E.PULL_PHYS].astype(np.float64)
        self._envf = e
        self.frame = int(e[B.E_FRAME])
        self.cur_timestep = int(e[B.E_T])
        self.target = np.array([e[B.E_TARGET_X], e[B.E_TARGET_Y], 0.0])
        self.starting_robot_pos = np.array([e[B.E_START_X], e[B.E_START_Y], 0.0])
        self.starting_ep_pos = np.array([e[B.E_SEP_X], e[B.E_SEP_Y], e[B.E_SEP_Z]])
        self.robot_pos = np.array([e[B.E_ROBOT_X], e[B.E_ROBOT_Y], 0.0])
        self.highLevelDegTarget = float(e[B.E_HLDEG])
        self.lowTargetScore = float(e[B.E_LOW_TARGET_SCORE])
        self.highTargetScore = float(e[B.E_HIGH_TARGET_SCORE])
        if with_terms:
            t = row[E.PULL_TERMS].astype(np.float64)
            for k, i in _REWARD_ATTRS.items():
                setattr(self, k, float(t[i]))
        return e

    def _pull(self, obs=None):
        return self._mirror(self._env.pull(obs)[0])

    def _push_env_words(self, **words):
        e = self._envf.astype(np.float32).copy()
        for k, v in words.items():
            e[getattr(B, k)] = v
        self._env.set_state(None, e[None, :])
        self._envf = e.astype(np.float64)

    def _peek_deg(self):
        """The integer the reference would draw with `self.rng.integers(-180, 180)` IF the target is reached this
        step (REF low_level_env.py:240-245, 416-417): drawn on a copy of the generator and committed afterwards only
        when the kernel reports that the target switched, so `self.rng` stays in step with the reference's."""
        saved = copy.deepcopy(self.rng.bit_generator.state)
        deg = int(self.rng.integers(-180, 180))
        after = copy.deepcopy(self.rng.bit_generator.state)
        self.rng.bit_generator.state = saved
        return deg, after

    def _after_target_logic(self, old_target, after_state):
        """Commit the peeked draw and apply the usePredefinedTarget list (REF low_level_env.py:419-429) when the
        kernel switched the target this step."""
        switched = not np.allclose(self.target[:2], old_target[:2], atol=0, rtol=0)
        if not switched:
            return
        self.rng.bit_generator.state = after_state
        if self.usePredefinedTarget:
            self.predefinedTargetIndex = (self.predefinedTargetIndex + 1) % len(self.predefinedTarget)
            new = np.asarray(self.predefinedTarget[self.predefinedTargetIndex], dtype=np.float64)
            start = self.starting_robot_pos  # = the old target (Q8)
            score = -float(np.linalg.norm(new[:2] - start[:2]))
            words = dict(E_TARGET_X=new[0], E_TARGET_Y=new[1])
            if self._env.mode == 0:
                # lowTargetScore restart + the per-step heading / walk-target refresh on the new target
                h = float(np.arctan2(new[1] - self.robot_pos[1], new[0] - self.robot_pos[0]))
                words.update(E_LOW_TARGET_SCORE=score, E_HLDEG=h, E_WALK_X=self.robot_pos[0] + np.cos(h) * 10,
                             E_WALK_Y=self.robot_pos[1] + np.sin(h) * 10)
                self.lowTargetScore, self.highLevelDegTarget = score, h
            else:
                words.update(E_HIGH_TARGET_SCORE=score)
                self.highTargetScore = score
            self._push_env_words(**words)
            self.target = np.array([new[0], new[1], 0.0])

    # ---- the reference's individual score methods, callable at any time on the current state (param_check.py calls
    # calcEndPointScore / calcJointScore between steps, REF param_check.py:43-60).  They are evaluated by the same
    # kernel as a step: the env state is saved, one step with the physics skipped runs, the terms are read, and the
    # state is put back.
    def _peek_terms(self, action=None):
        phys, envf = self._env.get_state()
        tmp = envf.clone()
        tmp[:, B.E_HIGH_PENDING] = 0.0
        self._env.set_state(None, tmp)
        self._env.set_forced_target_deg(np.array([0], np.int64))     # no draw is consumed by the peek
        a = np.zeros((1, 17), np.float32) if action is None else np.asarray(action, np.float32).reshape(1, 17)
        _, _, _, terms = self._env.step(a, physics=False)
        t = terms[0].cpu().numpy().astype(np.float64)
        self._env.set_forced_target_deg(None)
        self._env.set_state(phys, envf)
        return t

    def calcJointScore(self, useExp=False):          # REF low_level_env.py:325-341
        v = self._peek_terms()[0]
        return float(v) if useExp else float(np.log(max(v, 1e-300)) / 4.0)

    def calcJointVelScore(self, useExp=False):       # REF low_level_env.py:343-359
        v = self._peek_terms()[1]
        return float(v) if useExp else float(2.0 * np.log(max(v, 1e-300)))

    def calcBodyPostureScore(self, useExp=False):    # REF low_level_env.py:405-410
        v = self._peek_terms()[6]
        return float(v) if useExp else float(np.log(max(v, 1e-300)))

    def calcAliveReward(self):                       # REF low_level_env.py:384-387
        return float(self._peek_terms()[5])

    def calcJointLimitCost(self):                    # REF low_level_env.py:396-397
        return float(self._peek_terms()[4])

    def calcElectricityCost(self, action):           # REF low_level_env.py:389-394
        return float(self._peek_terms(action)[3])

    def calcEndPointScore(self, useExp=False):       # REF low_level_env.py:361-382
        s = float(self._env.endpoint_score()[0].item())
        return s if useExp else float(np.log(max(s, 1e-300)) / 3.0)

    def setJointsOrientation(self, idx):             # REF low_level_env.py:205-216
        """Joint positions / velocities of reference frame `idx` (the three abdomen joints zero), base untouched."""
        name = self.reference_name if hasattr(self, "reference_name") else self.motion_list[self.selected_motion]
        c = load_clip(name)
        phys, _ = self._env.get_state()
        p = phys[0].cpu().numpy()
        p[13:30] = 0.0
        p[30:47] = 0.0
        joints = [6, 3, 5, 4, 10, 7, 9, 8, 12, 11, 13, 15, 14, 16]      # joint_map order -> joint slot
        cols = [3, 0, 1, 2, 7, 4, 5, 6, 8, 9, 10, 11, 12, 13]           # joint_map order -> CSV column
        for j, col in zip(joints, cols):
            p[13 + j] = c["pos"][idx, col]
            p[30 + j] = c["vel"][idx, col]
        self._env.set_state(p[None, :], None)

    def _first_target_xy(self):
        if self.usePredefinedTarget:
            self.predefinedTargetIndex = 0
            t = np.asarray(self.predefinedTarget[0], dtype=np.float32)
            return t[None, :2].copy()
        return None


class _CustomSceneShim(object):
    """The part of REF humanoid.py:68-144 `CustomScene` a driver touches through `env.flat_env.stadium_scene`: the
    heightfield data, its per-episode random regeneration and `replaceHeightfieldData` (REF env_vis_low.py:155-171).
    The collision shape itself is the handle's heightfield (ilrl_set_heightfield)."""
    numHeightfieldRows = 256
    numHeightfieldColumns = 256

    def __init__(self, owner):
        self._owner = owner
        self.heightfieldData = [0] * self.numHeightfieldRows * self.numHeightfieldColumns

    def _upload(self):
        self._owner._env.set_heightfield(self.heightfieldData, self.numHeightfieldRows, self.numHeightfieldColumns)

    def replaceHeightfieldData(self, newData):       # REF humanoid.py:76-86
        self.heightfieldData = list(newData)
        self._upload()

    def episode_restart(self, bullet_client=None):   # REF humanoid.py:88-124: 2 x 2 sample plateaus of U(0, 0.5), flat centre
        import random
        R, d = self.numHeightfieldRows, self.heightfieldData
        for j in range(self.numHeightfieldColumns // 2):
            for i in range(R // 2):
                h = random.uniform(0, 0.05) * 10
                d[2 * i + 2 * j * R] = d[2 * i + 1 + 2 * j * R] = d[2 * i + (2 * j + 1) * R] = d[2 * i + 1 + (2 * j + 1) * R] = h
        for j in (63, 64):
            for i in (63, 64):
                d[2 * i + 2 * j * R] = d[2 * i + 1 + 2 * j * R] = d[2 * i + (2 * j + 1) * R] = d[2 * i + 1 + (2 * j + 1) * R] = 0
        self._upload()


class _FlatEnvShim(object):
    def __init__(self, owner):
        self.stadium_scene = _CustomSceneShim(owner)


class LowLevelHumanoidEnv(_SingleEnv, _GymEnv):
    """REF low_level_env.py:36.  `useCustomEnv=True` steps on the reference's random heightfield terrain
    (REF humanoid.py:68-144: regenerated at every reset, replaceable through
    `env.flat_env.stadium_scene.replaceHeightfieldData`); `customRobot` is accepted and ignored (the kernels model
    `CustomHumanoidRobot` on `humanoid_symmetric_2.xml`, the robot every reference config passes)."""

    def __init__(self, reference_name="motion08_03", useCustomEnv=False, customRobot=None, device=0, seed=None,
                 self_collision=False):
        self.useCustomEnv = bool(useCustomEnv)
        self.reference_name = reference_name
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=[8 + 17 * 2 + 14 * 2])
        self.action_space = Box(low=-1, high=1, shape=[17])
        self.max_frame = load_clip(reference_name)["max_frame"]
        self._make("low", [reference_name], 0, device, seed, self_collision)
        if self.useCustomEnv:
            self.flat_env = _FlatEnvShim(self)

    def reset(self, resetYaw=0):
        # REF low_level_env.py:224-232: the start frame comes from the env's own generator
        return self.resetFromFrame(startFrame=int(self.rng.integers(0, self.max_frame - 5)), resetYaw=resetYaw,
                                   startFromRef=True, initVel=True)

    def resetFromFrame(self, startFrame=0, resetYaw=0, startFromRef=True, initVel=True):
        if not startFromRef:
            raise NotImplementedError("startFromRef=False (keep PyBullet's U(-0.1, 0.1) joint noise) is not part of "
                                      "the B200 path; every reference caller passes True")
        if self.useCustomEnv:   # flat_env.reset() -> CustomScene.episode_restart: a new terrain every episode
            self.flat_env.stadium_scene.episode_restart()
        xy = self._first_target_xy()
        deg = None if xy is not None else np.array([int(self.rng.integers(-180, 180))], np.int32)
        obs_dev = self._env.reset(start_frame=np.array([startFrame], np.int32), target_deg=deg,
                                  reset_yaw_deg=np.array([resetYaw], np.float32), target_xy=xy)
        self.initReward()
        self._pull(obs_dev)
        obs = self._row[BatchedHumanoidEnv.PULL_OBS].astype(np.float64)
        if not initVel:
            p = self._phys.astype(np.float32)
            p[7:10] = 0.0
            self._env.set_state(p[None, :], None)
        self.last_robotPos = self.robot_pos.copy()
        self.cur_obs = obs[:42].astype(np.float32)
        return obs

    def step(self, action, debug=False):
        return self.low_level_step(action, debug=debug)

    def low_level_step(self, action, debug=False):
        a = np.asarray(action, dtype=np.float32).reshape(17)
        assert np.isfinite(a).all()
        deg, after = self._peek_deg()
        old_target = self.target.copy()
        E = BatchedHumanoidEnv
        row = self._env.step_pull(a, deg)[0]      # one launch for the step, one for the packing, one transfer
        obs = row[E.PULL_OBS].astype(np.float64)
        reward = float(row[E.PULL_REWARD])
        done = bool(row[E.PULL_DONE])
        self._mirror(row, with_terms=True)
        self._after_target_logic(old_target, after)
        if debug:  # REF low_level_env.py:467-473: the debug rule ignores the distance test
            done = not (self.aliveReward > 0) or self.cur_timestep >= self.max_timestep
        self.cur_obs = obs[:42].astype(np.float32)
        return obs, reward, done, {}

    def getLowLevelObs(self):
        return np.hstack((self.cur_obs.astype(np.float64), self._obs_tail(self.reference_name, self.frame)))

    @staticmethod
    def _obs_tail(name, frame):
        c = load_clip(name)
        cols = [3, 0, 1, 2, 7, 4, 5, 6, 8, 9, 10, 11, 12, 13]  # joint_map order -> CSV column
        return np.stack([c["rel"][frame, cols], c["vel"][frame, cols]], 1).reshape(-1).astype(np.float64)



class HierarchicalHumanoidEnv(_SingleEnv, _MultiAgentEnv):
    """REF hier_env.py:39.  `selected_motion` indexes `motion_list` and may be reassigned between episodes, as the
    reference's evaluation scripts do (REF env_check_hier.py:93)."""

    def __init__(self, customRobot=None, device=0, seed=None, self_collision=False):
        self.motion_list = ["motion08_03", "motion09_03"]
        self.high_level_obs_space = Box(low=-np.inf, high=np.inf, shape=[8 + 17 * 2 + 2])
        self.high_level_act_space = Box(low=-1, high=1, shape=[2])
        self.low_level_obs_space = Box(low=-np.inf, high=np.inf, shape=[8 + 17 * 2 + 14 * 2])
        self.low_level_act_space = Box(low=-1, high=1, shape=[17])
        self._step_per_level = 5
        self.steps_remaining_at_level = self._step_per_level
        self.num_high_level_steps = 0
        self.max_frame = [load_clip(m)["max_frame"] for m in self.motion_list]
        self._selected_motion = 1
        self.selected_motion_frame = 0
        self.low_level_agent_id = "low_level_agent"
        self._make("hier", self.motion_list, self._selected_motion, device, seed, self_collision)

    _PO, _PH, _HACT = BatchedHumanoidEnv.PULL_OBS, BatchedHumanoidEnv.PULL_HIGH_OBS, 2   # packed-row columns / widths of the mode

    step_per_level = property(lambda self: self._step_per_level)

    @step_per_level.setter
    def step_per_level(self, v):
        self._step_per_level = int(v)
        self._env.set_config(step_per_level=int(v))

    @property
    def selected_motion(self):
        return self._selected_motion

    @selected_motion.setter
    def selected_motion(self, v):
        v = int(v)
        if not 0 <= v < len(self.motion_list):
            raise IndexError("selected_motion %d outside motion_list" % v)
        self._selected_motion = v
        self._env.set_clip_of_env(np.array([v], np.int32))

    def _mirror(self, row, with_terms=False):
        e = _SingleEnv._mirror(self, row, with_terms)
        self.selected_motion_frame = self.frame
        self.steps_remaining_at_level = int(e[B.E_STEPS_REMAINING])
        self.cumulative_driftScore = float(e[B.E_CUM_DRIFT])
        self.cumulative_aliveReward = float(e[B.E_CUM_ALIVE])
        return e

    def reset(self):
        # REF hier_env.py:235-243: both integers come from the env's own generator, start frame first
        sf = int(self.rng.integers(0, self.max_frame[self._selected_motion] - 5))
        yaw = int(self.rng.integers(-180, 180))
        return self.resetFromFrame(startFrame=sf, resetYaw=yaw, startFromRef=True, initVel=True)

    def resetFromFrame(self, startFrame=0, resetYaw=0, startFromRef=True, initVel=True):
        if not startFromRef:
            raise NotImplementedError("startFromRef=False is not part of the B200 path; every reference caller "
                                      "passes True")
        xy = self._first_target_xy()
        deg = None if xy is not None else np.array([int(self.rng.integers(-180, 180))], np.int32)
        hobs = self._env.reset(start_frame=np.array([startFrame], np.int32), target_deg=deg,
                               reset_yaw_deg=np.array([resetYaw], np.float32), target_xy=xy)
        hobs = hobs[0].cpu().numpy().astype(np.float64)   # (the reset writes the first high-level obs to the caller's buffer)
        self.initReward()
        self._pull()
        if not initVel:
            p = self._phys.astype(np.float32)
            p[7:10] = 0.0
            self._env.set_state(p[None, :], None)
        self.num_high_level_steps = 0
        self.low_level_agent_id = "low_level_agent"
        return {"high_level_agent": hobs}

    def step(self, action_dict, debug=False):
        assert len(action_dict) == 1, action_dict
        if "high_level_agent" in action_dict:
            return self.high_level_step(action_dict["high_level_agent"], debug=debug)
        return self.low_level_step(list(action_dict.values())[0], debug=debug)

    def _set_pending(self, flag):
        if bool(self._envf[B.E_HIGH_PENDING]) != bool(flag):
            self._push_env_words(E_HIGH_PENDING=1.0 if flag else 0.0)

    def high_level_step(self, action, debug=False):
        a = np.asarray(action, dtype=np.float32).reshape(self._HACT)
        self._set_pending(True)   # the reference applies a high-level action whenever one arrives
        obs_dev = self._env.high_step(a[None, :])
        self._pull(obs_dev)
        obs = self._row[self._PO].astype(np.float64)
        self.num_high_level_steps += 1
        return {self.low_level_agent_id: obs}, {self.low_level_agent_id: 0}, {"__all__": False}, {}

    def low_level_step(self, action, debug=False):
        a = np.asarray(action, dtype=np.float32).reshape(17)
        assert np.isfinite(a).all()
        self._set_pending(False)  # ... and a low-level action whenever one arrives
        deg, after = self._peek_deg()
        old_target = self.target.copy()
        E = BatchedHumanoidEnv
        row = self._env.step_pull(a, deg)[0]
        low_obs = row[self._PO].astype(np.float64)
        low_rew, high_rew = float(row[E.PULL_REWARD]), float(row[E.PULL_HIGH_REWARD])
        flags = int(row[E.PULL_HIGH_FLAGS])
        self._mirror(row, with_terms=True)
        self._after_target_logic(old_target, after)
        ended, high_present = bool(flags & 1), bool(flags & 2)
        o, r, d = {}, {}, {"__all__": False}
        if ended:
            d["__all__"] = True
            r["high_level_agent"] = high_rew
            o["high_level_agent"] = row[self._PH].astype(np.float64)
            o[self.low_level_agent_id] = low_obs
            r[self.low_level_agent_id] = low_rew
        elif high_present:
            r["high_level_agent"] = high_rew
            o["high_level_agent"] = row[self._PH].astype(np.float64)
        else:
            o = {self.low_level_agent_id: low_obs}
            r = {self.low_level_agent_id: low_rew}
        if debug and ended and self.aliveReward > 0 and self.cur_timestep < self.max_timestep:
            # REF hier_env.py:573-581: the debug rule ignores the distance test
            d["__all__"] = False
        return o, r, d, {}


class HierarchicalHumanoidEnv2(HierarchicalHumanoidEnv):
    """REF hier_env_2.py:39 (its class is also called HierarchicalHumanoidEnv): the variant whose high-level agent
    hands the low level 17 (joint position, joint velocity) targets instead of a heading.  36-d high action, 60-d high
    obs, 72-d low obs, step_per_level 20, skipFrame 5; the frame advances in reset and high_level_step only.
    Declared substitutions: clips from "Joints CSV With Hand" (the reference reads an unshipped
    Relative_Joints_CSV directory with the same file names) and the humanoid_symmetric_2.xml robot with the stock
    class's 44-entry state (the reference builds pybullet_envs' stock Humanoid on the unshipped humanoid_symmetric.xml)."""

    _PO, _PH, _HACT = BatchedHumanoidEnv.PULL_OBS2, BatchedHumanoidEnv.PULL_HIGH_OBS2, 36

    def __init__(self, device=0, seed=None, self_collision=False):
        self.motion_list = ["motion08_03", "motion09_03"]
        self.high_level_obs_space = Box(low=-np.inf, high=np.inf, shape=[2 + 42 + 8 * 2])
        self.high_level_act_space = Box(low=-1, high=1, shape=[2 + 17 * 2])
        self.low_level_obs_space = Box(low=-np.inf, high=np.inf, shape=[4 + 17 * 2 + 17 * 2])
        self.low_level_act_space = Box(low=-1, high=1, shape=[17])
        self._step_per_level = 20
        self.steps_remaining_at_level = self._step_per_level
        self.num_high_level_steps = 0
        self.max_frame = [load_clip(m)["max_frame"] for m in self.motion_list]
        self._selected_motion = 1
        self.selected_motion_frame = 0
        self.low_level_agent_id = "low_level_agent"
        self.jointTarget = [0] * 16                                          # REF hier_env_2.py:172
        self._make("hier2", self.motion_list, self._selected_motion, device, seed, self_collision)
        self._skip_frame = 5
        legs = [k for k in JOINT_MAP if "knee" in k or "hip" in k]          # REF hier_env_2.py:90-126
        self.joint_map = {k: JOINT_MAP[k] for k in legs}
        self.joint_weight = {k: JOINT_WEIGHT[k] for k in legs}
        self.joint_vel_weight = {k: 1 for k in legs}
        self.joint_weight_sum = sum(self.joint_weight.values())
        self.joint_vel_weight_sum = sum(self.joint_vel_weight.values())
        self.cur_obs = np.zeros(44, np.float32)

    def initReward(self):                                                    # REF hier_env_2.py:176-203
        HierarchicalHumanoidEnv.initReward(self)
        self.deltaJoints_low = self.deltaVelJoints_low = 0
        self.lowTargetScore = self.highTargetScore = -5
        self.cumulative_deltaJoints_low = self.cumulative_deltaVelJoints_low = 0

    def _mirror(self, row, with_terms=False):
        e = HierarchicalHumanoidEnv._mirror(self, row, with_terms)
        # MODE 2 keeps its two accumulators in the words the other modes use for lowTargetScore / cumulative_aliveReward
        self.cumulative_deltaVelJoints_low = float(e[B.E_LOW_TARGET_SCORE])
        self.cumulative_deltaJoints_low = float(e[B.E_CUM_ALIVE])
        self.lowTargetScore = -5
        self.cumulative_aliveReward = 0
        self.jointTarget = row[BatchedHumanoidEnv.PULL_JT].astype(np.float64)
        if with_terms:   # slots 2 / 8 of the terms row carry the low-level tracking scores in this mode
            t = row[BatchedHumanoidEnv.PULL_TERMS].astype(np.float64)
            self.deltaJoints_low, self.deltaVelJoints_low = float(t[2]), float(t[8])
            self.delta_lowTargetScore = self.deltaEndPoints = 0
        return e

    def reset(self, startFrame=None, startFromRef=True):                    # REF hier_env_2.py:254-262
        sf = int(self.rng.integers(0, self.max_frame[self._selected_motion] - 5)) if startFrame is None else startFrame
        yaw = int(self.rng.integers(-180, 180))
        return self.resetFromFrame(startFrame=sf, resetYaw=yaw, startFromRef=True)

    def resetFromFrame(self, startFrame=0, resetYaw=0, startFromRef=True):   # REF hier_env_2.py:277-352
        return HierarchicalHumanoidEnv.resetFromFrame(self, startFrame=startFrame, resetYaw=resetYaw,
                                                      startFromRef=startFromRef, initVel=True)
