"""Run the UNMODIFIED reference env code (/root/reference/low_level_env.py, hier_env.py, humanoid.py, math_util.py)
in this container, where pybullet / pybullet_envs / gym / ray are not installed.

TEST INFRASTRUCTURE ONLY (used by oracle/gen_golden.py and by tests that are skipped when /root/reference is
absent, i.e. on the GPU box).

How: the third-party modules the reference imports are replaced in sys.modules by small stand-ins, HOME is pointed at
a temp dir whose GitHub/TA is a symlink to the reference checkout (the CSV base path is hard-coded at
REF low_level_env.py:58 / hier_env.py:61), and `HumanoidBulletEnv` becomes `FakeFlatEnv`: an array-backed state
(47 numbers) exposing the jdict / parts / robot / scene surface the reference touches.  `scene.global_step()` calls
the C oracle's physics (oracle/ilrl_oracle.c) or nothing (injected-state mode).  `WalkerBase.calc_state` is restated
here in numpy (second, independent restatement of the upstream function: it cross-checks the C one), with its own
numpy/scipy forward kinematics.  Everything else that runs — reward terms, frame indexing, target logic, obs assembly,
termination, the multi-agent protocol, and `CustomHumanoidRobot.apply_action` / `robot_specific_reset` — is the
reference's own code, byte for byte.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
from scipy.spatial.transform import Rotation as R

REF = os.environ.get("ILRL_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
with open(os.path.join(_ROOT, "imitation-learning-rl_b200", "data", "model.json")) as _f:
    MODEL = json.load(_f)


def available():
    return os.path.exists(os.path.join(REF, "low_level_env.py"))


# ----------------------------------------------------------------------------- python FK (independent of the C one)
def py_fk(phys):
    """-> (body_pos[15,3], body_rot[15] (scipy Rotation), anchor_pos[17,3])"""
    nb, nj = MODEL["nb"], MODEL["nj"]
    bp = np.array(MODEL["body_pos"]).reshape(nb, 3)
    bq = np.array(MODEL["body_quat"]).reshape(nb, 4)
    ja = np.array(MODEL["joint_anchor"]).reshape(nj, 3)
    jx = np.array(MODEL["joint_axis"]).reshape(nj, 3)
    q = phys[13:30]
    rot = [None] * nb
    pos = np.zeros((nb, 3))
    anc = np.zeros((nj, 3))
    for b in range(nb):
        if b == 0:
            rot[0] = R.from_quat(phys[3:7])
            pos[0] = phys[0:3]
            continue
        p = MODEL["body_parent"][b]
        rc = rot[p] * R.from_quat(bq[b])
        oc = pos[p] + rot[p].apply(bp[b])
        for j in range(nj):
            if MODEL["joint_body"][j] != b:
                continue
            anc[j] = oc + rc.apply(ja[j])
            rn = rc * R.from_rotvec(jx[j] * q[j])
            oc = anc[j] - rn.apply(ja[j])
            rc = rn
        rot[b] = rc
        pos[b] = oc
    return pos, rot, anc


def bullet_rpy(q):
    """pybullet.getEulerFromQuaternion"""
    x, y, z, w = q
    sarg = -2 * (x * z - w * y)
    if sarg <= -0.99999:
        return 0.0, -0.5 * np.pi, 2 * np.arctan2(x, -y)
    if sarg >= 0.99999:
        return 0.0, 0.5 * np.pi, 2 * np.arctan2(-x, y)
    return (np.arctan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z), np.arcsin(sarg),
            np.arctan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z))


# ----------------------------------------------------------------------------- array-backed fake of the Bullet surface
class _Pose:
    def __init__(self, xyz, quat):
        self._xyz, self._q = np.array(xyz, dtype=np.float64), np.array(quat, dtype=np.float64)

    def xyz(self):
        return self._xyz

    def rpy(self):
        return bullet_rpy(self._q)

    def orientation(self):
        return self._q


class FakePart:
    def __init__(self, env, kind, idx):
        self.env, self.kind, self.idx = env, kind, idx

    def get_position(self):
        if self.kind == "floor":
            return np.zeros(3)
        pos, _, anc = self.env.fk_cached()
        return (pos[self.idx] if self.kind == "body" else anc[self.idx]).copy()

    current_position = get_position

    def pose(self):
        if self.kind == "body" and self.idx == 0:
            return _Pose(self.env.phys[0:3], self.env.phys[3:7])
        return _Pose(self.get_position(), [0, 0, 0, 1])

    def speed(self):
        assert self.kind == "body" and self.idx == 0
        return self.env.phys[7:10].copy()

    def reset_position(self, p):
        self.env.phys[0:3] = p

    def reset_orientation(self, q):
        self.env.phys[3:7] = q

    def reset_velocity(self, linearVelocity=(0, 0, 0), angularVelocity=(0, 0, 0)):
        self.env.phys[7:10] = linearVelocity
        self.env.phys[10:13] = angularVelocity


class FakeJoint:
    def __init__(self, env, idx, name):
        self.env, self.idx, self.joint_name = env, idx, name
        self.lowerLimit, self.upperLimit = MODEL["joint_lo"][idx], MODEL["joint_hi"][idx]
        self.torque = 0.0

    def get_position(self):
        return float(self.env.phys[13 + self.idx])

    def get_velocity(self):
        return float(self.env.phys[30 + self.idx])

    def get_state(self):
        return self.get_position(), self.get_velocity()

    def current_position(self):
        return self.get_state()

    def current_relative_position(self):  # upstream robot_bases.Joint
        pos, vel = self.get_state()
        pos_mid = 0.5 * (self.lowerLimit + self.upperLimit)
        return (2 * (pos - pos_mid) / (self.upperLimit - self.lowerLimit), 0.1 * vel)

    def set_state(self, x, vx):
        self.env.phys[13 + self.idx] = x
        self.env.phys[30 + self.idx] = vx

    reset_current_position = set_state
    reset_position = set_state

    def set_motor_torque(self, t):
        self.torque = float(t)

    set_torque = set_motor_torque


class _Scene:
    def __init__(self, env):
        self.env = env

    def global_step(self):
        self.env.global_step()

    def actor_introduce(self, robot):
        pass


class _Box:
    def __init__(self, low=None, high=None, shape=None, dtype=np.float32):
        self.low, self.high = low, high
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.dtype = dtype


class FakeFlatEnv:
    """Stand-in for pybullet_envs.gym_locomotion_envs.HumanoidBulletEnv(robot=...)."""
    physics = "oracle"  # "oracle" -> C oracle physics; "none" -> global_step is a no-op (injected-state mode)
    default_robot = None  # class used for HumanoidBulletEnv() without robot= (hier_env_2.py:45); None -> CustomHumanoidRobot

    def __init__(self, robot=None, render=False):
        from oracle import oracle as O
        self._O = O
        self.phys = O.default_phys()
        self.robot = robot if robot is not None else (FakeFlatEnv.default_robot or sys.modules["humanoid"].CustomHumanoidRobot)()
        self.action_space = _Box(-np.ones(17, np.float32), np.ones(17, np.float32))
        self.observation_space = _Box(-np.inf * np.ones(44, np.float32), np.inf * np.ones(44, np.float32))
        self.jdict = {n: FakeJoint(self, i, n) for i, n in enumerate(MODEL["joint_name"])}
        self.ordered_joints = [self.jdict[n] for n in MODEL["joint_name"]]
        # `parts` in pybullet link order: base, then per body its joint links then the body link
        self.parts = {"torso": FakePart(self, "body", 0)}
        for b in range(1, MODEL["nb"]):
            for j in range(MODEL["nj"]):
                if MODEL["joint_body"][j] == b:
                    self.parts["link0_%d" % self._link_index(j)] = FakePart(self, "anchor", j)
            self.parts[MODEL["body_name"][b]] = FakePart(self, "body", b)
        self.scene = _Scene(self)
        self.walk_target_x, self.walk_target_y = 1e3, 0
        r = self.robot
        r.jdict, r.parts, r.ordered_joints = self.jdict, self.parts, self.ordered_joints
        r.robot_body = self.parts["torso"]
        r.scene = self.scene
        self._resets = 0

    @staticmethod
    def _link_index(j):
        """pybullet link index of joint j's `link0_N` (bodies and joints numbered in creation order, base = 0...)"""
        n = 0
        for b in range(MODEL["nb"]):
            if b > 0:
                pass
            js = [k for k in range(MODEL["nj"]) if MODEL["joint_body"][k] == b]
            n += 1  # the body itself
            for k in js:
                if k == j:
                    return n
                n += 1
        raise KeyError(j)

    def fk_cached(self):
        key = self.phys.tobytes()
        if getattr(self, "_fk_key", None) != key:
            self._fk_key, self._fk_val = key, py_fk(self.phys)
        return self._fk_val

    def reset(self):
        self.phys[:] = self._O.default_phys()
        self.robot.robot_specific_reset(None)  # the reference's CustomHumanoidRobot method
        self.parts["floor"] = FakePart(self, "floor", 0)  # WalkerBaseBulletEnv.reset -> addToScene(ground plane)
        self._resets += 1
        return self.robot.calc_state()

    def global_step(self):
        if self.physics == "none":
            return
        tau = np.array([j.torque for j in self.ordered_joints])
        self.phys[:] = self._O.physics_step(self.phys, tau)

    def close(self):
        pass

    def render(self, mode="human"):
        return np.zeros((1, 1, 3), np.uint8)


class WalkerBase:
    """Stand-in for pybullet_envs.robot_locomotors.WalkerBase; calc_state restated from the upstream source."""
    foot_list = []
    reset_rng = None  # persistent generator for the joint noise of robot_specific_reset (None: a fresh seeded one per reset)

    def __init__(self, fn, robot_name, action_dim, obs_dim, power):
        self.model_xml, self.robot_name, self.power = fn, robot_name, power
        self.walk_target_x, self.walk_target_y = 1e3, 0
        self.initial_z = None
        self.body_xyz = [0, 0, 0]

    def robot_specific_reset(self, bullet_client):
        rng = WalkerBase.reset_rng or np.random.default_rng(12345)
        for j in self.ordered_joints:
            j.reset_current_position(rng.uniform(low=-0.1, high=0.1), 0)
        self.feet = [self.parts[f] for f in self.foot_list]
        self.feet_contact = np.array([0.0 for f in self.foot_list], dtype=np.float32)
        self.scene.actor_introduce(self)
        self.initial_z = None

    def calc_state(self):
        j = np.array([j.current_relative_position() for j in self.ordered_joints], dtype=np.float32).flatten()
        self.joint_speeds = j[1::2]
        self.joints_at_limit = np.count_nonzero(np.abs(j[0::2]) > 0.99)
        body_pose = self.robot_body.pose()
        parts_xyz = np.array([p.pose().xyz() for p in self.parts.values()]).flatten()
        self.body_xyz = (parts_xyz[0::3].mean(), parts_xyz[1::3].mean(), body_pose.xyz()[2])
        self.body_real_xyz = body_pose.xyz()
        self.body_rpy = body_pose.rpy()
        z = self.body_xyz[2]
        if self.initial_z is None:
            self.initial_z = z
        r, p, yaw = self.body_rpy
        self.walk_target_theta = np.arctan2(self.walk_target_y - self.body_xyz[1], self.walk_target_x - self.body_xyz[0])
        self.walk_target_dist = np.linalg.norm([self.walk_target_y - self.body_xyz[1], self.walk_target_x - self.body_xyz[0]])
        angle_to_target = self.walk_target_theta - yaw
        rot_speed = np.array([[np.cos(-yaw), -np.sin(-yaw), 0], [np.sin(-yaw), np.cos(-yaw), 0], [0, 0, 1]])
        vx, vy, vz = np.dot(rot_speed, self.robot_body.speed())
        more = np.array([z - self.initial_z, np.sin(angle_to_target), np.cos(angle_to_target), 0.3 * vx, 0.3 * vy,
                         0.3 * vz, r, p], dtype=np.float32)
        return np.clip(np.concatenate([more] + [j] + [self.feet_contact]), -5, +5)


_installed = False


def install():
    """Put the stand-ins into sys.modules and make the reference importable.  Idempotent."""
    global _installed
    if _installed:
        return
    assert available(), "reference checkout not found at %s" % REF

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class Env:
        pass

    class MultiAgentEnv:
        pass

    class Scene:
        def __init__(self, *a, **k):
            pass

    class MJCFBaseBulletEnv:
        def __init__(self, *a, **k):
            pass

    spaces = mod("gym.spaces", Box=_Box, Discrete=object, Tuple=object)
    mod("gym", Env=Env, spaces=spaces)
    mod("pybullet", addUserDebugLine=lambda *a, **k: 0, GUI=1, DIRECT=2)
    mod("pybullet_data", getDataPath=lambda: "/nonexistent")
    pe = mod("pybullet_envs")
    pe.gym_locomotion_envs = mod("pybullet_envs.gym_locomotion_envs", HumanoidBulletEnv=FakeFlatEnv)
    pe.robot_locomotors = mod("pybullet_envs.robot_locomotors", WalkerBase=WalkerBase)
    pe.env_bases = mod("pybullet_envs.env_bases", MJCFBaseBulletEnv=MJCFBaseBulletEnv)
    pe.scene_abstract = mod("pybullet_envs.scene_abstract", Scene=Scene)
    pe.robot_bases = mod("pybullet_envs.robot_bases", BodyPart=FakePart)
    ray = mod("ray")
    ray.rllib = mod("ray.rllib")
    ray.rllib.env = mod("ray.rllib.env", MultiAgentEnv=MultiAgentEnv)

    home = tempfile.mkdtemp(prefix="ilrl_home_")
    ta = os.path.join(home, "GitHub", "TA")
    os.makedirs(ta)
    for name in os.listdir(REF):
        os.symlink(os.path.join(REF, name), os.path.join(ta, name))
    # DECLARED DATA SUBSTITUTION: hier_env_2.py:64 reads ~/GitHub/TA/Relative_Joints_CSV, which the
    # reference repo does not ship; the four file names it opens per motion are exactly those of "Joints CSV With Hand".
    os.symlink(os.path.join(REF, "Joints CSV With Hand"), os.path.join(ta, "Relative_Joints_CSV"))
    os.environ["HOME"] = home
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if _ROOT not in sys.path:
        sys.path.insert(0, _ROOT)
    _installed = True


class LoggingRng:
    """Replaces env.rng: same draws as a seeded numpy Generator, every integers() result recorded."""

    def __init__(self, seed):
        self.g = np.random.default_rng(seed)
        self.log = []
        self.forced = []

    def integers(self, lo, hi=None):
        v = int(self.forced.pop(0)) if self.forced else int(self.g.integers(lo, hi))
        self.log.append(v)
        return v


def make_low_env(clip="motion09_03", seed=0, physics="oracle"):
    install()
    import humanoid
    import low_level_env
    FakeFlatEnv.physics = physics
    env = low_level_env.LowLevelHumanoidEnv(reference_name=clip, customRobot=humanoid.CustomHumanoidRobot())
    # declared divergence: motion13_13's velocity table has 120 rows for 220 position rows, so the
    # reference raises IndexError beyond frame 119; the frame range is clamped to the rows that exist.
    env.max_frame = min(env.max_frame, len(env.joints_vel_df))
    env.rng = LoggingRng(seed)
    return env


def make_hier_env(seed=0, physics="oracle"):
    install()
    import hier_env
    import humanoid
    FakeFlatEnv.physics = physics
    env = hier_env.HierarchicalHumanoidEnv(customRobot=humanoid.CustomHumanoidRobot())
    env.rng = LoggingRng(seed)
    return env


def env_words(env, hier=False):
    """The reference env's bookkeeping attributes in the flat ILRL_E_* layout (ilrl_constants.h)."""
    e = np.zeros(26)
    e[0] = env.selected_motion_frame if hier else env.frame
    e[1] = 0
    e[2] = env.cur_timestep
    e[3:5] = env.target[:2]
    e[5:7] = env.starting_robot_pos[:2]
    e[7:10] = env.starting_ep_pos[:3]
    e[10:12] = env.robot_pos[:2]
    e[12] = env.highLevelDegTarget
    e[13] = env.flat_env.robot.walk_target_x
    e[14] = env.flat_env.robot.walk_target_y
    e[15] = env.lowTargetScore
    e[16] = env.deltaJoints
    e[17] = env.deltaVelJoints
    e[18] = env.bodyPostureScore
    if getattr(env, "cur_obs", None) is not None:
        e[19:21] = env.cur_obs[1:3]
    if hier:
        e[21] = env.steps_remaining_at_level
        e[22] = env.cumulative_driftScore
        e[23] = env.highTargetScore
        e[24] = env.cumulative_aliveReward
    return e


def set_env_words(env, e, hier=False):
    """Inverse of env_words (injected-state mode)."""
    if hier:
        env.selected_motion_frame = int(e[0])
    else:
        env.frame = int(e[0])
    env.cur_timestep = int(e[2])
    env.target = np.array([e[3], e[4], 0.0])
    env.starting_robot_pos = np.array([e[5], e[6], 0.0])
    env.starting_ep_pos = np.array([e[7], e[8], e[9]])
    env.robot_pos = np.array([e[10], e[11], 0.0])
    env.highLevelDegTarget = float(e[12])
    env.setWalkTarget(float(e[13]), float(e[14]))
    env.lowTargetScore = float(e[15])
    env.deltaJoints = float(e[16])
    env.deltaVelJoints = float(e[17])
    env.bodyPostureScore = float(e[18])
    if hier:
        env.steps_remaining_at_level = int(e[21])
        env.cumulative_driftScore = float(e[22])
        env.highTargetScore = float(e[23])
        env.cumulative_aliveReward = float(e[24])


def terms_of(env):
    """The 12 attributes RewardLogCallback reads, in ILRL_T_* order."""
    return np.array([env.deltaJoints, env.deltaVelJoints, env.delta_lowTargetScore, env.electricityScore,
                     env.jointLimitScore, env.aliveReward, env.bodyPostureScore, env.lowTargetScore,
                     env.deltaEndPoints, env.highTargetScore, env.driftScore, env.delta_highTargetScore],
                    dtype=np.float64)


# ----------------------------------------------------------------------------- hier_env_2.py
def make_hier2_env(seed=0, physics="oracle"):
    """The unmodified REF hier_env_2.HierarchicalHumanoidEnv under two DECLARED SUBSTITUTIONS:
    data  - ~/GitHub/TA/Relative_Joints_CSV (not shipped) -> "Joints CSV With Hand" (same four file names per motion);
    robot - `HumanoidBulletEnv()` (REF hier_env_2.py:45) builds pybullet_envs' stock `Humanoid` on humanoid_symmetric.xml
            (not shipped either).  The stand-in is the reference's CustomHumanoidRobot (humanoid_symmetric_2.xml, same
            motor order / gears / power / initial_z as the stock class) with the stock class's foot_list, so calc_state
            returns the stock 44 entries; the two feet_contact entries stay 0 because hier_env_2 never calls
            flat_env.step(), the only place upstream refreshes them."""
    install()
    import humanoid
    import hier_env_2

    class StockHumanoidStandIn(humanoid.CustomHumanoidRobot):
        foot_list = ["right_foot", "left_foot"]

    FakeFlatEnv.physics = physics
    FakeFlatEnv.default_robot = StockHumanoidStandIn
    WalkerBase.reset_rng = np.random.default_rng(seed + 777)   # the arm joints keep this noise (hier_env_2.py:214-252)
    try:
        env = hier_env_2.HierarchicalHumanoidEnv()
    finally:
        FakeFlatEnv.default_robot = None
    env.rng = LoggingRng(seed)
    return env


def env_words2(env):
    """hier_env_2 bookkeeping in the ILRL_E_* layout; MODE 2 re-uses two words (ilrl_constants.h ILRL_E2_*)."""
    e = np.zeros(26)
    e[0] = env.selected_motion_frame
    e[2] = env.cur_timestep
    e[3:5] = env.target[:2]
    e[5:7] = env.starting_robot_pos[:2]
    e[7:10] = env.starting_ep_pos[:3]
    e[10:12] = env.robot_pos[:2]
    e[12] = env.highLevelDegTarget
    e[13] = env.flat_env.robot.walk_target_x
    e[14] = env.flat_env.robot.walk_target_y
    e[15] = env.cumulative_deltaVelJoints_low
    e[16] = env.deltaJoints
    e[17] = env.deltaVelJoints
    e[18] = env.bodyPostureScore
    if getattr(env, "cur_obs", None) is not None:
        e[19:21] = env.cur_obs[1:3]
    e[21] = env.steps_remaining_at_level
    e[22] = env.cumulative_driftScore
    e[23] = env.highTargetScore
    e[24] = env.cumulative_deltaJoints_low
    return e


def set_env_words2(env, e):
    env.selected_motion_frame = int(e[0])
    env.cur_timestep = int(e[2])
    env.target = np.array([e[3], e[4], 0.0])
    env.starting_robot_pos = np.array([e[5], e[6], 0.0])
    env.starting_ep_pos = np.array([e[7], e[8], e[9]])
    env.robot_pos = np.array([e[10], e[11], 0.0])
    env.highLevelDegTarget = float(e[12])
    env.setWalkTarget(float(e[13]), float(e[14]))
    env.cumulative_deltaVelJoints_low = float(e[15])
    env.deltaJoints = float(e[16])
    env.deltaVelJoints = float(e[17])
    env.bodyPostureScore = float(e[18])
    env.steps_remaining_at_level = int(e[21])
    env.cumulative_driftScore = float(e[22])
    env.highTargetScore = float(e[23])
    env.cumulative_deltaJoints_low = float(e[24])


def terms_of2(env):
    """ILRL_T_* row in MODE 2: slots 2 and 8 (delta_lowTargetScore / deltaEndPoints, constant 0 in hier_env_2) carry
    deltaJoints_low / deltaVelJoints_low."""
    return np.array([env.deltaJoints, env.deltaVelJoints, env.deltaJoints_low, env.electricityScore,
                     env.jointLimitScore, env.aliveReward, env.bodyPostureScore, env.lowTargetScore,
                     env.deltaVelJoints_low, env.highTargetScore, env.driftScore, env.delta_highTargetScore],
                    dtype=np.float64)
