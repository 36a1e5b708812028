"""RLlib-shaped batched faces of the path.

ray is not part of this image, so the classes are duck-typed to the ray 1.2.0 interfaces the reference trains
through (`ray.rllib.env.VectorEnv` for the single-agent env, `ray.rllib.env.BaseEnv` for the two-agent one) and
subclass them when ray is importable.  A creator registered with `register_env` may return these instead of the
N = 1 classes; RLlib then drives thousands of envs per worker through ONE kernel launch per step.

  LowLevelVectorEnv   vector_reset / reset_at / vector_step / get_unwrapped      (REF train_config.py:13-15, 29)
  HierBaseEnv         poll / send_actions / try_reset / get_unwrapped / stop     (REF train_config.py:18-27, 261)

`get_unwrapped()[0]` exposes the attributes `RewardLogCallback` reads per step (REF custom_callback.py:41-81).
"""
import numpy as np
import torch

from . import batched_env as B
from .batched_env import BatchedHumanoidEnv
from .ref_api import Box, _REWARD_ATTRS

try:
    from ray.rllib.env import BaseEnv as _BaseEnv, VectorEnv as _VectorEnv
except Exception:  # pragma: no cover - ray is not part of this image
    class _VectorEnv(object):
        pass

    class _BaseEnv(object):
        pass

HIGH, LOW = "high_level_agent", "low_level_agent"


def policy_mapping_fn(agent_id):
    """REF train_config.py:23-27."""
    return "low_level_policy" if agent_id.startswith("low_level_") else "high_level_policy"


class _EnvAttrView(object):
    """What `base_env.get_unwrapped()[i]` has to look like for RewardLogCallback: the reward attributes of env i
    after the last step, read from the step kernel's `terms` row (host copy refreshed once per step)."""

    def __init__(self, owner, index):
        self._o, self._i = owner, index
        self.baseReward = 0

    def __getattr__(self, name):
        if name in _REWARD_ATTRS:
            return float(self._o._terms_host[self._i, _REWARD_ATTRS[name]])
        if name == "robot_pos":
            xy = self._o._robot_xy()[self._i]       # one device->host copy per step for ALL envs, on first use
            return np.array([xy[0], xy[1], 0.0])
        raise AttributeError(name)



class _RobotPosCache(object):
    """`robot_pos` of every env, fetched from the handle at most once per step (RewardLogCallback reads it for each
    env after each step, REF custom_callback.py:43-46: one [N,28] device->host copy instead of N)."""

    def _robot_xy(self):
        if self._robot_cache is None:
            envf = self.env.get_state()[1]
            self._robot_cache = envf[:, [B.E_ROBOT_X, B.E_ROBOT_Y]].cpu().numpy().astype(np.float64)
        return self._robot_cache


class LowLevelVectorEnv(_RobotPosCache, _VectorEnv):
    """N `LowLevelHumanoidEnv`s as one RLlib VectorEnv.  Done envs are re-initialised by ONE masked reset launch
    right after the step; `reset_at(i)` (which RLlib calls for every done env) returns that env's new first obs."""

    def __init__(self, num_envs, reference_name="motion09_03", device=0, seed=0, self_collision=False):
        self.env = BatchedHumanoidEnv(num_envs, "low", clips=[reference_name], device=device, seed=seed, auto_reset=False,
                                      self_collision=self_collision)
        self.num_envs = int(num_envs)
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=[70])
        self.action_space = Box(low=-1, high=1, shape=[17])
        # pinned host buffers: ilrl_step_host reads the actions from / writes the results to them in place (zero-copy)
        pin = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype).pin_memory().numpy()  # noqa: E731
        self._act_host, self._obs_host = pin(self.num_envs, 17), pin(self.num_envs, 70)
        self._rew_host, self._done_host = pin(self.num_envs), pin(self.num_envs, dtype=torch.uint8)
        self._terms_host = pin(self.num_envs, B.TERM_WORDS)
        self._next_obs = np.zeros((self.num_envs, 70), np.float32)
        self._robot_cache = None
        self._views = [_EnvAttrView(self, i) for i in range(self.num_envs)]

    def vector_reset(self):
        obs = self.env.reset().cpu().numpy()
        self._next_obs[:] = obs
        return [o for o in obs.astype(np.float64)]

    def reset_at(self, index):
        return self._next_obs[index].astype(np.float64)

    def vector_step(self, actions):
        self._act_host[:] = np.asarray(actions, dtype=np.float32).reshape(self.num_envs, 17)
        assert np.isfinite(self._act_host).all()
        self.env.step_host(self._act_host, self._obs_host, self._rew_host, self._done_host, self._terms_host)
        self._robot_cache = None
        done_h = self._done_host.astype(bool)
        obs64 = self._obs_host.astype(np.float64)
        if done_h.any():  # one masked reset for every env that finished; their first obs waits for reset_at()
            mask = torch.from_numpy(self._done_host).to(self.env.device)
            self._next_obs[done_h] = self.env.reset(mask=mask).cpu().numpy()[done_h]
        return ([o for o in obs64], [float(r) for r in self._rew_host], [bool(d) for d in done_h],
                [{} for _ in range(self.num_envs)])

    def get_unwrapped(self):
        return self._views

    def close(self):
        self.env.close()


class HierBaseEnv(_RobotPosCache, _BaseEnv):
    """N `HierarchicalHumanoidEnv`s as one RLlib BaseEnv (async poll / send_actions protocol, agent ids
    "high_level_agent" / "low_level_agent").  Each env is either waiting for a heading from the high-level agent or
    for torques from the low-level one; both kinds advance in the same call: rows of agents that do not act are NaN
    and the kernels skip them (include/ilrl.h).  variant "hier2" = N envs of REF hier_env_2.py (36-d high action carrying
    the joint targets, 60-d / 72-d observations)."""

    def __init__(self, num_envs, device=0, seed=0, motion_list=("motion08_03", "motion09_03"), selected_motion=1,
                 variant="hier", self_collision=False):
        self.num_envs = int(num_envs)
        self.env = BatchedHumanoidEnv(num_envs, variant, clips=list(motion_list),
                                      clip_of_env=np.full(num_envs, selected_motion, np.int32), device=device,
                                      seed=seed, auto_reset=False, self_collision=self_collision)
        self.high_level_obs_space = Box(low=-np.inf, high=np.inf, shape=[self.env.hobs_w])
        self.high_level_act_space = Box(low=-1, high=1, shape=[self.env.hact_w])
        self.low_level_obs_space = Box(low=-np.inf, high=np.inf, shape=[self.env.obs_w])
        self.low_level_act_space = Box(low=-1, high=1, shape=[17])
        self._terms_host = np.zeros((self.num_envs, B.TERM_WORDS), np.float32)
        self._robot_cache = None
        self._views = [_EnvAttrView(self, i) for i in range(self.num_envs)]
        self._pending = None  # what the next poll() returns
        hobs = self.env.reset().cpu().numpy().astype(np.float64)
        self._pending = ({i: {HIGH: hobs[i]} for i in range(self.num_envs)}, {i: {} for i in range(self.num_envs)},
                         {i: {"__all__": False} for i in range(self.num_envs)})

    def poll(self):
        obs, rew, done = self._pending
        self._pending = ({}, {}, {})
        infos = {i: {a: {} for a in o} for i, o in obs.items()}
        return obs, rew, done, infos, {}

    def send_actions(self, action_dict):
        n = self.num_envs
        low = np.full((n, 17), np.nan, np.float32)
        high = np.full((n, self.env.hact_w), np.nan, np.float32)
        for i, acts in action_dict.items():
            assert len(acts) == 1, acts
            (agent, a), = acts.items()
            a = np.asarray(a, dtype=np.float32)
            assert np.isfinite(a).all()
            if agent == HIGH:
                high[i] = a
            else:
                low[i] = a
        dev = self.env.device
        self._robot_cache = None
        got_high = ~np.isnan(high[:, 0])
        got_low = ~np.isnan(low[:, 0])
        obs, rew, done = {}, {}, {}
        if got_high.any():
            lo = self.env.high_step(torch.from_numpy(high).to(dev)).cpu().numpy().astype(np.float64)
            for i in np.nonzero(got_high)[0]:
                obs[int(i)] = {LOW: lo[i]}
                rew[int(i)] = {LOW: 0}
                done[int(i)] = {"__all__": False}
        if got_low.any():
            o, r, d, terms = self.env.step(torch.from_numpy(low).to(dev))
            ho, hr, hf = self.env.high_readout()
            o, r, ho, hr, hf = [t.cpu().numpy() for t in (o, r, ho, hr, hf)]
            self._terms_host[got_low] = terms.cpu().numpy()[got_low]
            for i in np.nonzero(got_low)[0]:
                i = int(i)
                f = int(hf[i])
                if f & 1:    # episode over: both agents get their last obs / reward (Q19)
                    obs[i] = {HIGH: ho[i].astype(np.float64), LOW: o[i].astype(np.float64)}
                    rew[i] = {HIGH: float(hr[i]), LOW: float(r[i])}
                    done[i] = {"__all__": True}
                elif f & 2:  # level boundary: the high-level agent acts next
                    obs[i] = {HIGH: ho[i].astype(np.float64)}
                    rew[i] = {HIGH: float(hr[i])}
                    done[i] = {"__all__": False}
                else:
                    obs[i] = {LOW: o[i].astype(np.float64)}
                    rew[i] = {LOW: float(r[i])}
                    done[i] = {"__all__": False}
        self._pending = (obs, rew, done)

    def try_reset(self, env_id=None):
        mask = np.zeros(self.num_envs, np.uint8)
        mask[env_id] = 1
        hobs = self.env.reset(mask=mask).cpu().numpy().astype(np.float64)
        return {HIGH: hobs[env_id]}

    def get_unwrapped(self):
        return self._views

    def stop(self):
        self.env.close()
