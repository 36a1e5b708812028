"""On-device rollout collection for the low-level env (BASELINE cfg 5).

The reference collects PPO sample batches with RLlib rollout workers: per env step a TF policy forward on the CPU
worker, `env.step`, and at the end of the fragment `compute_advantages` on the host (REF train_config.py:91-113:
2 x 256 tanh MLP, free log-std, gamma 0.99, lambda 0.9).  Here the whole fragment stays on the GPU:

    for t in 0..T-1:   mean, value = policy(obs)              torch (cuBLAS GEMMs: plain library GEMMs, plumbing)
                       action = mean + exp(log_std) * eps      torch Philox
                       obs, reward, done = env.step(clip(a))   ONE fused CUDA kernel (csrc/)
    advantages, value_targets = GAE(rewards, values, dones)    CUDA kernel (ilrl_gae)

and the T-step loop is captured once into a CUDA graph, so an iteration is one graph launch: no Python, no launch
gaps, observations and actions never leave HBM.  Output columns follow RLlib's SampleBatch names.
"""
import ctypes as C

import torch

from . import _lib
from .batched_env import ACT_LOW, OBS_LOW, BatchedHumanoidEnv, _ptr


class GaussianMLPPolicy(torch.nn.Module):
    """RLlib's default fully-connected model for the reference's low-level policy: separate policy and value
    branches (vf_share_layers False), `fcnet_hiddens` [256, 256], tanh, state-independent log-std."""

    def __init__(self, obs_dim=OBS_LOW, act_dim=ACT_LOW, hiddens=(256, 256)):
        super().__init__()

        def mlp(out):
            layers, d = [], obs_dim
            for h in hiddens:
                layers += [torch.nn.Linear(d, h), torch.nn.Tanh()]
                d = h
            return torch.nn.Sequential(*layers, torch.nn.Linear(d, out))
        self.pi, self.vf = mlp(act_dim), mlp(1)
        self.log_std = torch.nn.Parameter(torch.zeros(act_dim))

    def forward(self, obs):
        return self.pi(obs), self.vf(obs).squeeze(-1)


def gae(rewards, values, dones, gamma, lam, stream=None):
    """advantages, value_targets = GAE over a step-major rollout.  rewards [T,N] f32, values [T+1,N] f32,
    dones [T,N] uint8; all contiguous CUDA tensors."""
    T, n = rewards.shape
    assert values.shape == (T + 1, n) and dones.shape == (T, n) and dones.dtype == torch.uint8
    assert rewards.is_cuda and rewards.is_contiguous() and values.is_contiguous() and dones.is_contiguous()
    adv, ret = torch.empty_like(rewards), torch.empty_like(rewards)
    st = C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream if stream is None else stream)
    rc = _lib.lib().ilrl_gae(_ptr(rewards), _ptr(values), _ptr(dones), gamma, lam, _ptr(adv), _ptr(ret), T, n, st)
    if rc != 0:
        raise _lib.IlrlError("ilrl_gae failed (%d)" % rc)
    return adv, ret


class RolloutCollector:
    """Collects [T, N] fragments from a BatchedHumanoidEnv (mode "low", auto_reset=True) with a policy living on the
    same GPU.  `collect()` returns a dict of step-major device tensors named like RLlib's SampleBatch columns."""

    def __init__(self, env, policy=None, horizon=8, gamma=0.99, lam=0.9, seed=0, use_graph=True, autocast_dtype=None):
        assert isinstance(env, BatchedHumanoidEnv) and env.mode == 0
        self.env, self.T, self.gamma, self.lam = env, int(horizon), float(gamma), float(lam)
        self.autocast_dtype = autocast_dtype  # e.g. torch.bfloat16: policy GEMMs on the tensor cores (inference only)
        dev, n, T = env.device, env.num_envs, self.T
        self.policy = (policy if policy is not None else GaussianMLPPolicy()).to(dev).eval()
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)
        f = dict(device=dev, dtype=torch.float32)
        self.cur_obs = torch.zeros(n, OBS_LOW, **f)
        self.buf = {
            "obs": torch.zeros(T, n, OBS_LOW, **f), "new_obs": torch.zeros(T, n, OBS_LOW, **f),
            "actions": torch.zeros(T, n, ACT_LOW, **f), "rewards": torch.zeros(T, n, **f),
            "dones": torch.zeros(T, n, device=dev, dtype=torch.uint8), "action_logp": torch.zeros(T, n, **f),
            "vf_preds": torch.zeros(T + 1, n, **f),
        }
        self._noise = torch.zeros(T, n, ACT_LOW, **f)
        self._graph = None
        self._use_graph = use_graph
        self.cur_obs.copy_(env.reset())

    @torch.no_grad()
    def _loop(self):
        b, env = self.buf, self.env
        std = self.policy.log_std.exp()
        def forward(obs):
            if self.autocast_dtype is None:
                return self.policy(obs)
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                m, v = self.policy(obs)
            return m.float(), v.float()
        for t in range(self.T):
            mean, v = forward(self.cur_obs)
            a = mean + std * self._noise[t]
            b["obs"][t].copy_(self.cur_obs)
            b["actions"][t].copy_(a)
            b["vf_preds"][t].copy_(v)
            b["action_logp"][t].copy_((-0.5 * self._noise[t] ** 2 - self.policy.log_std - 0.9189385332046727).sum(-1))
            # RLlib clips actions to the Box before env.step (clip_actions=True); the batch keeps the raw sample
            obs, rew, done, _ = env.step(a.clamp(-1.0, 1.0))
            b["new_obs"][t].copy_(obs)
            b["rewards"][t].copy_(rew)
            b["dones"][t].copy_(done)
            self.cur_obs.copy_(obs)
        b["vf_preds"][self.T].copy_(forward(self.cur_obs)[1])

    @torch.no_grad()
    def collect(self):
        """One fragment of T steps for every env.  The noise is drawn up front (one Philox launch), the T-step loop
        is a CUDA graph replay after the first call."""
        self._noise.normal_(generator=self.gen)
        if not self._use_graph:
            self._loop()
        elif self._graph is None:
            s = torch.cuda.Stream(device=self.env.device)
            s.wait_stream(torch.cuda.current_stream(self.env.device))
            with torch.cuda.stream(s):   # warm-up on a side stream (cuBLAS workspaces, lazy module loads)
                saved = [t.clone() for t in self.env.get_state()] + [self.cur_obs.clone()]
                self._loop()
                self.env.set_state(saved[0], saved[1])
                self.cur_obs.copy_(saved[2])
            torch.cuda.current_stream(self.env.device).wait_stream(s)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._loop()
            self._graph.replay()
        else:
            self._graph.replay()
        b = self.buf
        adv, vt = gae(b["rewards"], b["vf_preds"], b["dones"], self.gamma, self.lam)
        out = {k: b[k] for k in ("obs", "new_obs", "actions", "rewards", "dones", "action_logp")}
        out["vf_preds"] = b["vf_preds"][:self.T]
        out["advantages"], out["value_targets"] = adv, vt
        return out
