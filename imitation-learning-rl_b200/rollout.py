"""On-device rollout collection for the low-level env (BASELINE cfg 5).

The reference collects PPO sample batches with RLlib rollout workers: per env step a TF policy forward on the CPU
worker, `env.step`, and at the end of the fragment `compute_advantages` on the host (REF train_config.py:91-113:
2 x 256 tanh MLP, free log-std, gamma 0.99, lambda 0.9).  Here the whole fragment stays on the GPU:

    for t in 0..T-1:   mean, value = policy(obs)               }  ONE tcgen05 kernel (csrc/ilrl_policy.cu): six GEMMs on
                       action = mean + exp(log_std) * eps       }  the tensor cores, tanh / sampling / log-prob in the
                       logp, clip(action)                       }  epilogues, hidden activations never leave the SM
                       obs, reward, done = env.step(clip(a))   ONE fused CUDA kernel (csrc/ilrl_capi.cu), writing
                                                               into the [T, N] rollout buffers in place
    advantages, value_targets = GAE(rewards, values, dones)    CUDA kernel (ilrl_gae)

and the T-step loop (2 launches per step) is captured once into a CUDA graph, so an iteration is one graph launch: no
Python, no launch gaps, observations and actions never leave HBM.  The noise eps is drawn by torch's Philox, one
launch per iteration.  `fused=False` keeps the torch module (cuBLAS) as the model.  Output columns follow RLlib's
SampleBatch names.
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


class FusedPolicy:
    """The GaussianMLPPolicy forward + action sampling as ONE tcgen05 kernel (csrc/ilrl_policy.cu): bf16 operands, fp32
    accumulation, hidden activations kept on chip.  `repack()` after the parameters change (optimizer step)."""

    def __init__(self, policy):
        lin = lambda seq: [m for m in seq if isinstance(m, torch.nn.Linear)]  # noqa: E731
        self.policy, self._pi, self._vf = policy, lin(policy.pi), lin(policy.vf)
        assert len(self._pi) == 3 and len(self._vf) == 3, "two hidden layers"
        self.obs_dim, self.act_dim = self._pi[0].in_features, self._pi[2].out_features
        assert self._pi[0].out_features == 256 and self._pi[1].out_features == 256 and self._vf[2].out_features == 1
        assert self._vf[0].out_features == 256 and self._vf[1].out_features == 256 and self._vf[0].in_features == self.obs_dim
        assert self.obs_dim <= 80 and self.act_dim <= 32
        self.L = _lib.lib()
        self.device = self._pi[0].weight.device
        assert self.device.type == "cuda", "the fused policy is a CUDA kernel: move the policy to the GPU first"
        self.blob = torch.empty(int(self.L.ilrl_policy_blob_bytes()), dtype=torch.uint8, device=self.device)
        self.repack()

    def _st(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def repack(self):
        ps = []
        for net in (self._pi, self._vf):
            for m in net:
                ps += [m.weight.detach().float().contiguous(), m.bias.detach().float().contiguous()]
        ps.append(self.policy.log_std.detach().float().contiguous())
        with torch.cuda.device(self.device):   # the handle-less C entry points launch on the CURRENT device
            rc = self.L.ilrl_policy_pack(*[_ptr(p) for p in ps], self.obs_dim, self.act_dim, _ptr(self.blob), self._st())
        if rc != 0:
            raise _lib.IlrlError("ilrl_policy_pack failed (%d)" % rc)
        self._keep = ps   # alive until the pack kernel has certainly run

    def step(self, obs, noise=None, action=None, action_clipped=None, logp=None, value=None):
        """Outputs are written into the given contiguous fp32 CUDA tensors (None: not produced)."""
        n = obs.shape[0]
        assert obs.is_cuda and obs.dtype == torch.float32 and obs.is_contiguous() and obs.shape == (n, self.obs_dim)
        for t, shape in ((noise, (n, self.act_dim)), (action, (n, self.act_dim)), (action_clipped, (n, self.act_dim)),
                         (logp, (n,)), (value, (n,))):
            assert t is None or (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shape)
        p = lambda t: None if t is None else _ptr(t)  # noqa: E731
        with torch.cuda.device(self.device):
            rc = self.L.ilrl_policy_step(_ptr(self.blob), _ptr(obs), p(noise), p(action), p(action_clipped), p(logp), p(value),
                                         self.obs_dim, self.act_dim, n, self._st())
        if rc != 0:
            raise _lib.IlrlError("ilrl_policy_step failed (%d)" % rc)


def gae(rewards, values, dones, gamma, lam, stream=None):
    """advantages, value_targets = GAE over a step-major rollout.  rewards [T,N] f32, values [T+1,N] f32,
    dones [T,N] uint8; all contiguous CUDA tensors."""
    T, n = rewards.shape
    assert values.shape == (T + 1, n) and dones.shape == (T, n) and dones.dtype == torch.uint8
    assert rewards.is_cuda and rewards.is_contiguous() and values.is_contiguous() and dones.is_contiguous()
    adv, ret = torch.empty_like(rewards), torch.empty_like(rewards)
    st = C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream if stream is None else stream)
    with torch.cuda.device(rewards.device):
        rc = _lib.lib().ilrl_gae(_ptr(rewards), _ptr(values), _ptr(dones), gamma, lam, _ptr(adv), _ptr(ret), T, n, st)
    if rc != 0:
        raise _lib.IlrlError("ilrl_gae failed (%d)" % rc)
    return adv, ret


class RolloutCollector:
    """Collects [T, N] fragments from a BatchedHumanoidEnv (mode "low", auto_reset=True) with a policy living on the
    same GPU.  `collect()` returns a dict of step-major device tensors named like RLlib's SampleBatch columns.

    `action_logp` (fused=True) is the log-density of the sampled action under the KERNEL's policy: bf16 operands and
    hidden activations, fp32 accumulation.  A learner that recomputes the old log-probability with the fp32 torch
    module gets importance ratios that differ from 1 by up to exp(sum_j z_j d_j) (d = the two means' difference in
    units of sigma; |logp difference| <= 0.3 on random-init weights, tests/test_gpu_rollout.py): use the stored
    `action_logp` as the old log-probability (what RLlib's PPO does with the SampleBatch column), or collect with
    fused=False when the learner must reproduce it exactly."""

    def __init__(self, env, policy=None, horizon=8, gamma=0.99, lam=0.9, seed=0, use_graph=True, autocast_dtype=None,
                 fused=True):
        assert isinstance(env, BatchedHumanoidEnv) and env.mode == 0
        self.env, self.T, self.gamma, self.lam = env, int(horizon), float(gamma), float(lam)
        self.autocast_dtype = autocast_dtype  # torch path only, e.g. torch.bfloat16
        dev, n, T = env.device, env.num_envs, self.T
        self.policy = (policy if policy is not None else GaussianMLPPolicy()).to(dev).eval()
        # fused=True: model forward + sampling + log-prob are ONE tcgen05 kernel per step (FusedPolicy) and the env
        # kernel writes into the rollout buffers directly: 2 launches per step.  fused=False: the torch module.
        self.fused = FusedPolicy(self.policy) if fused else None
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)
        f = dict(device=dev, dtype=torch.float32)
        self._obs_all = torch.zeros(T + 1, n, OBS_LOW, **f)   # obs = rows 0..T-1, new_obs = rows 1..T
        self.cur_obs = self._obs_all[T]                       # the observation the next fragment starts from
        self.buf = {
            "obs": self._obs_all[:T], "new_obs": self._obs_all[1:],
            "actions": torch.zeros(T, n, ACT_LOW, **f), "rewards": torch.zeros(T, n, **f),
            "dones": torch.zeros(T, n, device=dev, dtype=torch.uint8), "action_logp": torch.zeros(T, n, **f),
            "vf_preds": torch.zeros(T + 1, n, **f),
        }
        self._noise = torch.zeros(T, n, ACT_LOW, **f)
        self._clipped = torch.zeros(n, ACT_LOW, **f)
        # RLlib's `t` / `eps_id` columns (ilrl_episode_columns): carries continue from fragment to fragment
        self._t_carry = torch.zeros(n, device=dev, dtype=torch.int32)
        self._eps_carry = torch.zeros(n, device=dev, dtype=torch.int32)
        self._t_col = torch.zeros(T, n, device=dev, dtype=torch.int32)
        self._eps_col = torch.zeros(T, n, device=dev, dtype=torch.int64)
        self._id_base = 0   # global id of env 0 of this shard (set it when a batch is sharded over ranks)
        self._graph = None
        self._warm = False
        self._use_graph = use_graph
        self.cur_obs.copy_(env.reset())

    def policy_updated(self):
        """Call after the policy's parameters changed (optimizer step): re-packs the fused kernel's weights."""
        if self.fused is not None:
            self.fused.repack()

    @torch.no_grad()
    def _loop(self):
        b, env, T = self.buf, self.env, self.T
        self._obs_all[0].copy_(self._obs_all[T])
        if self.fused is not None:
            for t in range(T):
                self.fused.step(self._obs_all[t], self._noise[t], b["actions"][t], self._clipped, b["action_logp"][t],
                                b["vf_preds"][t])
                # RLlib clips actions to the Box before env.step (clip_actions=True); the batch keeps the raw sample
                env.step_into(self._clipped, self._obs_all[t + 1], b["rewards"][t], b["dones"][t])
            self.fused.step(self._obs_all[T], value=b["vf_preds"][T])
            return
        std = self.policy.log_std.exp()
        def forward(obs):
            if self.autocast_dtype is None:
                return self.policy(obs)
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                m, v = self.policy(obs)
            return m.float(), v.float()
        for t in range(T):
            mean, v = forward(self._obs_all[t])
            a = mean + std * self._noise[t]
            b["actions"][t].copy_(a)
            b["vf_preds"][t].copy_(v)
            b["action_logp"][t].copy_((-0.5 * self._noise[t] ** 2 - self.policy.log_std - 0.9189385332046727).sum(-1))
            self._clipped.copy_(a.clamp(-1.0, 1.0))
            env.step_into(self._clipped, self._obs_all[t + 1], b["rewards"][t], b["dones"][t])
        b["vf_preds"][T].copy_(forward(self._obs_all[T])[1])

    @torch.no_grad()
    def collect(self):
        """One fragment of T steps for every env.  The noise is drawn up front (one Philox launch), the T-step loop
        is a CUDA graph replay after the first call."""
        self._noise.normal_(generator=self.gen)
        if not self._use_graph:
            self._loop()
        elif not self._warm:
            # The first fragment runs eagerly, on a side stream (cuBLAS workspaces, lazy module loads), and COUNTS: nothing
            # is rolled back, so the envs' Philox draw counters and the statistics accumulators stay exactly what an
            # ungraphed collector would have (use_graph=True and use_graph=False give the same trajectories).
            s = torch.cuda.Stream(device=self.env.device)
            s.wait_stream(torch.cuda.current_stream(self.env.device))
            with torch.cuda.stream(s):
                self._loop()
            torch.cuda.current_stream(self.env.device).wait_stream(s)
            self._warm = True
        elif self._graph is None:
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
        st = C.c_void_p(torch.cuda.current_stream(self.env.device).cuda_stream)
        with torch.cuda.device(self.env.device):
            rc = _lib.lib().ilrl_episode_columns(_ptr(b["dones"]), _ptr(self._t_carry), _ptr(self._eps_carry), _ptr(self._t_col),
                                                 _ptr(self._eps_col), self.T, self.env.num_envs, int(self._id_base), st)
        if rc != 0:
            raise _lib.IlrlError("ilrl_episode_columns failed (%d)" % rc)
        out["t"], out["eps_id"] = self._t_col, self._eps_col
        return out


def gae_decisions(rewards, flags, values, gamma, lam):
    """GAE of the high-level agent (irregular decision times).  rewards / flags / values [T+1, N]: the
    `high_readout` rows taken before every tick and once after the last, and the value estimates of those
    observations.  -> advantages, value_targets [T, N] f32 and valid [T, N] uint8 (see include/ilrl.h)."""
    T, n = rewards.shape[0] - 1, rewards.shape[1]
    assert flags.shape == (T + 1, n) and values.shape == (T + 1, n) and flags.dtype == torch.uint8
    assert rewards.is_cuda and rewards.is_contiguous() and flags.is_contiguous() and values.is_contiguous()
    adv, ret = torch.empty(T, n, device=rewards.device), torch.empty(T, n, device=rewards.device)
    valid = torch.empty(T, n, device=rewards.device, dtype=torch.uint8)
    st = C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)
    with torch.cuda.device(rewards.device):
        rc = _lib.lib().ilrl_gae_decisions(_ptr(rewards), _ptr(flags), _ptr(values), gamma, lam, _ptr(adv), _ptr(ret),
                                           _ptr(valid), T, n, st)
    if rc != 0:
        raise _lib.IlrlError("ilrl_gae_decisions failed (%d)" % rc)
    return adv, ret, valid


class HierRolloutCollector:
    """On-device fragments of the hierarchical env (REF hier_env.py; train_config.py:260-321: two policies,
    "high_level_policy" 44-256-256-2 and "low_level_policy" 70-256-256-17, both trained by PPO).

    Every tick (5 launches, all T ticks replayed as one CUDA graph):
        high_readout   -> high obs / reward / flags of every env        (K4 readout)
        high policy    -> heading action for every env                  (tcgen05 policy kernel; used where an env waits)
        high_step      -> envs waiting for a heading take it, their low-level obs row is refreshed
        low policy     -> torques                                       (tcgen05 policy kernel)
        step           -> one low-level env step of every env, into the buffers in place
    The env is created with auto_reset=True: an env whose episode ends is reset inside the step and waits for a
    heading at the next tick.  `collect()` returns {"low": {...}, "high": {...}} of step-major device tensors with
    RLlib's SampleBatch column names; the high-level columns are dense [T, N] with `valid` marking the ticks at which
    an env actually took a high-level decision whose outcome lies inside the fragment (the high-level agent acts once
    per `step_per_level` low-level steps, at times that differ per env)."""

    def __init__(self, env, high_policy=None, low_policy=None, horizon=10, gamma=0.99, lam=0.9, seed=0, use_graph=True):
        assert isinstance(env, BatchedHumanoidEnv) and env.mode == 1, "needs a hier-mode env"
        from .batched_env import ACT_HIGH, OBS_HIGH
        self.env, self.T, self.gamma, self.lam = env, int(horizon), float(gamma), float(lam)
        dev, n, T = env.device, env.num_envs, self.T
        self.high_policy = (high_policy if high_policy is not None else GaussianMLPPolicy(OBS_HIGH, ACT_HIGH)).to(dev).eval()
        self.low_policy = (low_policy if low_policy is not None else GaussianMLPPolicy(OBS_LOW, ACT_LOW)).to(dev).eval()
        self.fused_high, self.fused_low = FusedPolicy(self.high_policy), FusedPolicy(self.low_policy)
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)
        f = dict(device=dev, dtype=torch.float32)
        u8 = dict(device=dev, dtype=torch.uint8)
        self._lobs = torch.zeros(T + 1, n, OBS_LOW, **f)
        self._hobs = torch.zeros(T + 1, n, OBS_HIGH, **f)
        self.low = {"obs": self._lobs[:T], "new_obs": self._lobs[1:], "actions": torch.zeros(T, n, ACT_LOW, **f),
                    "rewards": torch.zeros(T, n, **f), "dones": torch.zeros(T, n, **u8),
                    "action_logp": torch.zeros(T, n, **f), "vf_preds": torch.zeros(T + 1, n, **f)}
        self.high = {"obs": self._hobs[:T], "actions": torch.zeros(T, n, ACT_HIGH, **f),
                     "action_logp": torch.zeros(T, n, **f), "vf_preds": torch.zeros(T + 1, n, **f),
                     "readout_rewards": torch.zeros(T + 1, n, **f), "flags": torch.zeros(T + 1, n, **u8)}
        self._noise_l, self._noise_h = torch.zeros(T, n, ACT_LOW, **f), torch.zeros(T, n, ACT_HIGH, **f)
        self._clip_l, self._clip_h = torch.zeros(n, ACT_LOW, **f), torch.zeros(n, ACT_HIGH, **f)
        self._graph, self._use_graph = None, use_graph
        env.reset()   # every env now waits for its first heading; its low-level obs row is written by high_step

    def policies_updated(self):
        self.fused_high.repack()
        self.fused_low.repack()

    @torch.no_grad()
    def _loop(self):
        env, T, lo, hi = self.env, self.T, self.low, self.high
        self._lobs[0].copy_(self._lobs[T])
        for t in range(T):
            env.high_readout_into(self._hobs[t], hi["readout_rewards"][t], hi["flags"][t])
            self.fused_high.step(self._hobs[t], self._noise_h[t], hi["actions"][t], self._clip_h, hi["action_logp"][t],
                                 hi["vf_preds"][t])
            env.high_step_into(self._clip_h, self._lobs[t])
            self.fused_low.step(self._lobs[t], self._noise_l[t], lo["actions"][t], self._clip_l, lo["action_logp"][t],
                                lo["vf_preds"][t])
            env.step_into(self._clip_l, self._lobs[t + 1], lo["rewards"][t], lo["dones"][t])
        env.high_readout_into(self._hobs[T], hi["readout_rewards"][T], hi["flags"][T])
        self.fused_high.step(self._hobs[T], value=hi["vf_preds"][T])
        self.fused_low.step(self._lobs[T], value=lo["vf_preds"][T])

    @torch.no_grad()
    def collect(self):
        self._noise_l.normal_(generator=self.gen)
        self._noise_h.normal_(generator=self.gen)
        if not self._use_graph:
            self._loop()
        elif self._graph is None:
            # the first fragment runs eagerly (it is also the warm-up); capturing afterwards executes nothing, so the
            # env advances by exactly one fragment per collect() with or without the graph
            self._loop()
            torch.cuda.synchronize(self.env.device)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._loop()
        else:
            self._graph.replay()
        lo, hi, T = self.low, self.high, self.T
        adv, vt = gae(lo["rewards"], lo["vf_preds"], lo["dones"], self.gamma, self.lam)
        low = {k: lo[k] for k in ("obs", "new_obs", "actions", "rewards", "dones", "action_logp")}
        low["vf_preds"], low["advantages"], low["value_targets"] = lo["vf_preds"][:T], adv, vt
        hadv, hvt, valid = gae_decisions(hi["readout_rewards"], hi["flags"], hi["vf_preds"], self.gamma, self.lam)
        high = {k: hi[k] for k in ("obs", "actions", "action_logp")}
        high["vf_preds"], high["advantages"], high["value_targets"], high["valid"] = hi["vf_preds"][:T], hadv, hvt, valid
        high["decided"] = (hi["flags"][:T] & 4) != 0     # every decision, with or without its outcome in the fragment
        high["readout_rewards"], high["flags"] = hi["readout_rewards"], hi["flags"]
        return {"low": low, "high": high}
