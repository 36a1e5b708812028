"""Multi-GPU side of the path: env sharding and the ONLY collective — the sum of the 16-float statistics vector.

Envs are independent: rank r of W owns a contiguous block of env ids and its own handle; there
is no data-path exchange.  Once per rollout iteration every rank reads its device-accumulated statistics
(`BatchedHumanoidEnv.stats()`, filled by the step kernel's warp reductions) and the ranks all-reduce them
(NCCL over NVLink on the GPU box; gloo in the CPU tests).  The names below are the ones the reference's
`RewardLogCallback` writes into `episode.custom_metrics` (REF custom_callback.py:83-125).
"""
import numpy as np

STATS_WORDS = 16
# stats16 = {episodes, sum return, sum length, steps, sum reward, sums of terms[0..10]}
TERM_METRICS = ["deltaJoints", "deltaVelJoints", "delta_lowTargetScore", "electricityScore", "jointLimitScore",
                "aliveReward", "bodyPostureScore", "lowTargetScore", "deltaEndPoints", "highTargetScore", "driftScore"]
CLIP_ORDER = ["motion02_04", "motion08_03", "motion09_03", "motion13_13"]


def shard_envs(total_envs, world_size, rank):
    """Contiguous block of env ids owned by `rank`: (first_id, count).  Blocks differ by at most one env."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, rem = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def clip_of_env(first_id, count, num_clips):
    """BASELINE cfg 4 / SURVEY 8d: env id -> clip index, `clip = env_id mod num_clips` on GLOBAL env ids."""
    return ((np.arange(count, dtype=np.int64) + int(first_id)) % int(num_clips)).astype(np.int32)


def allreduce_stats(stats, group=None):
    """Sum the statistics vector over all ranks, in place; returns it.  No-op without an initialised process group.
    `stats` is a torch tensor of 16 floats on the rank's device (cuda for NCCL, cpu for gloo)."""
    import torch.distributed as dist
    if stats.numel() != STATS_WORDS:
        raise ValueError("statistics vector must hold %d floats, got %d" % (STATS_WORDS, stats.numel()))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def summarize(stats):
    """Means in the reference's vocabulary: `episode_reward_mean`, `episode_len_mean` (RLlib) and the per-step means
    of the reward terms under RewardLogCallback's names."""
    s = np.asarray(stats.detach().cpu().numpy() if hasattr(stats, "detach") else stats, dtype=np.float64)
    episodes, steps = s[0], s[3]
    out = {"episodes_this_iter": int(round(episodes)), "timesteps_this_iter": int(round(steps)),
           "episode_reward_mean": s[1] / episodes if episodes > 0 else float("nan"),
           "episode_len_mean": s[2] / episodes if episodes > 0 else float("nan"),
           "step_reward_mean": s[4] / steps if steps > 0 else float("nan")}
    for i, name in enumerate(TERM_METRICS):
        out["custom_metrics/%s_mean" % name] = s[5 + i] / steps if steps > 0 else float("nan")
    return out
