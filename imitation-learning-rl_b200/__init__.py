"""imitation-learning-rl_b200 — B200-native batched humanoid imitation env (one hot path of
AdityaPutraS/Imitation-Learning-RL): env step / reset + imitation reward + observation as sm_100a CUDA kernels behind
the C ABI of include/ilrl.h.  Import name: `ilrl_b200` (see ilrl_b200.py at the repo root; the directory name has a
hyphen).  The CUDA library is required; nothing here falls back to a CPU implementation."""
from . import _build, _lib  # noqa: F401
from .batched_env import BatchedHumanoidEnv  # noqa: F401
from .clips import CLIP_NAMES, load_clip  # noqa: F401
from .ref_api import HierarchicalHumanoidEnv, HierarchicalHumanoidEnv2, LowLevelHumanoidEnv  # noqa: F401
from .rllib_adapters import HierBaseEnv, LowLevelVectorEnv, policy_mapping_fn  # noqa: F401
from . import stats  # noqa: F401
from .rollout import FusedPolicy, GaussianMLPPolicy, HierRolloutCollector, RolloutCollector, gae, gae_decisions  # noqa: F401

__all__ = ["BatchedHumanoidEnv", "CLIP_NAMES", "load_clip", "LowLevelHumanoidEnv", "HierarchicalHumanoidEnv", "HierarchicalHumanoidEnv2",
           "LowLevelVectorEnv", "HierBaseEnv", "policy_mapping_fn", "stats", "RolloutCollector", "HierRolloutCollector",
           "GaussianMLPPolicy", "FusedPolicy", "gae", "gae_decisions"]
