"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol include/ilrl.h declares, the product
fails loudly without a CUDA device and never reaches into oracle/, the sharding / statistics helpers, and the only
collective of the path (sum of the 16-float statistics vector) under a world-size-2 gloo group."""
import os
import re
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "imitation-learning-rl_b200")


def _header_symbols():
    h = open(os.path.join(ROOT, "include", "ilrl.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(ilrl_[a-z_0-9]+)\s*\(", h)))


def test_capi_exports_every_declared_symbol():
    import ctypes
    import ilrl_b200  # noqa: F401
    from ilrl_b200 import _lib
    L = _lib.lib()
    syms = _header_symbols()
    assert len(syms) >= 19, syms
    for s in syms:
        assert hasattr(L, s), "libilrl_b200.so does not export %s" % s
    assert sorted(_lib.SYMBOLS) == syms, (sorted(set(_lib.SYMBOLS) ^ set(syms)))
    assert isinstance(L, ctypes.CDLL)


def test_create_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import ctypes as C
    from ilrl_b200 import _lib
    from ilrl_b200.batched_env import BatchedHumanoidEnv
    L = _lib.lib()
    cfg = _lib.Config(device=0, num_envs=8, mode=0, auto_reset=1, seed=1, skip_frame=2, max_timestep=3000,
                      step_per_level=5, env_id_base=0)
    h = C.c_void_p()
    rc = L.ilrl_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and not h.value                     # ILRL_ERR_CUDA, no handle
    assert b"no CPU path" in L.ilrl_last_error(None)
    with pytest.raises(_lib.IlrlError):
        BatchedHumanoidEnv(8)
    from ilrl_b200 import LowLevelHumanoidEnv
    with pytest.raises(_lib.IlrlError):
        LowLevelHumanoidEnv("motion09_03")


def test_bad_arguments_are_rejected_before_any_device_work():
    import ctypes as C
    from ilrl_b200 import _lib
    L = _lib.lib()
    h = C.c_void_p()
    assert L.ilrl_create(None, C.byref(h)) == -1
    cfg = _lib.Config(device=0, num_envs=0, mode=0)
    assert L.ilrl_create(C.byref(cfg), C.byref(h)) == -1
    cfg = _lib.Config(device=0, num_envs=4, mode=7)
    assert L.ilrl_create(C.byref(cfg), C.byref(h)) == -1
    assert L.ilrl_step(None, None, None, None, None, None, None) == -1
    assert L.ilrl_launch_count(None) == 0


def test_policy_entry_points_without_a_device():
    """The fused policy kernel's C ABI: blob size is a pure function; argument errors come back before any CUDA call;
    the host wrapper refuses a CPU-resident policy instead of falling back to torch."""
    import ctypes as C
    from ilrl_b200 import GaussianMLPPolicy, _lib
    from ilrl_b200.rollout import FusedPolicy
    L = _lib.lib()
    per_net = (256 * 80 + 2 * 128 * 256 + 32 * 256) * 2       # bf16 operand images: W1 | W2 halves | W3
    assert L.ilrl_policy_blob_bytes() == 2 * per_net + 2 * (256 + 256 + 32) * 4 + 32 * 4
    x = C.c_void_p(16)   # never dereferenced: every call below fails validation first
    assert L.ilrl_policy_step(None, x, None, None, None, None, x, 70, 17, 8, None) == -1
    assert L.ilrl_policy_step(x, None, None, None, None, None, x, 70, 17, 8, None) == -1
    assert L.ilrl_policy_step(x, x, None, None, None, None, None, 70, 17, 8, None) == -1   # no output requested
    assert L.ilrl_policy_step(x, x, None, None, None, None, x, 81, 17, 8, None) == -1
    assert L.ilrl_policy_step(x, x, None, None, None, None, x, 70, 33, 8, None) == -1
    assert L.ilrl_policy_step(x, x, None, None, None, None, x, 70, 17, 0, None) == -1
    assert L.ilrl_policy_step(C.c_void_p(8), x, None, None, None, None, x, 70, 17, 8, None) == -1   # blob not 16-byte aligned
    assert L.ilrl_policy_pack(*([None] * 13), 70, 17, x, None) == -1
    with pytest.raises(AssertionError, match="GPU"):
        FusedPolicy(GaussianMLPPolicy())   # parameters on the CPU


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import, load or mention it."""
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "libilrl_oracle" not in src and "ilrl_oracle.c" not in src.replace("oracle/ilrl_oracle.c", ""), f


def test_missing_library_is_an_error_not_a_fallback(tmp_path, monkeypatch):
    from ilrl_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "SO", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.IlrlError):
        _lib.lib()


def test_clip_tables_and_max_frame_clamp():
    from ilrl_b200 import CLIP_NAMES, load_clip
    rows = {"motion02_04": (299, 298), "motion08_03": (126, 125), "motion09_03": (90, 89), "motion13_13": (220, 120)}
    for name in CLIP_NAMES:
        c = load_clip(name)
        assert (len(c["pos"]), len(c["vel"])) == rows[name]
        assert c["pos"].shape[1] == c["rel"].shape[1] == c["vel"].shape[1] == 14 and c["ep"].shape[1] == 27
        assert c["max_frame"] == min(rows[name][0] - 1, rows[name][1])      # 13_13: 120 (declared divergence)
    with pytest.raises(KeyError):
        load_clip("motion00_00")


def test_shard_envs_and_clip_assignment():
    from ilrl_b200 import stats
    for total, world in ((65536, 8), (65536, 2), (10, 4), (3, 8), (0, 2)):
        blocks = [stats.shard_envs(total, world, r) for r in range(world)]
        assert sum(c for _, c in blocks) == total
        assert all(blocks[r][0] + blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
        assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    assert stats.shard_envs(65536, 8, 3) == (24576, 8192)
    ids = np.concatenate([stats.clip_of_env(*stats.shard_envs(37, 4, r), 4) for r in range(4)])
    np.testing.assert_array_equal(ids, np.arange(37) % 4)
    with pytest.raises(ValueError):
        stats.shard_envs(8, 2, 2)


def test_summarize_uses_the_callback_names():
    from ilrl_b200 import stats
    s = np.zeros(16)
    s[:5] = [4, -20.0, 130, 200, -37.0]
    s[5:] = np.arange(11) * 200.0
    out = stats.summarize(s)
    assert out["episodes_this_iter"] == 4 and out["timesteps_this_iter"] == 200
    assert out["episode_reward_mean"] == -5.0 and out["episode_len_mean"] == 32.5
    assert out["custom_metrics/deltaVelJoints_mean"] == 1.0 and out["custom_metrics/driftScore_mean"] == 10.0
    empty = stats.summarize(np.zeros(16))
    assert np.isnan(empty["episode_reward_mean"]) and np.isnan(empty["step_reward_mean"])


def test_box_stand_in():
    from ilrl_b200.ref_api import Box
    b = Box(low=-1, high=1, shape=[17])
    assert tuple(b.shape) == (17,) and b.contains(b.sample())
    assert not b.contains(np.full(17, 2.0))
    assert tuple(Box(low=-np.inf, high=np.inf, shape=[70]).shape) == (70,)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import ilrl_b200  # noqa: F401
    from ilrl_b200 import stats
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        first, count = stats.shard_envs(37, world, rank)
        # what a rank's step kernel would have accumulated for its shard: one episode of return -i, length i per env id
        ids = np.arange(first, first + count, dtype=np.float64)
        local = torch.zeros(16, dtype=torch.float32)
        local[0] = count
        local[1] = float(-ids.sum())
        local[2] = float(ids.sum())
        local[3] = float(ids.sum())
        local[4] = float(-ids.sum())
        out = stats.allreduce_stats(local.clone())
        q.put((rank, out.numpy().tolist(), stats.summarize(out)))
    finally:
        dist.destroy_process_group()


def test_statistics_allreduce_world_size_2_gloo():
    """The N > 1 path on CPU: two ranks shard 37 envs, each contributes its statistics, the all-reduced vector is
    the whole-job total on both ranks."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tot = float(np.arange(37).sum())
    for rank, vec, summ in res:
        assert vec[0] == 37 and vec[1] == -tot and vec[2] == tot and vec[3] == tot
        assert summ["episodes_this_iter"] == 37 and abs(summ["episode_len_mean"] - tot / 37) < 1e-6


def test_allreduce_is_a_noop_without_a_group():
    import torch
    from ilrl_b200 import stats
    v = torch.arange(16, dtype=torch.float32)
    assert torch.equal(stats.allreduce_stats(v.clone()), v)
    with pytest.raises(ValueError):
        stats.allreduce_stats(torch.zeros(3))


def test_reference_named_tables_and_weights():
    """The DataFrames / dictionaries the reference env exposes as attributes (REF low_level_env.py:58-152): same
    column names as the CSV headers, same weights (python float sums 17.400000000000002 and 8.6)."""
    from ilrl_b200 import ref_api

    class View(ref_api._SingleEnv):
        reference_name = "motion09_03"
    v = View.__new__(View)
    v._df = None
    assert list(v.joints_df.columns) == ref_api.JOINT_COLS and v.joints_df.shape == (90, 14)
    assert v.joints_vel_df.shape == (89, 14) and v.joints_rel_df.shape == (90, 14)
    assert list(v.end_point_df.columns)[:3] == ["LeftLeg_Xposition", "LeftLeg_Yposition", "LeftLeg_Zposition"]
    assert v.end_point_df.shape == (90, 27)
    assert sum(ref_api.JOINT_WEIGHT.values()) == 17.400000000000002 and abs(sum(ref_api.JOINT_VEL_WEIGHT.values()) - 8.6) < 1e-12
    assert list(ref_api.JOINT_MAP) == list(ref_api.JOINT_WEIGHT) == list(ref_api.JOINT_VEL_WEIGHT)
    if os.path.isdir("/root/reference"):
        import pandas as pd
        ref = pd.read_csv("/root/reference/Joints CSV With Hand/motion09_03JointPosRad.csv")
        assert list(ref.columns) == ref_api.JOINT_COLS
        np.testing.assert_allclose(v.joints_df.values, ref.values, atol=2e-7)
        ep = pd.read_csv("/root/reference/Joints CSV With Hand/motion09_03JointVecFromHip.csv")
        assert list(ep.columns) == ref_api.EP_COLS
