"""Multi-GPU test (needs >= 2 CUDA devices; skipped on a single-GPU box): two ranks, one env shard each, no data-path
collective; the statistics vectors are all-reduced over NCCL and must equal the single-process total of the same
shards (envs are independent, so sharding cannot change any env's trajectory)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
    pytest.skip("needs two CUDA devices", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOTAL, STEPS = 3000, 40


def _rollout(first, count, device, seed_base, TOTAL=TOTAL, STEPS=STEPS):
    import ilrl_b200
    from ilrl_b200 import BatchedHumanoidEnv, stats
    clips = ilrl_b200.CLIP_NAMES
    env = BatchedHumanoidEnv(count, "low", clips=clips, clip_of_env=stats.clip_of_env(first, count, len(clips)),
                             device=device, seed=seed_base, auto_reset=True, env_id_base=first)
    env.reset(start_frame=(np.arange(first, first + count) * 7) % 60, target_deg=(np.arange(first, first + count) * 37) % 360 - 180)
    g = torch.Generator(device="cuda:%d" % device)
    for t in range(STEPS):
        g.manual_seed(1000 + t)
        a = torch.rand(TOTAL, 17, device="cuda:%d" % device, generator=g)[first:first + count] * 2 - 1
        env.step(a)
    st = env.stats()
    phys = env.get_state()[0].cpu()
    env.close()
    return st, phys


def _worker(rank, world, port, q, TOTAL=TOTAL, STEPS=STEPS):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import ilrl_b200  # noqa: F401
    from ilrl_b200 import stats
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        first, count = stats.shard_envs(TOTAL, world, rank)
        st, phys = _rollout(first, count, rank, 5, TOTAL, STEPS)
        total = stats.allreduce_stats(st.clone())
        q.put((rank, st.cpu().numpy(), total.cpu().numpy(), phys.numpy()))
    finally:
        dist.destroy_process_group()


def test_eight_rank_cfg4_sharding_matches_single_process():
    """BASELINE cfg 4 on a full box: 65536 multi-clip envs (clip = global env id mod 4) sharded 8192 per GPU over 8 NCCL
    ranks; every shard bit-identical to the single-GPU batch, all-reduced statistics equal to its totals."""
    if torch.cuda.device_count() < 8:
        pytest.skip("needs eight CUDA devices")
    _run_world(8, 65536, 12)


def test_two_rank_sharding_matches_single_process():
    _run_world(2, TOTAL, STEPS)


def _run_world(world, TOTAL, STEPS):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, TOTAL, STEPS)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    whole_st, whole_phys = _rollout(0, TOTAL, 0, 5, TOTAL, STEPS)
    whole_st = whole_st.cpu().numpy()
    # every rank holds the same all-reduced total, equal to the sum of the per-rank vectors
    for r in res[1:]:
        np.testing.assert_allclose(res[0][2], r[2], rtol=0, atol=0)
    np.testing.assert_allclose(res[0][2], sum(r[1] for r in res), rtol=1e-6)
    # counts are exact; sums agree to fp32 accumulation order; and every env's state is bit-identical to what the
    # single process computed for it (same seed, env_id_base = first global id of the shard)
    assert res[0][2][3] == TOTAL * STEPS and res[0][2][0] == whole_st[0]
    np.testing.assert_allclose(res[0][2], whole_st, rtol=2e-4, atol=1e-2)
    np.testing.assert_array_equal(np.concatenate([r[3] for r in res]), whole_phys.numpy())
