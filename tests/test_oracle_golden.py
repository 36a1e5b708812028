"""Pins oracle/ilrl_oracle.c (the CPU restatement) against the golden vectors recorded from the UNMODIFIED reference
Python (tests/golden/*.npz, made by oracle/gen_golden.py) and against the reference's notebook outputs.
CPU only.  Tolerances: both sides are float64 except `cur_obs` (float32 in both), so 1e-9 relative is demanded on
rewards/state and exact equality on frame indices, done flags and the float32 observation head."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")
RT, AT = 1e-9, 1e-9
ENV_LOW_WORDS = list(range(0, 21))
ENV_HIER_WORDS = list(range(0, 25))


def _load(path):
    with np.load(path) as z:  # NpzFile re-reads an array on every [] access: materialise once
        return {k: z[k] for k in z.files}


def _close(a, b, rt=RT, at=AT, what=""):
    np.testing.assert_allclose(a, b, rtol=rt, atol=at, err_msg=what)


def test_low_injected_states_match_reference():
    z = _load(os.path.join(G, "low_injected.npz"))
    envs = [O.OracleEnv(c, 0) for c in O.CLIPS]
    for i in range(len(z["clip"])):
        v = envs[z["clip"][i]]
        v.set(z["phys"][i], z["env_before"][i])
        deg = int(z["rand_deg"][i])
        obs, rew, done = v.low_step(z["action"][i], 0 if deg == -999 else deg, skip_physics=True)
        _, e, t = v.get()
        assert int(e[O.E_FRAME]) == int(z["env_after"][i][O.E_FRAME]), i       # frame index: bit-exact
        assert done == bool(z["done"][i]), i
        np.testing.assert_array_equal(obs[:42].astype(np.float32), z["obs"][i][:42].astype(np.float32))
        _close(obs[42:], z["obs"][i][42:], what="obs tail %d" % i)
        _close(rew, z["reward"][i], what="reward %d" % i)
        _close(e[ENV_LOW_WORDS], z["env_after"][i][ENV_LOW_WORDS], what="env words %d" % i)
        _close(t[:9], z["terms"][i][:9], what="terms %d" % i)


def test_endpoint_score_matches_reference():
    z = _load(os.path.join(G, "low_injected.npz"))
    envs = [O.OracleEnv(c, 0) for c in O.CLIPS]
    for i in range(0, len(z["clip"]), 7):
        v = envs[z["clip"][i]]
        v.set(z["phys"][i], z["env_before"][i])
        _close(v.endpoint_score(), z["endpoint_score"][i], rt=1e-9)


def test_low_trajectory_cfg1_replays():
    """BASELINE cfg 1: every step of the 1000-step random-action run, restarted from the recorded pre-step state."""
    z = _load(os.path.join(G, "low_traj_motion09_03.npz"))
    v = O.OracleEnv("motion09_03", 0)
    for i in range(len(z["reward"])):
        if z["reset_before"][i]:
            robs = v.reset(int(z["reset_start_frame"][i]), 0.0, int(z["reset_target_deg"][i]))
            _close(robs, z["reset_obs"][i], rt=1e-6, at=1e-7, what="reset obs %d" % i)
            p, e, _ = v.get()
            _close(p, z["phys_before"][i], what="reset phys %d" % i)
            _close(e[ENV_LOW_WORDS], z["env_before"][i][ENV_LOW_WORDS], what="reset env %d" % i)
        v.set(z["phys_before"][i], z["env_before"][i])
        deg = int(z["rand_deg"][i])
        obs, rew, done = v.low_step(z["action"][i], 0 if deg == -999 else deg)
        p, e, t = v.get()
        _close(p, z["phys_after"][i], what="phys %d" % i)
        assert int(e[O.E_FRAME]) == int(z["env_after"][i][O.E_FRAME])
        assert done == bool(z["done"][i])
        _close(obs, z["obs"][i], rt=1e-6, at=1e-7)
        _close(rew, z["reward"][i])
        _close(e[ENV_LOW_WORDS], z["env_after"][i][ENV_LOW_WORDS])


def _hier_check(v, z, i, skip_physics):
    kind = int(z["kind"][i])
    if kind == 1:
        lo = v.high_step(z["action"][i][:2])
        _close(lo, z["low_obs"][i], rt=1e-6, at=1e-7, what="high->low obs %d" % i)
    else:
        deg = int(z["rand_deg"][i])
        lo, lr, ho, hr, done, has_high = v.hier_low_step(z["action"][i], 0 if deg == -999 else deg, skip_physics)
        flags = int(z["flags"][i])
        assert done == bool(flags & 1) and has_high == bool(flags & 2), i
        if flags & 4:
            _close(lo, z["low_obs"][i], rt=1e-6, at=1e-7, what="low obs %d" % i)
            _close(lr, z["low_reward"][i], what="low reward %d" % i)
        if flags & 2:
            _close(ho, z["high_obs"][i], rt=1e-6, at=1e-7, what="high obs %d" % i)
            _close(hr, z["high_reward"][i], what="high reward %d" % i)
    _, e, t = v.get()
    assert int(e[O.E_FRAME]) == int(z["env_after"][i][O.E_FRAME])
    _close(e[ENV_HIER_WORDS], z["env_after"][i][ENV_HIER_WORDS], what="env words %d" % i)


def test_hier_injected_states_match_reference():
    z = _load(os.path.join(G, "hier_injected.npz"))
    v = O.OracleEnv("motion09_03", 1)
    for i in range(len(z["kind"])):
        e = z["env_before"][i].copy()
        e[O.E_OBS_SIN:O.E_OBS_COS + 1] = z["obs_sincos"][i]
        v.set(z["phys"][i], e)
        _hier_check(v, z, i, True)


def test_hier_trajectory_replays():
    z = _load(os.path.join(G, "hier_traj.npz"))
    v = O.OracleEnv("motion09_03", 1)
    for i in range(len(z["kind"])):
        v.set(z["phys_before"][i], z["env_before"][i])
        if int(z["kind"][i]) == 0:
            sf, yaw, deg = [int(x) for x in z["draws"][i]]
            ho = v.reset(sf, yaw, deg)
            _close(ho, z["high_obs"][i], rt=1e-6, at=1e-7, what="reset high obs %d" % i)
            p, e, _ = v.get()
            _close(p, z["phys_after"][i])
            _close(e[ENV_HIER_WORDS], z["env_after"][i][ENV_HIER_WORDS])
        else:
            _hier_check(v, z, i, False)
            p, _, _ = v.get()
            _close(p, z["phys_after"][i])


def test_reset_vectors():
    z = _load(os.path.join(G, "reset_vectors.npz"))
    for i in range(len(z["mode"])):
        mode = int(z["mode"][i])
        v = O.OracleEnv(O.CLIPS[int(z["clip"][i])], mode)
        p, e, _ = v.get()
        e[O.E_SEP_X:O.E_SEP_Z + 1] = z["sep_before"][i]
        v.set(p, e)
        obs = v.reset(int(z["start_frame"][i]), float(z["yaw"][i]), int(z["target_deg"][i]))
        n = 44 if mode else 70
        _close(obs, z["obs"][i][:n], rt=1e-6, at=1e-7)
        p, e, _ = v.get()
        _close(p, z["phys_after"][i])
        words = ENV_HIER_WORDS if mode else ENV_LOW_WORDS
        _close(e[words], z["env_after"][i][words])


def test_notebook_vectors():
    """recorded cell outputs of the reference's own notebook."""
    nb = json.load(open(os.path.join(G, "notebook_vectors.json")))
    obs78 = np.array(nb["reset_obs_frame0_78"])
    c = O.load_clip("motion09_03")
    m = O.load_model()
    # item 1: obs tail = interleave(rel[2], vel[2]) in joint_map order -> reset advances the frame by skipFrame=2
    tail = np.empty(28)
    tail[0::2] = c["rel"][2][m["map_col"]]
    tail[1::2] = c["vel"][2][m["map_col"]]
    np.testing.assert_allclose(obs78[-28:], tail, atol=5e-8)
    v = O.OracleEnv("motion09_03", 0)
    obs = v.reset(0, 0.0, 0)
    np.testing.assert_allclose(obs[42:], obs78[-28:], atol=5e-8)
    # item 2: joint slots in ordered_joints order (that notebook run still had 4 ankle joints: drop them)
    names21 = nb["jdict_order"]
    keep = [k for k, n in enumerate(names21) if "ankle" not in n]
    assert [names21[k] for k in keep] == m["joint_name"]
    relpos = obs78[8:8 + 42:2][keep]
    np.testing.assert_allclose(obs[8:42:2], relpos, atol=2e-6)
    # z - initial_z = 1.17 - 0.8
    assert abs(obs78[0] - 0.37) < 1e-7 and abs(obs[0] - 0.37) < 1e-6
    # item 3: CSV relative table == 2(q-mid)/(hi-lo) with the MJCF ranges
    lo, hi = np.array(m["joint_lo"]), np.array(m["joint_hi"])
    for name in O.CLIPS:
        cl = O.load_clip(name)
        for k in range(14):
            j = m["map_joint"][k]
            col = m["map_col"][k]
            rel = 2 * (cl["pos"][:, col] - 0.5 * (lo[j] + hi[j])) / (hi[j] - lo[j])
            np.testing.assert_allclose(rel, cl["rel"][:, col], atol=1e-6)
    # item 4: reset pose (yaw -45 deg): [0,0,1.17, 0,0,-0.38268343,0.92387953]
    pose = np.array(nb["reset_pose"])
    v.reset(0, 0.0, -45)
    p, _, _ = v.get()
    np.testing.assert_allclose(p[:7], pose, atol=1e-8)
    # item 5: 'floor' is one of robot.parts -> body_xyz is a 33-way mean
    assert nb["parts_keys_has_floor"] is True
    _, xyz, _, _, _ = O.calc_state(p, 0.0, 0.0)
    bo, ao, _ = O.fk(p)
    np.testing.assert_allclose(xyz[:2], (bo[:, :2].sum(0) + ao[:, :2].sum(0)) / 33.0, atol=1e-12)
    # ... and the part positions behind that mean against the SECOND, independent forward kinematics (numpy / scipy,
    # oracle/ref_shim.py: the one the unmodified reference Python runs on when the fixtures are recorded), on random poses
    from oracle import ref_shim as S
    from oracle.gen_golden import random_phys
    rng = np.random.default_rng(12)
    for _ in range(25):
        ph = random_phys(rng, upright=False)
        pos, _, anc = S.py_fk(ph)
        _, xyz, _, _, _ = O.calc_state(ph, 0.0, 0.0)
        bo, ao, _ = O.fk(ph)
        np.testing.assert_allclose(bo, pos, atol=1e-9)
        np.testing.assert_allclose(ao, anc, atol=1e-9)
        np.testing.assert_allclose(xyz[:2], (pos[:, :2].sum(0) + anc[:, :2].sum(0)) / 33.0, atol=1e-9)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference checkout exists only in the build container")
def test_golden_fixtures_regenerate_bit_identically_from_the_reference(tmp_path):
    """The committed fixtures ARE what the unmodified reference Python produces: oracle/gen_golden.py (which imports
    /root/reference through oracle/ref_shim.py) is re-run into a scratch directory and every array compared bit for bit.
    Runs only where the reference is mounted (never on the GPU box)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, ILRL_GOLDEN_OUT=str(tmp_path))
    subprocess.check_call([sys.executable, os.path.join(root, "oracle", "gen_golden.py")], env=env,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
    gold = os.path.join(root, "tests", "golden")
    names = sorted(f for f in os.listdir(gold))
    assert sorted(os.listdir(str(tmp_path))) == names
    for f in names:
        if f.endswith(".npz"):
            with np.load(os.path.join(gold, f)) as a, np.load(os.path.join(str(tmp_path), f)) as b:
                assert sorted(a.files) == sorted(b.files), f
                for k in a.files:
                    assert np.array_equal(a[k], b[k]), (f, k)
        else:
            assert open(os.path.join(gold, f)).read() == open(os.path.join(str(tmp_path), f)).read(), f


def test_hier2_numpy_oracle_matches_reference_vectors():
    """Row a18: oracle/hier2_np.py (the CPU restatement of hier_env_2's env logic) against the vectors recorded from the
    UNMODIFIED REF hier_env_2.py: high_level_step and low_level_step on 400 injected states each."""
    from oracle import hier2_np as H
    z = dict(np.load(os.path.join(G, "hier2_injected.npz")))
    clip = O.load_clip("motion09_03")
    kinds = z["kind"]
    n_high = n_low = 0
    for i in np.nonzero(kinds == 1)[0]:
        e = z["env_before"][i].copy(); e[19:21] = z["obs_sincos"][i]
        lo, jt = H.high_step(z["phys"][i], e, z["action"][i], clip)
        np.testing.assert_allclose(lo, z["low_obs"][i], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(jt, z["jt_after"][i], rtol=0, atol=0)
        assert int(e[0]) == int(z["env_after"][i][0])
        np.testing.assert_allclose(e[2:25], z["env_after"][i][2:25], rtol=1e-6, atol=1e-6)
        n_high += 1
    for i in np.nonzero(kinds == 2)[0]:
        e = z["env_before"][i].copy()
        r = H.low_step_no_physics(z["phys"][i], e, z["jt_before"][i], z["action"][i][:17], int(z["rand_deg"][i]), clip)
        assert r["flags"] == int(z["flags"][i]), i
        assert int(e[0]) == int(z["env_after"][i][0])
        np.testing.assert_allclose(e[2:25], z["env_after"][i][2:25], rtol=1e-6, atol=1e-5)
        if r["flags"] & 4:
            np.testing.assert_allclose(r["low_obs"], z["low_obs"][i], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(r["low_reward"], z["low_reward"][i], rtol=1e-6, atol=1e-6)
        if r["flags"] & 2:
            np.testing.assert_allclose(r["high_obs"], z["high_obs"][i], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(r["high_reward"], z["high_reward"][i], rtol=1e-5, atol=1e-5)
        n_low += 1
    assert n_high == 400 and n_low == 400
    # fixture sanity: all three protocol outcomes occur, and the reset records carry the stock robot's 60-word high obs
    assert set(np.unique(z["flags"][kinds == 2])) == {2, 4, 7}
    assert np.abs(z["high_obs"][kinds == 0][:, 44:]).max() > 0
