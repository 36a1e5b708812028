"""numpy restatement of the env logic of REF hier_env_2.py — TEST INFRASTRUCTURE ONLY.

The CPU oracle of MODE 2 (the CUDA path is csrc/ilrl_env.cuh `update_reward2` / `update_reward_high2` / ...).  It is
pinned by tests/golden/hier2_injected.npz, recorded from the UNMODIFIED reference under oracle/ref_shim.make_hier2_env
(tests/test_oracle_golden.py).  State layout: phys[47], env words e[26+] in ILRL_E_* order with the two MODE 2
aliases (e[15] = cumulative_deltaVelJoints_low, e[24] = cumulative_deltaJoints_low), jointTarget jt[34].
"""
import numpy as np

from . import oracle as O

M = O.load_model()
MAP_JOINT, MAP_COL, MAP_W = M["map_joint"][:8], M["map_col"][:8], M["map_w"][:8]   # the 8 leg joints (:90-114)
STEP_PER_LEVEL, SKIP_FRAME, TARGET_LEN = 20, 5, 5.0


def inc_frame(e, max_frame, robot_pos, inc=SKIP_FRAME):                      # REF hier_env_2.py:236-247
    old = int(e[0])
    e[0] = (old + inc) % (max_frame - 1)
    if e[0] <= old:
        e[7:10] = [robot_pos[0], robot_pos[1], 0.0]


def low_obs(cur, jt):                                                        # :354-372
    return np.hstack((cur[1:3], cur[6:8], cur[8:42], jt))


def high_obs(cur, e, yaw, clip):                                             # :374-406
    a_t = np.arctan2(e[4] - e[11], e[3] - e[10]) - yaw
    a_s = np.arctan2(e[6] - e[11], e[5] - e[10]) - yaw
    f = int(e[0])
    tail = np.array([[clip["rel"][f][c], clip["vel"][f][c]] for c in MAP_COL]).ravel()
    return np.hstack((cur[:1], [np.cos(a_t), np.sin(a_t)], [np.cos(a_s), np.sin(a_s)], cur[3:42], tail))


def high_step(phys, e, action36, clip):
    """-> (low obs, new jointTarget); e updated in place"""                  # :408-419, :699-741
    cur, xyz, _, _, _ = O.calc_state(phys, e[13], e[14])
    cur = cur.copy(); cur[1:3] = e[19:21]                                    # cur_obs predates this call
    jt = np.array(action36[2:], dtype=np.float64)
    e[10:12] = xyz[:2]
    e[21] = STEP_PER_LEVEL
    inc_frame(e, clip["max_frame"], xyz)
    return low_obs(cur, jt), jt


def low_step_no_physics(phys, e, jt, action, rand_deg, clip, max_timestep=3000):
    """low_level_step with scene.global_step() skipped.  -> dict(low_obs, low_reward, high_obs, high_reward, flags, terms);
    e updated in place."""                                                   # :408-419, :751-769
    stale = O.calc_state(phys, e[13], e[14])[1][:2].copy()                   # robot_pos from the PREVIOUS calc_state
    e[10:12] = stale
    e[21] -= 1
    cur, xyz, js, at_limit, (roll, pitch, yaw) = O.calc_state(phys, e[13], e[14])
    e[19:21] = cur[1:3]
    # updateReward :577-643
    rel, vel = cur[8:42:2].astype(np.float64), cur[9:42:2].astype(np.float64) / 0.1
    sj = np.exp(-np.abs(rel - jt[0::2]).sum() / 17 * 2)
    sv = np.exp(-np.abs(vel - jt[1::2]).sum() / 17 / 2)
    posture = np.exp(-(abs(yaw - e[12]) + abs(roll) + abs(pitch)))
    elec = -float(np.abs(action * js).mean()) - 0.1 * float(np.square(action).mean())
    limit = -0.1 * at_limit
    alive = 2.0 if cur[0] + 0.8 > 0.75 else -1.0
    a, b, p = e[5:7], e[3:5], e[10:12]                                       # calcDriftScore (math_util.py:20-27)
    t = np.clip(np.dot(p - a, b - a) / np.dot(b - a, b - a), 0, 1)
    e[22] += np.exp(-6 * np.linalg.norm(a + t * (b - a) - p))
    e[24] += sj; e[15] += sv; e[18] = posture
    reward = (0.1 * elec + 0.2 * limit + 0.1 * alive + 0.4 * posture) / 2
    # checkTarget :545-570
    if np.linalg.norm(e[10:12] - e[3:5]) <= 0.5:
        rr = yaw + np.deg2rad(rand_deg)
        new = e[10:12] + TARGET_LEN * np.array([np.cos(rr), np.sin(rr)])
        e[5:7] = e[3:5]; e[3:5] = new
        e[7:10] = [e[10], e[11], 0.0]
        e[23] = -np.linalg.norm(e[3:5] - e[5:7])
    e[12] = np.arctan2(e[4] - e[11], e[3] - e[10])
    e[13:15] = e[3:5]
    done = not (alive > 0 and np.linalg.norm(e[3:5] - e[10:12]) <= np.linalg.norm(e[3:5] - e[5:7]) + 1)
    e[2] += 1
    done = done or e[2] >= max_timestep
    out = dict(low_obs=low_obs(cur, jt), low_reward=reward, high_obs=None, high_reward=0.0,
               terms=dict(deltaJoints_low=sj, deltaVelJoints_low=sv, electricityScore=elec, jointLimitScore=limit,
                          aliveReward=alive, bodyPostureScore=posture))
    if done or e[21] <= 0:                                                   # updateRewardHigh :645-697
        d = STEP_PER_LEVEL - e[21]
        hs = -np.linalg.norm(e[3:5] - e[10:12])
        dhigh = (hs - e[23]) / 0.0165 / d
        e[23] = hs
        drift, djl, dvl = e[22] / d, e[24] / d, e[15] / d
        e[22] = e[24] = e[15] = 0.0
        f = int(e[0])
        q, qd = phys[13:30], phys[30:47]
        dj = sum(abs(q[j] - clip["pos"][f][c]) * w for j, c, w in zip(MAP_JOINT, MAP_COL, MAP_W))
        dv = sum(abs(qd[j] - clip["vel"][f][c]) for j, c in zip(MAP_JOINT, MAP_COL))
        e[16], e[17] = np.exp(4 * -dj / 16), np.exp(-dv / 8 / 2)
        posture = np.exp(-(abs(yaw - e[12]) + abs(roll) + abs(pitch)))
        e[18] = posture
        out["high_reward"] = (0.3 * dhigh + 0.2 * drift + e[16] + 0.3 * e[17] + 0.2 * posture + 0.6 * djl + 0.4 * dvl) / 3
        out["high_obs"] = high_obs(cur, e, yaw, clip)
        out["terms"].update(driftScore=drift, delta_highTargetScore=dhigh)
    out["flags"] = int(done) | (int(out["high_obs"] is not None) << 1) | (int(done or out["high_obs"] is None) << 2)
    return out
