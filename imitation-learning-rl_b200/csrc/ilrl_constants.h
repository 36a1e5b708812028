/* Named constants of the humanoid imitation hot path, each with its provenance.
 *
 * [REF f:l]   read in the reference (AdityaPutraS/Imitation-Learning-RL) at that file:line.
 * [BULLET]    restated from memory of the un-vendored third-party code the reference calls (pybullet / Bullet3
 *             btMultiBody*, pybullet_envs): NOT pinned by any reference test or fixture ("parity unpinned",
 *             DESIGN.md section 3).  Every such value lives here so that it can be re-validated in one place the
 *             first time a real PyBullet is importable.
 * Included by oracle/ilrl_oracle.c (double) and the CUDA sources under csrc/ (float); values only, no code. */
#ifndef ILRL_CONSTANTS_H
#define ILRL_CONSTANTS_H

/* ---- time stepping [BULLET pybullet_envs WalkerBaseBulletEnv.create_single_player_scene; REF humanoid.py:157-159] */
#define ILRL_FRAME_DT        0.0165          /* env step; also the reward "delta" divisor REF low_level_env.py:449-455 */
#define ILRL_SUBSTEPS        4
#define ILRL_GRAVITY         9.8
#define ILRL_SOLVER_ITERS    5               /* [BULLET] World.clean_everything numSolverIterations=5 */

/* ---- contact / limit rows [BULLET btMultiBodyConstraintSolver, btMultiBodyJointLimitConstraint] */
#define ILRL_CONTACT_ERP     0.9             /* setDefaultContactERP(0.9) -> m_erp2 */
#define ILRL_LIMIT_ERP       0.2             /* m_erp default; limit rows exist only while q is beyond the limit */
#define ILRL_CONTACT_BREAK   0.02            /* gContactBreakingThreshold: rows exist for distance < this */
#define ILRL_FRICTION        1.6             /* plane 0.8 (StadiumScene) x geom 2.0 (REF humanoid_symmetric_2.xml:5) */
#define ILRL_MAX_CONTACTS    8               /* kept contacts per env: the deepest ones, rows in sphere-table order */
#define ILRL_MAX_ROWS        (17 + 3 * ILRL_MAX_CONTACTS)
#define ILRL_MAX_COORD_VEL   100.0           /* btMultiBody m_maxCoordinateVelocity */

/* ---- velocity damping inside the forward dynamics [BULLET btMultiBody.cpp DAMPING_K*] */
#define ILRL_DAMP_K1_LIN     0.02
#define ILRL_DAMP_K2_LIN     0.04
#define ILRL_DAMP_K1_ANG     0.02
#define ILRL_DAMP_K2_ANG     0.04

/* ---- reference env constants */
#define ILRL_INITIAL_Z       0.8             /* REF humanoid.py:49 */
#define ILRL_RESET_Z         1.17            /* REF low_level_env.py:264, hier_env.py:266 */
#define ILRL_RESET_Z_HIER2   1.15            /* REF hier_env_2.py:295 */
#define ILRL_TARGET_LEN      5.0             /* REF low_level_env.py:155 */
#define ILRL_ALIVE_Z         0.75            /* REF low_level_env.py:387 */
#define ILRL_TARGET_REACHED  0.5             /* REF low_level_env.py:415 */
#define ILRL_DONE_MARGIN_LOW 3.0             /* REF low_level_env.py:469 */
#define ILRL_DONE_MARGIN_HI  1.0             /* REF hier_env.py:576 */
#define ILRL_JOINT_W_SUM     17.400000000000002  /* python sum() of REF low_level_env.py:103-119 */
#define ILRL_JOINT_WV_SUM    8.6             /* REF low_level_env.py:121-138 (8 + 6 x 0.1, python float sum) */
#define ILRL_EP_W_SUM        8.0             /* REF low_level_env.py:146-152 */
/* reward weights REF low_level_env.py:507, hier_env.py:607 */
#define ILRL_RW_JOINT 0.34
#define ILRL_RW_JVEL  0.1
#define ILRL_RW_TARGET 0.34
#define ILRL_RW_ELEC  0.034
#define ILRL_RW_LIMIT 0.15
#define ILRL_RW_ALIVE 0.034
#define ILRL_RW_POSTURE 0.1
/* hier_env_2 (MODE 2): low-level reward REF hier_env_2.py:625-641 = sum(r * w) / 2 over [delta_deltaJoints (never
 * written: 0), delta_deltaVelJoints (0), electricity, joint limit, alive, posture]; high-level reward :683-697 =
 * sum / 3 over [delta_highTargetScore, driftScore, deltaJoints, deltaVelJoints, bodyPostureScore, mean deltaJoints_low,
 * mean deltaVelJoints_low] */
#define ILRL_RW2_ELEC 0.1
#define ILRL_RW2_LIMIT 0.2
#define ILRL_RW2_ALIVE 0.1
#define ILRL_RW2_POSTURE 0.4
#define ILRL_RW2_HIGH {0.3, 0.2, 1.0, 0.3, 0.2, 0.6, 0.4}
#define ILRL_JOINT_W_SUM2    16.0            /* REF hier_env_2.py:104-114: the 8 leg joints, weights 3,1,3,1,3,1,3,1 */
#define ILRL_JOINT_WV_SUM2   8.0             /* REF hier_env_2.py:116-126 */

/* ---- flat state layouts shared by the C-ABI get/set_state calls, the oracle and the tests */
#define ILRL_PHYS_WORDS 47   /* base pos3, quat4 (x,y,z,w), lin vel3, ang vel3 (world), q[17], qd[17] */
/* Resident state in HBM: one row per env (env-major), rows 16-byte aligned so that the four lanes of an env move it with
 * 128-bit accesses and a row can be reached through the cost-grouping permutation without losing coalescing. */
#define ILRL_PHYS_STRIDE 48
#define ILRL_ENV_STRIDE 28
#define ILRL_JT_STRIDE 36
enum {
  ILRL_E_FRAME = 0, ILRL_E_CLIP, ILRL_E_T, ILRL_E_TARGET_X, ILRL_E_TARGET_Y, ILRL_E_START_X, ILRL_E_START_Y,
  ILRL_E_SEP_X, ILRL_E_SEP_Y, ILRL_E_SEP_Z, ILRL_E_ROBOT_X, ILRL_E_ROBOT_Y, ILRL_E_HLDEG, ILRL_E_WALK_X,
  ILRL_E_WALK_Y, ILRL_E_LOW_TARGET_SCORE, ILRL_E_JOINT_SCORE, ILRL_E_JVEL_SCORE, ILRL_E_POSTURE_SCORE,
  ILRL_E_OBS_SIN, ILRL_E_OBS_COS,                       /* cur_obs[1:3] of the last calc_state (hier low obs reuse) */
  ILRL_E_STEPS_REMAINING, ILRL_E_CUM_DRIFT, ILRL_E_HIGH_TARGET_SCORE, ILRL_E_CUM_ALIVE, ILRL_E_HIGH_PENDING,
  ILRL_E_EP_RETURN, ILRL_E_EP_LEN,                      /* running episode return / length (statistics only) */
  ILRL_ENV_WORDS /* = 28 */
};
/* MODE 2 (hier_env_2.py) keeps two accumulators the other modes do not have in words it does not use otherwise:
 * lowTargetScore is the constant -targetLen there (initReward, never updated) and there is no cumulative_aliveReward. */
#define ILRL_E2_CUM_DV_LOW ILRL_E_LOW_TARGET_SCORE   /* cumulative_deltaVelJoints_low REF hier_env_2.py:622 */
#define ILRL_E2_CUM_DJ_LOW ILRL_E_CUM_ALIVE          /* cumulative_deltaJoints_low    REF hier_env_2.py:621 */
/* per-step "terms" row (the attributes RewardLogCallback reads, REF custom_callback.py:43-80) */
enum {
  ILRL_T_JOINT = 0, ILRL_T_JVEL, ILRL_T_DLOWTARGET, ILRL_T_ELEC, ILRL_T_LIMIT, ILRL_T_ALIVE, ILRL_T_POSTURE,
  ILRL_T_LOWTARGET, ILRL_T_ENDPOINT /* deltaEndPoints: stays 0 in the reference step path */, ILRL_T_HIGHTARGET,
  ILRL_T_DRIFT, ILRL_T_DHIGHTARGET /* delta_highTargetScore (baseReward is always 0: host side) */,
  ILRL_TERM_WORDS /* = 12 */
};
/* MODE 2: delta_lowTargetScore and deltaEndPoints are constant 0 in hier_env_2.py; their slots carry the two low-level
 * tracking scores of calcJointPosVelLowScore (REF hier_env_2.py:461-474) */
#define ILRL_T2_DJ_LOW ILRL_T_DLOWTARGET
#define ILRL_T2_DV_LOW ILRL_T_ENDPOINT
#define ILRL_OBS_LOW  70
#define ILRL_OBS_HIGH 44
#define ILRL_ACT_LOW  17
#define ILRL_ACT_HIGH 2
#define ILRL_OBS_LOW2  72   /* hier_env_2: cur_obs[1:3], [6:8], [8:42] + jointTarget[34]  REF hier_env_2.py:369-372 */
#define ILRL_OBS_HIGH2 60   /* 44 + (rel, vel) of the 8 leg joints at the current frame   REF hier_env_2.py:374-406 */
#define ILRL_ACT_HIGH2 36   /* 2 unused + jointTarget[34]                                 REF hier_env_2.py:54-56, 731 */
#define ILRL_JT_WORDS  34
#define ILRL_STATS_WORDS 16
#endif
