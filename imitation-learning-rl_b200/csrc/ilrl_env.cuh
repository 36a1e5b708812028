// ilrl_env.cuh — per-env imitation-env logic fused around the physics substeps (device code, fp32).
//
// Restates, for one env held by one thread, the reference's own Python:
//   calc_state                      pybullet_envs WalkerBase.calc_state (un-vendored), REF humanoid.py:49 initial_z
//   reward terms / updateReward     REF low_level_env.py:325-410, 441-465   hier_env.py:368-467, 494-536
//   incFrame / checkTarget / done   REF low_level_env.py:218-222, 412-434, 467-473   hier_env.py:227-233, 469-487, 573-581
//   observations                    REF low_level_env.py:307-320   hier_env.py:321-353
//   resetFromFrame                  REF low_level_env.py:247-305   hier_env.py:255-319
//   high_level_step                 REF hier_env.py:538-571
// MODE 2 = the hier_env_2.py variant: joint-target tracking low level
//   incFrame :236-247, resetFromFrame :277-352, getLowLevelObs :354-372, getHighLevelObs :374-406,
//   calcJointScore / calcJointVelScore :421-459, calcJointPosVelLowScore :461-474, checkTarget :545-570,
//   updateReward :577-643, updateRewardHigh :645-697, high_level_step :699-741, low_level_step :751-769
// Reference quirks that are mirrored on purpose are tagged (Qn) as in SURVEY.md appendix B.
#pragma once
#include "ilrl_physics.cuh"

namespace ilrl {

__device__ constexpr int kMotorJoint[NJ] = ILRL_MOTOR_JOINT;
__device__ constexpr float kMotorGear[NJ] = ILRL_MOTOR_GEAR;
__device__ constexpr int kMapJoint[NMAP] = ILRL_MAP_JOINT;
__device__ constexpr int kMapCol[NMAP] = ILRL_MAP_COL;
__device__ constexpr float kMapW[NMAP] = ILRL_MAP_W;
__device__ constexpr float kMapWv[NMAP] = ILRL_MAP_WV;

constexpr int MAX_CLIPS = 8;
struct ClipDesc {
  const float* pos;  // [n,14] JointPosRad
  const float* rel;  // [n,14] JointPosRadRelative
  const float* vel;  // [n,14] JointSpeedRadSec
  const float* ep;   // [n,27] JointVecFromHip
  int n_pos, n_vel, max_frame, pad;
};

// ---- Philox4x32-10 (counter = env id / draw index, key = seed): the env's private random integers
__device__ __forceinline__ uint32_t philox_u32(uint64_t seed, uint32_t env, uint32_t draw) {
  uint32_t c0 = env, c1 = draw, c2 = 0x1f83d9abu, c3 = 0x5be0cd19u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c0;
}
// integer in [lo, hi)  (rng.integers(lo, hi))
__device__ __forceinline__ int rand_int(uint64_t seed, uint32_t env, uint32_t& ctr, int lo, int hi) {
  uint32_t r = philox_u32(seed, env, ctr++);
  return lo + (int)(((uint64_t)r * (uint32_t)(hi - lo)) >> 32);
}

struct Calc {
  float obs[42];
  float bx, by;  // body_xyz[0:2]: mean over the 33 `parts` (Q2)
  float roll, pitch, yaw;
  float js[NJ];  // joint_speeds
  int at_limit;
};

__device__ __forceinline__ float clip5(float v) { return fminf(fmaxf(v, -5.f), 5.f); }

// pybullet.getEulerFromQuaternion
__device__ __forceinline__ void quat_rpy(const float* q, float& roll, float& pitch, float& yaw) {
  float x = q[0], y = q[1], z = q[2], w = q[3];
  float sarg = -2.f * (x * z - w * y);
  if (sarg <= -0.99999f) { roll = 0.f; pitch = -1.5707963267948966f; yaw = 2.f * atan2f(x, -y); }
  else if (sarg >= 0.99999f) { roll = 0.f; pitch = 1.5707963267948966f; yaw = 2.f * atan2f(-x, y); }
  else {
    roll = atan2f(2.f * (y * z + w * x), w * w - x * x - y * y + z * z);
    pitch = asinf(sarg);
    yaw = atan2f(2.f * (x * y + w * z), w * w + x * x - y * y - z * z);
  }
}

// WalkerBase.calc_state.  sumx/sumy: sums of the 32 part origins relative to the torso (from fk()).
__device__ __forceinline__ void calc_state(const Phys& s, float sumx, float sumy, float wtx, float wty, Calc& c) {
  c.bx = (32.f * s.p[0] + sumx) * (1.f / 33.f);  // torso + 31 offsets + the floor at the origin
  c.by = (32.f * s.p[1] + sumy) * (1.f / 33.f);
  quat_rpy(s.quat, c.roll, c.pitch, c.yaw);
  float ang = atan2f(wty - c.by, wtx - c.bx) - c.yaw;
  float sy, cy;
  sincosf(-c.yaw, &sy, &cy);
  float sa, ca;
  sincosf(ang, &sa, &ca);
  c.obs[0] = clip5(s.p[2] - (float)ILRL_INITIAL_Z);
  c.obs[1] = sa; c.obs[2] = ca;
  c.obs[3] = clip5(0.3f * (cy * s.v[0] - sy * s.v[1]));
  c.obs[4] = clip5(0.3f * (sy * s.v[0] + cy * s.v[1]));
  c.obs[5] = clip5(0.3f * s.v[2]);
  c.obs[6] = clip5(c.roll); c.obs[7] = clip5(c.pitch);
  c.at_limit = 0;
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    const float mid = 0.5f * (kJointLo[j] + kJointHi[j]), isp = 2.f / (kJointHi[j] - kJointLo[j]);
    float rp = (s.q[j] - mid) * isp, rv = 0.1f * s.qd[j];
    c.js[j] = rv;
    c.at_limit += fabsf(rp) > 0.99f ? 1 : 0;
    c.obs[8 + 2 * j] = clip5(rp);
    c.obs[9 + 2 * j] = clip5(rv);
  }
}

__device__ __forceinline__ float hyp(float x, float y) { return sqrtf(x * x + y * y); }

// the per-env bookkeeping words, ILRL_E_* order
struct EnvW { float e[ILRL_ENV_WORDS]; };

template <int MODE = 0>
__device__ __forceinline__ void inc_frame(EnvW& w, const ClipDesc& c, int inc) {
  const int old = (int)w.e[ILRL_E_FRAME];
  int f = (old + inc) % (c.max_frame - 1);
  w.e[ILRL_E_FRAME] = (float)f;
  // (Q7) modes 0 / 1 re-anchor only on an exact wrap to 0; hier_env_2 whenever the frame did not grow (:246)
  if (MODE == 2 ? f <= old : f == 0) {
    w.e[ILRL_E_SEP_X] = w.e[ILRL_E_ROBOT_X]; w.e[ILRL_E_SEP_Y] = w.e[ILRL_E_ROBOT_Y]; w.e[ILRL_E_SEP_Z] = 0.f;
  }
}

// low obs row: cur_obs[42] + interleave(rel[frame], vel[frame]) in joint_map order (Q6, Q10)
__device__ __forceinline__ void write_low_obs(const float* cur42, const EnvW& w, const ClipDesc& c, float* out) {
#pragma unroll
  for (int i = 0; i < 42; i++) out[i] = cur42[i];
  const int f = (int)w.e[ILRL_E_FRAME];
  const float* rel = c.rel + f * 14;
  const float* vel = c.vel + f * 14;
#pragma unroll
  for (int m = 0; m < NMAP; m++) {
    out[42 + 2 * m] = __ldg(rel + kMapCol[m]);
    out[43 + 2 * m] = __ldg(vel + kMapCol[m]);
  }
}

// high obs row (Q15: cos first)
__device__ __forceinline__ void write_high_obs(const Calc& c, const EnvW& w, float* out) {
  float tt = atan2f(w.e[ILRL_E_TARGET_Y] - w.e[ILRL_E_ROBOT_Y], w.e[ILRL_E_TARGET_X] - w.e[ILRL_E_ROBOT_X]) - c.yaw;
  float ts = atan2f(w.e[ILRL_E_START_Y] - w.e[ILRL_E_ROBOT_Y], w.e[ILRL_E_START_X] - w.e[ILRL_E_ROBOT_X]) - c.yaw;
  float s1, c1, s2, c2;
  sincosf(tt, &s1, &c1);
  sincosf(ts, &s2, &c2);
  out[0] = c.obs[0]; out[1] = c1; out[2] = s1; out[3] = c2; out[4] = s2;
#pragma unroll
  for (int i = 3; i < 42; i++) out[2 + i] = c.obs[i];
}

// hier_env_2 low obs (72): cur_obs[1:3] | cur_obs[6:8] | cur_obs[8:42] | jointTarget[34]
__device__ __forceinline__ void write_low_obs2(const float* cur42, const float* jt, float* out) {
  out[0] = cur42[1]; out[1] = cur42[2]; out[2] = cur42[6]; out[3] = cur42[7];
#pragma unroll
  for (int i = 0; i < 34; i++) { out[4 + i] = cur42[8 + i]; out[38 + i] = jt[i]; }
}
// hier_env_2 high obs (60): the 44 of write_high_obs (cur_obs[3:-2] of the stock 44-entry state = cur_obs[3:42])
// + interleave(rel[frame], vel[frame]) of the 8 leg joints (the first 8 of joint_map order)
__device__ __forceinline__ void write_high_obs2(const Calc& c, const EnvW& w, const ClipDesc& cl, float* out) {
  write_high_obs(c, w, out);
  const int f = (int)w.e[ILRL_E_FRAME];
  const float* rel = cl.rel + f * 14;
  const float* vel = cl.vel + f * 14;
#pragma unroll
  for (int m = 0; m < 8; m++) {
    out[44 + 2 * m] = __ldg(rel + kMapCol[m]);
    out[45 + 2 * m] = __ldg(vel + kMapCol[m]);
  }
}

// hier_env_2 updateReward (:577-643).  jt = jointTarget[34] (pos, vel interleaved in ordered_joints order).
__device__ __forceinline__ float update_reward2(const Calc& c, EnvW& w, const float* jt, const float* action, float* terms) {
  float dj = 0.f, dv = 0.f;
#pragma unroll
  for (int i = 0; i < NJ; i++) {  // calcJointPosVelLowScore on the CLIPPED float32 cur_obs entries (:465-469)
    dj += fabsf(c.obs[8 + 2 * i] - jt[2 * i]);
    dv += fabsf(c.obs[9 + 2 * i] / 0.1f - jt[2 * i + 1]);
  }
  const float sj = expf((-dj / 17.f) * 2.f), sv = expf((-dv / 17.f) * 0.5f);
  const float posture = expf(-(fabsf(c.yaw - w.e[ILRL_E_HLDEG]) + fabsf(c.roll) + fabsf(c.pitch)));
  w.e[ILRL_E_POSTURE_SCORE] = posture;
  float run = 0.f, stall = 0.f;
#pragma unroll
  for (int i = 0; i < NJ; i++) { run += fabsf(action[i] * c.js[i]); stall += action[i] * action[i]; }
  const float elec = -(run * (1.f / NJ)) - 0.1f * (stall * (1.f / NJ));
  const float limit = -0.1f * (float)c.at_limit;
  const float alive = (c.obs[0] + (float)ILRL_INITIAL_Z) > (float)ILRL_ALIVE_Z ? 2.f : -1.f;
  {  // calcDriftScore, as in mode 1
    float lx = w.e[ILRL_E_TARGET_X] - w.e[ILRL_E_START_X], ly = w.e[ILRL_E_TARGET_Y] - w.e[ILRL_E_START_Y];
    float len2 = lx * lx + ly * ly;
    float t = ((w.e[ILRL_E_ROBOT_X] - w.e[ILRL_E_START_X]) * lx + (w.e[ILRL_E_ROBOT_Y] - w.e[ILRL_E_START_Y]) * ly) / len2;
    t = fminf(fmaxf(t, 0.f), 1.f);
    float px = w.e[ILRL_E_START_X] + t * lx, py = w.e[ILRL_E_START_Y] + t * ly;
    w.e[ILRL_E_CUM_DRIFT] += expf(-6.f * hyp(px - w.e[ILRL_E_ROBOT_X], py - w.e[ILRL_E_ROBOT_Y]));
  }
  w.e[ILRL_E2_CUM_DJ_LOW] += sj;
  w.e[ILRL_E2_CUM_DV_LOW] += sv;
  terms[ILRL_T_JOINT] = w.e[ILRL_E_JOINT_SCORE]; terms[ILRL_T_JVEL] = w.e[ILRL_E_JVEL_SCORE];  // attributes of the last high reward
  terms[ILRL_T2_DJ_LOW] = sj; terms[ILRL_T2_DV_LOW] = sv;
  terms[ILRL_T_ELEC] = elec; terms[ILRL_T_LIMIT] = limit; terms[ILRL_T_ALIVE] = alive; terms[ILRL_T_POSTURE] = posture;
  terms[ILRL_T_LOWTARGET] = -(float)ILRL_TARGET_LEN;
  // delta_deltaJoints / delta_deltaVelJoints (weights 1, 0.2) are never written by hier_env_2: they contribute 0
  return ((float)ILRL_RW2_ELEC * elec + (float)ILRL_RW2_LIMIT * limit + (float)ILRL_RW2_ALIVE * alive +
          (float)ILRL_RW2_POSTURE * posture) * 0.5f;
}

// hier_env_2 updateRewardHigh (:645-697): returns the high-level reward.  Uses steps_remaining AFTER its decrement.
__device__ __forceinline__ float update_reward_high2(const Phys& s, const Calc& c, EnvW& w, const ClipDesc& cl,
                                                     float* terms, float step_per_level) {
  const float hs = -hyp(w.e[ILRL_E_TARGET_X] - w.e[ILRL_E_ROBOT_X], w.e[ILRL_E_TARGET_Y] - w.e[ILRL_E_ROBOT_Y]);
  const float d = step_per_level - w.e[ILRL_E_STEPS_REMAINING];  // no +1 here (unlike Q14)
  const float dhigh = (hs - w.e[ILRL_E_HIGH_TARGET_SCORE]) / 0.0165f / d;
  w.e[ILRL_E_HIGH_TARGET_SCORE] = hs;
  const float drift = w.e[ILRL_E_CUM_DRIFT] / d;
  const float djl = w.e[ILRL_E2_CUM_DJ_LOW] / d, dvl = w.e[ILRL_E2_CUM_DV_LOW] / d;
  w.e[ILRL_E_CUM_DRIFT] = 0.f; w.e[ILRL_E2_CUM_DJ_LOW] = 0.f; w.e[ILRL_E2_CUM_DV_LOW] = 0.f;
  const int f = (int)w.e[ILRL_E_FRAME];
  const float* pos = cl.pos + f * 14;
  const float* vel = cl.vel + f * 14;
  float dj = 0.f, dv = 0.f;
#pragma unroll
  for (int m = 0; m < 8; m++) {  // the 8 leg joints: weights 3,1,3,1,... / 16 and 1 / 8
    dj += fabsf(s.q[kMapJoint[m]] - __ldg(pos + kMapCol[m])) * kMapW[m];
    dv += fabsf(s.qd[kMapJoint[m]] - __ldg(vel + kMapCol[m])) * kMapWv[m];
  }
  const float js = expf(4.f * (-dj / (float)ILRL_JOINT_W_SUM2)), jv = expf((-dv / (float)ILRL_JOINT_WV_SUM2) * 0.5f);
  const float posture = expf(-(fabsf(c.yaw - w.e[ILRL_E_HLDEG]) + fabsf(c.roll) + fabsf(c.pitch)));
  w.e[ILRL_E_JOINT_SCORE] = js; w.e[ILRL_E_JVEL_SCORE] = jv; w.e[ILRL_E_POSTURE_SCORE] = posture;
  terms[ILRL_T_JOINT] = js; terms[ILRL_T_JVEL] = jv; terms[ILRL_T_POSTURE] = posture;
  terms[ILRL_T_DHIGHTARGET] = dhigh; terms[ILRL_T_DRIFT] = drift;
  const float wh[7] = ILRL_RW2_HIGH;
  return (dhigh * wh[0] + drift * wh[1] + js * wh[2] + jv * wh[3] + posture * wh[4] + djl * wh[5] + dvl * wh[6]) / 3.f;
}

// updateReward + weighted sum.  `action` is the raw (unclipped) policy output (Q4); slots are cross-paired (Q3).
template <int MODE>
__device__ __forceinline__ float update_reward(const Phys& s, const Calc& c, EnvW& w, const ClipDesc& cl,
                                               const float* action, float* terms) {
  const int f = (int)w.e[ILRL_E_FRAME];
  const float* pos = cl.pos + f * 14;
  const float* vel = cl.vel + f * 14;
  float dj = 0.f, dv = 0.f;
#pragma unroll
  for (int m = 0; m < NMAP; m++) {
    dj += fabsf(s.q[kMapJoint[m]] - __ldg(pos + kMapCol[m])) * kMapW[m];
    dv += fabsf(s.qd[kMapJoint[m]] - __ldg(vel + kMapCol[m])) * kMapWv[m];
  }
  float joint_score = expf(4.f * (-dj / (float)ILRL_JOINT_W_SUM));
  float jvel_score = expf((-dv / (float)ILRL_JOINT_WV_SUM) * 0.5f);
  float low_target = MODE == 1 ? 0.f : -hyp(w.e[ILRL_E_TARGET_X] - w.e[ILRL_E_ROBOT_X], w.e[ILRL_E_TARGET_Y] - w.e[ILRL_E_ROBOT_Y]);
  float posture = expf(-(fabsf(c.yaw - w.e[ILRL_E_HLDEG]) + fabsf(c.roll) + fabsf(c.pitch)));  // (Q9) no wrap
  float d_low = (low_target - w.e[ILRL_E_LOW_TARGET_SCORE]) / 0.0165f * 0.1f;                   // (Q5)
  w.e[ILRL_E_JOINT_SCORE] = joint_score; w.e[ILRL_E_JVEL_SCORE] = jvel_score;
  w.e[ILRL_E_LOW_TARGET_SCORE] = low_target; w.e[ILRL_E_POSTURE_SCORE] = posture;
  float run = 0.f, stall = 0.f;
#pragma unroll
  for (int i = 0; i < NJ; i++) { run += fabsf(action[i] * c.js[i]); stall += action[i] * action[i]; }
  float elec = -(run * (1.f / NJ)) - 0.1f * (stall * (1.f / NJ));
  float limit = -0.1f * (float)c.at_limit;
  float alive = (c.obs[0] + (float)ILRL_INITIAL_Z) > (float)ILRL_ALIVE_Z ? 2.f : -1.f;  // (Q11)
  terms[ILRL_T_JOINT] = joint_score; terms[ILRL_T_JVEL] = jvel_score; terms[ILRL_T_DLOWTARGET] = d_low;
  terms[ILRL_T_ELEC] = elec; terms[ILRL_T_LIMIT] = limit; terms[ILRL_T_ALIVE] = alive; terms[ILRL_T_POSTURE] = posture;
  terms[ILRL_T_ENDPOINT] = 0.f;
  if (MODE == 1) {
    w.e[ILRL_E_CUM_ALIVE] += alive;
    // calcDriftScore: distance of robot_pos to the segment starting_robot_pos -> target (math_util.py:20-27)
    float lx = w.e[ILRL_E_TARGET_X] - w.e[ILRL_E_START_X], ly = w.e[ILRL_E_TARGET_Y] - w.e[ILRL_E_START_Y];
    float len2 = lx * lx + ly * ly;
    float t = ((w.e[ILRL_E_ROBOT_X] - w.e[ILRL_E_START_X]) * lx + (w.e[ILRL_E_ROBOT_Y] - w.e[ILRL_E_START_Y]) * ly) / len2;
    t = fminf(fmaxf(t, 0.f), 1.f);
    float px = w.e[ILRL_E_START_X] + t * lx, py = w.e[ILRL_E_START_Y] + t * ly;
    w.e[ILRL_E_CUM_DRIFT] += expf(-6.f * hyp(px - w.e[ILRL_E_ROBOT_X], py - w.e[ILRL_E_ROBOT_Y]));
  }
  return (float)ILRL_RW_JOINT * joint_score + (float)ILRL_RW_JVEL * jvel_score + (float)ILRL_RW_TARGET * d_low +
         (float)ILRL_RW_ELEC * elec + (float)ILRL_RW_LIMIT * limit + (float)ILRL_RW_ALIVE * alive +
         (float)ILRL_RW_POSTURE * posture;
}

template <int MODE>
__device__ __forceinline__ void check_target(const Calc& c, EnvW& w, int rand_deg) {
  float dist = hyp(w.e[ILRL_E_ROBOT_X] - w.e[ILRL_E_TARGET_X], w.e[ILRL_E_ROBOT_Y] - w.e[ILRL_E_TARGET_Y]);
  if (dist <= (float)ILRL_TARGET_REACHED) {  // (Q8, Q12)
    float rr = c.yaw + (float)rand_deg * 0.017453292519943295f, sr, cr;
    sincosf(rr, &sr, &cr);
    float nx = w.e[ILRL_E_ROBOT_X] + cr * (float)ILRL_TARGET_LEN, ny = w.e[ILRL_E_ROBOT_Y] + sr * (float)ILRL_TARGET_LEN;
    w.e[ILRL_E_START_X] = w.e[ILRL_E_TARGET_X]; w.e[ILRL_E_START_Y] = w.e[ILRL_E_TARGET_Y];
    w.e[ILRL_E_TARGET_X] = nx; w.e[ILRL_E_TARGET_Y] = ny;
    float sc = -hyp(nx - w.e[ILRL_E_START_X], ny - w.e[ILRL_E_START_Y]);
    if (MODE >= 1) w.e[ILRL_E_HIGH_TARGET_SCORE] = sc;
    else w.e[ILRL_E_LOW_TARGET_SCORE] = sc;
    if (MODE == 2) {  // REF hier_env_2.py:562
      w.e[ILRL_E_SEP_X] = w.e[ILRL_E_ROBOT_X]; w.e[ILRL_E_SEP_Y] = w.e[ILRL_E_ROBOT_Y]; w.e[ILRL_E_SEP_Z] = 0.f;
    }
  }
  if (MODE == 2) {  // REF hier_env_2.py:567-570: heading and walk target follow the target every step
    w.e[ILRL_E_HLDEG] = atan2f(w.e[ILRL_E_TARGET_Y] - w.e[ILRL_E_ROBOT_Y], w.e[ILRL_E_TARGET_X] - w.e[ILRL_E_ROBOT_X]);
    w.e[ILRL_E_WALK_X] = w.e[ILRL_E_TARGET_X]; w.e[ILRL_E_WALK_Y] = w.e[ILRL_E_TARGET_Y];
  }
  if (MODE == 0) {
    float h = atan2f(w.e[ILRL_E_TARGET_Y] - w.e[ILRL_E_ROBOT_Y], w.e[ILRL_E_TARGET_X] - w.e[ILRL_E_ROBOT_X]), sh, ch;
    sincosf(h, &sh, &ch);
    w.e[ILRL_E_HLDEG] = h;
    w.e[ILRL_E_WALK_X] = w.e[ILRL_E_ROBOT_X] + ch * 10.f;
    w.e[ILRL_E_WALK_Y] = w.e[ILRL_E_ROBOT_Y] + sh * 10.f;
  }
}

template <int MODE>
__device__ __forceinline__ bool check_done(const EnvW& w, float alive) {
  const float margin = MODE >= 1 ? (float)ILRL_DONE_MARGIN_HI : (float)ILRL_DONE_MARGIN_LOW;
  bool near = hyp(w.e[ILRL_E_TARGET_X] - w.e[ILRL_E_ROBOT_X], w.e[ILRL_E_TARGET_Y] - w.e[ILRL_E_ROBOT_Y]) <=
              hyp(w.e[ILRL_E_TARGET_X] - w.e[ILRL_E_START_X], w.e[ILRL_E_TARGET_Y] - w.e[ILRL_E_START_Y]) + margin;
  return !(alive > 0.f && near);
}

__device__ __forceinline__ void update_reward_high(EnvW& w, float* terms, float step_per_level) {
  float hs = -hyp(w.e[ILRL_E_TARGET_X] - w.e[ILRL_E_ROBOT_X], w.e[ILRL_E_TARGET_Y] - w.e[ILRL_E_ROBOT_Y]);
  float d = step_per_level - w.e[ILRL_E_STEPS_REMAINING] + 1.f;  // (Q14)
  terms[ILRL_T_DHIGHTARGET] = (hs - w.e[ILRL_E_HIGH_TARGET_SCORE]) / 0.0165f / d;
  w.e[ILRL_E_HIGH_TARGET_SCORE] = hs;
  terms[ILRL_T_DRIFT] = w.e[ILRL_E_CUM_DRIFT] / d;
  w.e[ILRL_E_CUM_DRIFT] = 0.f;
}

__device__ __forceinline__ void rotz(float rad, const float* v, float* o) {
  float s, c;
  sincosf(rad, &s, &c);
  o[0] = c * v[0] - s * v[1]; o[1] = s * v[0] + c * v[1]; o[2] = v[2];
}

// resetFromFrame, in two phases around the forward kinematics of the reset pose (the caller runs its own FK between
// them: thread-per-env fk() in the reset kernel, the 4-lane qfk() in the fused step kernel).
// yaw_deg: low = caller's resetYaw (rotates the body only), hier = reset()'s own draw, folded into the heading (Q16).
struct ResetCtx { float sep[3]; float rot; int start_frame; };

// noise17 (MODE 2 only): what WalkerBase.robot_specific_reset left in the joints (uniform(-0.1, 0.1), velocity 0);
// hier_env_2's setJointsOrientation (:214-234) overwrites the abdomen and the 8 leg joints only, the arms keep it.
template <int MODE>
__device__ __forceinline__ void reset_pose(Phys& s, EnvW& w, const ClipDesc& cl, int start_frame, float yaw_deg,
                                           int target_deg, ResetCtx& rx, const float* target_xy = nullptr,
                                           const float* noise17 = nullptr) {
  const float D2R = 0.017453292519943295f;
  rx.sep[0] = w.e[ILRL_E_SEP_X]; rx.sep[1] = w.e[ILRL_E_SEP_Y]; rx.sep[2] = w.e[ILRL_E_SEP_Z];
  rx.start_frame = start_frame;
  float clip_id = w.e[ILRL_E_CLIP];
#pragma unroll
  for (int i = 0; i < ILRL_ENV_WORDS; i++) w.e[i] = 0.f;
  w.e[ILRL_E_CLIP] = clip_id;
  // degToTarget = rad2deg(atan2(target)): for a drawn target it is the integer draw itself (exact in fp32); an
  // explicit target (usePredefinedTarget, REF low_level_env.py:253-255) gives whatever its direction is
  float deg_to_target;
  if (target_xy) {
    w.e[ILRL_E_TARGET_X] = target_xy[0]; w.e[ILRL_E_TARGET_Y] = target_xy[1];
    deg_to_target = atan2f(target_xy[1], target_xy[0]) * 57.29577951308232f;
  } else {
    float st, ct;
    sincosf((float)target_deg * D2R, &st, &ct);
    w.e[ILRL_E_TARGET_X] = ct * (float)ILRL_TARGET_LEN;
    w.e[ILRL_E_TARGET_Y] = st * (float)ILRL_TARGET_LEN;
    deg_to_target = (float)target_deg;
  }
  w.e[ILRL_E_FRAME] = (float)start_frame;
#pragma unroll
  for (int j = 0; j < NJ; j++) { s.q[j] = (MODE == 2 && noise17 && j >= 3) ? noise17[j] : 0.f; s.qd[j] = 0.f; }
  const float* pos = cl.pos + start_frame * 14;
  const float* vel = cl.vel + start_frame * 14;
#pragma unroll
  for (int m = 0; m < (MODE == 2 ? 8 : NMAP); m++) {
    s.q[kMapJoint[m]] = __ldg(pos + kMapCol[m]);
    s.qd[kMapJoint[m]] = __ldg(vel + kMapCol[m]);
  }
  s.p[0] = 0.f; s.p[1] = 0.f; s.p[2] = MODE == 2 ? (float)ILRL_RESET_Z_HIER2 : (float)ILRL_RESET_Z;
  float body_deg;
  if (MODE >= 1) { deg_to_target += yaw_deg; body_deg = deg_to_target; }
  else body_deg = deg_to_target + yaw_deg;
  float sd, cd;
  sincosf(deg_to_target, &sd, &cd);  // (Q1) degrees fed to cos/sin as radians
  const float wlen = MODE == 2 ? (float)ILRL_TARGET_LEN : 1000.f;  // REF hier_env_2.py:304-305
  w.e[ILRL_E_WALK_X] = cd * wlen; w.e[ILRL_E_WALK_Y] = sd * wlen;
  float sh, ch;
  sincosf(0.5f * body_deg * D2R, &sh, &ch);
  s.quat[0] = 0.f; s.quat[1] = 0.f; s.quat[2] = sh; s.quat[3] = ch;
  w.e[ILRL_E_HLDEG] = deg_to_target * D2R;
  s.w[0] = s.w[1] = s.w[2] = 0.f;
  rx.rot = deg_to_target * D2R;
}

// rfx, rfy: origin of the right_foot body relative to the torso origin; sumx, sumy: sums of the 31 part offsets.
template <int MODE>
__device__ __forceinline__ void reset_finish(Phys& s, EnvW& w, const ClipDesc& cl, const ResetCtx& rx, float rfx,
                                             float rfy, float sumx, float sumy, float step_per_level, int skip_frame,
                                             Calc& c) {
  const int start_frame = rx.start_frame;
  const float rot = rx.rot;
  const float* ep0 = cl.ep + start_frame * 27;
  if (MODE == 0) {
    const float* ep1 = cl.ep + ((start_frame + skip_frame) % cl.max_frame) * 27;  // REF low_level_env.py:278
    float rf[3] = {__ldg(ep0 + 9), __ldg(ep0 + 10), __ldg(ep0 + 11)}, rfr[3];  // RightFoot
    rotz(rot, rf, rfr);
    w.e[ILRL_E_SEP_X] = s.p[0] + rfx - rfr[0];
    w.e[ILRL_E_SEP_Y] = s.p[1] + rfy - rfr[1];
    w.e[ILRL_E_SEP_Z] = 0.f;
    float d[3] = {__ldg(ep1 + 6) - __ldg(ep0 + 6), __ldg(ep1 + 7) - __ldg(ep0 + 7), __ldg(ep1 + 8) - __ldg(ep0 + 8)}, dr[3];
    rotz(rot, d, dr);  // RightLeg displacement over skipFrame frames
#pragma unroll
    for (int i = 0; i < 3; i++) s.v[i] = (dr[i] / 0.0165f) / 1.2f;
  } else {
    const float* ep1 = cl.ep + (start_frame + 1) * 27;
    if (MODE == 2) {  // REF hier_env_2.py:325-335: anchored on the right foot as in the low-level env
      float rf[3] = {__ldg(ep0 + 9), __ldg(ep0 + 10), __ldg(ep0 + 11)}, rfr[3];
      rotz(rot, rf, rfr);
      w.e[ILRL_E_SEP_X] = s.p[0] + rfx - rfr[0];
      w.e[ILRL_E_SEP_Y] = s.p[1] + rfy - rfr[1];
      w.e[ILRL_E_SEP_Z] = 0.f;
    } else {
      w.e[ILRL_E_SEP_X] = rx.sep[0]; w.e[ILRL_E_SEP_Y] = rx.sep[1]; w.e[ILRL_E_SEP_Z] = rx.sep[2];
    }
    float d[3] = {__ldg(ep1 + 6) - __ldg(ep0 + 6), __ldg(ep1 + 7) - __ldg(ep0 + 7), __ldg(ep1 + 8) - __ldg(ep0 + 8)}, dr[3];
    rotz(rot, d, dr);
#pragma unroll
    for (int i = 0; i < 3; i++) s.v[i] = dr[i] / 0.0165f;
    w.e[ILRL_E_HIGH_TARGET_SCORE] = -(float)ILRL_TARGET_LEN;
    w.e[ILRL_E_STEPS_REMAINING] = step_per_level;
    w.e[ILRL_E_HIGH_PENDING] = 1.f;
  }
  inc_frame<MODE>(w, cl, skip_frame);
  calc_state(s, sumx, sumy, w.e[ILRL_E_WALK_X], w.e[ILRL_E_WALK_Y], c);
  w.e[ILRL_E_OBS_SIN] = c.obs[1]; w.e[ILRL_E_OBS_COS] = c.obs[2];
}

// the 14 joints hier_env_2's reset leaves at WalkerBase's noise: uniform(-0.1, 0.1) from the env's Philox stream
// (joints 0..2, the abdomen, are overwritten with 0; the legs by the clip)
__device__ __forceinline__ void draw_reset_noise(uint64_t seed, uint32_t env, uint32_t& ctr, float* noise17) {
#pragma unroll
  for (int j = 0; j < NJ; j++)
    noise17[j] = j < 11 ? 0.f : -0.1f + 0.2f * ((float)(philox_u32(seed, env, ctr++) >> 8) * (1.f / 16777216.f));
}

// thread-per-env form (reset kernel).  Leaves FK of the reset pose in k and the calc_state result in c.
template <int MODE>
__device__ __forceinline__ void reset_env(Phys& s, EnvW& w, const ClipDesc& cl, int start_frame, float yaw_deg,
                                          int target_deg, float step_per_level, int skip_frame, Work& k, Calc& c,
                                          const float* target_xy = nullptr, const float* noise17 = nullptr) {
  ResetCtx rx;
  reset_pose<MODE>(s, w, cl, start_frame, yaw_deg, target_deg, rx, target_xy, noise17);
  fk(s, k);
  reset_finish<MODE>(s, w, cl, rx, k.o[5][0], k.o[5][1], k.sumx, k.sumy, step_per_level, skip_frame, c);  // body 5 = right_foot
}

// calcEndPointScore(useExp=True)
__device__ __forceinline__ float endpoint_score(const Phys& s, const Work& k, const EnvW& w, const ClipDesc& cl) {
  const float* ep = cl.ep + (int)w.e[ILRL_E_FRAME] * 27;
  // link0_11 (right knee anchor) -> RightLeg, right_foot -> RightFoot, link0_18 -> LeftLeg, left_foot -> LeftFoot
  const int col[4] = {6, 9, 0, 3};
  const float wg[4] = {1.f, 3.f, 1.f, 3.f};
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    V3 part;
    if (i == 0 || i == 2) {  // origin of the knee's joint link = the knee anchor, given in the shin frame
      const int b = i == 0 ? 4 : 7, j = i == 0 ? 6 : 10;
      part = ld3(k.o[b]) + mv(k.R[b], mk(kJointAnchor[3 * j], kJointAnchor[3 * j + 1], kJointAnchor[3 * j + 2]));
    } else {
      part = ld3(k.o[i == 1 ? 5 : 8]);
    }
    float v[3] = {__ldg(ep + col[i]), __ldg(ep + col[i] + 1), __ldg(ep + col[i] + 2)}, r[3];
    rotz(w.e[ILRL_E_HLDEG], v, r);
    float dx = w.e[ILRL_E_SEP_X] + r[0] - (s.p[0] + part.x), dy = w.e[ILRL_E_SEP_Y] + r[1] - (s.p[1] + part.y),
          dz = w.e[ILRL_E_SEP_Z] + r[2] - (s.p[2] + part.z);
    acc += sqrtf(dx * dx + dy * dy + dz * dz) * wg[i];
  }
  return expf(3.f * (-acc / (float)ILRL_EP_W_SUM));
}

}  // namespace ilrl
