// ilrl_chain.cuh — third kernel generation of the physics substep: FOUR LANES PER ENV, every per-link computation a
// ROLLED loop over a 7-link chain, all per-link data in shared memory.
//
// Why (profiles/r1_v2_step_kernel_ncu.md): the second generation kept each limb in registers and therefore had to be
// fully unrolled: 12 k SASS instructions (190 KB) per substep, re-fetched from L2 four times per env step; 36 % of the
// stall samples were `no_instruction`.  Here
//   * lane r of a quad walks the chain  spine(3 links, identical in all four lanes) + limb r (4 slots; the arms have a
//     leading dummy slot with a zero axis).  For the legs that chain IS the path torso -> foot; the arms restart from
//     the torso at chain index 3.  Forward kinematics / velocities / body inertias, the inward articulated-inertia
//     pass and the outward acceleration pass are three `#pragma unroll 1` loops over the chain index whose bodies are
//     a few hundred instructions each: one substep is ~4 k instructions and its loops run out of the instruction caches;
//   * per-link records (S, cJ, U, 1/D, u, q, qd, tau, nu) live in shared memory, thread-minor (conflict-free); the
//     spine's are stored once per env.  They double as the publication that lets ANY lane build constraint rows;
//   * the limbs' articulated inertias meet at the pelvis / torso through two rounds of __shfl_xor over 27 words;
//   * constraint rows are built three at a time (one contact = normal + 2 friction directions, or up to three joint
//     limits) by ONE routine with 3-way instruction-level parallelism, dealt round-robin to the lanes, and stored as
//     40-word records (128-bit loads, env stride = 4 mod 32 words: the 8 envs of a warp hit 8 different bank groups);
//   * projected Gauss-Seidel keeps the velocity change distributed like the state (base + spine replicated, limb
//     private); a row's limb part belongs to one lane, so the 4-lane reduction is a single broadcast shuffle.
// Row order, row formulas and constants are exactly those of the oracle (oracle/ilrl_oracle.c) and of the earlier
// generations (git history): limits in joint order, then contacts in sphere-table order, 5 sweeps.
#pragma once
#include <type_traits>

#include "ilrl_env.cuh"

namespace ilrl {
namespace chain {

#ifndef ILRL_QE
#define ILRL_QE 16
#endif
constexpr int QE = ILRL_QE;            // envs per CTA
constexpr int QT = 4 * QE;             // threads per CTA
constexpr int NL = 7;                  // chain links per lane: 3 spine + 4 limb slots
constexpr int RW = 40;                 // words per stored row
#ifndef ILRL_MIN_RSM
#define ILRL_MIN_RSM 2   // = the row budget of LayoutDense4
#endif
constexpr int MIN_RSM = ILRL_MIN_RSM;  // smallest on-chip row budget of any layout
constexpr int GROWS = MAXROWS - MIN_RSM;  // rows per env in the global overflow scratch
// link record: 24 words in 6 float4 (S | cJ | U dinv u | q qd tau nu), one contiguous 28-word slot per thread: a
// 28-word stride puts the float4 of 8 consecutive threads in 8 different bank groups (conflict-free LDS.128 / STS.128)
enum { W_S = 0, W_CJ = 6, W_U = 12, W_DINV = 18, W_UU = 19, W_Q = 20, W_QD = 21, W_TAU = 22, W_NU = 23, LKW = 28 };
// body record words: rigid inertia about the reference point (A 6, m*c 3, m 1) + bias force 6
constexpr int RECW = 16;
// stored row words
enum { R_RB = 0 /*resp base 6*/, R_RS = 6 /*resp spine 3*/, R_JS = 9 /*J spine 3*/, R_JB = 12 /*J base 6*/, R_RHS = 18,
       R_DINV = 19, R_RL = 20 /*resp limbs 4x4*/, R_JL = 36 /*J of the row's limb 4*/ };

// ---- compile-time model tables in chain form
struct LinkC {
  float pre[3];   // translation (parent body frame) applied before the joint: the body position when the link starts a body
  float an[3];    // joint anchor (body frame)
  float ax[3];    // joint axis (body frame); zero for a dummy slot
  float lo, hi;   // limits (dummy: never violated)
  float gear;     // torque per unit of clipped action
  int j;          // global joint index, -1 = dummy slot
  int motor;      // action slot that drives the joint (humanoid.py:28-37), -1 = none
  int rot;        // 1: the body that starts here carries the fixed rotation kQ
  int nbody;      // rigid bodies completed by this link
  int pad[2];     // 18 words: a chain of 4 is 72 words = 8 (mod 32), so the four roles of a quad - which read the same
                  // field of four different chains in one instruction - hit four different banks
};
struct BodyC {
  float off[3];      // body origin relative to the link's body origin (0, or the fixed child's position)
  float m, ix, iy, iz;
  int nsph;
  float sph[2][4];   // centre (body frame), radius
  int sidx[2];       // global sphere index
  float reach;       // max over the body's spheres of |centre| + radius: no sphere can touch the ground while
                     // origin_z - reach >= the breaking distance
  int pad[3];        // 22 words: a chain's 4 records are 88 words = 24 (mod 32) (bank spread over the roles, as LinkC)
};
struct Tables {
  LinkC lc[5][4];        // chains 0..3 = limbs (right leg, left leg, right arm, left arm), 4 = spine (3 links)
  BodyC bc[5][2][2];     // [chain][slot][k]: limb slot 0 = link 2 (body A), slot 1 = link 3 (bodies B, E);
                         //                    spine slot 0 = link 1 (lwaist), slot 1 = link 2 (pelvis)
  int sphL[NS], sphC[NS];  // chain (limb 0..3, -1 = spine/torso) and chain index (-1 = torso) carrying sphere g
  int jL[NJ], jC[NJ];      // the same for joint j
  float Q[9];              // fixed rotation of lwaist / pelvis in their parent frame
  float torso_reach;       // as BodyC::reach, for the 5 torso spheres
  constexpr Tables() : lc(), bc(), sphL(), sphC(), jL(), jC(), Q(), torso_reach(0.f) {
    constexpr int jb[NJ] = ILRL_JOINT_BODY;
    constexpr double bp[NB * 3] = ILRL_BODY_POS;
    constexpr double bq[NB * 4] = ILRL_BODY_QUAT;
    constexpr double bm[NB] = ILRL_BODY_MASS;
    constexpr double bi[NB * 3] = ILRL_BODY_INERTIA;
    constexpr double ja[NJ * 3] = ILRL_JOINT_ANCHOR;
    constexpr double jx[NJ * 3] = ILRL_JOINT_AXIS;
    constexpr double jlo[NJ] = ILRL_JOINT_LO;
    constexpr double jhi[NJ] = ILRL_JOINT_HI;
    constexpr int sb[NS] = ILRL_SPHERE_BODY;
    constexpr int sl[NS] = ILRL_SPHERE_LINK;
    constexpr double sc[NS * 3] = ILRL_SPHERE_C;
    constexpr double sr[NS] = ILRL_SPHERE_R;
    constexpr int mj[NJ] = ILRL_MOTOR_JOINT;
    constexpr double mg[NJ] = ILRL_MOTOR_GEAR;
    constexpr int first[4] = {3, 7, 11, 14};
    constexpr int count[4] = {4, 4, 3, 3};
    int body_of[5][2][2] = {};
    for (int q = 0; q < 5; q++) {
      for (int k = 0; k < 4; k++) {
        LinkC& o = lc[q][k];
        int j = -1;
        if (q < 4) { const int lead = 4 - count[q]; j = k < lead ? -1 : first[q] + (k - lead); }
        else if (k < 3) j = k;
        o.j = j; o.motor = -1; o.gear = 0.f; o.rot = 0; o.nbody = 0; o.pad[0] = o.pad[1] = 0;
        o.lo = -1e30f; o.hi = 1e30f;
        for (int i = 0; i < 3; i++) { o.pre[i] = 0.f; o.an[i] = 0.f; o.ax[i] = 0.f; }
        if (q == 4 && k == 3) continue;
        const int jr = j >= 0 ? j : first[q];          // a dummy slot borrows the anchor of the body's first joint
        for (int i = 0; i < 3; i++) { o.an[i] = (float)ja[3 * jr + i]; if (j >= 0) o.ax[i] = (float)jx[3 * j + i]; }
        if (j >= 0) {
          o.lo = (float)jlo[j]; o.hi = (float)jhi[j];
          for (int m = 0; m < NJ; m++) if (mj[m] == j) { o.motor = m; o.gear = (float)mg[m]; }
        }
        // does this link start a new body?  (first link of the chain, or the body of its joint differs from the previous link's)
        const int b = jb[jr];
        bool starts = k == 0;
        if (k > 0) { const int jp = lc[q][k - 1].j >= 0 ? lc[q][k - 1].j : first[q]; starts = jb[jp] != b; }
        if (starts) {
          for (int i = 0; i < 3; i++) o.pre[i] = (float)bp[3 * b + i];
          o.rot = (bq[4 * b] != 0.0 || bq[4 * b + 1] != 0.0 || bq[4 * b + 2] != 0.0) ? 1 : 0;
        }
      }
      // bodies completed by links
      if (q < 4) {
        const int bA = jb[first[q]], bB = jb[first[q] + count[q] - 1], bE = bB + 1;
        lc[q][2].nbody = 1; lc[q][3].nbody = 2;
        body_of[q][0][0] = bA; body_of[q][0][1] = -1; body_of[q][1][0] = bB; body_of[q][1][1] = bE;
      } else {
        lc[q][1].nbody = 1; lc[q][2].nbody = 1;
        body_of[q][0][0] = 1; body_of[q][0][1] = -1; body_of[q][1][0] = 2; body_of[q][1][1] = -1;
      }
      for (int s = 0; s < 2; s++)
        for (int k = 0; k < 2; k++) {
          BodyC& o = bc[q][s][k];
          const int b = body_of[q][s][k];
          o.nsph = 0; o.m = 0.f; o.ix = o.iy = o.iz = 0.f; o.pad[0] = o.pad[1] = o.pad[2] = 0; o.reach = 0.f;
          for (int i = 0; i < 3; i++) o.off[i] = 0.f;
          for (int t = 0; t < 2; t++) { o.sidx[t] = 0; for (int i = 0; i < 4; i++) o.sph[t][i] = 0.f; }
          if (b < 0) continue;
          if (k == 1) for (int i = 0; i < 3; i++) o.off[i] = (float)bp[3 * b + i];   // fixed child of the slot's first body
          o.m = (float)bm[b]; o.ix = (float)bi[3 * b]; o.iy = (float)bi[3 * b + 1]; o.iz = (float)bi[3 * b + 2];
          for (int g = 0; g < NS; g++)
            if (sb[g] == b) {
              const int t = o.nsph++;
              o.sidx[t] = g;
              for (int i = 0; i < 3; i++) o.sph[t][i] = (float)sc[3 * g + i];
              o.sph[t][3] = (float)sr[g];
              // |c| + r, rounded up (1-norm bound on the centre: no sqrt in a constexpr constructor, and it is safe)
              double n1 = 0.0;
              for (int i = 0; i < 3; i++) n1 += sc[3 * g + i] < 0 ? -sc[3 * g + i] : sc[3 * g + i];
              const float rch = (float)(n1 + sr[g]) * 1.0001f;
              if (rch > o.reach) o.reach = rch;
            }
        }
    }
    for (int j = 0; j < NJ; j++) {
      if (j < 3) { jL[j] = -1; jC[j] = j; }
      else
        for (int q = 0; q < 4; q++)
          for (int k = 0; k < 4; k++) if (lc[q][k].j == j) { jL[j] = q; jC[j] = 3 + k; }
    }
    for (int g = 0; g < NS; g++) {
      if (sl[g] < 0) { sphL[g] = -1; sphC[g] = -1; } else { sphL[g] = jL[sl[g]]; sphC[g] = jC[sl[g]]; }
    }
    for (int g = NS - 5; g < NS; g++) {
      double n1 = 0.0;
      for (int i = 0; i < 3; i++) n1 += sc[3 * g + i] < 0 ? -sc[3 * g + i] : sc[3 * g + i];
      const float rch = (float)(n1 + sr[g]) * 1.0001f;
      if (rch > torso_reach) torso_reach = rch;
    }
    // kQ from the quaternion of body 1 (bodies 1 and 2 carry the same one)
    const double x = bq[4], y = bq[5], z = bq[6], w = bq[7];
    const double s = 2.0 / (x * x + y * y + z * z + w * w);
    Q[0] = (float)(1 - s * (y * y + z * z)); Q[1] = (float)(s * (x * y - w * z)); Q[2] = (float)(s * (x * z + w * y));
    Q[3] = (float)(s * (x * y + w * z)); Q[4] = (float)(1 - s * (x * x + z * z)); Q[5] = (float)(s * (y * z - w * x));
    Q[6] = (float)(s * (x * z - w * y)); Q[7] = (float)(s * (y * z + w * x)); Q[8] = (float)(1 - s * (x * x + y * y));
  }
};
__device__ constexpr Tables kTables{};
constexpr int TABLE_WORDS = sizeof(Tables) / 4;
static_assert(sizeof(Tables) % 4 == 0, "table copy is word-wise");

// ---- shared memory of one CTA, in three layouts chosen by the host from the batch size (ilrl_create)
//   LayoutSmall : up to one wave of CTAs at 2 per SM (<= 4736 envs on 148 SMs): everything on chip - 20 rows per env,
//                 padded (conflict-free) body records, model tables in shared memory.  108 KB.
//   LayoutLarge : larger batches, where resident warps per SM are what limits throughput: 8 rows per env on chip (the
//                 rest in the L2-resident global scratch), unpadded body records, model tables read through L1.
//                 73 KB -> 3 CTAs per SM (7104 envs per wave).
//   LayoutDense4: 2 rows per env on chip, body-view lane stride 32 (4-way bank conflicts on its 4 float4 accesses per
//                 substep).  55.5 KB -> 4 CTAs per SM (9472 envs per wave): a tile is ~1.35x slower than LayoutLarge's,
//                 so it wins exactly where it saves a round of tiles (8192 envs: 1 instead of 2; 16384: 2 instead of 3).
// In all, the body records (written by the FK phase, consumed by the inward pass) share their storage with what
// only the row phase uses (response scratch, multipliers, row owners) and with the action tile (consumed before the
// first substep).
template <int RSM_, int BRW_, bool TSM_, int LANE_>
struct Layout {
  static constexpr int RSM = RSM_;               // constraint rows per env kept in shared memory
  static constexpr int ROWSTRIDE = RSM_ * RW + 4;  // env stride in words: = 4 (mod 32)
  static constexpr int BRW = BRW_;               // words per body record (20: room for padding; 16: dense)
  static constexpr bool TSM = TSM_;              // model tables staged in shared memory
  static constexpr int LANE = LANE_;             // lane stride of the body view in the per-env scratch block (words)
};
#ifndef ILRL_LARGE_RSM  // (overridable for layout experiments)
#define ILRL_LARGE_RSM 8
#define ILRL_LARGE_BRW 16
#define ILRL_LARGE_TSM false
#endif
static_assert(ILRL_LARGE_RSM >= MIN_RSM, "the overflow scratch holds MAXROWS - MIN_RSM rows per env");
#ifndef ILRL_SMALL_RSM
#define ILRL_SMALL_RSM 20   // 16 -> 20: +3 % at 4096 envs (fewer envs overflow, and a warp with both kinds runs two loop variants)
#endif
using LayoutSmall = Layout<ILRL_SMALL_RSM, 20, true, 40>;
using LayoutLarge = Layout<ILRL_LARGE_RSM, ILRL_LARGE_BRW, ILRL_LARGE_TSM, 40>;
using LayoutDense4 = Layout<2, 16, false, 32>;

// Per-env scratch block, seen in two ways that are never live at the same time within an env:
//   body view (FK phase -> inward pass): body records of the four lanes (2 each, lane stride LANE words) + the spine's 2
//   row view  (row phase)              : response scratch u[lane][3 impulses][7 links], multipliers, action tile
//                                        (consumed before the first substep), row owners (bytes)
// The block is PER ENV because warps of a CTA run unsynchronised: one env's row phase must never touch another env's
// body records.  Env stride = 4 (mod 32) words and lane stride 40 (LayoutSmall / LayoutLarge): the float4 accesses of the 2 envs x 4 lanes of a
// quarter-warp fall in 8 different bank groups, and scalar accesses of the 8 envs of a warp in 8 different banks.
constexpr int SCR_SU = 0, SCR_LAM = 4 * 3 * NL, SCR_ACT = SCR_LAM + MAXROWS, SCR_ROWL = SCR_ACT + NJ;  // row view (words)
constexpr int SCR_ROWVIEW = SCR_ROWL + (MAXROWS + 3) / 4;
template <int BRW, int LANE> constexpr int scr_words() {
  int w = 4 * LANE + 2 * BRW;                      // body view
  if (w < SCR_ROWVIEW) w = SCR_ROWVIEW;
  while (w % 32 != 4) w++;
  return w;
}

struct NoTables {};
template <class LY>
struct __align__(16) SmemT {
  static constexpr int RSM = LY::RSM, ROWSTRIDE = LY::ROWSTRIDE, BRW = LY::BRW, ES = scr_words<LY::BRW, LY::LANE>();
  static constexpr bool TSM = LY::TSM;
  static constexpr int SCR_LANE = LY::LANE;
  static_assert(2 * BRW <= SCR_LANE, "two body records per lane");
  // --- float4-accessed arrays first (every size below is a multiple of 16 bytes)
  float rows[QE][ROWSTRIDE];   // stored rows; after the substeps the first 71 words of an env's block stage its obs row
  float lk[4][QT][LKW];        // limb link records, per thread
  float sp[3][QE][LKW];        // spine link records, per env
  float scr[QE][ES];           // per-env scratch block (see above)
  // --- scalar-accessed
  float L0[21][QE];            // Cholesky factor of the base articulated inertia
  float sph[NS][3][QE];        // contact candidates: x, y, z - r relative to the torso origin (distance = base z + that)
  typename std::conditional<TSM, Tables, NoTables>::type T;
  static_assert((sizeof(float) * QE * ROWSTRIDE) % 16 == 0 && (sizeof(float) * LKW) % 16 == 0 && BRW % 4 == 0 && ES % 4 == 0,
                "float4 alignment of the shared-memory records");
  static_assert(ROWSTRIDE >= 71, "the obs row is staged in the env's row block");
  // body view
  __device__ __forceinline__ float* rl(int e, int lane, int slot) { return &scr[e][lane * SCR_LANE + slot * BRW]; }
  __device__ __forceinline__ float* rs(int e, int slot) { return &scr[e][4 * SCR_LANE + slot * BRW]; }
  // row view
  __device__ __forceinline__ float* su(int e, int lane, int i) { return &scr[e][SCR_SU + (lane * 3 + i) * NL]; }
  __device__ __forceinline__ float* lam(int e) { return &scr[e][SCR_LAM]; }
  __device__ __forceinline__ float* act(int e) { return &scr[e][SCR_ACT]; }
  __device__ __forceinline__ signed char* rowL(int e) { return reinterpret_cast<signed char*>(&scr[e][SCR_ROWL]); }
};
// the model tables as the kernel sees them: the CTA's shared-memory copy, or the global object through L1
template <class SM>
__device__ __forceinline__ const Tables& tables(const SM& sm) {
  if constexpr (SM::TSM) return sm.T; else return kTables;
}

// replicated floating-base state of one env
struct Base { float p[3], quat[4], v[3], w[3]; };

__device__ __forceinline__ float qsum(float v, unsigned qm) {
  v += __shfl_xor_sync(qm, v, 1);
  v += __shfl_xor_sync(qm, v, 2);
  return v;
}
__device__ __forceinline__ SV neg(SV a) { SV r; r.a = mk(-a.a.x, -a.a.y, -a.a.z); r.l = mk(-a.l.x, -a.l.y, -a.l.z); return r; }
__device__ __forceinline__ SV svzero() { SV r; r.a = r.l = mk(0, 0, 0); return r; }
__device__ __forceinline__ V3 rd3(const float* p) { return mk(p[0], p[1], p[2]); }
__device__ __forceinline__ float rcp_or_zero(float d) { return d > 1e-9f ? __frcp_rn(d) : 0.f; }

// sin/cos of a joint angle.  Joint angles stay within a few radians of their limits (|q| << 100), so the Payne-Hanek
// large-argument path of sincosf() would be dead weight in the instruction stream: two-constant Cody-Waite reduction
// by pi/2 + the minimax polynomials of the usual fast path, max error 7e-8 for |q| < 100 (checked on the host).
__device__ __forceinline__ void sincos_joint(float x, float& s, float& c) {
  const float k = rintf(x * 0.636619772f);
  float r = fmaf(k, -1.57079601e+00f, x);
  r = fmaf(k, -3.13916473e-07f, r);
  r = fmaf(k, -5.39030253e-15f, r);
  const int q = (int)k;
  const float r2 = r * r;
  float sp = fmaf(r2, -1.95152959e-4f, 8.33216087e-3f);
  sp = fmaf(sp, r2, -1.66666546e-1f);
  sp = fmaf(sp * r2, r, r);
  float cp = fmaf(r2, 2.44331571e-5f, -1.38873163e-3f);
  cp = fmaf(cp, r2, 4.16666457e-2f);
  cp = fmaf(cp, r2, -0.5f);
  cp = fmaf(cp, r2, 1.0f);
  const float a = (q & 1) ? cp : sp, b = (q & 1) ? sp : cp;
  s = (q & 2) ? -a : a;
  c = ((q + 1) & 2) ? -b : b;
}

// position of the n-th (0-based) set bit of m (m has more than n bits set)
__device__ __forceinline__ int nth_set_bit(uint32_t m, int n) {
#pragma unroll 1
  for (int i = 0; i < n; i++) m &= m - 1;
  return __ffs(m) - 1;
}

// link record of chain index c of THIS lane (spine: the env's shared record)
template <class SM>
__device__ __forceinline__ float* link_rec(SM& sm, int c, int e, int tid) {
  return c < 3 ? &sm.sp[c][e][0] : &sm.lk[c - 3][tid][0];
}
// link record of chain index c of limb L of this env (L < 0 or c < 3: spine)
template <class SM>
__device__ __forceinline__ const float* link_rec_of(const SM& sm, int L, int c, int e, int qb) {
  return c < 3 ? &sm.sp[c][e][0] : &sm.lk[c - 3][qb + L][0];
}
// record field access in whole float4 (word layout: S 0..5 | cJ 6..11 | U 12..17, dinv 18, u 19 | q qd tau nu 20..23)
__device__ __forceinline__ void ld_SU(const float* rec, SV& S, SV& U, float& dinv) {
  const float4* r = reinterpret_cast<const float4*>(rec);
  const float4 a = r[0], b = r[1], c = r[3], d = r[4];
  S.a = mk(a.x, a.y, a.z); S.l = mk(a.w, b.x, b.y);
  U.a = mk(c.x, c.y, c.z); U.l = mk(c.w, d.x, d.y);
  dinv = d.z;
}
__device__ __forceinline__ void ld_ScJ(const float* rec, SV& S, SV& cJ) {
  const float4* r = reinterpret_cast<const float4*>(rec);
  const float4 a = r[0], b = r[1], c = r[2];
  S.a = mk(a.x, a.y, a.z); S.l = mk(a.w, b.x, b.y);
  cJ.a = mk(b.z, b.w, c.x); cJ.l = mk(c.y, c.z, c.w);
}
__device__ __forceinline__ void ld_all(const float* rec, SV& S, SV& cJ, SV& U, float& dinv, float& u) {
  const float4* r = reinterpret_cast<const float4*>(rec);
  const float4 a = r[0], b = r[1], c = r[2], d = r[3], f = r[4];
  S.a = mk(a.x, a.y, a.z); S.l = mk(a.w, b.x, b.y);
  cJ.a = mk(b.z, b.w, c.x); cJ.l = mk(c.y, c.z, c.w);
  U.a = mk(d.x, d.y, d.z); U.l = mk(d.w, f.x, f.y);
  dinv = f.z; u = f.w;
}
__device__ __forceinline__ void st_ScJ(float* rec, SV S, SV cJ) {
  float4* r = reinterpret_cast<float4*>(rec);
  r[0] = make_float4(S.a.x, S.a.y, S.a.z, S.l.x);
  r[1] = make_float4(S.l.y, S.l.z, cJ.a.x, cJ.a.y);
  r[2] = make_float4(cJ.a.z, cJ.l.x, cJ.l.y, cJ.l.z);
}
__device__ __forceinline__ void st_Udu(float* rec, SV U, float dinv, float u) {
  float4* r = reinterpret_cast<float4*>(rec);
  r[3] = make_float4(U.a.x, U.a.y, U.a.z, U.l.x);
  r[4] = make_float4(U.l.y, U.l.z, dinv, u);
}

// rigid inertia (about the reference point, world axes) and bias force of one body, ACCUMULATED into rec[16]
__device__ __forceinline__ void rigid_rec(const float* R, V3 c, float m, float ix, float iy, float iz, SV V, float* rec) {
  float Ic[6];
  Ic[0] = ix * R[0] * R[0] + iy * R[1] * R[1] + iz * R[2] * R[2];
  Ic[1] = ix * R[0] * R[3] + iy * R[1] * R[4] + iz * R[2] * R[5];
  Ic[2] = ix * R[0] * R[6] + iy * R[1] * R[7] + iz * R[2] * R[8];
  Ic[3] = ix * R[3] * R[3] + iy * R[4] * R[4] + iz * R[5] * R[5];
  Ic[4] = ix * R[3] * R[6] + iy * R[4] * R[7] + iz * R[5] * R[8];
  Ic[5] = ix * R[6] * R[6] + iy * R[7] * R[7] + iz * R[8] * R[8];
  const float cc = dot(c, c);
  rec[0] += Ic[0] + m * (cc - c.x * c.x); rec[1] += Ic[1] - m * c.x * c.y; rec[2] += Ic[2] - m * c.x * c.z;
  rec[3] += Ic[3] + m * (cc - c.y * c.y); rec[4] += Ic[4] - m * c.y * c.z; rec[5] += Ic[5] + m * (cc - c.z * c.z);
  rec[6] += m * c.x; rec[7] += m * c.y; rec[8] += m * c.z; rec[9] += m;
  // momentum about the reference point and its velocity-product rate; gravity + Bullet's damping as external forces
  V3 vc = V.l + cross(V.a, c);
  V3 Iw = symv(Ic, V.a);
  SV h; h.l = m * vc; h.a = Iw + cross(c, h.l);
  SV pb = crf(V, h);
  const float vn = sqrtf(dot(vc, vc)), wn = sqrtf(dot(V.a, V.a));
  const float kl = m * ((float)ILRL_DAMP_K1_LIN + (float)ILRL_DAMP_K2_LIN * vn);
  const float ka = (float)ILRL_DAMP_K1_ANG + (float)ILRL_DAMP_K2_ANG * wn;
  V3 F = mk(-kl * vc.x, -kl * vc.y, -kl * vc.z - m * (float)ILRL_GRAVITY);
  V3 N = mk(-ka * Iw.x, -ka * Iw.y, -ka * Iw.z);
  pb.a = pb.a - (N + cross(c, F));
  pb.l = pb.l - F;
  rec[10] += pb.a.x; rec[11] += pb.a.y; rec[12] += pb.a.z; rec[13] += pb.l.x; rec[14] += pb.l.y; rec[15] += pb.l.z;
}

struct IP { Inertia I; SV p; };  // articulated inertia + bias force (27 words)

__device__ __forceinline__ void ip_add_rec(IP& x, const float* r) {
#pragma unroll
  for (int i = 0; i < 6; i++) x.I.A[i] += r[i];
  const float mx = r[6], my = r[7], mz = r[8], m = r[9];
  x.I.B[1] -= mz; x.I.B[2] += my; x.I.B[3] += mz; x.I.B[5] -= mx; x.I.B[6] -= my; x.I.B[7] += mx;
  x.I.C[0] += m; x.I.C[3] += m; x.I.C[5] += m;
  x.p.a.x += r[10]; x.p.a.y += r[11]; x.p.a.z += r[12];
  x.p.l.x += r[13]; x.p.l.y += r[14]; x.p.l.z += r[15];
}
__device__ __forceinline__ void ip_add_rec4(IP& x, const float* rec) {  // 16-byte aligned record in shared memory
  const float4* r4 = reinterpret_cast<const float4*>(rec);
  const float4 a = r4[0], b = r4[1], c = r4[2], d = r4[3];
  const float r[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
  ip_add_rec(x, r);
}
__device__ __forceinline__ void ip_shfl_add(IP& x, unsigned qm, int lane_mask) {
#pragma unroll
  for (int i = 0; i < 6; i++) { x.I.A[i] += __shfl_xor_sync(qm, x.I.A[i], lane_mask); x.I.C[i] += __shfl_xor_sync(qm, x.I.C[i], lane_mask); }
#pragma unroll
  for (int i = 0; i < 9; i++) x.I.B[i] += __shfl_xor_sync(qm, x.I.B[i], lane_mask);
  x.p.a.x += __shfl_xor_sync(qm, x.p.a.x, lane_mask); x.p.a.y += __shfl_xor_sync(qm, x.p.a.y, lane_mask);
  x.p.a.z += __shfl_xor_sync(qm, x.p.a.z, lane_mask); x.p.l.x += __shfl_xor_sync(qm, x.p.l.x, lane_mask);
  x.p.l.y += __shfl_xor_sync(qm, x.p.l.y, lane_mask); x.p.l.z += __shfl_xor_sync(qm, x.p.l.z, lane_mask);
}
__device__ __forceinline__ IP ip_shfl_get(const IP& x, unsigned qm, int lane_mask) {
  IP y;
#pragma unroll
  for (int i = 0; i < 6; i++) { y.I.A[i] = __shfl_xor_sync(qm, x.I.A[i], lane_mask); y.I.C[i] = __shfl_xor_sync(qm, x.I.C[i], lane_mask); }
#pragma unroll
  for (int i = 0; i < 9; i++) y.I.B[i] = __shfl_xor_sync(qm, x.I.B[i], lane_mask);
  y.p.a.x = __shfl_xor_sync(qm, x.p.a.x, lane_mask); y.p.a.y = __shfl_xor_sync(qm, x.p.a.y, lane_mask);
  y.p.a.z = __shfl_xor_sync(qm, x.p.a.z, lane_mask); y.p.l.x = __shfl_xor_sync(qm, x.p.l.x, lane_mask);
  y.p.l.y = __shfl_xor_sync(qm, x.p.l.y, lane_mask); y.p.l.z = __shfl_xor_sync(qm, x.p.l.z, lane_mask);
  return y;
}

// Cholesky factor of the base inertia, stored per env in shared memory (diagonal entries hold 1 / L_jj)
__device__ __forceinline__ void chol6_to_smem(const Inertia& I, float* L /* [21][QE], this env's column */) {
  float Lr[21];
  chol6(I, Lr);
#pragma unroll
  for (int i = 0; i < 21; i++) L[i * QE] = Lr[i];
}
__device__ __forceinline__ SV chol6_solve_smem(const float* L, SV b) {
  float y[6] = {b.a.x, b.a.y, b.a.z, b.l.x, b.l.y, b.l.z};
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float t = y[i];
#pragma unroll
    for (int c = 0; c < i; c++) t -= L[(i * (i + 1) / 2 + c) * QE] * y[c];
    y[i] = t * L[(i * (i + 1) / 2 + i) * QE];
  }
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    float t = y[i];
#pragma unroll
    for (int c = i + 1; c < 6; c++) t -= L[(c * (c + 1) / 2 + i) * QE] * y[c];
    y[i] = t * L[(i * (i + 1) / 2 + i) * QE];
  }
  SV r; r.a = mk(y[0], y[1], y[2]); r.l = mk(y[3], y[4], y[5]);
  return r;
}

// ---- phase A: forward kinematics, link velocities, velocity-product accelerations, body records, contact candidates.
// Returns the candidate mask of the spheres this lane saw, the sums of part origins (spine part replicated in ssx/ssy,
// limb part in sx/sy) and the origin of the lane's end body.
struct FkOut { uint32_t act; float sx, sy, ssx, ssy, ex, ey; };

// FULL = false: pose only (part-origin sums and the end-body origin), nothing is written to shared memory.
template <bool FULL, class SM>
__device__ __forceinline__ void fk_phase(const Base& b, SM& sm, int e, int tid, int role, FkOut& o) {
  const Tables& T = tables(sm);
  float R0[9];
  quat2mat(b.quat[0], b.quat[1], b.quat[2], b.quat[3], R0);
  SV V0; V0.a = rd3(b.w); V0.l = rd3(b.v);
  uint32_t act = 0;
  // torso spheres
  // (a body whose origin is higher than its reach + the breaking distance cannot have a candidate: skip its spheres)
  if (FULL && b.p[2] - T.torso_reach < (float)ILRL_CONTACT_BREAK) {
#pragma unroll 1
    for (int g = NS - 5; g < NS; g++) {
      V3 c = mv(R0, mk(kSphereC[3 * g], kSphereC[3 * g + 1], kSphereC[3 * g + 2]));
      const float rad = kSphereR[g], d = b.p[2] + (c.z - rad);   // same association as where it is recomputed
      if (d < (float)ILRL_CONTACT_BREAK) {
        act |= 1u << g;
        float* sp = &sm.sph[g][0][e];
        sp[0] = c.x; sp[QE] = c.y; sp[2 * QE] = c.z - rad;
      }
    }
  }
  float Rc[9];
#pragma unroll
  for (int i = 0; i < 9; i++) Rc[i] = R0[i];
  V3 oc = mk(0.f, 0.f, 0.f);
  SV Vp = V0;
  float sx = 0.f, sy = 0.f, ssx = 0.f, ssy = 0.f, ex = 0.f, ey = 0.f;
#pragma unroll 1
  for (int c = 0; c < NL; c++) {
    if (c == 3 && role >= 2) {  // the arms hang off the torso
#pragma unroll
      for (int i = 0; i < 9; i++) Rc[i] = R0[i];
      oc = mk(0.f, 0.f, 0.f);
      Vp = V0;
    }
    const LinkC& L = c < 3 ? T.lc[4][c] : T.lc[role][c - 3];
    float* rec = link_rec(sm, c, e, tid);
    oc = oc + mv(Rc, rd3(L.pre));
    if (L.rot) {
      float Rn[9];
      mm(Rc, T.Q, Rn);
#pragma unroll
      for (int i = 0; i < 9; i++) Rc[i] = Rn[i];
    }
    const V3 an = rd3(L.an), ax = rd3(L.ax);
    const V3 rw = oc + mv(Rc, an);
    SV S;
    S.a = mv(Rc, ax);
    S.l = cross(rw, S.a);
    {
      float sn, cs;
      sincos_joint(rec[W_Q], sn, cs);
      const float t = 1.f - cs;
      float Rj[9], Rn[9];
      Rj[0] = t * ax.x * ax.x + cs;        Rj[1] = t * ax.x * ax.y - sn * ax.z; Rj[2] = t * ax.x * ax.z + sn * ax.y;
      Rj[3] = t * ax.x * ax.y + sn * ax.z; Rj[4] = t * ax.y * ax.y + cs;        Rj[5] = t * ax.y * ax.z - sn * ax.x;
      Rj[6] = t * ax.x * ax.z - sn * ax.y; Rj[7] = t * ax.y * ax.z + sn * ax.x; Rj[8] = t * ax.z * ax.z + cs;
      mm(Rc, Rj, Rn);
      oc = rw - mv(Rn, an);
#pragma unroll
      for (int i = 0; i < 9; i++) Rc[i] = Rn[i];
    }
    if (FULL) {
      SV X = rec[W_QD] * S;
      st_ScJ(rec, S, crm(Vp, X));
      Vp = Vp + X;
    }
    if (c < 3) { ssx += rw.x; ssy += rw.y; }
    else if (L.j >= 0) { sx += rw.x; sy += rw.y; }
    const int nb = L.nbody;
    if (nb > 0) {
      float rcd[RECW];
#pragma unroll
      for (int i = 0; i < RECW; i++) rcd[i] = 0.f;
      const int slot = (c == 2 || c == 6) ? 1 : 0;
#pragma unroll 1
      for (int k = 0; k < nb; k++) {
        const BodyC& B = T.bc[c < 3 ? 4 : role][slot][k];
        const V3 ob = oc + mv(Rc, rd3(B.off));
        if (FULL) rigid_rec(Rc, ob, B.m, B.ix, B.iy, B.iz, Vp, rcd);
        if (c < 3) { ssx += ob.x; ssy += ob.y; } else { sx += ob.x; sy += ob.y; }
        ex = ob.x; ey = ob.y;
        if (FULL && b.p[2] + ob.z - B.reach < (float)ILRL_CONTACT_BREAK) {
#pragma unroll 1
          for (int t = 0; t < B.nsph; t++) {
            const int g = B.sidx[t];
            V3 cs_ = ob + mv(Rc, rd3(B.sph[t]));
            const float rad = B.sph[t][3], d = b.p[2] + (cs_.z - rad);
            if (d < (float)ILRL_CONTACT_BREAK) {
              act |= 1u << g;
              float* sp = &sm.sph[g][0][e];
              sp[0] = cs_.x; sp[QE] = cs_.y; sp[2 * QE] = cs_.z - rad;
            }
          }
        }
      }
      if (FULL) {
        float4* br = reinterpret_cast<float4*>(c < 3 ? sm.rs(e, slot) : sm.rl(e, role, slot));
#pragma unroll
        for (int i = 0; i < RECW / 4; i++) br[i] = make_float4(rcd[4 * i], rcd[4 * i + 1], rcd[4 * i + 2], rcd[4 * i + 3]);
      }
    }
  }
  o.act = act; o.sx = sx; o.sy = sy; o.ssx = ssx; o.ssy = ssy; o.ex = ex; o.ey = ey;
}

// sums of the 31 part offsets (relative to the torso) and the right-foot origin, for calc_state / resetFromFrame
template <class SM>
__device__ __forceinline__ void pose_sums(const Base& b, SM& sm, int e, int tid, int role, unsigned qm, float& sumx,
                                          float& sumy, float& rfx, float& rfy) {
  FkOut o;
  __syncwarp(qm);
  fk_phase<false>(b, sm, e, tid, role, o);
  sumx = o.ssx + qsum(o.sx, qm);
  sumy = o.ssy + qsum(o.sy, qm);
  const int l0 = (tid & 31) & ~3;
  rfx = __shfl_sync(qm, o.ex, l0);
  rfy = __shfl_sync(qm, o.ey, l0);
  __syncwarp(qm);
}

// ---- constraint-row construction: responses of the generalized velocities to up to three unit impulses
struct Imp {
  int L, c;      // chain that carries the impulse: limb (-1 = spine / torso) and chain index (-1 = torso)
  int jl;        // 1: generalized impulse `dir` on the joint of (L, c); 0: spatial force F on the body after (L, c)
  float dir;
  SV F;
  float* row;    // where the row is stored (null: slot unused)
};

// inward walk of one impulse from its link to the base.  Leaves u of the visited links in su (chain-local), writes the
// row's Jacobian chain entries, returns the force arriving at the base and rv = J . nu (chain part).
template <class SM>
__device__ __forceinline__ SV walk_in(const SM& sm, const Imp& im, float* su, int e, int qb, float& rv) {
  SV pf = svzero();
  rv = 0.f;
  if (!im.row) return pf;
  int c = im.c;
  if (im.jl) {
    const float* rec = link_rec_of(sm, im.L, c, e, qb);
    SV S, U;
    float di;
    ld_SU(rec, S, U, di);
    su[c] = im.dir;
    pf = (im.dir * di) * U;
    rv = im.dir * rec[W_NU];
    im.row[c < 3 ? R_JS + c : R_JL + c - 3] = im.dir;
    c--;
  } else {
    pf = neg(im.F);
  }
#pragma unroll 1
  for (; c >= 0; c--) {
    if (c == 2 && im.L >= 2) break;  // arms attach to the torso
    const float* rec = link_rec_of(sm, im.L, c, e, qb);
    SV S, U;
    float di;
    ld_SU(rec, S, U, di);
    const float u = -sdot(S, pf);
    su[c] = u;
    pf = pf + (u * di) * U;
    if (!im.jl) {
      const float Jl = sdot(S, im.F);
      rv += Jl * rec[W_NU];
      im.row[c < 3 ? R_JS + c : R_JL + c - 3] = Jl;
    }
  }
  return pf;
}
// J . resp over the chain entries of a finished row, and restore the scratch to zero
__device__ __forceinline__ float walk_dd(const Imp& im, float* su) {
  float dd = 0.f;
  if (!im.row) return dd;
#pragma unroll 1
  for (int c = im.c; c >= 0; c--) {
    if (c == 2 && im.L >= 2) break;
    su[c] = 0.f;
    dd += c < 3 ? im.row[R_JS + c] * im.row[R_RS + c] : im.row[R_JL + c - 3] * im.row[R_RL + 4 * im.L + c - 3];
  }
  return dd;
}

// the three impulses of ONE contact share their chain: one walk, link records loaded once, 3-way ILP
template <class SM>
__device__ __forceinline__ void walk_in3(const SM& sm, const Imp* im, float* su0, float* su1, float* su2, int e, int qb,
                                         SV* pf, float* rv) {
#pragma unroll
  for (int i = 0; i < 3; i++) { pf[i] = neg(im[i].F); rv[i] = 0.f; }
  const int L = im[0].L;
#pragma unroll 1
  for (int c = im[0].c; c >= 0; c--) {
    if (c == 2 && L >= 2) break;  // arms attach to the torso
    const float* rec = link_rec_of(sm, L, c, e, qb);
    SV S, U;
    float di;
    ld_SU(rec, S, U, di);
    const float nu = rec[W_NU];
    const int jw = c < 3 ? R_JS + c : R_JL + c - 3;
    const float u0 = -sdot(S, pf[0]), u1 = -sdot(S, pf[1]), u2 = -sdot(S, pf[2]);
    const float J0 = sdot(S, im[0].F), J1 = sdot(S, im[1].F), J2 = sdot(S, im[2].F);
    su0[c] = u0; su1[c] = u1; su2[c] = u2;
    pf[0] = pf[0] + (u0 * di) * U; pf[1] = pf[1] + (u1 * di) * U; pf[2] = pf[2] + (u2 * di) * U;
    rv[0] += J0 * nu; rv[1] += J1 * nu; rv[2] += J2 * nu;
    im[0].row[jw] = J0; im[1].row[jw] = J1; im[2].row[jw] = J2;
  }
}

// contact = true: im[0..2] are the normal and the two friction directions of one contact (all rows used, same chain)
// e, qb: the env whose rows are built (any env of this warp); su_e, su_lane: the scratch slot of the EXECUTING lane
template <class SM>
__device__ __forceinline__ void responses3(SM& sm, Imp* im, const float* nub, int e, int su_e, int su_lane, int qb, float idt,
                                           const float* pos /* [3] position term of each row */, bool contact) {
  float* su0 = sm.su(su_e, su_lane, 0);
  float* su1 = sm.su(su_e, su_lane, 1);
  float* su2 = sm.su(su_e, su_lane, 2);
  // clear the rows
#pragma unroll
  for (int i = 0; i < 3; i++)
    if (im[i].row) {
      float4* r4 = reinterpret_cast<float4*>(im[i].row);
#pragma unroll 1
      for (int t = 0; t < RW / 4; t++) r4[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  float rv[3];
  SV pf[3];
  if (contact) walk_in3(sm, im, su0, su1, su2, e, qb, pf, rv);
  else {
    pf[0] = walk_in(sm, im[0], su0, e, qb, rv[0]);
    pf[1] = walk_in(sm, im[1], su1, e, qb, rv[1]);
    pf[2] = walk_in(sm, im[2], su2, e, qb, rv[2]);
  }
  // base part: J_base = F (contact rows), response of the base
  SV ap[3], a0[3], apel[3];
  float dd[3];
  const float* L0 = &sm.L0[0][e];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    a0[i] = chol6_solve_smem(L0, neg(pf[i]));
    ap[i] = a0[i];
    dd[i] = 0.f;
    if (im[i].row) {
      float* r = im[i].row;
      r[R_RB + 0] = a0[i].a.x; r[R_RB + 1] = a0[i].a.y; r[R_RB + 2] = a0[i].a.z;
      r[R_RB + 3] = a0[i].l.x; r[R_RB + 4] = a0[i].l.y; r[R_RB + 5] = a0[i].l.z;
      if (!im[i].jl) {
        const SV F = im[i].F;
        r[R_JB + 0] = F.a.x; r[R_JB + 1] = F.a.y; r[R_JB + 2] = F.a.z; r[R_JB + 3] = F.l.x; r[R_JB + 4] = F.l.y; r[R_JB + 5] = F.l.z;
        rv[i] += F.a.x * nub[0] + F.a.y * nub[1] + F.a.z * nub[2] + F.l.x * nub[3] + F.l.y * nub[4] + F.l.z * nub[5];
        dd[i] = sdot(a0[i], F);
      }
    }
  }
  // outward sweep: spine, then the four limbs
#pragma unroll 1
  for (int c = 0; c < 3; c++) {
    SV S, U;
    float di;
    ld_SU(&sm.sp[c][e][0], S, U, di);
    const float u0 = su0[c], u1 = su1[c], u2 = su2[c];
    const float q0 = di * (u0 - sdot(ap[0], U)), q1 = di * (u1 - sdot(ap[1], U)), q2 = di * (u2 - sdot(ap[2], U));
    ap[0] = ap[0] + q0 * S; ap[1] = ap[1] + q1 * S; ap[2] = ap[2] + q2 * S;
    if (im[0].row) im[0].row[R_RS + c] = q0;
    if (im[1].row) im[1].row[R_RS + c] = q1;
    if (im[2].row) im[2].row[R_RS + c] = q2;
  }
#pragma unroll
  for (int i = 0; i < 3; i++) apel[i] = ap[i];
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 3; i++) ap[i] = r < 2 ? apel[i] : a0[i];
    const bool m0 = im[0].L == r, m1 = im[1].L == r, m2 = im[2].L == r;
#pragma unroll 1
    for (int k = r < 2 ? 0 : 1; k < 4; k++) {  // the arms' leading slot is a dummy: its response stays 0
      SV S, U;
      float di;
      ld_SU(&sm.lk[k][qb + r][0], S, U, di);
      const float u0 = m0 ? su0[3 + k] : 0.f, u1 = m1 ? su1[3 + k] : 0.f, u2 = m2 ? su2[3 + k] : 0.f;
      const float q0 = di * (u0 - sdot(ap[0], U)), q1 = di * (u1 - sdot(ap[1], U)), q2 = di * (u2 - sdot(ap[2], U));
      ap[0] = ap[0] + q0 * S; ap[1] = ap[1] + q1 * S; ap[2] = ap[2] + q2 * S;
      if (im[0].row) im[0].row[R_RL + 4 * r + k] = q0;
      if (im[1].row) im[1].row[R_RL + 4 * r + k] = q1;
      if (im[2].row) im[2].row[R_RL + 4 * r + k] = q2;
    }
  }
  dd[0] += walk_dd(im[0], su0);
  dd[1] += walk_dd(im[1], su1);
  dd[2] += walk_dd(im[2], su2);
#pragma unroll
  for (int i = 0; i < 3; i++)
    if (im[i].row) {
      const float di = __frcp_rn(dd[i]);
      im[i].row[R_DINV] = di;
      im[i].row[R_RHS] = (pos[i] - rv[i]) * di;
    }
}

// ---- projected Gauss-Seidel pieces.  dvb: base (replicated), dvc: chain (spine replicated, limb private)
struct RowRegs { float4 a, b, c, d, e, rl, jl; };  // 20 shared words, own limb response, the row's limb Jacobian
template <class P4>
__device__ __forceinline__ void row_load(P4 rp, int role, RowRegs& r) {
  r.a = rp[0]; r.b = rp[1]; r.c = rp[2]; r.d = rp[3]; r.e = rp[4];
  r.rl = rp[5 + role];
  r.jl = rp[9];
}
__device__ __forceinline__ float row_jdot(const RowRegs& r, const float* dvb, const float* dvc, bool mine, int src, unsigned qm) {
  // words: a = RB0..3, b = RB4,5 RS0,1, c = RS2 JS0..2, d = JB0..3, e = JB4,5 rhs dinv
  // three short chains instead of one 9-deep one: this dot product sits on the Gauss-Seidel critical path
  const float r0 = fmaf(r.d.z, dvb[2], fmaf(r.d.y, dvb[1], r.d.x * dvb[0]));
  const float r1 = fmaf(r.e.y, dvb[5], fmaf(r.e.x, dvb[4], r.d.w * dvb[3]));
  const float r2 = fmaf(r.c.w, dvc[2], fmaf(r.c.z, dvc[1], r.c.y * dvc[0]));
  const float own = mine ? fmaf(r.jl.y, dvc[4], r.jl.x * dvc[3]) + fmaf(r.jl.w, dvc[6], r.jl.z * dvc[5]) : 0.f;
  return (r0 + r1) + (r2 + __shfl_sync(qm, own, src));
}
__device__ __forceinline__ void row_axpy(const RowRegs& r, float a, float* dvb, float* dvc) {
  dvb[0] += a * r.a.x; dvb[1] += a * r.a.y; dvb[2] += a * r.a.z; dvb[3] += a * r.a.w; dvb[4] += a * r.b.x; dvb[5] += a * r.b.y;
  dvc[0] += a * r.b.z; dvc[1] += a * r.b.w; dvc[2] += a * r.c.x;
  dvc[3] += a * r.rl.x; dvc[4] += a * r.rl.y; dvc[5] += a * r.rl.z; dvc[6] += a * r.rl.w;
}
// OVER = false: the env has no row beyond the shared-memory budget (the common case): plain shared-memory loads, no
// predicated-off global loads in the instruction stream (those would take issue slots in every row evaluation).
template <bool OVER, class SM>
__device__ __forceinline__ void row_fetch(const SM& sm, const float* gscr, int e, int r, int role, RowRegs& rr) {
  if (!OVER || r < SM::RSM) row_load(reinterpret_cast<const float4*>(&sm.rows[e][r * RW]), role, rr);
  else row_load(reinterpret_cast<const float4*>(gscr + (size_t)(r - SM::RSM) * RW), role, rr);
}

// 5 projected-Gauss-Seidel sweeps on the velocity change (dvb: base, dvc: chain) of the env of this quad
template <bool OVER, class SM>
__device__ __forceinline__ void pgs_sweeps(SM& sm, const float* gscr, int e, int role, int qb, unsigned qm, int nlim,
                                           int ncon, float* dvb, float* dvc) {
#pragma unroll 1
  for (int itn = 0; itn < ILRL_SOLVER_ITERS; itn++) {
    // limits, then contact normals.  The next row (independent of dv) is fetched while this one is applied:
    // two register buffers used alternately (the loop is unrolled by two so that no copies are needed).
    const int nfirst = nlim + ncon;
    auto row_of = [&](int k) { return k < nlim ? k : nlim + 3 * (k - nlim); };
    auto fetch = [&](int k, RowRegs& rr, int& L, float& lam) {
      const int r = row_of(k);
      row_fetch<OVER>(sm, gscr, e, r, role, rr);
      L = sm.rowL(e)[r]; lam = sm.lam(e)[r];
    };
    auto apply = [&](int k, const RowRegs& rr, int L, float lam) {
      const float nl = fmaxf(lam + rr.e.z - row_jdot(rr, dvb, dvc, role == L, qb + (L & 3), qm) * rr.e.w, 0.f);
      sm.lam(e)[row_of(k)] = nl;
      row_axpy(rr, nl - lam, dvb, dvc);
    };
    RowRegs ra, rb;
    int La = 0, Lb = 0;
    float lama = 0.f, lamb = 0.f;
    fetch(0, ra, La, lama);  // nrows > 0 implies nfirst > 0
#pragma unroll 1
    for (int k = 0; k < nfirst; k += 2) {
      if (k + 1 < nfirst) fetch(k + 1, rb, Lb, lamb);
      apply(k, ra, La, lama);
      if (k + 1 < nfirst) {
        if (k + 2 < nfirst) fetch(k + 2, ra, La, lama);
        apply(k + 1, rb, Lb, lamb);
      }
    }
#pragma unroll 1
    for (int c = 0; c < ncon; c++) {  // friction pairs, cone re-projected on the current normal impulse
      const int rn = nlim + 3 * c;
      const float ln = sm.lam(e)[rn];
      if (!(ln > 0.f)) continue;
      RowRegs r1, r2;
      row_fetch<OVER>(sm, gscr, e, rn + 1, role, r1);
      row_fetch<OVER>(sm, gscr, e, rn + 2, role, r2);
      const int L = sm.rowL(e)[rn];
      const float lim_f = (float)ILRL_FRICTION * ln;
      const float l1 = sm.lam(e)[rn + 1], l2 = sm.lam(e)[rn + 2];
      float s1 = l1 + r1.e.z - row_jdot(r1, dvb, dvc, role == L, qb + (L & 3), qm) * r1.e.w;
      float s2 = l2 + r2.e.z - row_jdot(r2, dvb, dvc, role == L, qb + (L & 3), qm) * r2.e.w;
      const float n2 = s1 * s1 + s2 * s2;
      if (n2 > lim_f * lim_f) { const float sc = lim_f * rsqrtf(n2); s1 *= sc; s2 *= sc; }
      sm.lam(e)[rn + 1] = s1; sm.lam(e)[rn + 2] = s2;
      row_axpy(r1, s1 - l1, dvb, dvc);
      row_axpy(r2, s2 - l2, dvb, dvc);
    }
  }
}

// ---- one substep of dt for the env of this quad.  Joint state / torques live in the link records.
template <class SM>
__device__ __forceinline__ void substep(Base& b, SM& sm, float* gscr, int e, int tid, int role, unsigned qm, unsigned wm,
                                        float dt) {
  constexpr int RSM = SM::RSM;
  const int qb = tid & ~3;
  const Tables& T = tables(sm);
  // ---- phase A
  FkOut fo;
  fk_phase<true>(b, sm, e, tid, role, fo);
  uint32_t act = fo.act, lim = 0;
  // ---- phase B: inward pass
  SV a0;
  {
    IP x, arms;
    x.I.zero(); x.p = svzero();
    arms.I.zero(); arms.p = svzero();
#pragma unroll 1
    for (int c = NL - 1; c >= 0; c--) {
      if (c == 2) {
        // the limbs are done: legs (lanes 0,1) meet at the pelvis, arms (lanes 2,3) at the torso
        ip_shfl_add(x, qm, 1);
        IP y = ip_shfl_get(x, qm, 2);
        if (role < 2) arms = y; else { arms = x; x = y; }
      }
      float* rec = link_rec(sm, c, e, tid);
      if (c == 1 || c == 2 || c == 5 || c == 6) {
        const int slot = (c == 2 || c == 6) ? 1 : 0;
        ip_add_rec4(x, c < 3 ? sm.rs(e, slot) : sm.rl(e, role, slot));
      }
      SV S, cJ;
      ld_ScJ(rec, S, cJ);
      const SV U = imul(x.I, S);
      const float dinv = rcp_or_zero(sdot(S, U));  // a dummy slot has S = 0
      const float u = rec[W_TAU] - sdot(S, x.p);
      st_Udu(rec, U, dinv, u);
      downdate(x.I, U, dinv);
      x.p = x.p + imul(x.I, cJ) + (u * dinv) * U;
    }
    // floating base: torso + spine/legs + arms
    float rcd[RECW];
#pragma unroll
    for (int i = 0; i < RECW; i++) rcd[i] = 0.f;
    float R0[9];
    quat2mat(b.quat[0], b.quat[1], b.quat[2], b.quat[3], R0);
    SV V0; V0.a = rd3(b.w); V0.l = rd3(b.v);
    rigid_rec(R0, mk(0.f, 0.f, 0.f), kBodyMass[0], kBodyInertia[0], kBodyInertia[1], kBodyInertia[2], V0, rcd);
    ip_add_rec(x, rcd);
    x.I.add(arms.I); x.p = x.p + arms.p;
    chol6_to_smem(x.I, &sm.L0[0][e]);
    __syncwarp(qm);
    a0 = chol6_solve_smem(&sm.L0[0][e], neg(x.p));
  }
  // ---- phase C: outward pass -> unconstrained new velocities; violated limits
  float nub[6];
  {
    const float M = (float)ILRL_MAX_COORD_VEL;
    V3 w = rd3(b.w), v = rd3(b.v);
    V3 lin = a0.l + cross(w, v);  // classical acceleration of the torso origin
    nub[0] = clampf(b.w[0] + dt * a0.a.x, -M, M); nub[1] = clampf(b.w[1] + dt * a0.a.y, -M, M);
    nub[2] = clampf(b.w[2] + dt * a0.a.z, -M, M); nub[3] = clampf(b.v[0] + dt * lin.x, -M, M);
    nub[4] = clampf(b.v[1] + dt * lin.y, -M, M);  nub[5] = clampf(b.v[2] + dt * lin.z, -M, M);
    SV ap = a0;
#pragma unroll 1
    for (int c = 0; c < NL; c++) {
      if (c == 3 && role >= 2) ap = a0;
      float* rec = link_rec(sm, c, e, tid);
      const LinkC& L = c < 3 ? T.lc[4][c] : T.lc[role][c - 3];
      SV S, cJ, U;
      float di, uu;
      ld_all(rec, S, cJ, U, di, uu);
      const float4 st4 = reinterpret_cast<const float4*>(rec)[5];  // q, qd, tau, nu
      const SV ad = ap + cJ;
      const float qa = di * (uu - sdot(ad, U));
      rec[W_NU] = clampf(st4.y + dt * qa, -M, M);
      ap = ad + qa * S;
      const float q = st4.x;
      if (L.j >= 0 && (q - L.lo <= 0.f || L.hi - q <= 0.f)) lim |= 1u << L.j;
    }
  }
  lim |= __shfl_xor_sync(qm, lim, 1); lim |= __shfl_xor_sync(qm, lim, 2);
  act |= __shfl_xor_sync(qm, act, 1); act |= __shfl_xor_sync(qm, act, 2);
  __syncwarp(qm);
  int nact = __popc(act);
  while (nact > MAXC) {  // keep the deepest MAXC (ties: drop the highest index); replicated in the quad
    int worst = -1;
    float wd = -1e30f;
#pragma unroll 1
    for (int g = 0; g < NS; g++)
      if ((act >> g) & 1u) {
        const float d = b.p[2] + sm.sph[g][2][e];
        if (d >= wd) { wd = d; worst = g; }
      }
    act &= ~(1u << worst);
    nact--;
  }
  const int nlim = __popc(lim), ncon = nact, nrows = nlim + 3 * ncon;
  float dvb[6], dvc[NL];
#pragma unroll
  for (int i = 0; i < 6; i++) dvb[i] = 0.f;
#pragma unroll
  for (int i = 0; i < NL; i++) dvc[i] = 0.f;
  // ---- build the rows three at a time (one contact, or up to three limits, per item).  The items of the warp's 8 envs
  // form ONE pool that is dealt to all its lanes: an env with many rows is helped by the lanes of envs with few (every
  // input of an item is in shared memory or one shuffle away), so a warp needs ceil(items / lanes) rounds of the
  // builder instead of max over its envs of ceil(items / 4).  wm = lanes of this warp that execute the substep.
  {
    const float idt = 1.0f / dt;
    const int lane = tid & 31, q_own = lane >> 2;
    const int nlg_own = (nlim + 2) / 3, nit_own = nlg_own + ncon;
    int cnt[8], total = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int c = __shfl_sync(wm, nit_own, 4 * q);
      cnt[q] = ((wm >> (4 * q)) & 1u) ? c : 0;   // a quad that does not step has no items (its lanes are not here)
      total += cnt[q];
    }
    if (total > 0) {
      {  // the response scratch shares its storage with the body records of the FK phase: clear this lane's part
        float* z = sm.su(e, role, 0);
#pragma unroll
        for (int i = 0; i < 3 * NL; i++) z[i] = 0.f;
      }
      const int nlanes = __popc(wm), rank = __popc(wm & ((1u << lane) - 1u));
#pragma unroll 1
      for (int g0 = 0; g0 < total; g0 += nlanes) {
        // item g -> (quad q, item k of that quad's env)
        int k = g0 + rank, q = 0;
        const bool has = k < total;
#pragma unroll
        for (int t = 0; t < 7; t++)
          if (q == t && k >= cnt[t]) { k -= cnt[t]; q = t + 1; }
        if (!has) { q = q_own; k = 0; }
        const int src = 4 * q;
        const uint32_t lim_q = __shfl_sync(wm, lim, src), act_q = __shfl_sync(wm, act, src);
        const int nlim_q = __shfl_sync(wm, nlim, src);
        const float bz_q = __shfl_sync(wm, b.p[2], src);
        float nub_q[6];
#pragma unroll
        for (int i = 0; i < 6; i++) nub_q[i] = __shfl_sync(wm, nub[i], src);
        if (has) {
          const int e_q = (e & ~7) + q, qb_q = (tid & ~31) + 4 * q, nlg_q = (nlim_q + 2) / 3;
          float* gscr_q = gscr + (ptrdiff_t)(q - q_own) * (GROWS * RW);
          const int it = k;
          Imp im[3];
          float pos[3] = {0.f, 0.f, 0.f};
          int r0;
          if (it < nlg_q) {
            r0 = 3 * it;
#pragma unroll
            for (int i = 0; i < 3; i++) {
              im[i].row = nullptr; im[i].L = -1; im[i].c = -1; im[i].jl = 1; im[i].dir = 0.f; im[i].F = svzero();
              if (r0 + i < nlim_q) {
                const int j = nth_set_bit(lim_q, r0 + i);
                im[i].L = T.jL[j]; im[i].c = T.jC[j];
                const float q_ = link_rec_of(sm, im[i].L, im[i].c, e_q, qb_q)[W_Q];
                float pen;
                if (q_ - kJointLo[j] <= 0.f) { pen = q_ - kJointLo[j]; im[i].dir = 1.f; } else { pen = kJointHi[j] - q_; im[i].dir = -1.f; }
                pos[i] = -pen * (float)ILRL_LIMIT_ERP * idt;
              }
            }
          } else {
            const int ci = it - nlg_q;
            r0 = nlim_q + 3 * ci;
            const int g = nth_set_bit(act_q, ci);
            const float* sp = &sm.sph[g][0][e_q];
            const V3 xx = mk(sp[0], sp[QE], sp[2 * QE]);
            const float dist = bz_q + sp[2 * QE];
            pos[0] = dist > 0.f ? -dist * idt : -dist * (float)ILRL_CONTACT_ERP * idt;
#pragma unroll
            for (int i = 0; i < 3; i++) {
              im[i].L = T.sphL[g]; im[i].c = T.sphC[g]; im[i].jl = 0; im[i].dir = 0.f;
              // normal (0,0,1), tangents btPlaneSpace1 -> (0,-1,0), (1,0,0)
              im[i].F.l = i == 0 ? mk(0.f, 0.f, 1.f) : (i == 1 ? mk(0.f, -1.f, 0.f) : mk(1.f, 0.f, 0.f));
              im[i].F.a = cross(xx, im[i].F.l);
            }
          }
#pragma unroll
          for (int i = 0; i < 3; i++) {
            const int r = r0 + i;
            const bool used = it >= nlg_q || r < nlim_q;
            im[i].row = !used ? nullptr : (r < RSM ? &sm.rows[e_q][r * RW] : gscr_q + (size_t)(r - RSM) * RW);
            if (used) { sm.lam(e_q)[r] = 0.f; sm.rowL(e_q)[r] = (signed char)(im[i].c >= 3 ? im[i].L : -1); }
          }
          responses3(sm, im, nub_q, e_q, e, role, qb_q, idt, pos, it >= nlg_q);
        }
      }
      __syncwarp(wm);   // rows of an env may have been written by lanes of another quad
    }
  }
  if (nrows > 0) {
    // ---- projected Gauss-Seidel on the velocity change
    // One loop variant per layout wherever both would be common: the quads of a warp that took different variants
    // run them one after the other (+4..9 % in the dense layouts from dropping the split).  Only the on-chip layout,
    // where overflow rows are rare, keeps the loop without the overflow test for the envs that fit (+1.5 % there).
    if (RSM >= 16 && nrows <= RSM) pgs_sweeps<false>(sm, gscr, e, role, qb, qm, nlim, ncon, dvb, dvc);
    else pgs_sweeps<true>(sm, gscr, e, role, qb, qm, nlim, ncon, dvb, dvc);
    __syncwarp(qm);
  }
  // ---- integrate (exponential map on the torso quaternion, as btMultiBody::stepPositionsMultiDof)
  {
    const float M = (float)ILRL_MAX_COORD_VEL;
    float nu[6];
#pragma unroll
    for (int i = 0; i < 6; i++) nu[i] = nrows > 0 ? clampf(nub[i] + dvb[i], -M, M) : nub[i];
#pragma unroll
    for (int i = 0; i < 3; i++) { b.w[i] = nu[i]; b.v[i] = nu[3 + i]; b.p[i] += dt * nu[3 + i]; }
    const float wn = sqrtf(nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]);
    float sc, cw, sh;
    sincos_joint(0.5f * wn * dt, sh, cw);
    if (wn < 1e-3f) sc = 0.5f * dt - dt * dt * dt * 0.020833333333f * wn * wn;
    else sc = sh / wn;
    const float dx = nu[0] * sc, dy = nu[1] * sc, dz = nu[2] * sc;
    const float x_ = b.quat[0], y_ = b.quat[1], z_ = b.quat[2], w_ = b.quat[3];
    const float nx = cw * x_ + dx * w_ + dy * z_ - dz * y_;
    const float ny = cw * y_ - dx * z_ + dy * w_ + dz * x_;
    const float nz = cw * z_ + dx * y_ - dy * x_ + dz * w_;
    const float nw = cw * w_ - dx * x_ - dy * y_ - dz * z_;
    const float inv = rsqrtf(nx * nx + ny * ny + nz * nz + nw * nw);
    b.quat[0] = nx * inv; b.quat[1] = ny * inv; b.quat[2] = nz * inv; b.quat[3] = nw * inv;
#pragma unroll
    for (int c = 0; c < NL; c++) {
      float* rec = link_rec(sm, c, e, tid);
      const float nuc = rec[W_NU];
      const float qd = nrows > 0 ? clampf(nuc + dvc[c], -M, M) : nuc;
      if (c >= 3 || role == 0) {  // the spine records are shared by the quad: one writer for the read-modify-write
        rec[W_QD] = qd;
        rec[W_Q] += dt * qd;
      }
    }
  }
  __syncwarp(qm);
}

// ---- state movement between HBM (SoA phys[47][n]), the link records and a replicated full Phys
__device__ __forceinline__ void load_base(const float* phys, int n, int i, Base& b) {
  const float* p = phys + i;
#pragma unroll
  for (int k = 0; k < 3; k++) { b.p[k] = p[k * n]; b.v[k] = p[(7 + k) * n]; b.w[k] = p[(10 + k) * n]; }
#pragma unroll
  for (int k = 0; k < 4; k++) b.quat[k] = p[(3 + k) * n];
}
// joint state of this lane's chain from HBM into the link records (spine: every lane writes the same values)
template <class SM>
__device__ __forceinline__ void load_links(const float* phys, int n, int i, SM& sm, int e, int tid, int role) {
  const float* p = phys + i;
#pragma unroll
  for (int c = 0; c < NL; c++) {
    float* rec = link_rec(sm, c, e, tid);
    const int j = c < 3 ? c : tables(sm).lc[role][c - 3].j;
    rec[W_Q] = j >= 0 ? p[(13 + j) * n] : 0.f;
    rec[W_QD] = j >= 0 ? p[(30 + j) * n] : 0.f;
  }
}
// link records -> full Phys in every lane
template <class SM>
__device__ __forceinline__ void gather(const Base& b, const SM& sm, int e, int qb, unsigned qm, Phys& ps) {
  __syncwarp(qm);
#pragma unroll
  for (int j = 0; j < NJ; j++) {
    constexpr Tables T{};
    const int L = T.jL[j], c = T.jC[j];
    const float* rec = c < 3 ? &sm.sp[c][e][0] : &sm.lk[c - 3][qb + L][0];
    ps.q[j] = rec[W_Q];
    ps.qd[j] = rec[W_QD];
  }
#pragma unroll
  for (int k = 0; k < 3; k++) { ps.p[k] = b.p[k]; ps.v[k] = b.v[k]; ps.w[k] = b.w[k]; }
#pragma unroll
  for (int k = 0; k < 4; k++) ps.quat[k] = b.quat[k];
}
// full (replicated) Phys -> base registers + link records
template <class SM>
__device__ __forceinline__ void scatter(const Phys& ps, SM& sm, int e, int qb, int role, unsigned qm, Base& b) {
  __syncwarp(qm);
  if (role == 0) {
#pragma unroll
    for (int j = 0; j < NJ; j++) {
      constexpr Tables T{};
      const int L = T.jL[j], c = T.jC[j];
      float* rec = c < 3 ? &sm.sp[c][e][0] : &sm.lk[c - 3][qb + L][0];
      rec[W_Q] = ps.q[j];
      rec[W_QD] = ps.qd[j];
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) { b.p[k] = ps.p[k]; b.v[k] = ps.v[k]; b.w[k] = ps.w[k]; }
#pragma unroll
  for (int k = 0; k < 4; k++) b.quat[k] = ps.quat[k];
  __syncwarp(qm);
}
// store a replicated Phys: the 47 words are dealt to the 4 lanes
__device__ __forceinline__ void store_phys(float* phys, int n, int i, int role, const Phys& ps) {
  float* p = phys + i;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (((0 + k) & 3) == role) p[(0 + k) * n] = ps.p[k];
    if (((7 + k) & 3) == role) p[(7 + k) * n] = ps.v[k];
    if (((10 + k) & 3) == role) p[(10 + k) * n] = ps.w[k];
  }
#pragma unroll
  for (int k = 0; k < 4; k++) if (((3 + k) & 3) == role) p[(3 + k) * n] = ps.quat[k];
#pragma unroll
  for (int k = 0; k < NJ; k++) {
    if (((13 + k) & 3) == role) p[(13 + k) * n] = ps.q[k];
    if (((30 + k) & 3) == role) p[(30 + k) * n] = ps.qd[k];
  }
}

}  // namespace chain
}  // namespace ilrl
