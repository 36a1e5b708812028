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
constexpr int RW = 16;                 // words per stored row
#ifndef ILRL_MIN_RSM
#define ILRL_MIN_RSM 4   // = the row budget of LayoutDense4
#endif
constexpr int MIN_RSM = ILRL_MIN_RSM;  // smallest on-chip row budget of any layout
constexpr int GROWS = MAXROWS - MIN_RSM;  // rows per env in the global overflow scratch
// link record: 25 words (S | cJ | U dinv u | q qd tau nu | sqrt(dinv)), one contiguous 28-word slot per thread: a
// 28-word stride puts the float4 of 8 consecutive threads in 8 different bank groups (conflict-free LDS.128 / STS.128)
enum { W_S = 0, W_CJ = 6, W_U = 12, W_DINV = 18, W_UU = 19, W_Q = 20, W_QD = 21, W_TAU = 22, W_NU = 23, W_SQD = 24, LKW = 28 };
// body record words: rigid inertia about the reference point (A 6, m*c 3, m 1) + bias force 6
constexpr int RECW = 16;
// stored row words (whitened rows, see "constraint rows" below)
enum { R_ZB = 0 /*base 6*/, R_RHS = 6, R_DINV = 7, R_ZS = 8 /*spine 3*/, R_LIMB = 11, R_ZL = 12 /*the row's limb 4*/ };

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
//   LayoutSmall : up to one wave of CTAs at 2 per SM (<= 4736 envs on 148 SMs): everything on chip - every possible
//                 row of an env (41, no overflow path at all), padded (conflict-free) body records, model tables in
//                 shared memory.
//   LayoutLarge : larger batches, where resident warps per SM are what limits throughput: 16 rows per env on chip (the
//                 rest in the L2-resident global scratch), unpadded body records, model tables read through L1.
//   LayoutDense4: 4 rows per env on chip, body-view lane stride 32 (4-way bank conflicts on its 4 float4 accesses per
//                 substep): 4 CTAs per SM.
// In all, the body records (written by the FK phase, consumed by the inward pass) share their storage with what
// only the row phase uses (multipliers) and with the action tile (consumed before the first substep).
constexpr int NSELF = ILRL_NSELF, MAXSELF = 8, WRW = 24;   // self-collision: pairs, kept contacts per env, words per wide row
constexpr int SELF_GP = NS * 3 + 1;                          // sphere centres of an env (+1: env stride 88 = 24 mod 32 words)
constexpr int SELF_WR = 3 * MAXSELF * WRW;                   // wide-row words per env
template <int RSM_, int BRW_, bool TSM_, int LANE_, bool SELF_ = false>
struct Layout {
  static constexpr bool SELF = SELF_;            // self-collision instantiations: sphere centres + candidate masks on chip
  static constexpr int RSM = RSM_;               // constraint rows per env kept in shared memory
  static constexpr int ROWSTRIDE = RSM_ * RW + 4;  // env stride in words: = 4 (mod 32)
  static constexpr int BRW = BRW_;               // words per body record (20: room for padding; 16: dense)
  static constexpr bool TSM = TSM_;              // model tables staged in shared memory
#ifndef ILRL_HOT_TABLES
#define ILRL_HOT_TABLES 1
#endif
  // layouts that leave the tables in global memory still stage the per-link part (what the FK pass reads for every link
  // of every substep, 1.5 KB) when it fits without costing a resident CTA (not next to the self-collision arrays)
  static constexpr bool HSM = !TSM_ && !SELF_ && ILRL_HOT_TABLES;
  static constexpr int LANE = LANE_;             // lane stride of the body view in the per-env scratch block (words)
};
#ifndef ILRL_LARGE_RSM  // (overridable for layout experiments)
#define ILRL_LARGE_RSM 16
#define ILRL_LARGE_BRW 16
#define ILRL_LARGE_TSM false
#endif
static_assert(ILRL_LARGE_RSM >= MIN_RSM, "the overflow scratch holds MAXROWS - MIN_RSM rows per env");
#ifndef ILRL_SMALL_RSM
#define ILRL_SMALL_RSM 42   // >= MAXROWS: no env of this layout ever has an overflow row
#endif
using LayoutSmall = Layout<ILRL_SMALL_RSM, 20, true, 40>;
using LayoutLarge = Layout<ILRL_LARGE_RSM, ILRL_LARGE_BRW, ILRL_LARGE_TSM, 40>;
using LayoutDense4 = Layout<ILRL_MIN_RSM, 16, false, 32>;
using LayoutSelf = Layout<ILRL_LARGE_RSM, ILRL_LARGE_BRW, ILRL_LARGE_TSM, 40, true>;   // LayoutLarge + the self-collision arrays

// Per-env scratch block, seen in two ways that are never live at the same time within an env:
//   body view (FK phase -> inward pass): body records of the four lanes (2 each, lane stride LANE words) + the spine's 2
//   row view  (row phase)              : multipliers, action tile (consumed before the first substep)
// The block is PER ENV because warps of a CTA run unsynchronised: one env's row phase must never touch another env's
// body records.  Env stride = 4 (mod 32) words and lane stride 40 (LayoutSmall / LayoutLarge): the float4 accesses of the 2 envs x 4 lanes of a
// quarter-warp fall in 8 different bank groups, and scalar accesses of the 8 envs of a warp in 8 different banks.
constexpr int SCR_LAM = 0, SCR_ACT = SCR_LAM + MAXROWS;  // row view (words)
constexpr int SCR_ROWVIEW = SCR_ACT + NJ;
template <int BRW, int LANE> constexpr int scr_words() {
  int w = 4 * LANE + 2 * BRW;                      // body view
  if (w < SCR_ROWVIEW) w = SCR_ROWVIEW;
  while (w % 32 != 4) w++;
  return w;
}

struct NoTables {};
struct NoHot {};
struct NoSelf {};
struct alignas(16) HotTables { LinkC lc[5][4]; float Q[9]; float pad[3]; };   // Tables::lc + Tables::Q
constexpr int HOT_LC_WORDS = sizeof(LinkC) * 20 / 4;
struct alignas(16) SelfSm {
  float gp[QE][SELF_GP];      // sphere centres relative to the torso (world axes), published by the FK pass
  uint32_t mask[QE][4];       // candidate pairs (66 bits), set by the pair tests of the warp's pool
  uint4 pairs[NSELF + 2];     // per pair: sphere indices a0 | a1 << 8 | b0 << 16 | b1 << 24, 4 x squared broad-phase reach,
                              // sum of the two radii + the breaking distance, -
  float wr[QE][SELF_WR + 4];  // the wide rows (env stride = 4 mod 32 words): 36 KB per CTA, 2 CTAs per SM
};
template <class LY>
struct __align__(16) SmemT {
  static constexpr int RSM = LY::RSM, ROWSTRIDE = LY::ROWSTRIDE, BRW = LY::BRW, ES = scr_words<LY::BRW, LY::LANE>();
  static constexpr bool TSM = LY::TSM, HSM = LY::HSM;
  static constexpr bool SELF = LY::SELF;
  static constexpr int SCR_LANE = LY::LANE;
  static_assert(2 * BRW <= SCR_LANE, "two body records per lane");
  // --- float4-accessed arrays first (every size below is a multiple of 16 bytes)
  float rows[QE][ROWSTRIDE];   // stored rows
  float lk[4][QT][LKW];        // limb link records, per thread
  float sp[3][QE][LKW];        // spine link records, per env
  float scr[QE][ES];           // per-env scratch block (see above); after the substeps its first 70 words stage the obs row
  // --- scalar-accessed
  float L0[21][QE];            // Cholesky factor of the base articulated inertia
  float sph[NS][3][QE];        // contact candidates: x, y, z - r relative to the torso origin (distance = base z + that)
  typename std::conditional<TSM, Tables, NoTables>::type T;
  typename std::conditional<HSM, HotTables, NoHot>::type H;
  typename std::conditional<LY::SELF, SelfSm, NoSelf>::type S;
  static_assert((sizeof(float) * QE * ROWSTRIDE) % 16 == 0 && (sizeof(float) * LKW) % 16 == 0 && BRW % 4 == 0 && ES % 4 == 0,
                "float4 alignment of the shared-memory records");
  static_assert(ES >= 71, "the obs row is staged in the env's scratch block");
  static_assert(RSM % 2 == 0, "env stride of the rows = 4 (mod 32) words");
  // body view
  __device__ __forceinline__ float* rl(int e, int lane, int slot) { return &scr[e][lane * SCR_LANE + slot * BRW]; }
  __device__ __forceinline__ float* rs(int e, int slot) { return &scr[e][4 * SCR_LANE + slot * BRW]; }
  // row view
  __device__ __forceinline__ float* lam(int e) { return &scr[e][SCR_LAM]; }
  __device__ __forceinline__ float* act(int e) { return &scr[e][SCR_ACT]; }
};
// the model tables as the kernel sees them: the CTA's shared-memory copy, or the global object through L1
template <class SM>
__device__ __forceinline__ const Tables& tables(const SM& sm) {
  if constexpr (SM::TSM) return sm.T; else return kTables;
}

// the per-link constants (record k of chain ch) and the fixed lwaist / pelvis rotation: shared memory in every layout
// but the self-collision one
template <class SM>
__device__ __forceinline__ const LinkC& link_c(const SM& sm, int ch, int k) {
  if constexpr (SM::HSM) return sm.H.lc[ch][k]; else return tables(sm).lc[ch][k];
}
template <class SM>
__device__ __forceinline__ const float* fixed_rot(const SM& sm) {
  if constexpr (SM::HSM) return sm.H.Q; else return tables(sm).Q;
}

constexpr unsigned FULLMASK = 0xffffffffu;

// Phase timer of the measurement build (-DILRL_PROF, tools/warp_profile.py): lane 0 of every warp accumulates clock64()
// intervals per phase.  In the product build it is an empty object and every call vanishes.
enum { PF_HEAD = 0, PF_FK, PF_INWARD, PF_OUTWARD, PF_ROWS, PF_PGS, PF_INTEG, PF_TAIL, PF_BARRIER, PF_TOTAL,
       PF_MAXROWS, PF_T_POSE, PF_T_REWARD, PF_T_OBS, PF_T_STORE, PF_WORDS = 16 };  // (PF_T_*: parts of the tail; PF_TAIL = its rest)
#ifdef ILRL_PROF
struct Prof {
  long long* p = nullptr;
  long long t = 0;
  __device__ __forceinline__ void start() { t = clock64(); }
  __device__ __forceinline__ void mark(int k) { if (p) { const long long n = clock64(); p[k] += n - t; t = n; } }
  __device__ __forceinline__ void maxv(int k, long long v) { if (p && v > p[k]) p[k] = v; }
};
#else
struct Prof {
  __device__ __forceinline__ void start() {}
  __device__ __forceinline__ void mark(int) {}
  __device__ __forceinline__ void maxv(int, long long) {}
};
#endif

// replicated floating-base state of one env
struct Base { float p[3], quat[4], v[3], w[3]; };

// Heightfield terrain of the handle (REF humanoid.py:68-144 CustomScene).  [BULLET, restated]
// btHeightfieldTerrainShape: sample (i, j) at x = i - (rows - 1) / 2, y = j - (cols - 1) / 2, world z = h[i + j * rows]
// + zoff; cells cut along the diagonal (i+1, j) - (i, j+1).  Contact model (declared): a candidate sphere meets the
// plane of the triangle under its centre.  hmax / nzmin (largest world height, smallest normal z of any triangle) make
// the per-body early-outs conservative.  The kernels that take a Terrain are separate instantiations (TERR = true):
// the flat-ground kernels carry none of this.
struct Terrain { const float* h; int rows, cols; float zoff, hmax, nzmin; };
__device__ __forceinline__ float terrain_sample(const Terrain& t, float x, float y, V3& n) {
  const float fx = x + 0.5f * (float)(t.rows - 1), fy = y + 0.5f * (float)(t.cols - 1);
  const int i = min(max((int)floorf(fx), 0), t.rows - 2), j = min(max((int)floorf(fy), 0), t.cols - 2);
  const float u = fminf(fmaxf(fx - (float)i, 0.f), 1.f), v = fminf(fmaxf(fy - (float)j, 0.f), 1.f);
  const float* p = t.h + i + j * t.rows;
  const float h00 = __ldg(p), h10 = __ldg(p + 1), h01 = __ldg(p + t.rows), h11 = __ldg(p + t.rows + 1);
  float gx, gy, h;
  if (u + v <= 1.f) { gx = h10 - h00; gy = h01 - h00; h = h00 + u * gx + v * gy; }
  else { gx = h11 - h01; gy = h11 - h10; h = h11 - (1.f - u) * gx - (1.f - v) * gy; }
  const float il = rsqrtf(1.f + gx * gx + gy * gy);
  n = mk(-gx * il, -gy * il, il);
  return h + t.zoff;
}
// btPlaneSpace1 for a normal with |n.z| > 0.707 (heightfield slopes below 45 degrees; checked when the terrain is set)
__device__ __forceinline__ void plane_space_up(V3 n, V3& p, V3& q) {
  const float a = n.y * n.y + n.z * n.z, k = rsqrtf(a);
  p = mk(0.f, -n.z * k, n.y * k);
  q = mk(a * k, -n.x * p.z, n.x * p.y);
}

__device__ __forceinline__ float qsum(float v, unsigned qm) {
  v += __shfl_xor_sync(qm, v, 1);
  v += __shfl_xor_sync(qm, v, 2);
  return v;
}
__device__ __forceinline__ SV neg(SV a) { SV r; r.a = mk(-a.a.x, -a.a.y, -a.a.z); r.l = mk(-a.l.x, -a.l.y, -a.l.z); return r; }
__device__ __forceinline__ SV svzero() { SV r; r.a = r.l = mk(0, 0, 0); return r; }
__device__ __forceinline__ V3 rd3(const float* p) { return mk(p[0], p[1], p[2]); }
__device__ __forceinline__ V3 rd3g(const float* p) { return mk(__ldcg(p), __ldcg(p + 1), __ldcg(p + 2)); }  // L2 (written by other lanes)
__device__ __forceinline__ float rcp_or_zero(float d) { return d > 1e-9f ? __frcp_rn(d) : 0.f; }

// ---- self-collision (REF humanoid.py:13 `self_collision = True`).  Only the SELFC instantiations of the
// kernels carry any of it.  [BULLET, restated] URDF_USE_SELF_COLLISION | URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS: every
// pair of geoms collides unless one body is an ancestor of the other: 66 pairs of the 14 single-geom bodies below the
// torso (capsules; spheres are zero-length capsules), friction 2.0 x 2.0, rows after the ground contacts, at most
// MAXSELF contacts per env (the deepest).  Same model as the oracle (ilrl_oracle_set_self_collision).
// A contact between bodies A and B is a row on base + spine + (up to) TWO limbs: 17 whitened numbers, stored as a
// 24-word "wide" row in shared memory (SelfSm::wr):
//   zb[0..3] | zb[4] zb[5] rhs dinv | zs[0..2] limbs (LA | LB << 8; 15 = none) | zlA[0..3] | zlB[0..3] | - - - lambda
__device__ constexpr int kSelfA0[NSELF] = ILRL_SELF_A0;
__device__ constexpr int kSelfA1[NSELF] = ILRL_SELF_A1;
__device__ constexpr int kSelfB0[NSELF] = ILRL_SELF_B0;
__device__ constexpr int kSelfB1[NSELF] = ILRL_SELF_B1;
__device__ constexpr float kSelfReach[NSELF] = ILRL_SELF_REACH;   // broad phase: no contact while the axis midpoints are farther apart
constexpr float SELF_FRICTION = 4.0f;   // geom friction 2.0 (REF humanoid_symmetric_2.xml:5) x 2.0
// (sphere centres, candidate masks and the wide rows live in shared memory: SelfSm)
struct SelfC { int unused; };

// closest points of two segments (Ericson 5.1.9), as the oracle's seg_seg
__device__ __forceinline__ void seg_seg_f(V3 p1, V3 q1, V3 p2, V3 q2, V3& c1, V3& c2) {
  const V3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
  const float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), EPS = 1e-12f;
  float s, t;
  if (a <= EPS && e <= EPS) { s = t = 0.f; }
  else if (a <= EPS) { s = 0.f; t = fminf(fmaxf(f / e, 0.f), 1.f); }
  else {
    const float c = dot(d1, r);
    if (e <= EPS) { t = 0.f; s = fminf(fmaxf(-c / a, 0.f), 1.f); }
    else {
      const float b = dot(d1, d2), den = a * e - b * b;
      s = den > EPS ? fminf(fmaxf((b * f - c * e) / den, 0.f), 1.f) : 0.f;
      t = (b * s + f) / e;
      if (t < 0.f) { t = 0.f; s = fminf(fmaxf(-c / a, 0.f), 1.f); }
      else if (t > 1.f) { t = 1.f; s = fminf(fmaxf((b - c) / a, 0.f), 1.f); }
    }
  }
  c1 = p1 + s * d1; c2 = p2 + t * d2;
}
// pair p of an env whose sphere centres are at gp: closest points on the two axes, distance between the surfaces
__device__ __forceinline__ float self_pair(const float* gp, int p, V3& c1, V3& c2) {
  const int a0 = kSelfA0[p], a1 = kSelfA1[p], b0 = kSelfB0[p], b1 = kSelfB1[p];
  seg_seg_f(rd3(gp + 3 * a0), rd3(gp + 3 * a1), rd3(gp + 3 * b0), rd3(gp + 3 * b1), c1, c2);
  const V3 d = c1 - c2;
  return sqrtf(dot(d, d)) - kSphereR[a0] - kSphereR[b0];
}
// the same with the broad phase in front: a pair whose axis midpoints are farther apart than its reach cannot touch
__device__ __forceinline__ bool self_pair_active(const float* gp, int p) {
  const int a0 = kSelfA0[p], a1 = kSelfA1[p], b0 = kSelfB0[p], b1 = kSelfB1[p];
  const V3 pa0 = rd3(gp + 3 * a0), pa1 = rd3(gp + 3 * a1), pb0 = rd3(gp + 3 * b0), pb1 = rd3(gp + 3 * b1);
  const V3 dm = 0.5f * ((pa0 + pa1) - (pb0 + pb1));
  const float reach = kSelfReach[p];
  if (dot(dm, dm) > reach * reach) return false;
  V3 c1, c2;
  seg_seg_f(pa0, pa1, pb0, pb1, c1, c2);
  const V3 d = c1 - c2;
  return sqrtf(dot(d, d)) - kSphereR[a0] - kSphereR[b0] < (float)ILRL_CONTACT_BREAK;
}

// btPlaneSpace1 (any unit normal)
__device__ __forceinline__ void plane_space(V3 n, V3& p, V3& q) {
  if (fabsf(n.z) > 0.7071067811865475244f) {
    const float a = n.y * n.y + n.z * n.z, k = rsqrtf(a);
    p = mk(0.f, -n.z * k, n.y * k);
    q = mk(a * k, -n.x * p.z, n.x * p.y);
  } else {
    const float a = n.x * n.x + n.y * n.y, k = rsqrtf(a);
    p = mk(-n.y * k, n.x * k, 0.f);
    q = mk(-n.z * p.y, n.z * p.x, a * k);
  }
}

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
struct FkOut { uint32_t act, lim; float sx, sy, ssx, ssy, ex, ey; };

// FULL = false: pose only (part-origin sums and the end-body origin), nothing is written to shared memory.
template <bool FULL, bool TERR = false, bool SELFC = false, class SM>
__device__ __forceinline__ void fk_phase(const Base& b, SM& sm, int e, int tid, int role, FkOut& o,
                                         const Terrain* terr = nullptr, float* gpe = nullptr) {
  const Tables& T = tables(sm);
  float R0[9];
  quat2mat(b.quat[0], b.quat[1], b.quat[2], b.quat[3], R0);
  SV V0; V0.a = rd3(b.w); V0.l = rd3(b.v);
  uint32_t act = 0, lim = 0;
  // torso spheres
  // (a body whose origin is higher than its reach + the breaking distance cannot have a candidate: skip its spheres)
  // TERR: candidate words = centre x, y (relative to the torso) and distance - base z (so that, as on flat ground,
  // distance = base z + word 2 wherever it is needed again)
  if (FULL && (TERR ? (b.p[2] - terr->hmax) * terr->nzmin - T.torso_reach : b.p[2] - T.torso_reach) < (float)ILRL_CONTACT_BREAK) {
#pragma unroll 1
    for (int g = NS - 5; g < NS; g++) {
      V3 c = mv(R0, mk(kSphereC[3 * g], kSphereC[3 * g + 1], kSphereC[3 * g + 2]));
      const float rad = kSphereR[g];
      float d, w2;
      if constexpr (TERR) {
        V3 n;
        const float hgt = terrain_sample(*terr, b.p[0] + c.x, b.p[1] + c.y, n);
        d = (b.p[2] + c.z - hgt) * n.z - rad; w2 = d - b.p[2];
      } else { w2 = c.z - rad; d = b.p[2] + w2; }   // same association as where it is recomputed
      if (d < (float)ILRL_CONTACT_BREAK) {
        act |= 1u << g;
        float* sp = &sm.sph[g][0][e];
        sp[0] = c.x; sp[QE] = c.y; sp[2 * QE] = w2;
      }
    }
  }
  float Rc[9];
#pragma unroll
  for (int i = 0; i < 9; i++) Rc[i] = R0[i];
  V3 oc = mk(0.f, 0.f, 0.f);
  SV Vp = V0;
  float sx = 0.f, sy = 0.f, ssx = 0.f, ssy = 0.f, ex = 0.f, ey = 0.f;
#pragma unroll 1   // (the pose-only pass as well: unrolled, its three inlined copies cost more instruction fetch than they save, -3 %)
  for (int c = 0; c < NL; c++) {
    if (c == 3 && role >= 2) {  // the arms hang off the torso
#pragma unroll
      for (int i = 0; i < 9; i++) Rc[i] = R0[i];
      oc = mk(0.f, 0.f, 0.f);
      Vp = V0;
    }
    const LinkC& L = c < 3 ? link_c(sm, 4, c) : link_c(sm, role, c - 3);
    float* rec = link_rec(sm, c, e, tid);
    oc = oc + mv(Rc, rd3(L.pre));
    if (L.rot) {
      float Rn[9];
      mm(Rc, fixed_rot(sm), Rn);
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
      const float q = rec[W_Q];
      if (FULL && L.j >= 0 && (q - L.lo <= 0.f || L.hi - q <= 0.f)) lim |= 1u << L.j;   // violated joint limit
      sincos_joint(q, sn, cs);
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
        // (SELFC: the sphere centres of every body are published for the pair tests, so the per-body early-out of the
        // ground test - an optimisation that never changes which spheres pass - is not taken)
        const bool near_ground = (TERR ? (b.p[2] + ob.z - terr->hmax) * terr->nzmin - B.reach : b.p[2] + ob.z - B.reach) < (float)ILRL_CONTACT_BREAK;
        if (FULL && (SELFC || near_ground)) {
#pragma unroll 1
          for (int t = 0; t < B.nsph; t++) {
            const int g = B.sidx[t];
            V3 cs_ = ob + mv(Rc, rd3(B.sph[t]));
            const float rad = B.sph[t][3];
            if (SELFC && gpe) { gpe[3 * g] = cs_.x; gpe[3 * g + 1] = cs_.y; gpe[3 * g + 2] = cs_.z; }
            if (SELFC && !near_ground) continue;
            float d, w2;
            if constexpr (TERR) {
              V3 n;
              const float hgt = terrain_sample(*terr, b.p[0] + cs_.x, b.p[1] + cs_.y, n);
              d = (b.p[2] + cs_.z - hgt) * n.z - rad; w2 = d - b.p[2];
            } else { w2 = cs_.z - rad; d = b.p[2] + w2; }
            if (d < (float)ILRL_CONTACT_BREAK) {
              act |= 1u << g;
              float* sp = &sm.sph[g][0][e];
              sp[0] = cs_.x; sp[QE] = cs_.y; sp[2 * QE] = w2;
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
  o.act = act; o.lim = lim; o.sx = sx; o.sy = sy; o.ssx = ssx; o.ssy = ssy; o.ex = ex; o.ey = ey;
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

// ---- constraint rows in WHITENED form.
// With the articulated-body quantities of the inward pass (U_c = IA_c S_c, d_c = S_c . U_c, base factor IA_0 = L0 L0^T)
// the inverse joint-space inertia factorises as  M^-1 = W^T W,  W = D^-1/2 (inward sweep)  (innovations factorisation;
// checked numerically against a composite-Jacobian M by tools/check_innovations_identity.py).  A constraint row J_i
// therefore only needs its own inward sweep  z_i = W J_i^T : the innovations u_c of the links between the row's link
// and the base, scaled by sqrt(1/d_c), and  L0^-1 (force arriving at the base)  - 13 numbers with the sparsity of J_i
// itself (base 6, spine 3, the row's limb 4).  J_i M^-1 J_k^T = z_i . z_k, so projected Gauss-Seidel runs on
// z = sum_k z_k lambda_k  (J_i . dv = z_i . z) and the velocity change  dv = W^T z  is recovered by ONE outward sweep
// per substep, distributed over the quad like the forward dynamics.  Compared with storing M^-1 J^T per row (the
// previous generation: an outward sweep over all 19 links per row, 40-word rows) a row costs a walk of <= 7 links,
// is 16 words, and a Gauss-Seidel evaluation is 13 + 13 multiply-adds.  Same row order, formulas and constants as the
// oracle (limits in joint order, contact normals, friction pairs; 5 sweeps).
//
// stored row (4 float4): zb[0..3] | zb[4] zb[5] rhs dinv | zs[0..2] limb | zl[0..3]      (limb: -1 = spine / torso)

// y = L^-1 b  /  x = L^-T y  with the packed factor of this env in shared memory (diagonal entries hold 1 / L_jj)
__device__ __forceinline__ void fwd_subst(const float* L, SV b, float* y) {
  y[0] = b.a.x; y[1] = b.a.y; y[2] = b.a.z; y[3] = b.l.x; y[4] = b.l.y; y[5] = b.l.z;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float t = y[i];
#pragma unroll
    for (int c = 0; c < i; c++) t -= L[(i * (i + 1) / 2 + c) * QE] * y[c];
    y[i] = t * L[(i * (i + 1) / 2 + i) * QE];
  }
}
__device__ __forceinline__ void bwd_subst(const float* L, float* y) {
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    float t = y[i];
#pragma unroll
    for (int c = i + 1; c < 6; c++) t -= L[(c * (c + 1) / 2 + i) * QE] * y[c];
    y[i] = t * L[(i * (i + 1) / 2 + i) * QE];
  }
}
__device__ __forceinline__ void ld_SUq(const float* rec, SV& S, SV& U, float& dinv, float& sqd) {
  ld_SU(rec, S, U, dinv);
  sqd = rec[W_SQD];
}

// row of a violated joint limit on joint j of env e (any env of this warp): generalized impulse `dir` on the joint
template <class SM>
__device__ __forceinline__ void build_limit_row(const SM& sm, const Tables& T, int j, int e, int qb, float idt, float* row) {
  const int L = T.jL[j];
  int c = T.jC[j];
  const float* rec0 = link_rec_of(sm, L, c, e, qb);
  const float q_ = rec0[W_Q], nu = rec0[W_NU];
  float pen, dir;
  const LinkC& K = c < 3 ? link_c(sm, 4, c) : link_c(sm, L, c - 3);   // (the joint's limits from the staged per-link table)
  const float lo = K.lo, hi = K.hi;
  if (q_ - lo <= 0.f) { pen = q_ - lo; dir = 1.f; } else { pen = hi - q_; dir = -1.f; }
  float4* r4 = reinterpret_cast<float4*>(row);
  r4[2] = make_float4(0.f, 0.f, 0.f, __int_as_float(c >= 3 ? L : -1));
  r4[3] = make_float4(0.f, 0.f, 0.f, 0.f);
  SV pf = svzero();
  float dd = 0.f, f = dir;
#pragma unroll 1
  for (; c >= 0; c--) {
    if (c == 2 && L >= 2) break;  // arms attach to the torso
    SV S, U;
    float di, sq;
    ld_SUq(link_rec_of(sm, L, c, e, qb), S, U, di, sq);
    const float u = f - sdot(S, pf);
    f = 0.f;
    const float z = u * sq;
    row[c < 3 ? R_ZS + c : R_ZL + c - 3] = z;
    dd = fmaf(z, z, dd);
    pf = pf + (u * di) * U;
  }
  float y[6];
  fwd_subst(&sm.L0[0][e], neg(pf), y);
#pragma unroll
  for (int i = 0; i < 6; i++) dd = fmaf(y[i], y[i], dd);
  const float dinv = __frcp_rn(dd);
  const float pos = -pen * (float)ILRL_LIMIT_ERP * idt;
  r4[0] = make_float4(y[0], y[1], y[2], y[3]);
  r4[1] = make_float4(y[4], y[5], (pos - dir * nu) * dinv, dinv);
}

// the three rows (normal, two friction directions) of the ground contact of sphere g of env e: one walk, link records
// loaded once, 3-way instruction-level parallelism.  xx: contact point relative to the torso origin, dist: its distance
// along the normal nrm; t1, t2 = btPlaneSpace1(nrm) (flat ground: (0,0,1), (0,-1,0), (1,0,0), folded at compile time).
template <class SM>
__device__ __forceinline__ void build_contact_rows(const SM& sm, const Tables& T, int g, int e, int qb, float idt, V3 xx,
                                                   float dist, const float* nub, float* row0, float* row1, float* row2,
                                                   V3 nrm, V3 t1, V3 t2) {
  const int L = T.sphL[g];
  int c = T.sphC[g];
  float* rows[3] = {row0, row1, row2};
  // normal (0,0,1), tangents btPlaneSpace1 -> (0,-1,0), (1,0,0); spatial force of a unit impulse at the contact point
  SV F[3], pf[3];
  F[0].l = nrm; F[1].l = t1; F[2].l = t2;
  float dd[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    F[i].a = cross(xx, F[i].l);
    pf[i] = neg(F[i]);
    dd[i] = 0.f;
    float4* r4 = reinterpret_cast<float4*>(rows[i]);
    r4[2] = make_float4(0.f, 0.f, 0.f, __int_as_float(c >= 3 ? L : -1));
    r4[3] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  SV Vb;  // unconstrained new velocity of the contact body: J . nu = F . Vb
  Vb.a = mk(nub[0], nub[1], nub[2]); Vb.l = mk(nub[3], nub[4], nub[5]);
#pragma unroll 1
  for (; c >= 0; c--) {
    if (c == 2 && L >= 2) break;
    const float* rec = link_rec_of(sm, L, c, e, qb);
    SV S, U;
    float di, sq;
    ld_SUq(rec, S, U, di, sq);
    Vb = Vb + rec[W_NU] * S;
    const int zw = c < 3 ? R_ZS + c : R_ZL + c - 3;
    const float u0 = -sdot(S, pf[0]), u1 = -sdot(S, pf[1]), u2 = -sdot(S, pf[2]);
    const float z0 = u0 * sq, z1 = u1 * sq, z2 = u2 * sq;
    row0[zw] = z0; row1[zw] = z1; row2[zw] = z2;
    dd[0] = fmaf(z0, z0, dd[0]); dd[1] = fmaf(z1, z1, dd[1]); dd[2] = fmaf(z2, z2, dd[2]);
    pf[0] = pf[0] + (u0 * di) * U; pf[1] = pf[1] + (u1 * di) * U; pf[2] = pf[2] + (u2 * di) * U;
  }
  const float* L0 = &sm.L0[0][e];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float y[6];
    fwd_subst(L0, neg(pf[i]), y);
    float d = dd[i];
#pragma unroll
    for (int k = 0; k < 6; k++) d = fmaf(y[k], y[k], d);
    const float dinv = __frcp_rn(d);
    const float pos = i == 0 ? (dist > 0.f ? -dist * idt : -dist * (float)ILRL_CONTACT_ERP * idt) : 0.f;
    float4* r4 = reinterpret_cast<float4*>(rows[i]);
    r4[0] = make_float4(y[0], y[1], y[2], y[3]);
    r4[1] = make_float4(y[4], y[5], (pos - sdot(Vb, F[i])) * dinv, dinv);
  }
}

// the three WIDE rows (normal, two friction directions) of self-contact pair p of env e: +F on body A at xa, -F on body
// B at xb.  Each side walks its own limb inward; the legs' remainders (and the forces on lwaist / pelvis) enter the
// spine walk, the arms' remainders go straight to the base.  gp: the env's sphere centres; wrow: 3 x WRW words.
template <class SM>
__device__ __forceinline__ void build_self_rows(const SM& sm, const Tables& T, int p, int e, int qb, float idt,
                                                const float* gp, const float* nub, float* wrow) {
  V3 c1, c2;
  const float dist = self_pair(gp, p, c1, c2);
  V3 n = c1 - c2;
  float len = sqrtf(dot(n, n));
  if (len < 1e-9f) { n = mk(0.f, 0.f, 1.f); len = 1.f; }
  n = (1.f / len) * n;   // from body B towards body A
  const int ga = kSelfA0[p], gb = kSelfB0[p];
  const V3 xa = c1 - kSphereR[ga] * n, xb = c2 + kSphereR[gb] * n;
  V3 dir[3];
  dir[0] = n;
  plane_space(n, dir[1], dir[2]);
  const int LA = T.sphL[ga], CA = T.sphC[ga], LB = T.sphL[gb], CB = T.sphC[gb];
  SV FA[3], FB[3], pfS[3], pfB0[3];   // pfS: force travelling down the spine, pfB0: force arriving at the base directly
  // (the whitened spine / limb entries go straight into the stored row, words 8..10 and 12..19: the walks below index
  // them with run-time link numbers, and as register arrays they lived in local memory)
  const int lab = (LA < 0 ? 15 : LA) | ((LB < 0 ? 15 : LB) << 8);
  float dd[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    FA[i].l = dir[i]; FA[i].a = cross(xa, dir[i]);
    FB[i].l = dir[i]; FB[i].a = cross(xb, dir[i]);
    pfS[i] = svzero(); pfB0[i] = svzero(); dd[i] = 0.f;
    float4* r4 = reinterpret_cast<float4*>(wrow + i * WRW);
    r4[2] = make_float4(0.f, 0.f, 0.f, __int_as_float(lab));
    r4[3] = make_float4(0.f, 0.f, 0.f, 0.f);
    r4[4] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  SV Vs[2];   // unconstrained new velocity of the two bodies
  Vs[0].a = mk(nub[0], nub[1], nub[2]); Vs[0].l = mk(nub[3], nub[4], nub[5]);
  Vs[1] = Vs[0];
  int enter[2];   // spine index at which the side's force enters the spine walk (-1: never, an arm)
#pragma unroll
  for (int side = 0; side < 2; side++) {
    const int L = side ? LB : LA, C = side ? CB : CA;
    SV pf[3];
#pragma unroll
    for (int i = 0; i < 3; i++) pf[i] = side ? FB[i] : neg(FA[i]);
    if (L >= 0) {
#pragma unroll 1
      for (int c = C; c >= 3; c--) {
        SV S, U;
        float di, sq;
        const float* rec = link_rec_of(sm, L, c, e, qb);
        ld_SUq(rec, S, U, di, sq);
        Vs[side] = Vs[side] + rec[W_NU] * S;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const float u = -sdot(S, pf[i]), z = u * sq;
          wrow[i * WRW + 12 + 4 * side + (c - 3)] = z;
          dd[i] = fmaf(z, z, dd[i]);
          pf[i] = pf[i] + (u * di) * U;
        }
      }
      enter[side] = L < 2 ? 2 : -1;
    } else {
      enter[side] = C;   // lwaist (1) or pelvis (2)
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      if (enter[side] == 2) pfS[i] = pfS[i] + pf[i];
      else if (enter[side] < 0) pfB0[i] = pfB0[i] + pf[i];
    }
  }
  // spine walk 2 -> 0; a force on lwaist joins at link 1
#pragma unroll 1
  for (int c = 2; c >= 0; c--) {
    if (c == 1) {
#pragma unroll
      for (int side = 0; side < 2; side++)
        if (enter[side] == 1) {
#pragma unroll
          for (int i = 0; i < 3; i++) pfS[i] = pfS[i] + (side ? FB[i] : neg(FA[i]));
        }
    }
    SV S, U;
    float di, sq;
    const float* rec = link_rec_of(sm, -1, c, e, qb);
    ld_SUq(rec, S, U, di, sq);
    const float nu = rec[W_NU];
#pragma unroll
    for (int side = 0; side < 2; side++)
      if (enter[side] >= c) Vs[side] = Vs[side] + nu * S;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float u = -sdot(S, pfS[i]), z = u * sq;
      wrow[i * WRW + 8 + c] = z;
      dd[i] = fmaf(z, z, dd[i]);
      pfS[i] = pfS[i] + (u * di) * U;
    }
  }
  const float* L0 = &sm.L0[0][e];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float y[6];
    fwd_subst(L0, neg(pfS[i] + pfB0[i]), y);
    float d = dd[i];
#pragma unroll
    for (int k = 0; k < 6; k++) d = fmaf(y[k], y[k], d);
    const float dinv = __frcp_rn(d);
    const float pos = i == 0 ? (dist > 0.f ? -dist * idt : -dist * (float)ILRL_CONTACT_ERP * idt) : 0.f;
    const float rv = sdot(Vs[0], FA[i]) - sdot(Vs[1], FB[i]);
    float4* r4 = reinterpret_cast<float4*>(wrow + i * WRW);
    r4[0] = make_float4(y[0], y[1], y[2], y[3]);
    r4[1] = make_float4(y[4], y[5], (pos - rv) * dinv, dinv);
    r4[5] = make_float4(0.f, 0.f, 0.f, 0.f);   // (word 23: the row's multiplier)
  }
}
// wide-row Gauss-Seidel pieces (the quad sums the limb parts: a wide row has two owner lanes)
struct WRow { float4 a, b, c, d, f, g; };
__device__ __forceinline__ void wrow_load(const float* w, WRow& r) {
  const float4* p = reinterpret_cast<const float4*>(w);
  r.a = p[0]; r.b = p[1]; r.c = p[2]; r.d = p[3]; r.f = p[4]; r.g = p[5];
}
__device__ __forceinline__ float wrow_resid(const WRow& r, const float* zb, const float* zc, int role) {
  const int lab = __float_as_int(r.c.w), LA = lab & 15, LB = (lab >> 8) & 15;
  float own = 0.f;
  if (role == LA) own = fmaf(r.d.y, zc[4], r.d.x * zc[3]) + fmaf(r.d.w, zc[6], r.d.z * zc[5]);
  if (role == LB) own += fmaf(r.f.y, zc[4], r.f.x * zc[3]) + fmaf(r.f.w, zc[6], r.f.z * zc[5]);
  own += __shfl_xor_sync(FULLMASK, own, 1);
  own += __shfl_xor_sync(FULLMASK, own, 2);
  const float r0 = fmaf(r.a.z, zb[2], fmaf(r.a.y, zb[1], r.a.x * zb[0]));
  const float r1 = fmaf(r.b.y, zb[5], fmaf(r.b.x, zb[4], r.a.w * zb[3]));
  const float r2 = fmaf(r.c.z, zc[2], fmaf(r.c.y, zc[1], r.c.x * zc[0]));
  return fmaf(-r.b.w, ((r0 + r1) + r2) + own, r.g.w + r.b.z);   // lambda + rhs - dinv (z_row . z)
}
__device__ __forceinline__ void wrow_axpy(const WRow& r, float a, int role, float* zb, float* zc) {
  const int lab = __float_as_int(r.c.w), LA = lab & 15, LB = (lab >> 8) & 15;
  zb[0] += a * r.a.x; zb[1] += a * r.a.y; zb[2] += a * r.a.z; zb[3] += a * r.a.w; zb[4] += a * r.b.x; zb[5] += a * r.b.y;
  zc[0] += a * r.c.x; zc[1] += a * r.c.y; zc[2] += a * r.c.z;
  const float aa = role == LA ? a : 0.f, ab = role == LB ? a : 0.f;
  zc[3] += aa * r.d.x + ab * r.f.x; zc[4] += aa * r.d.y + ab * r.f.y; zc[5] += aa * r.d.z + ab * r.f.z; zc[6] += aa * r.d.w + ab * r.f.w;
}
// one sweep's self-contact normals / friction pairs (warp-uniform loops; a quad past its own count re-evaluates its
// row 0 - finite: zero at allocation or a real row - with a zero step)
__device__ __forceinline__ void wide_normals(float* wr, int nself, int ns_w, int role, float* zb, float* zc) {
#pragma unroll 1
  for (int k = 0; k < ns_w; k++) {
    const bool live = k < nself;
    float* w = wr + (live ? 3 * k : 0) * WRW;
    WRow r;
    wrow_load(w, r);
    const float nl = fmaxf(wrow_resid(r, zb, zc, role), 0.f);
    if (live && role == 0) w[23] = nl;
    wrow_axpy(r, live ? nl - r.g.w : 0.f, role, zb, zc);
  }
}
__device__ __forceinline__ void wide_friction(float* wr, int nself, int ns_w, int role, float* zb, float* zc) {
#pragma unroll 1
  for (int k = 0; k < ns_w; k++) {
    const bool in = k < nself;
    const float ln = in ? wr[3 * k * WRW + 23] : 0.f;
    const bool live = in && ln > 0.f;
    if (!__any_sync(FULLMASK, live)) continue;
    float* w1 = wr + (live ? 3 * k + 1 : 0) * WRW;
    float* w2 = wr + (live ? 3 * k + 2 : 0) * WRW;
    WRow r1, r2;
    wrow_load(w1, r1);
    wrow_load(w2, r2);
    const float lim_f = SELF_FRICTION * ln;
    float s1 = wrow_resid(r1, zb, zc, role), s2 = wrow_resid(r2, zb, zc, role);
    const float n2 = s1 * s1 + s2 * s2;
    if (n2 > lim_f * lim_f) { const float sc = lim_f * rsqrtf(n2); s1 *= sc; s2 *= sc; }
    if (live && role == 0) { w1[23] = s1; w2[23] = s2; }
    wrow_axpy(r1, live ? s1 - r1.g.w : 0.f, role, zb, zc);
    wrow_axpy(r2, live ? s2 - r2.g.w : 0.f, role, zb, zc);
  }
}

// ---- projected Gauss-Seidel on the whitened impulse sum z.  zb: base, zc: chain (spine replicated, limb private)
struct RowRegs { float4 a, b, c, d; };
template <class P4>
__device__ __forceinline__ void row_load(P4 rp, RowRegs& r) { r.a = rp[0]; r.b = rp[1]; r.c = rp[2]; r.d = rp[3]; }
// lam + rhs - (z_row . z) * dinv.  This sits on the Gauss-Seidel critical path: the limb part (owned by one lane of the
// quad) is requested first, the replicated base / spine part is folded into the constant while the shuffle is in
// flight (three short chains instead of one 9-deep one), one multiply-add remains after it.
__device__ __forceinline__ float row_resid(const RowRegs& r, float lam, const float* zb, const float* zc, bool mine, int src,
                                           unsigned qm) {
  const float own = mine ? fmaf(r.d.y, zc[4], r.d.x * zc[3]) + fmaf(r.d.w, zc[6], r.d.z * zc[5]) : 0.f;
  const float part = __shfl_sync(qm, own, src);
  const float r0 = fmaf(r.a.z, zb[2], fmaf(r.a.y, zb[1], r.a.x * zb[0]));
  const float r1 = fmaf(r.b.y, zb[5], fmaf(r.b.x, zb[4], r.a.w * zb[3]));
  const float r2 = fmaf(r.c.z, zc[2], fmaf(r.c.y, zc[1], r.c.x * zc[0]));
  const float t = fmaf(-r.b.w, (r0 + r1) + r2, lam + r.b.z);
  return fmaf(-r.b.w, part, t);
}
__device__ __forceinline__ void row_axpy(const RowRegs& r, float a, bool mine, float* zb, float* zc) {
  zb[0] += a * r.a.x; zb[1] += a * r.a.y; zb[2] += a * r.a.z; zb[3] += a * r.a.w; zb[4] += a * r.b.x; zb[5] += a * r.b.y;
  zc[0] += a * r.c.x; zc[1] += a * r.c.y; zc[2] += a * r.c.z;
  const float al = mine ? a : 0.f;
  zc[3] += al * r.d.x; zc[4] += al * r.d.y; zc[5] += al * r.d.z; zc[6] += al * r.d.w;
}
// OVER = false: the env has no row beyond the shared-memory budget (the common case): plain shared-memory loads, no
// predicated-off global loads in the instruction stream (those would take issue slots in every row evaluation).
template <bool OVER, class SM>
__device__ __forceinline__ void row_fetch(const SM& sm, const float* gscr, int e, int r, RowRegs& rr) {
  if (!OVER || r < SM::RSM) row_load(reinterpret_cast<const float4*>(&sm.rows[e][r * RW]), rr);
  else row_load(reinterpret_cast<const float4*>(gscr + (size_t)(r - SM::RSM) * RW), rr);
}

// 5 projected-Gauss-Seidel sweeps for the env of this quad.  Row order in storage = evaluation order of a sweep:
// limits [0, nlim), contact normals [nlim, nlim + ncon), then the friction pair of contact c at nlim + ncon + 2c.
// Every lane of the warp is here and every shuffle carries the constant full mask (a run-time quad mask costs a
// MATCH / REDUX / VOTE / branch preamble in front of every shuffle: in-order issue put ~100 cycles of it on the
// critical path of each row evaluation), so the loops run to the WARP's largest count and a lane past its own count
// evaluates its row 0 (finite: zeroed at tile start or a real row) with a zero step.
template <bool OVER, bool SELFC = false, class SM>
__device__ __forceinline__ void pgs_sweeps(SM& sm, const float* gscr, int e, int role, int qb, int nlim, int ncon,
                                           float* zb, float* zc, float* wr = nullptr, int nself = 0) {
  constexpr bool FULL = true;
  const int ns_w = SELFC ? __reduce_max_sync(FULLMASK, nself) : 0;
  constexpr unsigned m = FULLMASK;
  float* lamv = sm.lam(e);
  const int nfirst = nlim + ncon;
  const int nf_w = __reduce_max_sync(FULLMASK, nfirst);
  const int nc_w = __reduce_max_sync(FULLMASK, ncon);
#pragma unroll 1
  for (int itn = 0; itn < ILRL_SOLVER_ITERS; itn++) {
    // limits, then contact normals.  The next row (independent of z) is fetched while this one is applied:
    // two register buffers used alternately (the loop is unrolled by two so that no copies are needed).
    auto fetch = [&](int k, RowRegs& rr, float& lam) {
      const bool live = !FULL || k < nfirst;
      const int r = live ? k : 0;
      row_fetch<OVER>(sm, gscr, e, r, rr);
      lam = lamv[r];
    };
    auto apply = [&](int k, const RowRegs& rr, float lam) {
      const bool live = !FULL || k < nfirst;
      const int L = __float_as_int(rr.c.w);
      const bool mine = role == L;
      const float nl = fmaxf(row_resid(rr, lam, zb, zc, mine, qb + (L & 3), m), 0.f);
      if (live) lamv[k] = nl;
      row_axpy(rr, live ? nl - lam : 0.f, mine, zb, zc);
    };
    RowRegs ra, rb;
    float lama = 0.f, lamb = 0.f;
    if (nf_w > 0) fetch(0, ra, lama);
#pragma unroll 1
    for (int k = 0; k < nf_w; k += 2) {
      if (k + 1 < nf_w) fetch(k + 1, rb, lamb);
      apply(k, ra, lama);
      if (k + 1 < nf_w) {
        if (k + 2 < nf_w) fetch(k + 2, ra, lama);
        apply(k + 1, rb, lamb);
      }
    }
    if constexpr (SELFC) wide_normals(wr, nself, ns_w, role, zb, zc);   // self-contact normals follow the ground normals
#pragma unroll 1
    for (int c = 0; c < nc_w; c++) {  // friction pairs, cone re-projected on the current normal impulse
      const bool in = !FULL || c < ncon;
      const float ln = in ? lamv[nlim + c] : 0.f;
      const bool live = in && ln > 0.f;
      if (!__any_sync(FULLMASK, live)) continue;
      const int r1i = live ? nfirst + 2 * c : 0, r2i = live ? r1i + 1 : 0;
      RowRegs r1, r2;
      row_fetch<OVER>(sm, gscr, e, r1i, r1);
      row_fetch<OVER>(sm, gscr, e, r2i, r2);
      const int L = __float_as_int(r1.c.w);
      const bool mine = role == L;
      const float lim_f = (float)ILRL_FRICTION * ln;
      const float l1 = lamv[r1i], l2 = lamv[r2i];
      float s1 = row_resid(r1, l1, zb, zc, mine, qb + (L & 3), m);
      float s2 = row_resid(r2, l2, zb, zc, mine, qb + (L & 3), m);
      const float n2 = s1 * s1 + s2 * s2;
      if (n2 > lim_f * lim_f) { const float sc = lim_f * rsqrtf(n2); s1 *= sc; s2 *= sc; }
      if (live) { lamv[r1i] = s1; lamv[r2i] = s2; }
      row_axpy(r1, live ? s1 - l1 : 0.f, mine, zb, zc);
      row_axpy(r2, live ? s2 - l2 : 0.f, mine, zb, zc);
    }
    if constexpr (SELFC) wide_friction(wr, nself, ns_w, role, zb, zc);
  }
}

// ---- one substep of dt for the env of this quad.  Joint state / torques live in the link records.
// ALL 32 lanes of a warp execute it, so that every shuffle carries the compile-time full mask (see pgs_sweeps) and an
// env computes bit-identical results whoever its neighbours are: the lanes of an env that does not step (beyond the
// end of the batch, waiting for a high-level action, skipped by a NaN action) run on a benign dummy state
// (`steps` = false: they produce no constraint rows and their results are never stored), and they still help to build
// the rows of the warp's other envs.  gscr_tile: overflow-row scratch of env 0 of this CTA's tile.
// SELFC: sc_tile = the self-collision scratch of env 0 of this CTA's tile (gp / sd / wr already offset to it).
template <bool TERR = false, bool SELFC = false, class SM>
__device__ __forceinline__ int substep(Base& b, SM& sm, float* gscr_tile, int e, int tid, int role, bool steps, float dt,
                                       Prof& pf, const Terrain* terr = nullptr, const SelfC* sc_tile = nullptr) {
  constexpr int RSM = SM::RSM;
  constexpr unsigned qm = FULLMASK, wm = FULLMASK;
  float* gscr = gscr_tile + (size_t)e * (GROWS * RW);
  const int qb = tid & ~3;
  const Tables& T = tables(sm);
  // ---- phase A
  FkOut fo;
  float* gpe = nullptr;
  if constexpr (SELFC) { if (steps) gpe = &sm.S.gp[e][0]; }
  fk_phase<true, TERR, SELFC>(b, sm, e, tid, role, fo, terr, gpe);
  uint32_t act = fo.act, lim = fo.lim;
  pf.mark(PF_FK);
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
      rec[W_SQD] = sqrtf(dinv);
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
  pf.mark(PF_INWARD);
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
      SV S, cJ, U;
      float di, uu;
      ld_all(rec, S, cJ, U, di, uu);
      const float4 st4 = reinterpret_cast<const float4*>(rec)[5];  // q, qd, tau, nu
      const SV ad = ap + cJ;
      const float qa = di * (uu - sdot(ad, U));
      rec[W_NU] = clampf(st4.y + dt * qa, -M, M);
      ap = ad + qa * S;
    }
  }
  lim |= __shfl_xor_sync(qm, lim, 1); lim |= __shfl_xor_sync(qm, lim, 2);
  act |= __shfl_xor_sync(qm, act, 1); act |= __shfl_xor_sync(qm, act, 2);
  if (!steps) { lim = 0; act = 0; }
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
  // ---- self-collision candidates: the 8 x 66 pair tests of the warp are one pool dealt to its 32 lanes
  uint32_t sm0 = 0u, sm1 = 0u, sm2 = 0u;
  int nself = 0;
  float* wr_e = nullptr;
  if constexpr (SELFC) {
    const int lane = tid & 31, e0 = e & ~7;
    const unsigned stepmask = __ballot_sync(FULLMASK, steps);
    if (role < 3) sm.S.mask[e][role] = 0u;
    __syncwarp();   // every lane's sphere centres (FK pass) and the cleared masks are visible to the warp
    // two items per lane and iteration: every load (pair record, 4 axis end points) in flight together, broad phase (axis
    // midpoints farther apart than the pair's reach: no contact possible), the segment test on the same registers
#pragma unroll 1
    for (int it0 = lane; it0 < 8 * NSELF; it0 += 64) {
      V3 A0[2], A1[2], B0[2], B1[2];
      float lim[2];
      bool near[2];
      int qq[2], pp[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const int it = it0 + 32 * u, itc = min(it, 8 * NSELF - 1);
        const int q = itc / NSELF, p = itc - q * NSELF;
        qq[u] = q; pp[u] = p;
        const uint4 pr = sm.S.pairs[p];
        const float* gp = &sm.S.gp[e0 + q][0];
        A0[u] = rd3(gp + 3 * (pr.x & 255u)); A1[u] = rd3(gp + 3 * ((pr.x >> 8) & 255u));
        B0[u] = rd3(gp + 3 * ((pr.x >> 16) & 255u)); B1[u] = rd3(gp + 3 * (pr.x >> 24));
        const V3 dm = (A0[u] + A1[u]) - (B0[u] + B1[u]);   // twice the midpoint difference
        lim[u] = __uint_as_float(pr.z);
        near[u] = it < 8 * NSELF && ((stepmask >> (4 * q)) & 1u) && dot(dm, dm) <= __uint_as_float(pr.y);
      }
#pragma unroll
      for (int u = 0; u < 2; u++)
        if (near[u]) {
          V3 c1, c2;
          seg_seg_f(A0[u], A1[u], B0[u], B1[u], c1, c2);
          const V3 d = c1 - c2;
          if (sqrtf(dot(d, d)) < lim[u]) atomicOr(&sm.S.mask[e0 + qq[u]][pp[u] >> 5], 1u << (pp[u] & 31));
        }
    }
    __syncwarp();
    if (steps) { sm0 = sm.S.mask[e][0]; sm1 = sm.S.mask[e][1]; sm2 = sm.S.mask[e][2]; }
    nself = __popc(sm0) + __popc(sm1) + __popc(sm2);
    while (nself > MAXSELF) {   // keep the deepest (ties: drop the later pair); replicated in the quad
      int worst = -1;
      float wd = -1e30f;
#pragma unroll 1
      for (int p = 0; p < NSELF; p++) {
        const uint32_t w = p < 32 ? sm0 : p < 64 ? sm1 : sm2;
        if ((w >> (p & 31)) & 1u) {
          V3 c1, c2;
          const float d = self_pair(&sm.S.gp[e][0], p, c1, c2);
          if (d >= wd) { wd = d; worst = p; }
        }
      }
      if (worst < 32) sm0 &= ~(1u << worst); else if (worst < 64) sm1 &= ~(1u << (worst - 32)); else sm2 &= ~(1u << (worst - 64));
      nself--;
    }
    wr_e = &sm.S.wr[e][0];
  }
  const int nlim = __popc(lim), ncon = nact, nrows = nlim + 3 * ncon + 3 * nself;
  pf.mark(PF_OUTWARD);
  pf.maxv(PF_MAXROWS, (long long)__reduce_max_sync(FULLMASK, nrows));
  float zb[6], zc[NL];   // whitened impulse sum: base (replicated), chain (spine replicated, limb private)
#pragma unroll
  for (int i = 0; i < 6; i++) zb[i] = 0.f;
#pragma unroll
  for (int i = 0; i < NL; i++) zc[i] = 0.f;
  // ---- build the rows.  The items (one violated limit = one row, one contact = three rows) of the warp's 8 envs form
  // ONE pool per kind that is dealt to all its lanes: an env with many rows is helped by the lanes of envs with few
  // (every input of an item is in shared memory or one shuffle away), so a warp needs ceil(items / lanes) rounds of a
  // builder instead of max over its envs of ceil(items / 4).
  {
    const float idt = 1.0f / dt;
    const int lane = tid & 31;
    constexpr int nlanes = 32;
    const int rank = lane;
    __syncwarp(wm);   // the link records and base factors of every env of the warp are complete
#pragma unroll 1
    for (int kind = 0; kind < (SELFC ? 3 : 2); kind++) {
      const int n_own = kind == 0 ? nlim : kind == 1 ? ncon : nself;
      int cnt[8], total = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        cnt[q] = __shfl_sync(wm, n_own, 4 * q);   // (a quad that does not step has no items)
        total += cnt[q];
      }
#pragma unroll 1
      for (int g0 = 0; g0 < total; g0 += nlanes) {
        // item g -> (quad q, item k of that quad's env)
        int k = g0 + rank, q = 0;
        const bool has = k < total;
#pragma unroll
        for (int t = 0; t < 7; t++)
          if (q == t && k >= cnt[t]) { k -= cnt[t]; q = t + 1; }
        if (!has) { q = lane >> 2; k = 0; }
        const int src = 4 * q;
        const uint32_t mask_q = __shfl_sync(wm, kind == 0 ? lim : kind == 1 ? act : sm0, src);
        const int nlim_q = __shfl_sync(wm, nlim, src);
        float bz_q = 0.f, bx_q = 0.f, by_q = 0.f, nub_q[6];
        int ncon_q = 0;
        uint32_t m1_q = 0u, m2_q = 0u;
        if (kind >= 1) {
          ncon_q = __shfl_sync(wm, ncon, src);
          bz_q = __shfl_sync(wm, b.p[2], src);
          if (TERR) { bx_q = __shfl_sync(wm, b.p[0], src); by_q = __shfl_sync(wm, b.p[1], src); }
#pragma unroll
          for (int i = 0; i < 6; i++) nub_q[i] = __shfl_sync(wm, nub[i], src);
          if (SELFC && kind == 2) { m1_q = __shfl_sync(wm, sm1, src); m2_q = __shfl_sync(wm, sm2, src); }
        }
        if (has) {
          const int e_q = (e & ~7) + q, qb_q = (tid & ~31) + 4 * q;
          float* gscr_q = gscr_tile + (size_t)e_q * (GROWS * RW);
          auto row_ptr = [&](int r) { return r < RSM ? &sm.rows[e_q][r * RW] : gscr_q + (size_t)(r - RSM) * RW; };
          if constexpr (SELFC) if (kind == 2) {
            // k-th set bit of the env's 66-bit pair mask
            int kk = k, pidx;
            const int c0 = __popc(mask_q), c1 = __popc(m1_q);
            if (kk < c0) pidx = nth_set_bit(mask_q, kk);
            else if (kk - c0 < c1) pidx = 32 + nth_set_bit(m1_q, kk - c0);
            else pidx = 64 + nth_set_bit(m2_q, kk - c0 - c1);
            build_self_rows(sm, T, pidx, e_q, qb_q, idt, &sm.S.gp[e_q][0], nub_q,
                            &sm.S.wr[e_q][3 * k * WRW]);
            continue;
          }
          const int g = nth_set_bit(mask_q, k);
          if (kind == 0) {
            sm.lam(e_q)[k] = 0.f;
            build_limit_row(sm, T, g, e_q, qb_q, idt, row_ptr(k));
          } else {
            const int r0 = nlim_q + k, r1 = nlim_q + ncon_q + 2 * k;   // normal | friction pair (storage = sweep order)
            sm.lam(e_q)[r0] = 0.f; sm.lam(e_q)[r1] = 0.f; sm.lam(e_q)[r1 + 1] = 0.f;
            const float* sp = &sm.sph[g][0][e_q];
            if constexpr (TERR) {
              // the plane under the sphere's centre again (L2-resident heightfield); centre height from the stored
              // distance: dist = (bz + cz - h) nz - r
              V3 n, t1, t2;
              const float cx = sp[0], cy = sp[QE], dist = bz_q + sp[2 * QE], rad = kSphereR[g];
              const float hgt = terrain_sample(*terr, bx_q + cx, by_q + cy, n);
              plane_space_up(n, t1, t2);
              const float cz = (dist + rad) / n.z + hgt - bz_q;
              build_contact_rows(sm, T, g, e_q, qb_q, idt, mk(cx - rad * n.x, cy - rad * n.y, cz - rad * n.z), dist, nub_q,
                                 row_ptr(r0), row_ptr(r1), row_ptr(r1 + 1), n, t1, t2);
            } else {
              build_contact_rows(sm, T, g, e_q, qb_q, idt, mk(sp[0], sp[QE], sp[2 * QE]), bz_q + sp[2 * QE], nub_q,
                                 row_ptr(r0), row_ptr(r1), row_ptr(r1 + 1), mk(0.f, 0.f, 1.f), mk(0.f, -1.f, 0.f),
                                 mk(1.f, 0.f, 0.f));
            }
          }
        }
      }
    }
    __syncwarp(wm);   // rows of an env may have been written by lanes of another quad
  }
  pf.mark(PF_ROWS);
  // ---- projected Gauss-Seidel on the whitened impulse sum (warp-uniform loops).  One loop variant per layout: the
  // on-chip layout holds every row in shared memory (no overflow variant at all), the dense layouts always take the
  // overflow-aware loads.
  // (A dense variant - Delassus matrix Z Z^T formed once per substep, sweeps on w = A lambda in registers, no shuffle -
  // was built and measured: 45.3M against 52.8M env-steps/s at 4096 envs; the per-slot blocks are if-converted and
  // every warp pays for all 20 slots, and it broke bit-identity between the layouts.  profiles/r2_warp_phases.txt.)
  if (__any_sync(FULLMASK, nrows > 0)) pgs_sweeps<(RSM < MAXROWS), SELFC>(sm, gscr, e, role, qb, nlim, ncon, zb, zc, wr_e, nself);
  if (nrows > 0) bwd_subst(&sm.L0[0][e], zb);   // velocity change of the base: dv_base = L0^-T z_base
  __syncwarp();
  pf.mark(PF_PGS);
  // ---- velocity change of the chain (outward sweep of dv = W^T z) and integration (exponential map on the torso
  // quaternion, as btMultiBody::stepPositionsMultiDof)
  {
    const float M = (float)ILRL_MAX_COORD_VEL;
    float nu[6];
#pragma unroll
    for (int i = 0; i < 6; i++) nu[i] = nrows > 0 ? clampf(nub[i] + zb[i], -M, M) : nub[i];
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
    SV a0v, ap;
    a0v.a = mk(zb[0], zb[1], zb[2]); a0v.l = mk(zb[3], zb[4], zb[5]);
    ap = a0v;
#pragma unroll
    for (int c = 0; c < NL; c++) {
      float* rec = link_rec(sm, c, e, tid);
      const float nuc = rec[W_NU];
      float qd = nuc;
      if (nrows > 0) {
        if (c == 3 && role >= 2) ap = a0v;   // the arms hang off the torso
        SV S, U;
        float di, sq;
        ld_SUq(rec, S, U, di, sq);
        const float dq = sq * zc[c] - di * sdot(ap, U);
        ap = ap + dq * S;
        qd = clampf(nuc + dq, -M, M);
      }
      if (c >= 3 || role == 0) {  // the spine records are shared by the quad: one writer for the read-modify-write
        rec[W_QD] = qd;
        rec[W_Q] += dt * qd;
      }
    }
  }
  __syncwarp(qm);
  pf.mark(PF_INTEG);
  return nrows;   // (the env's constraint rows of this substep: the cost key of the env-to-warp grouping)
}

// ---- state movement between HBM (one row of ILRL_PHYS_STRIDE words per env), the link records and a replicated full Phys
__device__ __forceinline__ void load_base(const float* phys, int n, int i, Base& b) {
  const float* p = phys + (size_t)i * ILRL_PHYS_STRIDE;
#pragma unroll
  for (int k = 0; k < 3; k++) { b.p[k] = p[k]; b.v[k] = p[7 + k]; b.w[k] = p[10 + k]; }
#pragma unroll
  for (int k = 0; k < 4; k++) b.quat[k] = p[3 + k];
}
// benign state for the lanes of an env that does not step: upright at rest, zero joint state and torques
template <class SM>
__device__ __forceinline__ void dummy_state(SM& sm, int e, int tid, Base& b) {
  b.p[0] = 0.f; b.p[1] = 0.f; b.p[2] = 1.4f;
  b.quat[0] = b.quat[1] = b.quat[2] = 0.f; b.quat[3] = 1.f;
#pragma unroll
  for (int k = 0; k < 3; k++) { b.v[k] = 0.f; b.w[k] = 0.f; }
#pragma unroll
  for (int c = 0; c < NL; c++) {
    float* rec = link_rec(sm, c, e, tid);
    rec[W_Q] = 0.f; rec[W_QD] = 0.f; rec[W_TAU] = 0.f;
  }
}
// joint state of this lane's chain from HBM into registers / from registers into the link records (spine: every lane
// writes the same values).  Two steps so that the loads of a tile's head are all in flight before the first is consumed.
template <class SM>
__device__ __forceinline__ void load_links(const float* phys, int n, int i, const SM& sm, int role, float* qv, float* qdv) {
  const float* p = phys + (size_t)i * ILRL_PHYS_STRIDE;
#pragma unroll
  for (int c = 0; c < NL; c++) {
    const int j = c < 3 ? c : link_c(sm, role, c - 3).j;
    qv[c] = j >= 0 ? p[13 + j] : 0.f;
    qdv[c] = j >= 0 ? p[30 + j] : 0.f;
  }
}
template <class SM>
__device__ __forceinline__ void store_links(SM& sm, int e, int tid, const float* qv, const float* qdv) {
#pragma unroll
  for (int c = 0; c < NL; c++) {
    float* rec = link_rec(sm, c, e, tid);
    rec[W_Q] = qv[c];
    rec[W_QD] = qdv[c];
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
  float* p = phys + (size_t)i * ILRL_PHYS_STRIDE;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (((0 + k) & 3) == role) p[0 + k] = ps.p[k];
    if (((7 + k) & 3) == role) p[7 + k] = ps.v[k];
    if (((10 + k) & 3) == role) p[10 + k] = ps.w[k];
  }
#pragma unroll
  for (int k = 0; k < 4; k++) if (((3 + k) & 3) == role) p[3 + k] = ps.quat[k];
#pragma unroll
  for (int k = 0; k < NJ; k++) {
    if (((13 + k) & 3) == role) p[13 + k] = ps.q[k];
    if (((30 + k) & 3) == role) p[30 + k] = ps.qd[k];
  }
}

}  // namespace chain
}  // namespace ilrl
