// ilrl_quad.cuh — second kernel generation: FOUR LANES PER ENV ("quad"), 8 envs per warp.
//
// Why (profiles/r1_v1_step_kernel_ncu.md): with one thread per env the headline workload (4096 envs) occupies 128
// warps of a 592-scheduler part, each issuing 0.32 IPC down one long dependent chain through 17 links, with the link
// scratch in local memory; at large N that scratch (9 KB / env) turns into DRAM traffic.  The quad kernel
//   * gives each of the four limbs (right leg, left leg, right arm, left arm) to one lane; the 3 spine links and the
//     floating base are computed redundantly by all four lanes (no broadcast, no divergence), so the sequential
//     chain is 3 + 4 links instead of 17 and everything of a limb lives in that lane's registers;
//   * combines the limbs where the tree joins (legs at the pelvis, arms + spine at the torso) with 2 rounds of
//     __shfl_xor over the 27 words of an articulated inertia + bias force;
//   * publishes per-link (S, U, 1/D) to shared memory once per substep, so that ANY lane can compute the response of
//     the generalized velocities to a unit impulse: constraint rows are dealt round-robin to the 4 lanes;
//   * runs the projected Gauss-Seidel sweeps with the velocity change distributed like the state (base + spine
//     replicated, limb slices private), one 4-lane reduction per row.
// Row order, row formulas and constants are exactly those of the one-thread-per-env kernel and of the oracle.
#pragma once
#include "ilrl_env.cuh"

namespace ilrl {
namespace quad {

#ifndef ILRL_QE
#define ILRL_QE 16
#endif
#ifndef ILRL_RSM
#define ILRL_RSM 16
#endif
constexpr int QE = ILRL_QE;   // envs per CTA
constexpr int QT = 4 * QE;    // threads per CTA
constexpr int RSM = ILRL_RSM; // constraint rows per env kept in shared memory (the rest overflow to global scratch)
constexpr int GROWS = MAXROWS - RSM;  // rows per env in the global overflow scratch
constexpr int SL = 25;        // slice layout of a generalized vector: base 6 | spine 3 | 4 limbs x 4 slots
constexpr int ROWW = 2 * SL;  // a stored row: response slices + Jacobian slices
constexpr int LW = 13;        // per-link record: S(6) U(6) 1/D

__device__ __forceinline__ int slice_of_joint(int j) { return j + 6 + (j >= 11) + (j >= 14); }

// ---- per-role (= per-limb) constants.  Slots 0..2 are the joints of body A (thigh / upper arm: the arms have a
// leading dummy slot with a zero axis), slot 3 is the joint of body B (shin / lower arm); E is the rigid end body.
struct Role {
  float attach[3];
  float ax[4][3];
  float ancA[3], posB[3], ancB[3], posE[3];
  float lo[4], hi[4];
  float mA, iA[3], mB, iB[3], mE, iE[3];
  float sph[5][4];   // local spheres 0 = E, 1,2 = B, 3,4 = A: centre (body frame), radius
  float gear[4];     // torque per unit of clipped action
  float mw[4], mwv[4];  // imitation reward weights of the slot's joint (REF low_level_env.py:103-137)
  int j[4];          // global joint index (-1: dummy)
  int motor[4];      // action slot that drives the joint (humanoid.py:28-37)
  int mcol[4];       // CSV column of the joint
  int mpos[4];       // position of the joint in joint_map order (obs tail)
  int sidx[5];       // global sphere index
  int pelvis;        // 1: the limb hangs off the pelvis (legs), 0: off the torso (arms)
  int nreal;         // real joints among slots 0..2
  int pad;
};
constexpr int ROLE_WORDS = sizeof(Role) / 4;

struct RoleTable {
  Role r[4];
  constexpr RoleTable() : r() {
    constexpr int jb[NJ] = ILRL_JOINT_BODY;
    constexpr double bp[NB * 3] = ILRL_BODY_POS;
    constexpr double bm[NB] = ILRL_BODY_MASS;
    constexpr double bi[NB * 3] = ILRL_BODY_INERTIA;
    constexpr int bpar[NB] = ILRL_BODY_PARENT;
    constexpr double ja[NJ * 3] = ILRL_JOINT_ANCHOR;
    constexpr double jx[NJ * 3] = ILRL_JOINT_AXIS;
    constexpr double jlo[NJ] = ILRL_JOINT_LO;
    constexpr double jhi[NJ] = ILRL_JOINT_HI;
    constexpr int sb[NS] = ILRL_SPHERE_BODY;
    constexpr double sc[NS * 3] = ILRL_SPHERE_C;
    constexpr double sr[NS] = ILRL_SPHERE_R;
    constexpr int mj[NJ] = ILRL_MOTOR_JOINT;
    constexpr double mg[NJ] = ILRL_MOTOR_GEAR;
    constexpr int pj[NMAP] = ILRL_MAP_JOINT;
    constexpr int pc[NMAP] = ILRL_MAP_COL;
    constexpr double pw[NMAP] = ILRL_MAP_W;
    constexpr double pwv[NMAP] = ILRL_MAP_WV;
    constexpr int first[4] = {3, 7, 11, 14};
    constexpr int count[4] = {4, 4, 3, 3};
    for (int q = 0; q < 4; q++) {
      Role& o = r[q];
      const int lead = 4 - count[q];  // dummy slots in front
      const int bA = jb[first[q]], bB = jb[first[q] + count[q] - 1], bE = bB + 1;
      o.pelvis = bpar[bA] == 2 ? 1 : 0;
      o.nreal = 3 - lead;
      o.pad = 0;
      for (int i = 0; i < 3; i++) {
        o.attach[i] = (float)bp[3 * bA + i]; o.posB[i] = (float)bp[3 * bB + i]; o.posE[i] = (float)bp[3 * bE + i];
        o.ancA[i] = (float)ja[3 * first[q] + i]; o.ancB[i] = (float)ja[3 * (first[q] + count[q] - 1) + i];
        o.iA[i] = (float)bi[3 * bA + i]; o.iB[i] = (float)bi[3 * bB + i]; o.iE[i] = (float)bi[3 * bE + i];
      }
      o.mA = (float)bm[bA]; o.mB = (float)bm[bB]; o.mE = (float)bm[bE];
      for (int k = 0; k < 4; k++) {
        const int j = k < lead ? -1 : first[q] + (k - lead);
        o.j[k] = j;
        o.motor[k] = -1; o.mcol[k] = 0; o.mpos[k] = -1; o.mw[k] = 0.f; o.mwv[k] = 0.f; o.gear[k] = 0.f;
        for (int i = 0; i < 3; i++) o.ax[k][i] = j < 0 ? 0.f : (float)jx[3 * j + i];
        o.lo[k] = j < 0 ? -1e30f : (float)jlo[j];
        o.hi[k] = j < 0 ? 1e30f : (float)jhi[j];
        if (j >= 0) {
          for (int m = 0; m < NJ; m++) if (mj[m] == j) { o.motor[k] = m; o.gear[k] = (float)mg[m]; }
          for (int m = 0; m < NMAP; m++) if (pj[m] == j) { o.mcol[k] = pc[m]; o.mpos[k] = m; o.mw[k] = (float)pw[m]; o.mwv[k] = (float)pwv[m]; }
        }
      }
      int nb_ = 0, na_ = 0;
      for (int s = 0; s < NS; s++) {
        int loc = -1;
        if (sb[s] == bE) loc = 0;
        else if (sb[s] == bB) loc = 1 + nb_++;
        else if (sb[s] == bA) loc = 3 + na_++;
        if (loc >= 0) {
          o.sidx[loc] = s;
          for (int i = 0; i < 3; i++) o.sph[loc][i] = (float)sc[3 * s + i];
          o.sph[loc][3] = (float)sr[s];
        }
      }
    }
  }
};
__device__ constexpr RoleTable kRoles{};

// ---- shared memory of one CTA
struct Smem {
  uint32_t role[4][ROLE_WORDS];
  float link[NJ][LW][QE];        // per-link S, U, 1/D of every env (env-minor: conflict-free across the 8 envs of a warp)
  float nu[QE][SL];              // unconstrained new velocities in slice layout
  float qj[QE][NJ];              // joint positions (limit rows may be built by any lane)
  float qdj[QE][NJ];             // joint velocities (gather / scatter between the quad layout and a full Phys)
  float sph[QE][NS][4];          // contact point (relative to the torso origin) and distance of every sphere
  float lam[QE][MAXROWS];
  float rhs[QE][MAXROWS];
  float dinv[QE][MAXROWS];
  float rows[QE][RSM][ROWW];
  float scr[QT][3][NJ];          // per-lane response scratch: u of the chain links for up to 3 impulses (kept zero)
  float act[QE][NJ];             // actions (motor order) / staging
  float obs[QE][71];
};

struct QState {
  float p[3], quat[4], v[3], w[3];  // replicated in the 4 lanes
  float qs[3], qds[3];              // spine joints (replicated)
  float ql[4], qdl[4];              // limb slots (private)
};

__device__ __forceinline__ float qsum(float v, unsigned qm) {
  v += __shfl_xor_sync(qm, v, 1);
  v += __shfl_xor_sync(qm, v, 2);
  return v;
}
__device__ __forceinline__ SV neg(SV a) { SV r; r.a = mk(-a.a.x, -a.a.y, -a.a.z); r.l = mk(-a.l.x, -a.l.y, -a.l.z); return r; }
__device__ __forceinline__ SV svzero() { SV r; r.a = r.l = mk(0, 0, 0); return r; }
__device__ __forceinline__ V3 rd3(const float* p) { return mk(p[0], p[1], p[2]); }

// one hinge: anchor/axis given in the current frame (Rc, oc); updates the frame, returns S and the anchor
__device__ __forceinline__ void joint_xf(float* Rc, V3& oc, V3 an, V3 ax, float q, SV& S, V3& rw) {
  rw = oc + mv(Rc, an);
  V3 aw = mv(Rc, ax);
  S.a = aw; S.l = cross(rw, aw);
  float sn, cs;
  sincosf(q, &sn, &cs);
  float t = 1.f - cs, Rj[9], Rn[9];
  Rj[0] = t * ax.x * ax.x + cs;        Rj[1] = t * ax.x * ax.y - sn * ax.z; Rj[2] = t * ax.x * ax.z + sn * ax.y;
  Rj[3] = t * ax.x * ax.y + sn * ax.z; Rj[4] = t * ax.y * ax.y + cs;        Rj[5] = t * ax.y * ax.z - sn * ax.x;
  Rj[6] = t * ax.x * ax.z - sn * ax.y; Rj[7] = t * ax.y * ax.z + sn * ax.x; Rj[8] = t * ax.z * ax.z + cs;
  mm(Rc, Rj, Rn);
  oc = rw - mv(Rn, an);
#pragma unroll
  for (int i = 0; i < 9; i++) Rc[i] = Rn[i];
}

// kinematics of one lane: torso + spine (replicated) + its limb
struct QKin {
  float R0[9], R1[9], R2[9], RA[9], RB[9];
  V3 o1, o2, oA, oB, oE;
  SV Ss[3], Sl[4];
  float sx, sy;  // partial sums of part origins: spine part (replicated) is in ssx/ssy, limb part in sx/sy
  float ssx, ssy;
};

__device__ __forceinline__ void qfk(const QState& s, const Role& rc, QKin& k) {
  quat2mat(s.quat[0], s.quat[1], s.quat[2], s.quat[3], k.R0);
  V3 rw;
  // lwaist (body 1): joints 0 (abdomen_z), 1 (abdomen_y)
  {
    float Q[9];
    quat2mat(kBodyQuat[4], kBodyQuat[5], kBodyQuat[6], kBodyQuat[7], Q);
    mm(k.R0, Q, k.R1);
    V3 oc = mv(k.R0, mk(kBodyPos[3], kBodyPos[4], kBodyPos[5]));
    joint_xf(k.R1, oc, mk(kJointAnchor[0], kJointAnchor[1], kJointAnchor[2]), mk(kJointAxis[0], kJointAxis[1], kJointAxis[2]), s.qs[0], k.Ss[0], rw);
    k.ssx = rw.x; k.ssy = rw.y;
    joint_xf(k.R1, oc, mk(kJointAnchor[3], kJointAnchor[4], kJointAnchor[5]), mk(kJointAxis[3], kJointAxis[4], kJointAxis[5]), s.qs[1], k.Ss[1], rw);
    k.ssx += rw.x + oc.x; k.ssy += rw.y + oc.y;
    k.o1 = oc;
  }
  // pelvis (body 2): joint 2 (abdomen_x)
  {
    float Q[9];
    quat2mat(kBodyQuat[8], kBodyQuat[9], kBodyQuat[10], kBodyQuat[11], Q);
    mm(k.R1, Q, k.R2);
    V3 oc = k.o1 + mv(k.R1, mk(kBodyPos[6], kBodyPos[7], kBodyPos[8]));
    joint_xf(k.R2, oc, mk(kJointAnchor[6], kJointAnchor[7], kJointAnchor[8]), mk(kJointAxis[6], kJointAxis[7], kJointAxis[8]), s.qs[2], k.Ss[2], rw);
    k.ssx += rw.x + oc.x; k.ssy += rw.y + oc.y;
    k.o2 = oc;
  }
  // limb: body A on its parent (pelvis or torso), 3 joint slots; body B, 1 joint slot; end body E
  {
    const bool pel = rc.pelvis != 0;
    V3 po = pel ? k.o2 : mk(0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 9; i++) k.RA[i] = pel ? k.R2[i] : k.R0[i];
    V3 oc = po + mv(k.RA, rd3(rc.attach));
    const V3 an = rd3(rc.ancA);
#pragma unroll
    for (int sl = 0; sl < 3; sl++) joint_xf(k.RA, oc, an, rd3(rc.ax[sl]), s.ql[sl], k.Sl[sl], rw);
    k.oA = oc;
    k.sx = (float)rc.nreal * rw.x + oc.x; k.sy = (float)rc.nreal * rw.y + oc.y;
#pragma unroll
    for (int i = 0; i < 9; i++) k.RB[i] = k.RA[i];
    oc = k.oA + mv(k.RA, rd3(rc.posB));
    joint_xf(k.RB, oc, rd3(rc.ancB), rd3(rc.ax[3]), s.ql[3], k.Sl[3], rw);
    k.oB = oc;
    k.oE = oc + mv(k.RB, rd3(rc.posE));
    k.sx += rw.x + oc.x + k.oE.x; k.sy += rw.y + oc.y + k.oE.y;
  }
}

struct IP { Inertia I; SV p; };  // articulated inertia + bias force (27 words)

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

// one inward ABA step on the running (IA, pA) of a chain; stores U, 1/D, u
__device__ __forceinline__ void inward(IP& x, SV S, SV cJ, float tau, SV& U, float& dinv, float& u) {
  U = imul(x.I, S);
  float D = sdot(S, U);
  dinv = D > 1e-9f ? 1.0f / D : 0.f;  // a dummy slot has S = 0
  u = tau - sdot(S, x.p);
  downdate(x.I, U, dinv);
  x.p = x.p + imul(x.I, cJ) + (u * dinv) * U;
}

__device__ __forceinline__ void st_link(Smem& sm, int j, int e, SV S, SV U, float dinv) {
  float* p = &sm.link[j][0][e];
  p[0 * QE] = S.a.x; p[1 * QE] = S.a.y; p[2 * QE] = S.a.z; p[3 * QE] = S.l.x; p[4 * QE] = S.l.y; p[5 * QE] = S.l.z;
  p[6 * QE] = U.a.x; p[7 * QE] = U.a.y; p[8 * QE] = U.a.z; p[9 * QE] = U.l.x; p[10 * QE] = U.l.y; p[11 * QE] = U.l.z;
  p[12 * QE] = dinv;
}
__device__ __forceinline__ SV ld_linkS(const Smem& sm, int j, int e) {
  const float* p = &sm.link[j][0][e];
  SV r; r.a = mk(p[0], p[QE], p[2 * QE]); r.l = mk(p[3 * QE], p[4 * QE], p[5 * QE]);
  return r;
}
__device__ __forceinline__ SV ld_linkU(const Smem& sm, int j, int e) {
  const float* p = &sm.link[j][6][e];
  SV r; r.a = mk(p[0], p[QE], p[2 * QE]); r.l = mk(p[3 * QE], p[4 * QE], p[5 * QE]);
  return r;
}

// pointer to the stored row r of env e (shared for r < RSM, global scratch beyond)
__device__ __forceinline__ float* row_ptr(Smem& sm, float* gscr, int e, int r) {
  return r < RSM ? &sm.rows[e][r][0] : gscr + (size_t)(r - RSM) * ROWW;
}

// Response of the generalized velocities to NR unit impulses applied at the same place: spatial forces F[i] on link
// `link` (-1 = torso), or (jl >= 0, NR = 1) a generalized impulse of sign dirl on joint jl.  Uses the per-link
// records of env e in shared memory and the base factor L0 (replicated).  Writes response + Jacobian slices into
// row[i] and returns dd[i] = J.resp, rv[i] = J.nu.  sm.scr[tid] must be all-zero on entry and is left all-zero.
template <int NR>
__device__ __forceinline__ void responses(Smem& sm, int e, int tid, const float* L0, int link, const SV* F, int jl,
                                          float dirl, float* const* row, float* dd, float* rv) {
  float(*su)[NJ] = sm.scr[tid];
  const float* nu = sm.nu[e];
  SV pf[NR];
#pragma unroll
  for (int i = 0; i < NR; i++) {
    for (int t = 0; t < ROWW; t++) row[i][t] = 0.f;
  }
  int l0;
  if (jl >= 0) {
    su[0][jl] = dirl;
    pf[0] = (dirl * sm.link[jl][12][e]) * ld_linkU(sm, jl, e);
    rv[0] = dirl * nu[slice_of_joint(jl)];
    row[0][SL + slice_of_joint(jl)] = dirl;
    l0 = kJointParent[jl];
  } else {
#pragma unroll
    for (int i = 0; i < NR; i++) {
      pf[i] = neg(F[i]);
      rv[i] = F[i].a.x * nu[0] + F[i].a.y * nu[1] + F[i].a.z * nu[2] + F[i].l.x * nu[3] + F[i].l.y * nu[4] + F[i].l.z * nu[5];
      float* Jr = row[i] + SL;
      Jr[0] = F[i].a.x; Jr[1] = F[i].a.y; Jr[2] = F[i].a.z; Jr[3] = F[i].l.x; Jr[4] = F[i].l.y; Jr[5] = F[i].l.z;
    }
    l0 = link;
  }
  // inward along the chain: u_l = -S_l . pf, pf += U_l u_l / D_l;  contact Jacobian entry J_l = S_l . F
  for (int l = l0; l >= 0; l = kJointParent[l]) {
    SV S = ld_linkS(sm, l, e), U = ld_linkU(sm, l, e);
    const float di = sm.link[l][12][e];
    const int sl = slice_of_joint(l);
#pragma unroll
    for (int i = 0; i < NR; i++) {
      float u = -sdot(S, pf[i]);
      su[i][l] = u;
      pf[i] = pf[i] + (u * di) * U;
      if (jl < 0) {
        float Jl = sdot(S, F[i]);
        rv[i] += Jl * nu[sl];
        row[i][SL + sl] = Jl;
      }
    }
  }
  SV a0[NR], acur[NR], apel[NR];
#pragma unroll
  for (int i = 0; i < NR; i++) {
    a0[i] = chol6_solve(L0, neg(pf[i]));
    acur[i] = a0[i]; apel[i] = a0[i];
    float* r = row[i];
    r[0] = a0[i].a.x; r[1] = a0[i].a.y; r[2] = a0[i].a.z; r[3] = a0[i].l.x; r[4] = a0[i].l.y; r[5] = a0[i].l.z;
    dd[i] = jl >= 0 ? 0.f : sdot(a0[i], F[i]);
  }
  // outward over every link
#pragma unroll 1
  for (int j = 0; j < NJ; j++) {
    SV S = ld_linkS(sm, j, e), U = ld_linkU(sm, j, e);
    const float di = sm.link[j][12][e];
    const int p = kJointParent[j], sl = slice_of_joint(j);
#pragma unroll
    for (int i = 0; i < NR; i++) {
      SV ap = p < 0 ? a0[i] : (p == kPelvisLink ? apel[i] : acur[i]);
      float qa = di * (su[i][j] - sdot(ap, U));
      row[i][sl] = qa;
      acur[i] = ap + qa * S;
      if (j == kPelvisLink) apel[i] = acur[i];
    }
  }
  // J.resp over the chain entries, and restore the scratch to zero
  if (jl >= 0) { dd[0] += dirl * row[0][slice_of_joint(jl)]; su[0][jl] = 0.f; }
  for (int l = l0; l >= 0; l = kJointParent[l]) {
    const int sl = slice_of_joint(l);
#pragma unroll
    for (int i = 0; i < NR; i++) {
      if (jl < 0) dd[i] += row[i][SL + sl] * row[i][sl];
      su[i][l] = 0.f;
    }
  }
}

// distributed generalized vector: base + spine replicated, own limb private
struct QVec { float b[6], s[3], l[4]; };

__device__ __forceinline__ float row_jdot(const float* row, const QVec& x, int role, unsigned qm) {
  const float* J = row + SL;
  float rep = J[0] * x.b[0] + J[1] * x.b[1] + J[2] * x.b[2] + J[3] * x.b[3] + J[4] * x.b[4] + J[5] * x.b[5] +
              J[6] * x.s[0] + J[7] * x.s[1] + J[8] * x.s[2];
  const float* Jl = J + 9 + 4 * role;
  float mine = Jl[0] * x.l[0] + Jl[1] * x.l[1] + Jl[2] * x.l[2] + Jl[3] * x.l[3];
  return rep + qsum(mine, qm);
}
__device__ __forceinline__ void row_axpy(const float* row, float a, QVec& x, int role) {
#pragma unroll
  for (int i = 0; i < 6; i++) x.b[i] += a * row[i];
#pragma unroll
  for (int i = 0; i < 3; i++) x.s[i] += a * row[6 + i];
  const float* rl = row + 9 + 4 * role;
#pragma unroll
  for (int i = 0; i < 4; i++) x.l[i] += a * rl[i];
}

// ---- one substep for the env of this quad.  tau_s: spine torques (replicated), tau_l: limb slot torques.
__device__ __forceinline__ void qsubstep(QState& s, const float* tau_s, const float* tau_l, const Role& rc, Smem& sm,
                                         float* gscr, int e, int tid, int role, unsigned qm, float dt) {
  QKin k;
  qfk(s, rc, k);
  // ---- velocities and velocity-product accelerations
  SV V0; V0.a = rd3(s.w); V0.l = rd3(s.v);
  SV Vs[3], cs_[3], Vl[4], cl[4];
  {
    SV Vp = V0;
#pragma unroll
    for (int i = 0; i < 3; i++) { SV X = s.qds[i] * k.Ss[i]; cs_[i] = crm(Vp, X); Vs[i] = Vp + X; Vp = Vs[i]; }
    Vp = rc.pelvis ? Vs[2] : V0;
#pragma unroll
    for (int i = 0; i < 4; i++) { SV X = s.qdl[i] * k.Sl[i]; cl[i] = crm(Vp, X); Vl[i] = Vp + X; Vp = Vl[i]; }
  }
  // ---- inward pass: limb (private), then pelvis/torso joins by quad shuffles, spine + base replicated
  SV Ul[4], Us[3];
  float dil[4], ul[4], dis[3], us[3];
  IP x, t;
  rigid_inertia_bias(k.RB, k.oB, rc.mB, rc.iB[0], rc.iB[1], rc.iB[2], Vl[3], x.I, x.p);
  rigid_inertia_bias(k.RB, k.oE, rc.mE, rc.iE[0], rc.iE[1], rc.iE[2], Vl[3], t.I, t.p);
  x.I.add(t.I); x.p = x.p + t.p;
  inward(x, k.Sl[3], cl[3], tau_l[3], Ul[3], dil[3], ul[3]);
  rigid_inertia_bias(k.RA, k.oA, rc.mA, rc.iA[0], rc.iA[1], rc.iA[2], Vl[2], t.I, t.p);
  x.I.add(t.I); x.p = x.p + t.p;
  inward(x, k.Sl[2], cl[2], tau_l[2], Ul[2], dil[2], ul[2]);
  inward(x, k.Sl[1], cl[1], tau_l[1], Ul[1], dil[1], ul[1]);
  inward(x, k.Sl[0], cl[0], tau_l[0], Ul[0], dil[0], ul[0]);
  // lanes 0,1 (legs) and lanes 2,3 (arms) pair up; then each pair fetches the other pair's sum
  ip_shfl_add(x, qm, 1);
  IP y = ip_shfl_get(x, qm, 2);
  IP legs, arms;
  if (role < 2) { legs = x; arms = y; } else { legs = y; arms = x; }
  // spine link 2 (abdomen_x) carries the pelvis and both legs
  rigid_inertia_bias(k.R2, k.o2, kBodyMass[2], kBodyInertia[6], kBodyInertia[7], kBodyInertia[8], Vs[2], t.I, t.p);
  legs.I.add(t.I); legs.p = legs.p + t.p;
  inward(legs, k.Ss[2], cs_[2], tau_s[2], Us[2], dis[2], us[2]);
  rigid_inertia_bias(k.R1, k.o1, kBodyMass[1], kBodyInertia[3], kBodyInertia[4], kBodyInertia[5], Vs[1], t.I, t.p);
  legs.I.add(t.I); legs.p = legs.p + t.p;
  inward(legs, k.Ss[1], cs_[1], tau_s[1], Us[1], dis[1], us[1]);
  inward(legs, k.Ss[0], cs_[0], tau_s[0], Us[0], dis[0], us[0]);
  rigid_inertia_bias(k.R0, mk(0.f, 0.f, 0.f), kBodyMass[0], kBodyInertia[0], kBodyInertia[1], kBodyInertia[2], V0, t.I, t.p);
  t.I.add(legs.I); t.I.add(arms.I); t.p = t.p + legs.p + arms.p;
  float L0[21];
  chol6(t.I, L0);
  SV a0 = chol6_solve(L0, neg(t.p));
  // ---- outward pass: accelerations -> unconstrained new velocities
  QVec nu;
  {
    V3 lin = a0.l + cross(V0.a, V0.l);
    nu.b[0] = s.w[0] + dt * a0.a.x; nu.b[1] = s.w[1] + dt * a0.a.y; nu.b[2] = s.w[2] + dt * a0.a.z;
    nu.b[3] = s.v[0] + dt * lin.x;  nu.b[4] = s.v[1] + dt * lin.y;  nu.b[5] = s.v[2] + dt * lin.z;
    SV ap = a0, apel;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      SV ad = ap + cs_[i];
      float qa = dis[i] * (us[i] - sdot(ad, Us[i]));
      nu.s[i] = s.qds[i] + dt * qa;
      ap = ad + qa * k.Ss[i];
    }
    apel = ap;
    ap = rc.pelvis ? apel : a0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      SV ad = ap + cl[i];
      float qa = dil[i] * (ul[i] - sdot(ad, Ul[i]));
      nu.l[i] = s.qdl[i] + dt * qa;
      ap = ad + qa * k.Sl[i];
    }
    const float M = (float)ILRL_MAX_COORD_VEL;
#pragma unroll
    for (int i = 0; i < 6; i++) nu.b[i] = clampf(nu.b[i], -M, M);
#pragma unroll
    for (int i = 0; i < 3; i++) nu.s[i] = clampf(nu.s[i], -M, M);
#pragma unroll
    for (int i = 0; i < 4; i++) nu.l[i] = clampf(nu.l[i], -M, M);
  }
  // ---- candidates: violated joint limits (17-bit mask), ground spheres (29-bit mask)
  uint32_t lim = 0, act = 0;
#pragma unroll
  for (int i = 0; i < 4; i++)
    if (rc.j[i] >= 0 && (s.ql[i] - rc.lo[i] <= 0.f || rc.hi[i] - s.ql[i] <= 0.f)) lim |= 1u << rc.j[i];
  if (role < 3) {
    float qv = role == 0 ? s.qs[0] : (role == 1 ? s.qs[1] : s.qs[2]);
    if (qv - kJointLo[role] <= 0.f || kJointHi[role] - qv <= 0.f) lim |= 1u << role;
  }
  // publish what any lane may need to build any row
  __syncwarp(qm);
#pragma unroll
  for (int i = 0; i < 4; i++)
    if (rc.j[i] >= 0) {
      st_link(sm, rc.j[i], e, k.Sl[i], Ul[i], dil[i]);
      sm.qj[e][rc.j[i]] = s.ql[i];
    }
#pragma unroll
  for (int i = 0; i < 4; i++) sm.nu[e][9 + 4 * role + i] = nu.l[i];
  if (role == 0) {
#pragma unroll
    for (int i = 0; i < 3; i++) { st_link(sm, i, e, k.Ss[i], Us[i], dis[i]); sm.qj[e][i] = s.qs[i]; sm.nu[e][6 + i] = nu.s[i]; }
#pragma unroll
    for (int i = 0; i < 6; i++) sm.nu[e][i] = nu.b[i];
  }
  // spheres: 5 of the limb, and (lane 0) the 9 of the trunk
  {
#pragma unroll
    for (int i = 0; i < 5; i++) {
      const float* R = i == 0 || i <= 2 ? k.RB : k.RA;
      V3 o = i == 0 ? k.oE : (i <= 2 ? k.oB : k.oA);
      V3 c = o + mv(R, rd3(rc.sph[i]));
      float rad = rc.sph[i][3], d = s.p[2] + c.z - rad;
      const int g = rc.sidx[i];
      if (d < (float)ILRL_CONTACT_BREAK) act |= 1u << g;
      float* sp = sm.sph[e][g];
      sp[0] = c.x; sp[1] = c.y; sp[2] = c.z - rad; sp[3] = d;
    }
    if (role == 0) {
#pragma unroll
      for (int g = 16; g < NS; g++) {
        if (g >= 20 && g < 24) continue;  // upper-arm spheres belong to the arm lanes
        const int b = kSphereBody[g];
        const float* R = b == 0 ? k.R0 : (b == 1 ? k.R1 : k.R2);
        V3 o = b == 0 ? mk(0.f, 0.f, 0.f) : (b == 1 ? k.o1 : k.o2);
        V3 c = o + mv(R, mk(kSphereC[3 * g], kSphereC[3 * g + 1], kSphereC[3 * g + 2]));
        float d = s.p[2] + c.z - kSphereR[g];
        if (d < (float)ILRL_CONTACT_BREAK) act |= 1u << g;
        float* sp = sm.sph[e][g];
        sp[0] = c.x; sp[1] = c.y; sp[2] = c.z - kSphereR[g]; sp[3] = d;
      }
    }
  }
  lim |= __shfl_xor_sync(qm, lim, 1); lim |= __shfl_xor_sync(qm, lim, 2);
  act |= __shfl_xor_sync(qm, act, 1); act |= __shfl_xor_sync(qm, act, 2);
  __syncwarp(qm);
  int nact = __popc(act);
  while (nact > MAXC) {  // keep the deepest MAXC (ties -> drop the highest index): replicated, reads shared memory
    int worst = -1;
    float wd = -1e30f;
    for (int g = 0; g < NS; g++)
      if (((act >> g) & 1u) && sm.sph[e][g][3] >= wd) { wd = sm.sph[e][g][3]; worst = g; }
    act &= ~(1u << worst);
    nact--;
  }
  const int nlim = __popc(lim), ncon = nact, nrows = nlim + 3 * ncon;
  if (nrows > 0) {
    const float idt = 1.0f / dt;
    // ---- build the rows, dealt round-robin to the 4 lanes: limit rows first, then one contact (3 rows) per item
    for (int it = role; it < nlim; it += 4) {
      const int j = __fns(lim, 0, it + 1);
      const float q = sm.qj[e][j];
      float pen, dir;
      if (q - kJointLo[j] <= 0.f) { pen = q - kJointLo[j]; dir = 1.f; } else { pen = kJointHi[j] - q; dir = -1.f; }
      float* row[1] = {row_ptr(sm, gscr, e, it)};
      float dd[1], rv[1];
      SV Fz[1]; Fz[0] = svzero();
      responses<1>(sm, e, tid, L0, -1, Fz, j, dir, row, dd, rv);
      float di = 1.0f / dd[0];
      sm.dinv[e][it] = di;
      sm.rhs[e][it] = (-pen * (float)ILRL_LIMIT_ERP * idt - rv[0]) * di;
      sm.lam[e][it] = 0.f;
    }
    for (int it = (role + 4 - (nlim & 3)) & 3; it < ncon; it += 4) {
      const int g = __fns(act, 0, it + 1);
      const float* sp = sm.sph[e][g];
      V3 xx = mk(sp[0], sp[1], sp[2]);
      const float dist = sp[3];
      SV F[3];
      F[0].l = mk(0.f, 0.f, 1.f); F[1].l = mk(0.f, -1.f, 0.f); F[2].l = mk(1.f, 0.f, 0.f);  // n, btPlaneSpace1 t1, t2
#pragma unroll
      for (int i = 0; i < 3; i++) F[i].a = cross(xx, F[i].l);
      const int r0 = nlim + 3 * it;
      float* row[3] = {row_ptr(sm, gscr, e, r0), row_ptr(sm, gscr, e, r0 + 1), row_ptr(sm, gscr, e, r0 + 2)};
      float dd[3], rv[3];
      responses<3>(sm, e, tid, L0, kSphereLink[g], F, -1, 0.f, row, dd, rv);
#pragma unroll
      for (int i = 0; i < 3; i++) {
        float di = 1.0f / dd[i];
        float pos = i == 0 ? (dist > 0.f ? -dist * idt : -dist * (float)ILRL_CONTACT_ERP * idt) : 0.f;
        sm.dinv[e][r0 + i] = di;
        sm.rhs[e][r0 + i] = (pos - rv[i]) * di;
        sm.lam[e][r0 + i] = 0.f;
      }
    }
    __syncwarp(qm);
    // ---- projected Gauss-Seidel, velocity change distributed like the state
    QVec dv;
#pragma unroll
    for (int i = 0; i < 6; i++) dv.b[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; i++) dv.s[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) dv.l[i] = 0.f;
#pragma unroll 1
    for (int itn = 0; itn < ILRL_SOLVER_ITERS; itn++) {
#pragma unroll 1
      for (int r = 0; r < nlim; r++) {
        const float* row = row_ptr(sm, gscr, e, r);
        float lam = sm.lam[e][r];
        float nl = fmaxf(lam + sm.rhs[e][r] - row_jdot(row, dv, role, qm) * sm.dinv[e][r], 0.f);
        sm.lam[e][r] = nl;
        row_axpy(row, nl - lam, dv, role);
      }
#pragma unroll 1
      for (int c = 0; c < ncon; c++) {
        const int r = nlim + 3 * c;
        const float* row = row_ptr(sm, gscr, e, r);
        float lam = sm.lam[e][r];
        float nl = fmaxf(lam + sm.rhs[e][r] - row_jdot(row, dv, role, qm) * sm.dinv[e][r], 0.f);
        sm.lam[e][r] = nl;
        row_axpy(row, nl - lam, dv, role);
      }
#pragma unroll 1
      for (int c = 0; c < ncon; c++) {
        const int rn = nlim + 3 * c;
        const float ln = sm.lam[e][rn];
        if (!(ln > 0.f)) continue;
        const float* r1 = row_ptr(sm, gscr, e, rn + 1);
        const float* r2 = row_ptr(sm, gscr, e, rn + 2);
        const float lim_f = (float)ILRL_FRICTION * ln;
        float l1 = sm.lam[e][rn + 1], l2 = sm.lam[e][rn + 2];
        float s1 = l1 + sm.rhs[e][rn + 1] - row_jdot(r1, dv, role, qm) * sm.dinv[e][rn + 1];
        float s2 = l2 + sm.rhs[e][rn + 2] - row_jdot(r2, dv, role, qm) * sm.dinv[e][rn + 2];
        float n2 = s1 * s1 + s2 * s2;
        if (n2 > lim_f * lim_f) { float sc = lim_f * rsqrtf(n2); s1 *= sc; s2 *= sc; }
        sm.lam[e][rn + 1] = s1; sm.lam[e][rn + 2] = s2;
        row_axpy(r1, s1 - l1, dv, role);
        row_axpy(r2, s2 - l2, dv, role);
      }
    }
    const float M = (float)ILRL_MAX_COORD_VEL;
#pragma unroll
    for (int i = 0; i < 6; i++) nu.b[i] = clampf(nu.b[i] + dv.b[i], -M, M);
#pragma unroll
    for (int i = 0; i < 3; i++) nu.s[i] = clampf(nu.s[i] + dv.s[i], -M, M);
#pragma unroll
    for (int i = 0; i < 4; i++) nu.l[i] = clampf(nu.l[i] + dv.l[i], -M, M);
    __syncwarp(qm);
  }
  // ---- integrate
#pragma unroll
  for (int i = 0; i < 3; i++) { s.w[i] = nu.b[i]; s.v[i] = nu.b[3 + i]; s.p[i] += dt * nu.b[3 + i]; }
  {
    float wn = sqrtf(nu.b[0] * nu.b[0] + nu.b[1] * nu.b[1] + nu.b[2] * nu.b[2]), sc, cw;
    if (wn < 1e-3f) sc = 0.5f * dt - dt * dt * dt * 0.020833333333f * wn * wn;
    else sc = sinf(0.5f * wn * dt) / wn;
    cw = cosf(0.5f * wn * dt);
    float dx = nu.b[0] * sc, dy = nu.b[1] * sc, dz = nu.b[2] * sc;
    float x_ = s.quat[0], y_ = s.quat[1], z_ = s.quat[2], w_ = s.quat[3];
    float nx = cw * x_ + dx * w_ + dy * z_ - dz * y_;
    float ny = cw * y_ - dx * z_ + dy * w_ + dz * x_;
    float nz = cw * z_ + dx * y_ - dy * x_ + dz * w_;
    float nw = cw * w_ - dx * x_ - dy * y_ - dz * z_;
    float inv = rsqrtf(nx * nx + ny * ny + nz * nz + nw * nw);
    s.quat[0] = nx * inv; s.quat[1] = ny * inv; s.quat[2] = nz * inv; s.quat[3] = nw * inv;
  }
#pragma unroll
  for (int i = 0; i < 3; i++) { s.qds[i] = nu.s[i]; s.qs[i] += dt * nu.s[i]; }
#pragma unroll
  for (int i = 0; i < 4; i++) { s.qdl[i] = nu.l[i]; s.ql[i] += dt * nu.l[i]; }
}

// ---- state movement between HBM (SoA phys[47][n]), the quad layout and a replicated full Phys
__device__ __forceinline__ void qload(const float* phys, int n, int i, const Role& rc, QState& s) {
  const float* p = phys + i;
#pragma unroll
  for (int k = 0; k < 3; k++) { s.p[k] = p[k * n]; s.v[k] = p[(7 + k) * n]; s.w[k] = p[(10 + k) * n]; }
#pragma unroll
  for (int k = 0; k < 4; k++) s.quat[k] = p[(3 + k) * n];
#pragma unroll
  for (int k = 0; k < 3; k++) { s.qs[k] = p[(13 + k) * n]; s.qds[k] = p[(30 + k) * n]; }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int j = rc.j[k];
    s.ql[k] = j >= 0 ? p[(13 + j) * n] : 0.f;
    s.qdl[k] = j >= 0 ? p[(30 + j) * n] : 0.f;
  }
}
// quad layout -> full Phys in every lane (through shared memory)
__device__ __forceinline__ void qgather(const QState& s, const Role& rc, Smem& sm, int e, int role, unsigned qm, Phys& ps) {
  __syncwarp(qm);
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (rc.j[k] >= 0) { sm.qj[e][rc.j[k]] = s.ql[k]; sm.qdj[e][rc.j[k]] = s.qdl[k]; }
  if (role == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) { sm.qj[e][k] = s.qs[k]; sm.qdj[e][k] = s.qds[k]; }
  }
  __syncwarp(qm);
#pragma unroll
  for (int j = 0; j < NJ; j++) { ps.q[j] = sm.qj[e][j]; ps.qd[j] = sm.qdj[e][j]; }
#pragma unroll
  for (int k = 0; k < 3; k++) { ps.p[k] = s.p[k]; ps.v[k] = s.v[k]; ps.w[k] = s.w[k]; }
#pragma unroll
  for (int k = 0; k < 4; k++) ps.quat[k] = s.quat[k];
}
// full (replicated) Phys -> quad layout
__device__ __forceinline__ void qscatter(const Phys& ps, const Role& rc, Smem& sm, int e, int role, unsigned qm, QState& s) {
  __syncwarp(qm);
  if (role == 0) {
#pragma unroll
    for (int j = 0; j < NJ; j++) { sm.qj[e][j] = ps.q[j]; sm.qdj[e][j] = ps.qd[j]; }
  }
  __syncwarp(qm);
#pragma unroll
  for (int k = 0; k < 3; k++) { s.p[k] = ps.p[k]; s.v[k] = ps.v[k]; s.w[k] = ps.w[k]; s.qs[k] = ps.q[k]; s.qds[k] = ps.qd[k]; }
#pragma unroll
  for (int k = 0; k < 4; k++) s.quat[k] = ps.quat[k];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int j = rc.j[k];
    s.ql[k] = j >= 0 ? sm.qj[e][j] : 0.f;
    s.qdl[k] = j >= 0 ? sm.qdj[e][j] : 0.f;
  }
}
// kinematics of the current pose: sums of the 31 part offsets and the right-foot origin (role 0's end body)
__device__ __forceinline__ void qpose_sums(const QState& s, const Role& rc, unsigned qm, float& sumx, float& sumy,
                                           float& rfx, float& rfy) {
  QKin k;
  qfk(s, rc, k);
  sumx = k.ssx + qsum(k.sx, qm);
  sumy = k.ssy + qsum(k.sy, qm);
  const int l0 = (threadIdx.x & 31) & ~3;
  rfx = __shfl_sync(qm, k.oE.x, l0);
  rfy = __shfl_sync(qm, k.oE.y, l0);
}
// store a replicated Phys: the 47 words are dealt to the 4 lanes
__device__ __forceinline__ void qstore_phys(float* phys, int n, int i, int role, const Phys& ps) {
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

}  // namespace quad
}  // namespace ilrl
