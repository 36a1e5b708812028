// ilrl_physics.cuh — per-env rigid-body substep for the humanoid of REF humanoid_symmetric_2.xml, sm_100a, fp32.
//
// Replaces, per env, what the reference reaches through `flat_env.scene.global_step()`
// (REF low_level_env.py:479, hier_env.py:589) -> pybullet stepSimulation -> Bullet btMultiBody:
//   forward dynamics by Featherstone's articulated-body algorithm (floating torso + 17 hinges), torque actuation,
//   Bullet's velocity damping, ground-plane contact with friction and violated-joint-limit rows solved by 5
//   projected-Gauss-Seidel sweeps on the velocities, semi-implicit integration.  Constants: ilrl_constants.h.
//
// Formulation (chosen for the GPU, not Bullet's link-local one): every spatial vector of one env is expressed in
// WORLD axes about ONE reference point, the torso origin at the start of the substep.  Parent->child spatial
// transforms are then the identity: the inward pass is plain additions and rank-1 downdates of symmetric 6x6
// matrices, the outward pass plain additions — no per-link 6x6 congruence transforms, and fp32 stays accurate
// however far the robot has walked because only offsets from the torso enter.
//
// This header holds the shared vocabulary: model tables, small vector / spatial-inertia helpers, the thread-per-env
// forward kinematics used by the service kernels (reset, high-level step, end-point score) and the base Cholesky.
// The substep itself (four lanes per env, rolled chain loops) lives in ilrl_chain.cuh.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "ilrl_constants.h"
#include "ilrl_model_data.h"

namespace ilrl {

constexpr int NB = ILRL_NB, NJ = ILRL_NJ, NS = ILRL_NS, NV = 6 + ILRL_NJ, NMAP = ILRL_NMAP;
constexpr int MAXC = ILRL_MAX_CONTACTS, MAXROWS = ILRL_MAX_ROWS;

// ---- model tables (generated from the reference MJCF by tools/gen_model.py)
__device__ constexpr int kBodyParent[NB] = ILRL_BODY_PARENT;
__device__ constexpr int kBodyLink[NB] = ILRL_BODY_LINK;
__device__ constexpr float kBodyPos[NB * 3] = ILRL_BODY_POS;
__device__ constexpr float kBodyQuat[NB * 4] = ILRL_BODY_QUAT;
__device__ constexpr float kBodyMass[NB] = ILRL_BODY_MASS;
__device__ constexpr float kBodyInertia[NB * 3] = ILRL_BODY_INERTIA;
__device__ constexpr int kJointBody[NJ] = ILRL_JOINT_BODY;
__device__ constexpr int kJointParent[NJ] = ILRL_JOINT_PARENT;
__device__ constexpr float kJointAnchor[NJ * 3] = ILRL_JOINT_ANCHOR;
__device__ constexpr float kJointAxis[NJ * 3] = ILRL_JOINT_AXIS;
__device__ constexpr float kJointLo[NJ] = ILRL_JOINT_LO;
__device__ constexpr float kJointHi[NJ] = ILRL_JOINT_HI;
__device__ constexpr int kSphereBody[NS] = ILRL_SPHERE_BODY;
__device__ constexpr int kSphereLink[NS] = ILRL_SPHERE_LINK;
__device__ constexpr float kSphereC[NS * 3] = ILRL_SPHERE_C;
__device__ constexpr float kSphereR[NS] = ILRL_SPHERE_R;

// ---- derived tree tables (compile-time) and the loop policy for per-link loops
// ILRL_UNROLL_LINKS = 1: loops over links/bodies are fully unrolled (tables fold into immediates, ~370 KB of SASS);
//                     0: they stay rolled and index the tables at run time (compact code that fits the instruction
//                        caches; the index is warp-uniform).  See DESIGN.md "Kernel generations".
#ifndef ILRL_UNROLL_LINKS
#define ILRL_UNROLL_LINKS 1
#endif
#if ILRL_UNROLL_LINKS
#define ILRL_LINK_LOOP _Pragma("unroll")
#else
#define ILRL_LINK_LOOP _Pragma("unroll 1")
#endif
struct TreeTables {
  int bodyJ0[NB], bodyNJ[NB];   // first joint / number of joints of a body
  int linkB0[NJ], linkB1[NJ];   // bodies carried by a link (-1: none)
  int leaf[NJ];                 // link has no child link
  int bodyIdent[NB];            // body frame has no fixed rotation relative to its parent
  constexpr TreeTables() : bodyJ0(), bodyNJ(), linkB0(), linkB1(), leaf(), bodyIdent() {
    constexpr int jb[NJ] = ILRL_JOINT_BODY;
    constexpr int jp[NJ] = ILRL_JOINT_PARENT;
    constexpr int bl[NB] = ILRL_BODY_LINK;
    constexpr double bq[NB * 4] = ILRL_BODY_QUAT;
    for (int b = 0; b < NB; b++) {
      bodyJ0[b] = 0; bodyNJ[b] = 0;
      for (int j = NJ - 1; j >= 0; j--) if (jb[j] == b) { bodyJ0[b] = j; bodyNJ[b]++; }
      bodyIdent[b] = (bq[4 * b] == 0.0 && bq[4 * b + 1] == 0.0 && bq[4 * b + 2] == 0.0) ? 1 : 0;
    }
    for (int j = 0; j < NJ; j++) {
      linkB0[j] = -1; linkB1[j] = -1; leaf[j] = 1;
      for (int b = 1; b < NB; b++) if (bl[b] == j) { if (linkB0[j] < 0) linkB0[j] = b; else linkB1[j] = b; }
      for (int c = 0; c < NJ; c++) if (jp[c] == j) leaf[j] = 0;
    }
  }
};
__device__ constexpr TreeTables kTree{};

// ---- small vector helpers
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
// row-major 3x3
__device__ __forceinline__ V3 mv(const float* R, V3 v) {
  return mk(R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z);
}
__device__ __forceinline__ V3 mtv(const float* R, V3 v) {
  return mk(R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z, R[2] * v.x + R[5] * v.y + R[8] * v.z);
}
__device__ __forceinline__ void mm(const float* A, const float* B, float* O) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) O[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
__device__ __forceinline__ void quat2mat(float x, float y, float z, float w, float* R) {
  float s = 2.0f / (x * x + y * y + z * z + w * w);
  R[0] = 1 - s * (y * y + z * z); R[1] = s * (x * y - w * z);     R[2] = s * (x * z + w * y);
  R[3] = s * (x * y + w * z);     R[4] = 1 - s * (x * x + z * z); R[5] = s * (y * z - w * x);
  R[6] = s * (x * z - w * y);     R[7] = s * (y * z + w * x);     R[8] = 1 - s * (x * x + y * y);
}

// spatial vectors: motion (w, v) / force (n, f), world axes, about the reference point
struct SV { V3 a, l; };  // angular part, linear part
__device__ __forceinline__ SV operator+(SV p, SV q) { SV r; r.a = p.a + q.a; r.l = p.l + q.l; return r; }
__device__ __forceinline__ SV operator*(float s, SV p) { SV r; r.a = s * p.a; r.l = s * p.l; return r; }
__device__ __forceinline__ float sdot(SV m, SV f) { return dot(m.a, f.a) + dot(m.l, f.l); }
__device__ __forceinline__ SV crm(SV v, SV x) { SV r; r.a = cross(v.a, x.a); r.l = cross(v.a, x.l) + cross(v.l, x.a); return r; }
__device__ __forceinline__ SV crf(SV v, SV f) { SV r; r.a = cross(v.a, f.a) + cross(v.l, f.l); r.l = cross(v.a, f.l); return r; }
__device__ __forceinline__ SV ldsv(const float* p) { SV r; r.a = ld3(p); r.l = ld3(p + 3); return r; }
__device__ __forceinline__ void stsv(float* p, SV v) { st3(p, v.a); st3(p + 3, v.l); }

// symmetric spatial inertia in 3x3 blocks: n = A w + B v, f = B^T w + C v.  A, C symmetric (xx,xy,xz,yy,yz,zz)
struct Inertia {
  float A[6], B[9], C[6];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 6; i++) { A[i] = 0; C[i] = 0; }
#pragma unroll
    for (int i = 0; i < 9; i++) B[i] = 0;
  }
  __device__ __forceinline__ void add(const Inertia& o) {
#pragma unroll
    for (int i = 0; i < 6; i++) { A[i] += o.A[i]; C[i] += o.C[i]; }
#pragma unroll
    for (int i = 0; i < 9; i++) B[i] += o.B[i];
  }
};
__device__ __forceinline__ V3 symv(const float* S, V3 v) {
  return mk(S[0] * v.x + S[1] * v.y + S[2] * v.z, S[1] * v.x + S[3] * v.y + S[4] * v.z, S[2] * v.x + S[4] * v.y + S[5] * v.z);
}
__device__ __forceinline__ SV imul(const Inertia& I, SV m) {
  SV r;
  r.a = symv(I.A, m.a) + mv(I.B, m.l);
  r.l = mtv(I.B, m.a) + symv(I.C, m.l);
  return r;
}
// I -= U U^T * dinv
__device__ __forceinline__ void downdate(Inertia& I, SV U, float dinv) {
  V3 ua = dinv * U.a, ul = dinv * U.l;
  I.A[0] -= ua.x * U.a.x; I.A[1] -= ua.x * U.a.y; I.A[2] -= ua.x * U.a.z;
  I.A[3] -= ua.y * U.a.y; I.A[4] -= ua.y * U.a.z; I.A[5] -= ua.z * U.a.z;
  I.B[0] -= ua.x * U.l.x; I.B[1] -= ua.x * U.l.y; I.B[2] -= ua.x * U.l.z;
  I.B[3] -= ua.y * U.l.x; I.B[4] -= ua.y * U.l.y; I.B[5] -= ua.y * U.l.z;
  I.B[6] -= ua.z * U.l.x; I.B[7] -= ua.z * U.l.y; I.B[8] -= ua.z * U.l.z;
  I.C[0] -= ul.x * U.l.x; I.C[1] -= ul.x * U.l.y; I.C[2] -= ul.x * U.l.z;
  I.C[3] -= ul.y * U.l.y; I.C[4] -= ul.y * U.l.z; I.C[5] -= ul.z * U.l.z;
}

// per-env physics state held by the owning thread
struct Phys {
  float p[3], quat[4], v[3], w[3];  // torso position, orientation (x,y,z,w), linear / angular velocity (world)
  float q[NJ], qd[NJ];
};

// per-thread scratch of one substep (local memory; every index below is warp-uniform)
struct Work {
  float R[NB][9];   // body rotations (world)
  float o[NB][3];   // body origins relative to the torso origin (world axes)
  float S[NJ][6];   // joint motion subspace (axis, anchor x axis)
  float V[NJ][6];   // link spatial velocities
  float cJ[NJ][6];  // velocity-product accelerations
  float U[NJ][6];   // IA * S
  float Dinv[NJ], u[NJ];
  float L0[21];     // Cholesky factor of the base articulated inertia (packed lower, row-major)
  float sumx, sumy; // sum of the 32 part origins (15 bodies + 17 joint anchors), relative to the torso
};

// ---- forward kinematics: body frames, joint subspaces, (optionally) link velocities and bias accelerations
__device__ __forceinline__ void fk(const Phys& s, Work& k) {
  quat2mat(s.quat[0], s.quat[1], s.quat[2], s.quat[3], k.R[0]);
  k.o[0][0] = k.o[0][1] = k.o[0][2] = 0.f;
  float sx = 0.f, sy = 0.f;
  ILRL_LINK_LOOP
  for (int b = 1; b < NB; b++) {
    const int p = kBodyParent[b];
    float Rc[9];
    if (kTree.bodyIdent[b]) {
#pragma unroll
      for (int i = 0; i < 9; i++) Rc[i] = k.R[p][i];
    } else {
      float Q[9];
      quat2mat(kBodyQuat[4 * b], kBodyQuat[4 * b + 1], kBodyQuat[4 * b + 2], kBodyQuat[4 * b + 3], Q);
      mm(k.R[p], Q, Rc);
    }
    V3 oc = ld3(k.o[p]) + mv(k.R[p], mk(kBodyPos[3 * b], kBodyPos[3 * b + 1], kBodyPos[3 * b + 2]));
    ILRL_LINK_LOOP
    for (int j = kTree.bodyJ0[b]; j < kTree.bodyJ0[b] + kTree.bodyNJ[b]; j++) {
      const V3 an = mk(kJointAnchor[3 * j], kJointAnchor[3 * j + 1], kJointAnchor[3 * j + 2]);
      const V3 ax = mk(kJointAxis[3 * j], kJointAxis[3 * j + 1], kJointAxis[3 * j + 2]);
      V3 rw = oc + mv(Rc, an);
      V3 aw = mv(Rc, ax);
      st3(k.S[j], aw);
      st3(k.S[j] + 3, cross(rw, aw));
      sx += rw.x; sy += rw.y;
      // Rn = Rc * Rot(ax, q)   (Rodrigues)
      float sn, cs;
      sincosf(s.q[j], &sn, &cs);
      float t = 1.f - cs, Rj[9], Rn[9];
      Rj[0] = t * ax.x * ax.x + cs;        Rj[1] = t * ax.x * ax.y - sn * ax.z; Rj[2] = t * ax.x * ax.z + sn * ax.y;
      Rj[3] = t * ax.x * ax.y + sn * ax.z; Rj[4] = t * ax.y * ax.y + cs;        Rj[5] = t * ax.y * ax.z - sn * ax.x;
      Rj[6] = t * ax.x * ax.z - sn * ax.y; Rj[7] = t * ax.y * ax.z + sn * ax.x; Rj[8] = t * ax.z * ax.z + cs;
      mm(Rc, Rj, Rn);
      oc = rw - mv(Rn, an);
#pragma unroll
      for (int i = 0; i < 9; i++) Rc[i] = Rn[i];
    }
#pragma unroll
    for (int i = 0; i < 9; i++) k.R[b][i] = Rc[i];
    st3(k.o[b], oc);
    sx += oc.x; sy += oc.y;
  }
  k.sumx = sx; k.sumy = sy;
}

// rigid-body spatial inertia (about the reference point, world axes) of body b and its bias force
// p = V x* (I V) - f_ext  with gravity and Bullet's velocity damping as external forces.
__device__ __forceinline__ void rigid_inertia_bias(const float* R, V3 c, float m, float ix, float iy, float iz, SV V,
                                                   Inertia& I, SV& pb) {
  // Ic = R diag(i) R^T
  float Ic[6];
  Ic[0] = ix * R[0] * R[0] + iy * R[1] * R[1] + iz * R[2] * R[2];
  Ic[1] = ix * R[0] * R[3] + iy * R[1] * R[4] + iz * R[2] * R[5];
  Ic[2] = ix * R[0] * R[6] + iy * R[1] * R[7] + iz * R[2] * R[8];
  Ic[3] = ix * R[3] * R[3] + iy * R[4] * R[4] + iz * R[5] * R[5];
  Ic[4] = ix * R[3] * R[6] + iy * R[4] * R[7] + iz * R[5] * R[8];
  Ic[5] = ix * R[6] * R[6] + iy * R[7] * R[7] + iz * R[8] * R[8];
  float cc = dot(c, c);
  I.A[0] = Ic[0] + m * (cc - c.x * c.x); I.A[1] = Ic[1] - m * c.x * c.y; I.A[2] = Ic[2] - m * c.x * c.z;
  I.A[3] = Ic[3] + m * (cc - c.y * c.y); I.A[4] = Ic[4] - m * c.y * c.z; I.A[5] = Ic[5] + m * (cc - c.z * c.z);
  I.B[0] = 0;        I.B[1] = -m * c.z; I.B[2] = m * c.y;
  I.B[3] = m * c.z;  I.B[4] = 0;        I.B[5] = -m * c.x;
  I.B[6] = -m * c.y; I.B[7] = m * c.x;  I.B[8] = 0;
  I.C[0] = m; I.C[1] = 0; I.C[2] = 0; I.C[3] = m; I.C[4] = 0; I.C[5] = m;
  // momentum about the reference point and its velocity-product rate
  V3 vc = V.l + cross(V.a, c);          // velocity of the body origin (= COM in Bullet's MJCF import)
  V3 Iw = symv(Ic, V.a);
  SV h; h.l = m * vc; h.a = Iw + cross(c, h.l);
  pb = crf(V, h);
  // external: gravity at the COM + Bullet's damping  -m v (K1 + K2|v|),  -Iw (K1 + K2|w|)
  float vn = sqrtf(dot(vc, vc)), wn = sqrtf(dot(V.a, V.a));
  float kl = m * ((float)ILRL_DAMP_K1_LIN + (float)ILRL_DAMP_K2_LIN * vn);
  float ka = (float)ILRL_DAMP_K1_ANG + (float)ILRL_DAMP_K2_ANG * wn;
  V3 F = mk(-kl * vc.x, -kl * vc.y, -kl * vc.z - m * (float)ILRL_GRAVITY);
  V3 N = mk(-ka * Iw.x, -ka * Iw.y, -ka * Iw.z);
  pb.a = pb.a - (N + cross(c, F));
  pb.l = pb.l - F;
}
__device__ __forceinline__ void body_inertia_bias(const Work& k, int b, SV V, Inertia& I, SV& pb) {
  rigid_inertia_bias(k.R[b], ld3(k.o[b]), kBodyMass[b], kBodyInertia[3 * b], kBodyInertia[3 * b + 1],
                     kBodyInertia[3 * b + 2], V, I, pb);
}

// Cholesky of the 6x6 SPD base inertia [[A,B],[B^T,C]] -> packed lower L (row-major: L[i*(i+1)/2 + j]), in place on the
// packed lower triangle, one column per template instance: every index is a compile-time constant, so the 21 numbers
// stay in registers (with run-time loop bounds the loops were left partly rolled and the arrays - 57 words - lived in
// local memory: 28 local stores and 13 dependent local loads per substep and lane).  Same operations in the same
// order as the textbook left-looking loop: identical results.
template <int J>
__device__ __forceinline__ void chol6_column(float (&a)[21]) {
  constexpr int rj = J * (J + 1) / 2;
  float sdiag = a[rj + J];
#pragma unroll
  for (int c = 0; c < J; c++) sdiag -= a[rj + c] * a[rj + c];
  const float inv = rsqrtf(fmaxf(sdiag, 1e-20f));
  a[rj + J] = inv;  // store the INVERSE of the diagonal
#pragma unroll
  for (int i = J + 1; i < 6; i++) {
    const int ri = i * (i + 1) / 2;
    float t = a[ri + J];
#pragma unroll
    for (int c = 0; c < J; c++) t -= a[ri + c] * a[rj + c];
    a[ri + J] = t * inv;
  }
}
__device__ __forceinline__ void chol6(const Inertia& I, float (&L)[21]) {
  // lower triangle of the symmetric matrix: rows 0-2 = A, rows 3-5 = [B^T | C]
  L[0] = I.A[0];
  L[1] = I.A[1]; L[2] = I.A[3];
  L[3] = I.A[2]; L[4] = I.A[4]; L[5] = I.A[5];
  L[6] = I.B[0]; L[7] = I.B[3]; L[8] = I.B[6]; L[9] = I.C[0];
  L[10] = I.B[1]; L[11] = I.B[4]; L[12] = I.B[7]; L[13] = I.C[1]; L[14] = I.C[3];
  L[15] = I.B[2]; L[16] = I.B[5]; L[17] = I.B[8]; L[18] = I.C[2]; L[19] = I.C[4]; L[20] = I.C[5];
  chol6_column<0>(L); chol6_column<1>(L); chol6_column<2>(L); chol6_column<3>(L); chol6_column<4>(L); chol6_column<5>(L);
}
// x = M^-1 b with the factor above (diagonal entries hold 1/L_jj)
__device__ __forceinline__ SV chol6_solve(const float* L, SV b) {
  float y[6] = {b.a.x, b.a.y, b.a.z, b.l.x, b.l.y, b.l.z};
#pragma unroll
  for (int i = 0; i < 6; i++) {
    float t = y[i];
#pragma unroll
    for (int c = 0; c < i; c++) t -= L[i * (i + 1) / 2 + c] * y[c];
    y[i] = t * L[i * (i + 1) / 2 + i];
  }
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    float t = y[i];
#pragma unroll
    for (int c = i + 1; c < 6; c++) t -= L[c * (c + 1) / 2 + i] * y[c];
    y[i] = t * L[i * (i + 1) / 2 + i];
  }
  SV r; r.a = mk(y[0], y[1], y[2]); r.l = mk(y[3], y[4], y[5]);
  return r;
}

constexpr int kPelvisLink = 2;

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

}  // namespace ilrl
