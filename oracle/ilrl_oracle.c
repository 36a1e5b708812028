/* oracle/ilrl_oracle.c — CPU restatement (plain C, double precision) of the reference's humanoid imitation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this file's shared object, and only as the checker / the CPU baseline.  The product path
 * (imitation-learning-rl_b200/) never links, imports or falls back to it.
 *
 * What it restates, and how each part is pinned:
 *   (A) env logic — reset / step / reward terms / frame indexing / target bookkeeping / observation / termination /
 *       hierarchical protocol of REF low_level_env.py:174-526 and REF hier_env.py:174-642 (+ math_util.py:20-27).
 *       PINNED: the .npz files under tests/golden/ hold traces of the UNMODIFIED reference Python (imported from /root/reference under
 *       oracle/ref_shim.py) and tests/test_oracle_golden.py replays them through this file.
 *   (B) calc_state / apply_action of the un-vendored pybullet_envs WalkerBase + REF humanoid.py:12-60 — restated
 *       from memory of the upstream source; pinned only by the reference notebook outputs listed in SURVEY.md §4
 *       (obs layout, joint order, relative-position formula, reset pose), checked in tests/test_oracle_golden.py.
 *   (C) the rigid-body step of the un-vendored Bullet (btMultiBody forward dynamics, plane contact + friction +
 *       joint-limit rows solved by 5 PGS sweeps, 4 substeps of 0.004125 s).  Restated from Bullet's published
 *       algorithm; NO reference test, fixture or golden vector pins a post-step state, and PyBullet is not
 *       installable in this image:                          ***  DYNAMICS: PARITY UNPINNED  ***
 *
 * Deliberately a DIFFERENT algorithm from the CUDA product for (C): the product runs Featherstone's O(n)
 * articulated-body algorithm in single precision with spatial vectors about the torso origin; this file builds the
 * dense 23x23 joint-space inertia matrix from per-body Jacobians, classical (non-spatial) bias accelerations and a
 * Cholesky solve, in double.  Agreement of the two is therefore a real check of the dynamics, not a tautology.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../imitation-learning-rl_b200/csrc/ilrl_constants.h"
#include "../imitation-learning-rl_b200/csrc/ilrl_model_data.h"

#define NB ILRL_NB
#define NJ ILRL_NJ
#define NS ILRL_NS
#define NV (6 + NJ)
#define MAXROWS (2 * NJ + 3 * NS)

static const int body_parent[NB] = ILRL_BODY_PARENT;
static const int body_link[NB] = ILRL_BODY_LINK;
static const double body_pos[NB * 3] = ILRL_BODY_POS;
static const double body_quat[NB * 4] = ILRL_BODY_QUAT;
static const double body_mass[NB] = ILRL_BODY_MASS;
static const double body_inertia[NB * 3] = ILRL_BODY_INERTIA;
static const int joint_body[NJ] = ILRL_JOINT_BODY;
static const int joint_parent[NJ] = ILRL_JOINT_PARENT;
static const double joint_anchor[NJ * 3] = ILRL_JOINT_ANCHOR;
static const double joint_axis[NJ * 3] = ILRL_JOINT_AXIS;
static const double joint_lo[NJ] = ILRL_JOINT_LO;
static const double joint_hi[NJ] = ILRL_JOINT_HI;
static const int sphere_body[NS] = ILRL_SPHERE_BODY;
static const double sphere_c[NS * 3] = ILRL_SPHERE_C;
static const double sphere_r[NS] = ILRL_SPHERE_R;
#define NG ILRL_NG
static const int geom_body[NG] = ILRL_GEOM_BODY;
static const double geom_p0[NG * 3] = ILRL_GEOM_P0;
static const double geom_p1[NG * 3] = ILRL_GEOM_P1;
static const double geom_r[NG] = ILRL_GEOM_R;
static const int motor_joint[NJ] = ILRL_MOTOR_JOINT;
static const double motor_gear[NJ] = ILRL_MOTOR_GEAR;
static const int map_joint[ILRL_NMAP] = ILRL_MAP_JOINT;
static const int map_col[ILRL_NMAP] = ILRL_MAP_COL;
static const double map_w[ILRL_NMAP] = ILRL_MAP_W;
static const double map_wv[ILRL_NMAP] = ILRL_MAP_WV;

/* ------------------------------------------------------------------ small vector helpers */
static void cross3(const double* a, const double* b, double* o) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void matvec3(const double* R, const double* v, double* o) {
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void matTvec3(const double* R, const double* v, double* o) {
  double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void matmul3(const double* A, const double* B, double* O) {
  double T[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
  memcpy(O, T, sizeof T);
}
static void quat2mat(const double* q /*x,y,z,w*/, double* R) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double n = x * x + y * y + z * z + w * w, s = 2.0 / n;
  R[0] = 1 - s * (y * y + z * z); R[1] = s * (x * y - w * z);     R[2] = s * (x * z + w * y);
  R[3] = s * (x * y + w * z);     R[4] = 1 - s * (x * x + z * z); R[5] = s * (y * z - w * x);
  R[6] = s * (x * z - w * y);     R[7] = s * (y * z + w * x);     R[8] = 1 - s * (x * x + y * y);
}
static void axisangle2mat(const double* a, double th, double* R) {
  double c = cos(th), s = sin(th), t = 1 - c, x = a[0], y = a[1], z = a[2];
  R[0] = t * x * x + c;     R[1] = t * x * y - s * z; R[2] = t * x * z + s * y;
  R[3] = t * x * y + s * z; R[4] = t * y * y + c;     R[5] = t * y * z - s * x;
  R[6] = t * x * z - s * y; R[7] = t * y * z + s * x; R[8] = t * z * z + c;
}

/* ------------------------------------------------------------------ kinematics */
typedef struct {
  double R[NB][9], o[NB][3];  /* world rotation / origin of every body frame                       */
  double aw[NJ][3], rw[NJ][3]; /* world hinge axis / anchor of every joint (origin of `link0_N`)    */
} orc_kin;

/* phys[47]: pos3 quat4(xyzw) lin3 ang3 q17 qd17 */
static void orc_fk(const double* phys, orc_kin* k) {
  const double* q = phys + 13;
  for (int b = 0; b < NB; b++) {
    double Rc[9], oc[3];
    if (b == 0) {
      quat2mat(phys + 3, Rc);
      memcpy(oc, phys, sizeof oc);
    } else {
      int p = body_parent[b];
      double Q[9], t[3];
      quat2mat(body_quat + 4 * b, Q);
      matmul3(k->R[p], Q, Rc);
      matvec3(k->R[p], body_pos + 3 * b, t);
      for (int i = 0; i < 3; i++) oc[i] = k->o[p][i] + t[i];
      for (int j = 0; j < NJ; j++) {
        if (joint_body[j] != b) continue; /* joints of a body compose in document order */
        double t2[3], Rj[9], Rn[9];
        matvec3(Rc, joint_anchor + 3 * j, t2);
        for (int i = 0; i < 3; i++) k->rw[j][i] = oc[i] + t2[i];
        matvec3(Rc, joint_axis + 3 * j, k->aw[j]);
        axisangle2mat(joint_axis + 3 * j, q[j], Rj);
        matmul3(Rc, Rj, Rn);
        matvec3(Rn, joint_anchor + 3 * j, t2);
        for (int i = 0; i < 3; i++) oc[i] = k->rw[j][i] - t2[i];
        memcpy(Rc, Rn, sizeof Rc);
      }
    }
    memcpy(k->R[b], Rc, sizeof Rc);
    memcpy(k->o[b], oc, sizeof oc);
  }
}

static int is_anc(int link, int j) { /* is joint j on the path from `link` up to the base? */
  for (int l = link; l >= 0; l = joint_parent[l])
    if (l == j) return 1;
  return 0;
}

/* Jacobian row(s) of a world point x rigidly attached to `link`: Jw (3xNV) and Jv (3xNV), column-major by dof */
static void point_jac(const orc_kin* k, const double* p0, int link, const double* x, double Jw[NV][3], double Jv[NV][3]) {
  memset(Jw, 0, sizeof(double) * NV * 3);
  memset(Jv, 0, sizeof(double) * NV * 3);
  double d[3] = {x[0] - p0[0], x[1] - p0[1], x[2] - p0[2]};
  for (int c = 0; c < 3; c++) {
    double e[3] = {0, 0, 0};
    e[c] = 1;
    Jw[c][c] = 1;
    cross3(e, d, Jv[c]); /* d(v)/d(omega_c) = e_c x d */
    Jv[3 + c][c] = 1;
  }
  for (int j = 0; j < NJ; j++) {
    if (!is_anc(link, j)) continue;
    double dj[3] = {x[0] - k->rw[j][0], x[1] - k->rw[j][1], x[2] - k->rw[j][2]};
    memcpy(Jw[6 + j], k->aw[j], sizeof(double) * 3);
    cross3(k->aw[j], dj, Jv[6 + j]);
  }
}

/* dense Cholesky M = L L^T (in place, lower), and solve */
static void chol(double* M, int n) {
  for (int j = 0; j < n; j++) {
    double s = M[j * n + j];
    for (int k = 0; k < j; k++) s -= M[j * n + k] * M[j * n + k];
    s = sqrt(s > 1e-300 ? s : 1e-300);
    M[j * n + j] = s;
    for (int i = j + 1; i < n; i++) {
      double t = M[i * n + j];
      for (int k = 0; k < j; k++) t -= M[i * n + k] * M[j * n + k];
      M[i * n + j] = t / s;
    }
  }
}
static void chol_solve(const double* L, int n, const double* b, double* x) {
  double y[NV];
  for (int i = 0; i < n; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[i * n + k] * y[k];
    y[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = y[i];
    for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
}

/* Joint-space inertia matrix M(q) (NVxNV, row-major) from per-body Jacobians. */
static void mass_matrix(const orc_kin* k, const double* p0, double* M) {
  memset(M, 0, sizeof(double) * NV * NV);
  for (int b = 0; b < NB; b++) {
    double Jw[NV][3], Jv[NV][3], IJ[NV][3];
    point_jac(k, p0, body_link[b], k->o[b], Jw, Jv);
    for (int c = 0; c < NV; c++) { /* Iw*Jw = R diag(I) R^T Jw */
      double t[3];
      matTvec3(k->R[b], Jw[c], t);
      for (int i = 0; i < 3; i++) t[i] *= body_inertia[3 * b + i];
      matvec3(k->R[b], t, IJ[c]);
    }
    for (int r = 0; r < NV; r++)
      for (int c = 0; c < NV; c++) M[r * NV + c] += dot3(Jw[r], IJ[c]) + body_mass[b] * dot3(Jv[r], Jv[c]);
  }
}

void ilrl_oracle_fk(const double* phys, double* body_o /*[15*3]*/, double* anchor_o /*[17*3]*/, double* body_R /*[15*9] or NULL*/) {
  orc_kin k;
  orc_fk(phys, &k);
  memcpy(body_o, k.o, sizeof k.o);
  memcpy(anchor_o, k.rw, sizeof k.rw);
  if (body_R) memcpy(body_R, k.R, sizeof k.R);
}

/* total mechanical energy (kinetic via M, potential via body heights) — used by conservation tests */
double ilrl_oracle_energy(const double* phys) {
  orc_kin k;
  orc_fk(phys, &k);
  double M[NV * NV], nu[NV], e = 0;
  mass_matrix(&k, phys, M);
  for (int i = 0; i < 3; i++) { nu[i] = phys[10 + i]; nu[3 + i] = phys[7 + i]; }
  for (int j = 0; j < NJ; j++) nu[6 + j] = phys[30 + j];
  for (int r = 0; r < NV; r++)
    for (int c = 0; c < NV; c++) e += 0.5 * nu[r] * M[r * NV + c] * nu[c];
  for (int b = 0; b < NB; b++) e += body_mass[b] * ILRL_GRAVITY * k.o[b][2];
  return e;
}

/* Model-sensitivity knobs (tools/model_sensitivity.py): the [BULLET] constants that no reference fixture pins can be
 * varied one at a time in THIS restatement to see what they do to the rollout statistics.  Defaults = ilrl_constants.h
 * (= what the CUDA product computes).  Not thread-safe: set before any thread steps. */
typedef struct {
  double limit_erp, contact_erp, friction, max_coord_vel, link_damp_scale;
  int solver_iters, limit_rows_always;
  double joint_damping[NJ], joint_armature[NJ], joint_stiffness[NJ];
} orc_params;
static orc_params g_par = {ILRL_LIMIT_ERP, ILRL_CONTACT_ERP, ILRL_FRICTION, ILRL_MAX_COORD_VEL, 1.0, ILRL_SOLVER_ITERS, 0, {0}, {0}, {0}};
/* p[0..6] = limit_erp, contact_erp, friction, max_coord_vel, link_damp_scale, solver_iters, limit_rows_always;
 * p[7..7+3*NJ) = joint damping, armature, stiffness (MJCF attributes; Bullet's importer is believed to ignore them).
 * NULL restores the defaults. */
void ilrl_oracle_set_params(const double* p) {
  orc_params d = {ILRL_LIMIT_ERP, ILRL_CONTACT_ERP, ILRL_FRICTION, ILRL_MAX_COORD_VEL, 1.0, ILRL_SOLVER_ITERS, 0, {0}, {0}, {0}};
  g_par = d;
  if (!p) return;
  g_par.limit_erp = p[0]; g_par.contact_erp = p[1]; g_par.friction = p[2]; g_par.max_coord_vel = p[3];
  g_par.link_damp_scale = p[4]; g_par.solver_iters = (int)p[5]; g_par.limit_rows_always = (int)p[6];
  for (int j = 0; j < NJ; j++) {
    g_par.joint_damping[j] = p[7 + j]; g_par.joint_armature[j] = p[7 + NJ + j]; g_par.joint_stiffness[j] = p[7 + 2 * NJ + j];
  }
}
/* Heightfield terrain (REF humanoid.py:68-144 CustomScene: 256 x 256 samples, meshScale 1, body at z = 0.25).
 * [BULLET, restated] btHeightfieldTerrainShape: sample (i, j) sits at x = i - (rows - 1) / 2, y = j - (cols - 1) / 2,
 * z = data[i + j * rows] - (min + max) / 2 (+ the body position); a cell is cut along the diagonal (i+1, j) - (i, j+1).
 * Contact model of this restatement (declared): every candidate sphere meets the PLANE of the triangle
 * under its centre (face contacts only; edge / vertex contacts of a real triangle mesh are not modelled).
 * zoff = body z - (min + max) / 2.  NULL = the flat ground plane z = 0.  Not thread-safe: set before stepping. */
static const double* g_hf = 0;
static int g_hf_rows = 0, g_hf_cols = 0;
static double g_hf_zoff = 0;
void ilrl_oracle_set_heightfield(const double* h, int rows, int cols, double zoff) {
  g_hf = h; g_hf_rows = rows; g_hf_cols = cols; g_hf_zoff = zoff;
}
/* ground height under (x, y) and the unit normal of its triangle */
static double terrain_sample(double x, double y, double* n) {
  if (!g_hf) { n[0] = n[1] = 0; n[2] = 1; return 0; }
  double fx = x + 0.5 * (g_hf_rows - 1), fy = y + 0.5 * (g_hf_cols - 1);
  int i = (int)floor(fx), j = (int)floor(fy);
  i = i < 0 ? 0 : (i > g_hf_rows - 2 ? g_hf_rows - 2 : i);
  j = j < 0 ? 0 : (j > g_hf_cols - 2 ? g_hf_cols - 2 : j);
  double u = fx - i, v = fy - j;
  u = u < 0 ? 0 : (u > 1 ? 1 : u);
  v = v < 0 ? 0 : (v > 1 ? 1 : v);
  const double h00 = g_hf[i + j * g_hf_rows], h10 = g_hf[i + 1 + j * g_hf_rows];
  const double h01 = g_hf[i + (j + 1) * g_hf_rows], h11 = g_hf[i + 1 + (j + 1) * g_hf_rows];
  double gx, gy, h;
  if (u + v <= 1) { gx = h10 - h00; gy = h01 - h00; h = h00 + u * gx + v * gy; }
  else { gx = h11 - h01; gy = h11 - h10; h = h11 - (1 - u) * gx - (1 - v) * gy; }
  const double il = 1.0 / sqrt(1 + gx * gx + gy * gy);
  n[0] = -gx * il; n[1] = -gy * il; n[2] = il;
  return h + g_hf_zoff;
}
/* btPlaneSpace1 */
static void plane_space(const double* n, double* p, double* q) {
  if (fabs(n[2]) > 0.7071067811865475244) {
    double a = n[1] * n[1] + n[2] * n[2], k = 1.0 / sqrt(a);
    p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
    q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
  } else {
    double a = n[0] * n[0] + n[1] * n[1], k = 1.0 / sqrt(a);
    p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
    q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
  }
}

/* Self-collision (REF humanoid.py:13 `self_collision = True`; the CUDA product does NOT model it: declared divergence,
 * DESIGN.md 4 / 7).  The oracle can switch it on to QUANTIFY what the omission does to the rollout statistics
 * (tools/model_sensitivity.py).  [BULLET, restated] pybullet_envs loads the MJCF with URDF_USE_SELF_COLLISION |
 * URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS: every pair of geoms collides unless one body is an ancestor of the
 * other; combined friction = product of the two geoms' friction (2.0 x 2.0).  Geoms are capsules (spheres = zero
 * length): closest points of the two axis segments, contact while distance < the breaking threshold, at most
 * MAX_SELF deepest pairs, rows (normal + 2 friction) after the ground contacts. */
#define MAX_SELF 8
static int g_selfcol = 0;
static long g_selfcol_contacts = 0, g_selfcol_substeps = 0, g_selfcol_hist[NG * NG];
static double g_selfcol_min_sep = 1e30; /* smallest distance between the AXES of an active pair since the last reset of the
                                           statistics: near 0 the contact normal is ill-conditioned (crossing capsules) */
void ilrl_oracle_set_self_collision(int on) {
  g_selfcol = on; g_selfcol_contacts = g_selfcol_substeps = 0; g_selfcol_min_sep = 1e30;
  memset(g_selfcol_hist, 0, sizeof g_selfcol_hist);
}
double ilrl_oracle_self_collision_min_sep(int reset) {
  const double v = g_selfcol_min_sep;
  if (reset) g_selfcol_min_sep = 1e30;
  return v;
}
void ilrl_oracle_self_collision_hist(long* out /* [NG * NG] contacts per geom pair */) { memcpy(out, g_selfcol_hist, sizeof g_selfcol_hist); }
void ilrl_oracle_self_collision_stats(long* contacts, long* substeps) { *contacts = g_selfcol_contacts; *substeps = g_selfcol_substeps; }
static int body_is_ancestor(int a, int b) { /* a is b or an ancestor of b */
  for (; b >= 0; b = body_parent[b]) if (b == a) return 1;
  return 0;
}
/* closest points of segments p1-q1 and p2-q2 (Ericson, Real-Time Collision Detection 5.1.9) */
static void seg_seg(const double* p1, const double* q1, const double* p2, const double* q2, double* c1, double* c2) {
  double d1[3], d2[3], r[3];
  for (int i = 0; i < 3; i++) { d1[i] = q1[i] - p1[i]; d2[i] = q2[i] - p2[i]; r[i] = p1[i] - p2[i]; }
  const double a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r), EPS = 1e-12;
  double s, t;
  if (a <= EPS && e <= EPS) { s = t = 0; }
  else if (a <= EPS) { s = 0; t = f / e; t = t < 0 ? 0 : (t > 1 ? 1 : t); }
  else {
    const double c = dot3(d1, r);
    if (e <= EPS) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
    else {
      const double b = dot3(d1, d2), den = a * e - b * b;
      s = den > EPS ? (b * f - c * e) / den : 0;
      s = s < 0 ? 0 : (s > 1 ? 1 : s);
      t = (b * s + f) / e;
      if (t < 0) { t = 0; s = -c / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
      else if (t > 1) { t = 1; s = (b - c) / a; s = s < 0 ? 0 : (s > 1 ? 1 : s); }
    }
  }
  for (int i = 0; i < 3; i++) { c1[i] = p1[i] + s * d1[i]; c2[i] = p2[i] + t * d2[i]; }
}

/* Diagnostics (tools/model_sensitivity.py): extreme-event counters over all substeps since the last reset:
 * [0] substeps, [1] max joint-limit overshoot (rad), [2] substeps with overshoot > 0.5 rad, [3] max |torso v_z change| in one
 * substep (m/s), [4] substeps with |dv_z| > 2 m/s, [5] max torso height (m) */
static double g_diag[8];
void ilrl_oracle_diag(double* out, int reset) {
  if (out) memcpy(out, g_diag, sizeof g_diag);
  if (reset) memset(g_diag, 0, sizeof g_diag);
}

/* Diagnostics (tools/rowstats.py): when set, every substep counts its (violated limits, kept contacts) pair in a
 * [18][9] histogram.  Not thread-safe; single-threaded tools only. */
static long* g_rowstats = 0;
void ilrl_oracle_set_rowstats(long* hist /* [18*9] or NULL */) { g_rowstats = hist; }

/* One Bullet substep.  flags: bit0 = no gravity, bit1 = no velocity damping, bit2 = no contacts, bit3 = no limits
 * (the flags exist only for the conservation / free-flight tests). */
static void orc_substep(double* phys, const double* tau, double dt, int flags) {
  orc_kin k;
  orc_fk(phys, &k);
  double* p0 = phys;
  double* q = phys + 13;
  double* qd = phys + 30;
  double nu[NV];
  for (int i = 0; i < 3; i++) { nu[i] = phys[10 + i]; nu[3 + i] = phys[7 + i]; }
  for (int j = 0; j < NJ; j++) nu[6 + j] = qd[j];

  /* ---- bias accelerations with nu_dot = 0: per joint link (frame origin at its anchor) */
  double w[NJ][3], al[NJ][3], aa[NJ][3];
  const double* w0 = nu;
  for (int j = 0; j < NJ; j++) {
    int p = joint_parent[j];
    double wp[3], alp[3] = {0, 0, 0}, ap[3], d[3], t[3], t2[3];
    if (p < 0) {
      memcpy(wp, w0, sizeof wp);
      for (int i = 0; i < 3; i++) d[i] = k.rw[j][i] - p0[i];
      cross3(wp, d, t); cross3(wp, t, ap);
    } else {
      memcpy(wp, w[p], sizeof wp); memcpy(alp, al[p], sizeof alp);
      for (int i = 0; i < 3; i++) d[i] = k.rw[j][i] - k.rw[p][i];
      cross3(wp, d, t); cross3(wp, t, t2); cross3(alp, d, t);
      for (int i = 0; i < 3; i++) ap[i] = aa[p][i] + t[i] + t2[i];
    }
    double rel[3] = {k.aw[j][0] * qd[j], k.aw[j][1] * qd[j], k.aw[j][2] * qd[j]};
    cross3(wp, rel, t);
    for (int i = 0; i < 3; i++) { w[j][i] = wp[i] + rel[i]; al[j][i] = alp[i] + t[i]; aa[j][i] = ap[i]; }
  }

  /* ---- assemble M, generalized bias h and applied forces Q */
  double M[NV * NV], rhs[NV];
  mass_matrix(&k, p0, M);
  for (int i = 0; i < 6; i++) rhs[i] = 0;
  for (int j = 0; j < NJ; j++) {   /* (joint damping / stiffness / armature: zero unless the sensitivity knobs are set) */
    rhs[6 + j] = tau[j] - g_par.joint_damping[j] * qd[j] - g_par.joint_stiffness[j] * q[j];
    M[(6 + j) * NV + 6 + j] += g_par.joint_armature[j];
  }
  for (int b = 0; b < NB; b++) {
    double Jw[NV][3], Jv[NV][3];
    int l = body_link[b];
    point_jac(&k, p0, l, k.o[b], Jw, Jv);
    double wb[3] = {0, 0, 0}, vb[3] = {0, 0, 0}, alb[3] = {0, 0, 0}, ab[3], d[3], t[3], t2[3];
    for (int c = 0; c < NV; c++)
      for (int i = 0; i < 3; i++) { wb[i] += Jw[c][i] * nu[c]; vb[i] += Jv[c][i] * nu[c]; }
    if (l < 0) {
      for (int i = 0; i < 3; i++) d[i] = k.o[b][i] - p0[i];
      cross3(w0, d, t); cross3(w0, t, ab);
    } else {
      memcpy(alb, al[l], sizeof alb);
      for (int i = 0; i < 3; i++) d[i] = k.o[b][i] - k.rw[l][i];
      cross3(w[l], d, t); cross3(w[l], t, t2); cross3(alb, d, t);
      for (int i = 0; i < 3; i++) ab[i] = aa[l][i] + t[i] + t2[i];
    }
    /* inertial torque  Iw*alpha + w x Iw*w  and force m*a */
    double Iw_al[3], Iw_w[3], wl[3], gy[3];
    matTvec3(k.R[b], alb, t); for (int i = 0; i < 3; i++) t[i] *= body_inertia[3 * b + i]; matvec3(k.R[b], t, Iw_al);
    matTvec3(k.R[b], wb, wl); for (int i = 0; i < 3; i++) t[i] = wl[i] * body_inertia[3 * b + i]; matvec3(k.R[b], t, Iw_w);
    cross3(wb, Iw_w, gy);
    /* applied: gravity + Bullet's velocity damping (force -m v (K1+K2|v|), torque -I w (K1+K2|w|)) */
    double f[3] = {0, 0, 0}, n[3] = {0, 0, 0};
    if (!(flags & 1)) f[2] -= body_mass[b] * ILRL_GRAVITY;
    if (!(flags & 2)) {
      double vn = sqrt(dot3(vb, vb)), wn = sqrt(dot3(wb, wb));
      for (int i = 0; i < 3; i++) {
        f[i] -= g_par.link_damp_scale * body_mass[b] * vb[i] * (ILRL_DAMP_K1_LIN + ILRL_DAMP_K2_LIN * vn);
        n[i] -= g_par.link_damp_scale * Iw_w[i] * (ILRL_DAMP_K1_ANG + ILRL_DAMP_K2_ANG * wn);
      }
    }
    for (int c = 0; c < NV; c++) {
      double tw = 0, tv = 0;
      for (int i = 0; i < 3; i++) {
        tw += Jw[c][i] * (n[i] - Iw_al[i] - gy[i]);
        tv += Jv[c][i] * (f[i] - body_mass[b] * ab[i]);
      }
      rhs[c] += tw + tv;
    }
  }
  double L[NV * NV], acc[NV];
  memcpy(L, M, sizeof M);
  chol(L, NV);
  chol_solve(L, NV, rhs, acc);
  const double vz_before = phys[9];
  for (int c = 0; c < NV; c++) {
    nu[c] += dt * acc[c];
    if (nu[c] > g_par.max_coord_vel) nu[c] = g_par.max_coord_vel;
    if (nu[c] < -g_par.max_coord_vel) nu[c] = -g_par.max_coord_vel;
  }

  /* ---- constraint rows: violated joint limits, then contact normals, then 2 friction rows per contact */
  /* flat ground: normal (0,0,1), tangents btPlaneSpace1((0,0,1)) = (0,-1,0), (1,0,0) */
  double J[MAXROWS][NV], Rsp[MAXROWS][NV], rrhs[MAXROWS], dinv[MAXROWS], lam[MAXROWS];
  int nlim = 0, ncon = 0;
  if (!(flags & 8))
    for (int j = 0; j < NJ; j++) {
      for (int side = 0; side < 2; side++) {
        double pen = side == 0 ? q[j] - joint_lo[j] : joint_hi[j] - q[j], dir = side == 0 ? 1 : -1;
        if (side == 1 && !g_par.limit_rows_always && q[j] - joint_lo[j] <= 0) continue; /* one row per joint, lower first */
        if (pen > 0 && !g_par.limit_rows_always) continue;
        int r = nlim++;
        memset(J[r], 0, sizeof J[r]);
        J[r][6 + j] = dir;
        /* violated: ERP push-out; (knob) not violated: speculative row, closing speed limited to distance / dt */
        rrhs[r] = pen > 0 ? -pen / dt : -pen * g_par.limit_erp / dt; /* positional part; velocity part added below */
        if (-pen > g_diag[1]) g_diag[1] = -pen;
        if (-pen > 0.5) g_diag[2] += 1;
      }
    }
  int crow[ILRL_MAX_CONTACTS + MAX_SELF];
  double cmu[ILRL_MAX_CONTACTS + MAX_SELF];
  if (!(flags & 4)) {
    /* candidate set: every sphere closer than the breaking threshold; if more than ILRL_MAX_CONTACTS, drop the
     * shallowest (largest distance, ties -> highest index) until the cap holds; rows then follow table order */
    double sdist[NS], sc[NS][3], sn[NS][3];
    int act[NS], nact = 0;
    for (int s = 0; s < NS; s++) {
      int b = sphere_body[s];
      matvec3(k.R[b], sphere_c + 3 * s, sc[s]);
      for (int i = 0; i < 3; i++) sc[s][i] += k.o[b][i];
      const double hgt = terrain_sample(sc[s][0], sc[s][1], sn[s]);
      sdist[s] = (sc[s][2] - hgt) * sn[s][2] - sphere_r[s];   /* distance of the sphere to the plane under its centre */
      act[s] = sdist[s] < ILRL_CONTACT_BREAK;
      nact += act[s];
    }
    while (nact > ILRL_MAX_CONTACTS) {
      int worst = -1;
      for (int s = 0; s < NS; s++)
        if (act[s] && (worst < 0 || sdist[s] >= sdist[worst])) worst = s;
      act[worst] = 0;
      nact--;
    }
    for (int s = 0; s < NS; s++) {
      if (!act[s]) continue;
      int b = sphere_body[s];
      double dist = sdist[s];
      double x[3] = {sc[s][0] - sphere_r[s] * sn[s][0], sc[s][1] - sphere_r[s] * sn[s][1], sc[s][2] - sphere_r[s] * sn[s][2]};
      double Jw[NV][3], Jv[NV][3], t1[3], t2[3];
      plane_space(sn[s], t1, t2);
      point_jac(&k, p0, body_link[b], x, Jw, Jv);
      int r = nlim + 3 * ncon;
      cmu[ncon] = g_par.friction;
      crow[ncon++] = r;
      for (int c2 = 0; c2 < NV; c2++) {
        J[r][c2] = dot3(Jv[c2], sn[s]);
        J[r + 1][c2] = dot3(Jv[c2], t1);
        J[r + 2][c2] = dot3(Jv[c2], t2);
      }
      /* Bullet: penetration > 0 -> speculative row (velocityError -= pen/dt), else ERP push-out */
      rrhs[r] = dist > 0 ? -dist / dt : -dist * g_par.contact_erp / dt;
      rrhs[r + 1] = rrhs[r + 2] = 0;
    }
  }
  if (g_selfcol && !(flags & 4)) {
    /* geom axis end points in the world, then every admissible pair */
    double w0[NG][3], w1[NG][3];
    for (int g = 0; g < NG; g++) {
      const int b = geom_body[g];
      matvec3(k.R[b], geom_p0 + 3 * g, w0[g]);
      matvec3(k.R[b], geom_p1 + 3 * g, w1[g]);
      for (int i = 0; i < 3; i++) { w0[g][i] += k.o[b][i]; w1[g][i] += k.o[b][i]; }
    }
    /* candidates in pair order; if more than MAX_SELF, drop the shallowest (largest distance, ties -> the later pair)
     * until the cap holds; rows follow pair order (the rule of the ground contacts) */
    enum { NPAIR = NG * (NG - 1) / 2 };
    int ca[NPAIR], cb[NPAIR], nc = 0;
    double cd[NPAIR], cc1[NPAIR][3], cc2[NPAIR][3];
    for (int ga = 0; ga < NG; ga++)
      for (int gb = ga + 1; gb < NG; gb++) {
        const int ba = geom_body[ga], bb = geom_body[gb];
        if (body_is_ancestor(ba, bb) || body_is_ancestor(bb, ba)) continue;
        double c1[3], c2[3];
        seg_seg(w0[ga], w1[ga], w0[gb], w1[gb], c1, c2);
        const double dx[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
        const double dist = sqrt(dot3(dx, dx)) - geom_r[ga] - geom_r[gb];
        if (!(dist < ILRL_CONTACT_BREAK)) continue;
        ca[nc] = ga; cb[nc] = gb; cd[nc] = dist;
        memcpy(cc1[nc], c1, sizeof c1); memcpy(cc2[nc], c2, sizeof c2);
        nc++;
      }
    while (nc > MAX_SELF) {
      int worst = 0;
      for (int q = 1; q < nc; q++) if (cd[q] >= cd[worst]) worst = q;
      for (int q = worst; q + 1 < nc; q++) {
        ca[q] = ca[q + 1]; cb[q] = cb[q + 1]; cd[q] = cd[q + 1];
        memcpy(cc1[q], cc1[q + 1], sizeof cc1[q]); memcpy(cc2[q], cc2[q + 1], sizeof cc2[q]);
      }
      nc--;
    }
    const int ns = nc;
    int pa[MAX_SELF], pb[MAX_SELF];
    double pd[MAX_SELF], pc1[MAX_SELF][3], pc2[MAX_SELF][3];
    for (int q = 0; q < ns; q++) {
      pa[q] = ca[q]; pb[q] = cb[q]; pd[q] = cd[q];
      memcpy(pc1[q], cc1[q], sizeof pc1[q]); memcpy(pc2[q], cc2[q], sizeof pc2[q]);
    }
    g_selfcol_substeps++;
    g_selfcol_contacts += ns;
    for (int q = 0; q < ns; q++) {
      g_selfcol_hist[pa[q] * NG + pb[q]]++;
      const double sep = pd[q] + geom_r[pa[q]] + geom_r[pb[q]];
      if (sep < g_selfcol_min_sep) g_selfcol_min_sep = sep;
    }
    for (int q = 0; q < ns; q++) {
      double n[3] = {pc1[q][0] - pc2[q][0], pc1[q][1] - pc2[q][1], pc1[q][2] - pc2[q][2]};
      double len = sqrt(dot3(n, n));
      if (len < 1e-9) { n[0] = 0; n[1] = 0; n[2] = 1; len = 1; }
      for (int i = 0; i < 3; i++) n[i] /= len; /* from geom b towards geom a */
      double xa[3], xb[3], t1[3], t2[3];
      for (int i = 0; i < 3; i++) { xa[i] = pc1[q][i] - geom_r[pa[q]] * n[i]; xb[i] = pc2[q][i] + geom_r[pb[q]] * n[i]; }
      plane_space(n, t1, t2);
      double Jwa[NV][3], Jva[NV][3], Jwb[NV][3], Jvb[NV][3];
      point_jac(&k, p0, body_link[geom_body[pa[q]]], xa, Jwa, Jva);
      point_jac(&k, p0, body_link[geom_body[pb[q]]], xb, Jwb, Jvb);
      int r = nlim + 3 * ncon;
      cmu[ncon] = 4.0; /* 2.0 x 2.0 */
      crow[ncon++] = r;
      for (int c2 = 0; c2 < NV; c2++) {
        double dv[3] = {Jva[c2][0] - Jvb[c2][0], Jva[c2][1] - Jvb[c2][1], Jva[c2][2] - Jvb[c2][2]};
        J[r][c2] = dot3(dv, n);
        J[r + 1][c2] = dot3(dv, t1);
        J[r + 2][c2] = dot3(dv, t2);
      }
      const double dist = pd[q];
      rrhs[r] = dist > 0 ? -dist / dt : -dist * g_par.contact_erp / dt;
      rrhs[r + 1] = rrhs[r + 2] = 0;
    }
  }
  int nrows = nlim + 3 * ncon;
  if (g_rowstats) { g_rowstats[(nlim > 17 ? 17 : nlim) * 9 + (ncon > 8 ? 8 : ncon)]++; }
  for (int r = 0; r < nrows; r++) {
    chol_solve(L, NV, J[r], Rsp[r]);
    double d = 0, rv = 0;
    for (int c = 0; c < NV; c++) { d += J[r][c] * Rsp[r][c]; rv += J[r][c] * nu[c]; }
    dinv[r] = 1.0 / d;
    rrhs[r] = (rrhs[r] - rv) * dinv[r];
    lam[r] = 0;
  }
  double dv[NV];
  memset(dv, 0, sizeof dv);
  for (int it = 0; it < g_par.solver_iters; it++) {
    for (int r = 0; r < nlim; r++) { /* joint limits: impulse >= 0 */
      double jd = 0;
      for (int c = 0; c < NV; c++) jd += J[r][c] * dv[c];
      double nl = lam[r] + rrhs[r] - jd * dinv[r];
      if (nl < 0) nl = 0;
      double dl = nl - lam[r];
      lam[r] = nl;
      for (int c = 0; c < NV; c++) dv[c] += dl * Rsp[r][c];
    }
    for (int ci = 0; ci < ncon; ci++) { /* contact normals */
      int r = crow[ci];
      double jd = 0;
      for (int c = 0; c < NV; c++) jd += J[r][c] * dv[c];
      double nl = lam[r] + rrhs[r] - jd * dinv[r];
      if (nl < 0) nl = 0;
      double dl = nl - lam[r];
      lam[r] = nl;
      for (int c = 0; c < NV; c++) dv[c] += dl * Rsp[r][c];
    }
    for (int ci = 0; ci < ncon; ci++) { /* friction pair, implicit cone (resolveConeFrictionConstraintRows) */
      int rn = crow[ci], r1 = rn + 1, r2 = rn + 2;
      if (!(lam[rn] > 0)) continue;
      double lim = cmu[ci] * lam[rn], jd1 = 0, jd2 = 0;
      for (int c = 0; c < NV; c++) { jd1 += J[r1][c] * dv[c]; jd2 += J[r2][c] * dv[c]; }
      double s1 = lam[r1] + rrhs[r1] - jd1 * dinv[r1], s2 = lam[r2] + rrhs[r2] - jd2 * dinv[r2];
      double n2 = s1 * s1 + s2 * s2;
      if (n2 > lim * lim) { double sc = lim / sqrt(n2); s1 *= sc; s2 *= sc; }
      double d1 = s1 - lam[r1], d2 = s2 - lam[r2];
      lam[r1] = s1; lam[r2] = s2;
      for (int c = 0; c < NV; c++) dv[c] += d1 * Rsp[r1][c] + d2 * Rsp[r2][c];
    }
  }
  for (int c = 0; c < NV; c++) {
    nu[c] += dv[c];
    if (nu[c] > g_par.max_coord_vel) nu[c] = g_par.max_coord_vel;
    if (nu[c] < -g_par.max_coord_vel) nu[c] = -g_par.max_coord_vel;
  }
  {
    double dvz = fabs(nu[5] - vz_before);
    g_diag[0] += 1;
    if (dvz > g_diag[3]) g_diag[3] = dvz;
    if (dvz > 2.0) g_diag[4] += 1;
    if (p0[2] > g_diag[5]) g_diag[5] = p0[2];
  }

  /* ---- integrate positions (btMultiBody::stepPositionsMultiDof: exponential map on the base quaternion) */
  for (int i = 0; i < 3; i++) { phys[10 + i] = nu[i]; phys[7 + i] = nu[3 + i]; p0[i] += dt * nu[3 + i]; }
  {
    double wn = sqrt(dot3(nu, nu)), sc, cw;
    if (wn < 1e-3) sc = 0.5 * dt - dt * dt * dt * 0.020833333333 * wn * wn;
    else sc = sin(0.5 * wn * dt) / wn;
    cw = cos(0.5 * wn * dt);
    double dx = nu[0] * sc, dy = nu[1] * sc, dz = nu[2] * sc;
    double* Q = phys + 3;
    double x = Q[0], y = Q[1], z = Q[2], ww = Q[3];
    /* dq (x) q */
    double nx = cw * x + dx * ww + dy * z - dz * y;
    double ny = cw * y - dx * z + dy * ww + dz * x;
    double nz = cw * z + dx * y - dy * x + dz * ww;
    double nw = cw * ww - dx * x - dy * y - dz * z;
    double n = sqrt(nx * nx + ny * ny + nz * nz + nw * nw);
    Q[0] = nx / n; Q[1] = ny / n; Q[2] = nz / n; Q[3] = nw / n;
  }
  for (int j = 0; j < NJ; j++) { qd[j] = nu[6 + j]; q[j] += dt * qd[j]; }
}

/* scene.global_step(): 4 substeps with the joint torques held (TORQUE_CONTROL).  tau in joint order. */
void ilrl_oracle_physics_step(double* phys, const double* tau, int flags) {
  for (int s = 0; s < ILRL_SUBSTEPS; s++) orc_substep(phys, tau, ILRL_FRAME_DT / ILRL_SUBSTEPS, flags);
}
void ilrl_oracle_substep(double* phys, const double* tau, double dt, int flags) { orc_substep(phys, tau, dt, flags); }

/* REF humanoid.py:54-60 apply_action: torque_i = gear_i * 0.41 * clip(a_i, -1, 1), motor order -> joint order */
void ilrl_oracle_action_to_torque(const double* action, double* tau) {
  for (int m = 0; m < NJ; m++) {
    double a = action[m] < -1 ? -1 : (action[m] > 1 ? 1 : action[m]);
    tau[motor_joint[m]] = motor_gear[m] * a;
  }
}

/* ------------------------------------------------------------------ calc_state (pybullet_envs WalkerBase) */
static void quat_rpy(const double* q, double* rpy) { /* pybullet getEulerFromQuaternion */
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double sarg = -2 * (x * z - w * y);
  if (sarg <= -0.99999) { rpy[0] = 0; rpy[1] = -0.5 * M_PI; rpy[2] = 2 * atan2(x, -y); }
  else if (sarg >= 0.99999) { rpy[0] = 0; rpy[1] = 0.5 * M_PI; rpy[2] = 2 * atan2(-x, y); }
  else {
    rpy[0] = atan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z);
    rpy[1] = asin(sarg);
    rpy[2] = atan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z);
  }
}

typedef struct {
  float obs[42];
  double body_xyz[3];
  float joint_speeds[NJ];
  int joints_at_limit;
  double rpy[3];
} orc_calc;

static float clip5(float v) { return v < -5.f ? -5.f : (v > 5.f ? 5.f : v); }

static void orc_calc_state(const double* phys, double wtx, double wty, orc_calc* o) {
  orc_kin k;
  orc_fk(phys, &k);
  /* body_xyz: xy mean over the 33 entries of `parts` (15 bodies + 17 joint links + the floor at 0), torso z */
  double sx = 0, sy = 0;
  for (int b = 0; b < NB; b++) { sx += k.o[b][0]; sy += k.o[b][1]; }
  for (int j = 0; j < NJ; j++) { sx += k.rw[j][0]; sy += k.rw[j][1]; }
  o->body_xyz[0] = sx / 33.0; o->body_xyz[1] = sy / 33.0; o->body_xyz[2] = phys[2];
  quat_rpy(phys + 3, o->rpy);
  double yaw = o->rpy[2];
  double theta = atan2(wty - o->body_xyz[1], wtx - o->body_xyz[0]);
  double ang = theta - yaw;
  double c = cos(-yaw), s = sin(-yaw);
  double vx = c * phys[7] - s * phys[8], vy = s * phys[7] + c * phys[8], vz = phys[9];
  float more[8] = {(float)(o->body_xyz[2] - ILRL_INITIAL_Z), (float)sin(ang), (float)cos(ang), (float)(0.3 * vx),
                   (float)(0.3 * vy), (float)(0.3 * vz), (float)o->rpy[0], (float)o->rpy[1]};
  for (int i = 0; i < 8; i++) o->obs[i] = clip5(more[i]);
  o->joints_at_limit = 0;
  for (int j = 0; j < NJ; j++) {
    double mid = 0.5 * (joint_lo[j] + joint_hi[j]);
    float rp = (float)(2 * (phys[13 + j] - mid) / (joint_hi[j] - joint_lo[j]));
    float rv = (float)(0.1 * phys[30 + j]);
    o->joint_speeds[j] = rv;
    if (fabsf(rp) > 0.99f) o->joints_at_limit++;
    o->obs[8 + 2 * j] = clip5(rp);
    o->obs[9 + 2 * j] = clip5(rv);
  }
}

void ilrl_oracle_calc_state(const double* phys, double wtx, double wty, float* obs42, double* body_xyz,
                            float* joint_speeds, int* at_limit, double* rpy) {
  orc_calc c;
  orc_calc_state(phys, wtx, wty, &c);
  memcpy(obs42, c.obs, sizeof c.obs);
  memcpy(body_xyz, c.body_xyz, sizeof c.body_xyz);
  memcpy(joint_speeds, c.joint_speeds, sizeof c.joint_speeds);
  *at_limit = c.joints_at_limit;
  memcpy(rpy, c.rpy, sizeof c.rpy);
}

/* ------------------------------------------------------------------ env logic (low level + hierarchical) */
typedef struct {
  const double *pos, *rel, *vel, *ep; /* [n,14] x3, [n,27] */
  int n_pos, n_vel, max_frame;
} orc_clip;

typedef struct {
  int mode; /* 0 = LowLevelHumanoidEnv, 1 = HierarchicalHumanoidEnv */
  orc_clip clip;
  double phys[ILRL_PHYS_WORDS];
  double e[ILRL_ENV_WORDS];
  double terms[ILRL_TERM_WORDS];
  orc_calc cs; /* last calc_state */
  int skip;    /* skipFrame (REF low_level_env.py:162; 0 = the reference's 2) */
} orc_env;

orc_env* ilrl_oracle_env_create(int mode, const double* pos, const double* rel, const double* vel, const double* ep,
                                int n_pos, int n_vel, int max_frame) {
  orc_env* v = (orc_env*)calloc(1, sizeof(orc_env));
  v->mode = mode;
  v->clip.pos = pos; v->clip.rel = rel; v->clip.vel = vel; v->clip.ep = ep;
  v->clip.n_pos = n_pos; v->clip.n_vel = n_vel;
  v->clip.max_frame = max_frame; /* REF low_level_env.py:80-82: len(joints_df) - 1 (clamped by the caller for 13_13) */
  v->phys[6] = 1;                /* identity quaternion */
  v->phys[2] = 1.4;
  return v;
}
void ilrl_oracle_env_destroy(orc_env* v) { free(v); }
/* the reference's drivers assign env.skipFrame; earlier versions of the env (the ones its shipped checkpoints were
 * trained on, REF Log/catatan_low_level.txt:596-640) ran with skipFrame 1 */
void ilrl_oracle_env_set_skip(orc_env* v, int skip) { v->skip = skip; }
void ilrl_oracle_env_get(const orc_env* v, double* phys, double* e, double* terms) {
  memcpy(phys, v->phys, sizeof v->phys);
  memcpy(e, v->e, sizeof v->e);
  if (terms) memcpy(terms, v->terms, sizeof v->terms);
}
void ilrl_oracle_env_set(orc_env* v, const double* phys, const double* e) {
  memcpy(v->phys, phys, sizeof v->phys);
  memcpy(v->e, e, sizeof v->e);
  orc_calc_state(v->phys, v->e[ILRL_E_WALK_X], v->e[ILRL_E_WALK_Y], &v->cs);
  /* the stored sin/cos are those of the calc_state that produced them, which may predate a walk-target change */
  v->cs.obs[1] = (float)v->e[ILRL_E_OBS_SIN];
  v->cs.obs[2] = (float)v->e[ILRL_E_OBS_COS];
}

static void rotz(double rad, const double* v, double* o) {
  double c = cos(rad), s = sin(rad);
  double x = c * v[0] - s * v[1], y = s * v[0] + c * v[1];
  o[0] = x; o[1] = y; o[2] = v[2];
}

static void inc_frame(orc_env* v, int inc) { /* REF low_level_env.py:218-222, hier_env.py:227-233 */
  int f = ((int)v->e[ILRL_E_FRAME] + inc) % (v->clip.max_frame - 1);
  v->e[ILRL_E_FRAME] = f;
  if (f == 0) {
    v->e[ILRL_E_SEP_X] = v->e[ILRL_E_ROBOT_X];
    v->e[ILRL_E_SEP_Y] = v->e[ILRL_E_ROBOT_Y];
    v->e[ILRL_E_SEP_Z] = 0;
  }
}

static void do_calc_state(orc_env* v) {
  orc_calc_state(v->phys, v->e[ILRL_E_WALK_X], v->e[ILRL_E_WALK_Y], &v->cs);
  v->e[ILRL_E_OBS_SIN] = v->cs.obs[1];
  v->e[ILRL_E_OBS_COS] = v->cs.obs[2];
}

void ilrl_oracle_low_obs(const orc_env* v, double* obs70) { /* REF low_level_env.py:307-320, hier_env.py:321-334 */
  int f = (int)v->e[ILRL_E_FRAME];
  for (int i = 0; i < 42; i++) obs70[i] = v->cs.obs[i];
  for (int m = 0; m < ILRL_NMAP; m++) {
    obs70[42 + 2 * m] = v->clip.rel[f * 14 + map_col[m]];
    obs70[43 + 2 * m] = v->clip.vel[f * 14 + map_col[m]];
  }
}

void ilrl_oracle_high_obs(const orc_env* v, double* obs44) { /* REF hier_env.py:336-353 */
  double yaw = v->cs.rpy[2];
  double tt = atan2(v->e[ILRL_E_TARGET_Y] - v->e[ILRL_E_ROBOT_Y], v->e[ILRL_E_TARGET_X] - v->e[ILRL_E_ROBOT_X]);
  double ts = atan2(v->e[ILRL_E_START_Y] - v->e[ILRL_E_ROBOT_Y], v->e[ILRL_E_START_X] - v->e[ILRL_E_ROBOT_X]);
  obs44[0] = v->cs.obs[0];
  obs44[1] = cos(tt - yaw); obs44[2] = sin(tt - yaw);
  obs44[3] = cos(ts - yaw); obs44[4] = sin(ts - yaw);
  for (int i = 3; i < 42; i++) obs44[2 + i] = v->cs.obs[i];
}

/* REF low_level_env.py:247-305 / hier_env.py:235-319.  target_deg: the rng.integers(-180,180) draw of getRandomVec.
 * reset_yaw_deg: low = caller argument (default 0); hier = the reset()'s rng.integers(-180,180) draw. */
void ilrl_oracle_env_reset(orc_env* v, int start_frame, double reset_yaw_deg, int target_deg, double* obs_out) {
  double* e = v->e;
  double* ph = v->phys;
  const orc_clip* c = &v->clip;
  int hier = v->mode == 1;
  double sep_keep[3] = {e[ILRL_E_SEP_X], e[ILRL_E_SEP_Y], e[ILRL_E_SEP_Z]};
  int clip_id = (int)e[ILRL_E_CLIP];
  memset(e, 0, sizeof v->e);
  memset(v->terms, 0, sizeof v->terms);
  e[ILRL_E_CLIP] = clip_id;
  double trad = target_deg * (M_PI / 180.0); /* np.deg2rad */
  e[ILRL_E_TARGET_X] = cos(trad) * ILRL_TARGET_LEN;
  e[ILRL_E_TARGET_Y] = sin(trad) * ILRL_TARGET_LEN;
  e[ILRL_E_FRAME] = start_frame;
  /* setJointsOrientation */
  for (int j = 0; j < NJ; j++) { ph[13 + j] = 0; ph[30 + j] = 0; }
  for (int m = 0; m < ILRL_NMAP; m++) {
    ph[13 + map_joint[m]] = c->pos[start_frame * 14 + map_col[m]];
    ph[30 + map_joint[m]] = c->vel[start_frame * 14 + map_col[m]];
  }
  ph[0] = 0; ph[1] = 0; ph[2] = ILRL_RESET_Z;
  double deg_to_target = atan2(e[ILRL_E_TARGET_Y], e[ILRL_E_TARGET_X]) * (180.0 / M_PI);
  double body_deg;
  if (hier) { deg_to_target += reset_yaw_deg; body_deg = deg_to_target; }
  else body_deg = deg_to_target + reset_yaw_deg;
  e[ILRL_E_WALK_X] = cos(deg_to_target) * 1000; /* degrees fed as radians: mirrored (Q1) */
  e[ILRL_E_WALK_Y] = sin(deg_to_target) * 1000;
  double half = 0.5 * body_deg * (M_PI / 180.0);
  ph[3] = 0; ph[4] = 0; ph[5] = sin(half); ph[6] = cos(half);
  e[ILRL_E_HLDEG] = deg_to_target * (M_PI / 180.0);
  for (int i = 7; i < 13; i++) ph[i] = 0;
  double rot = deg_to_target * (M_PI / 180.0);
  const double* ep0 = c->ep + start_frame * 27;
  if (!hier) {
    const double* ep1 = c->ep + ((start_frame + 2) % c->max_frame) * 27;
    orc_kin k;
    orc_fk(ph, &k);
    double rf[3] = {ep0[9], ep0[10], ep0[11]}, rfr[3]; /* RightFoot columns 9..11 */
    rotz(rot, rf, rfr);
    e[ILRL_E_SEP_X] = k.o[5][0] - rfr[0]; /* body 5 = right_foot */
    e[ILRL_E_SEP_Y] = k.o[5][1] - rfr[1];
    e[ILRL_E_SEP_Z] = 0;
    double a[3] = {ep0[6], ep0[7], ep0[8]}, b[3] = {ep1[6], ep1[7], ep1[8]}, ar[3], br[3]; /* RightLeg 6..8 */
    rotz(rot, a, ar); rotz(rot, b, br);
    for (int i = 0; i < 3; i++) ph[7 + i] = ((br[i] - ar[i]) / 0.0165) / 1.2;
  } else {
    const double* ep1 = c->ep + (start_frame + 1) * 27;
    e[ILRL_E_SEP_X] = sep_keep[0]; e[ILRL_E_SEP_Y] = sep_keep[1]; e[ILRL_E_SEP_Z] = sep_keep[2];
    double a[3] = {ep0[6], ep0[7], ep0[8]}, b[3] = {ep1[6], ep1[7], ep1[8]}, ar[3], br[3];
    rotz(rot, a, ar); rotz(rot, b, br);
    for (int i = 0; i < 3; i++) ph[7 + i] = (br[i] - ar[i]) / 0.0165;
    e[ILRL_E_HIGH_TARGET_SCORE] = -ILRL_TARGET_LEN;
    e[ILRL_E_STEPS_REMAINING] = 5;
    e[ILRL_E_HIGH_PENDING] = 1;
  }
  inc_frame(v, v->skip > 0 ? v->skip : 2);
  do_calc_state(v);
  if (obs_out) {
    if (hier) ilrl_oracle_high_obs(v, obs_out);
    else ilrl_oracle_low_obs(v, obs_out);
  }
}

static double norm2(double x, double y) { return sqrt(x * x + y * y); }

/* shared by low env and hier env: REF low_level_env.py:441-465 / hier_env.py:494-522 */
static double update_reward(orc_env* v, const double* action) {
  double* e = v->e;
  double* t = v->terms;
  const orc_clip* c = &v->clip;
  int f = (int)e[ILRL_E_FRAME];
  double dj = 0, dvv = 0;
  for (int m = 0; m < ILRL_NMAP; m++) {
    dj += fabs(v->phys[13 + map_joint[m]] - c->pos[f * 14 + map_col[m]]) * map_w[m];
    dvv += fabs(v->phys[30 + map_joint[m]] - c->vel[f * 14 + map_col[m]]) * map_wv[m];
  }
  double joint_score = exp(4 * (-dj / ILRL_JOINT_W_SUM));
  double jvel_score = exp((-dvv / ILRL_JOINT_WV_SUM) / 2);
  double low_target = v->mode == 1 ? 0.0 : -norm2(e[ILRL_E_TARGET_X] - e[ILRL_E_ROBOT_X], e[ILRL_E_TARGET_Y] - e[ILRL_E_ROBOT_Y]);
  double posture = exp(-(fabs(v->cs.rpy[2] - e[ILRL_E_HLDEG]) + fabs(v->cs.rpy[0]) + fabs(v->cs.rpy[1])));
  t[ILRL_T_DLOWTARGET] = (low_target - e[ILRL_E_LOW_TARGET_SCORE]) / 0.0165 * 0.1;
  e[ILRL_E_JOINT_SCORE] = joint_score; e[ILRL_E_JVEL_SCORE] = jvel_score;
  e[ILRL_E_LOW_TARGET_SCORE] = low_target; e[ILRL_E_POSTURE_SCORE] = posture;
  double run = 0, stall = 0;
  for (int i = 0; i < NJ; i++) { run += fabs(action[i] * (double)v->cs.joint_speeds[i]); stall += action[i] * action[i]; }
  t[ILRL_T_ELEC] = -1.0 * (run / NJ) + -0.1 * (stall / NJ);
  t[ILRL_T_LIMIT] = -0.1 * v->cs.joints_at_limit;
  float z = v->cs.obs[0] + (float)ILRL_INITIAL_Z;
  t[ILRL_T_ALIVE] = z > (float)ILRL_ALIVE_Z ? 2 : -1;
  t[ILRL_T_JOINT] = joint_score; t[ILRL_T_JVEL] = jvel_score; t[ILRL_T_POSTURE] = posture; t[ILRL_T_LOWTARGET] = low_target;
  if (v->mode == 1) {
    e[ILRL_E_CUM_ALIVE] += t[ILRL_T_ALIVE];
    /* calcDriftScore: math_util.projPointLineSegment(robot_pos, starting_robot_pos, target) */
    double lx = e[ILRL_E_TARGET_X] - e[ILRL_E_START_X], ly = e[ILRL_E_TARGET_Y] - e[ILRL_E_START_Y];
    double len = norm2(lx, ly);
    double tt = ((e[ILRL_E_ROBOT_X] - e[ILRL_E_START_X]) * lx + (e[ILRL_E_ROBOT_Y] - e[ILRL_E_START_Y]) * ly) / (len * len);
    tt = tt < 0 ? 0 : (tt > 1 ? 1 : tt);
    double px = e[ILRL_E_START_X] + tt * lx, py = e[ILRL_E_START_Y] + tt * ly;
    e[ILRL_E_CUM_DRIFT] += exp(-6 * norm2(px - e[ILRL_E_ROBOT_X], py - e[ILRL_E_ROBOT_Y]));
  }
  return ILRL_RW_JOINT * t[ILRL_T_JOINT] + ILRL_RW_JVEL * t[ILRL_T_JVEL] + ILRL_RW_TARGET * t[ILRL_T_DLOWTARGET] +
         ILRL_RW_ELEC * t[ILRL_T_ELEC] + ILRL_RW_LIMIT * t[ILRL_T_LIMIT] + ILRL_RW_ALIVE * t[ILRL_T_ALIVE] +
         ILRL_RW_POSTURE * t[ILRL_T_POSTURE];
}

static void check_target(orc_env* v, int rand_deg) { /* REF low_level_env.py:412-434 / hier_env.py:469-487 */
  double* e = v->e;
  double dist = norm2(e[ILRL_E_ROBOT_X] - e[ILRL_E_TARGET_X], e[ILRL_E_ROBOT_Y] - e[ILRL_E_TARGET_Y]);
  if (dist <= ILRL_TARGET_REACHED) {
    double rr = v->cs.rpy[2] + rand_deg * (M_PI / 180.0);
    double nx = e[ILRL_E_ROBOT_X] + cos(rr) * ILRL_TARGET_LEN, ny = e[ILRL_E_ROBOT_Y] + sin(rr) * ILRL_TARGET_LEN;
    e[ILRL_E_START_X] = e[ILRL_E_TARGET_X]; e[ILRL_E_START_Y] = e[ILRL_E_TARGET_Y];
    e[ILRL_E_TARGET_X] = nx; e[ILRL_E_TARGET_Y] = ny;
    double sc = -norm2(nx - e[ILRL_E_START_X], ny - e[ILRL_E_START_Y]);
    if (v->mode == 1) e[ILRL_E_HIGH_TARGET_SCORE] = sc;
    else e[ILRL_E_LOW_TARGET_SCORE] = sc;
  }
  if (v->mode == 0) {
    e[ILRL_E_HLDEG] = atan2(e[ILRL_E_TARGET_Y] - e[ILRL_E_ROBOT_Y], e[ILRL_E_TARGET_X] - e[ILRL_E_ROBOT_X]);
    e[ILRL_E_WALK_X] = e[ILRL_E_ROBOT_X] + cos(e[ILRL_E_HLDEG]) * 10;
    e[ILRL_E_WALK_Y] = e[ILRL_E_ROBOT_Y] + sin(e[ILRL_E_HLDEG]) * 10;
  }
}

static int check_done(const orc_env* v) { /* REF low_level_env.py:467-473 / hier_env.py:573-581 */
  const double* e = v->e;
  int alive = v->terms[ILRL_T_ALIVE] > 0;
  double margin = v->mode == 1 ? ILRL_DONE_MARGIN_HI : ILRL_DONE_MARGIN_LOW;
  int near = norm2(e[ILRL_E_TARGET_X] - e[ILRL_E_ROBOT_X], e[ILRL_E_TARGET_Y] - e[ILRL_E_ROBOT_Y]) <=
             norm2(e[ILRL_E_TARGET_X] - e[ILRL_E_START_X], e[ILRL_E_TARGET_Y] - e[ILRL_E_START_Y]) + margin;
  return !(alive && near);
}

static void update_reward_high(orc_env* v) { /* REF hier_env.py:524-536 */
  double* e = v->e;
  double hs = -norm2(e[ILRL_E_TARGET_X] - e[ILRL_E_ROBOT_X], e[ILRL_E_TARGET_Y] - e[ILRL_E_ROBOT_Y]);
  double d = 5 - e[ILRL_E_STEPS_REMAINING] + 1;
  v->terms[ILRL_T_DHIGHTARGET] = (hs - e[ILRL_E_HIGH_TARGET_SCORE]) / 0.0165 / d; /* delta_highTargetScore */
  e[ILRL_E_HIGH_TARGET_SCORE] = hs;
  v->terms[ILRL_T_DRIFT] = e[ILRL_E_CUM_DRIFT] / d;
  e[ILRL_E_CUM_DRIFT] = 0;
}

/* LowLevelHumanoidEnv.step — REF low_level_env.py:475-526.  phys_flags forwarded to the physics (0 in normal use).
 * If skip_physics != 0 the state is taken as already advanced (reward/obs-only harness, the K3 entry point). */
int ilrl_oracle_low_step(orc_env* v, const double* action, int rand_deg, int skip_physics, double* obs70,
                         double* reward) {
  double tau[NJ];
  ilrl_oracle_action_to_torque(action, tau);
  if (!skip_physics) ilrl_oracle_physics_step(v->phys, tau, 0);
  do_calc_state(v);
  v->e[ILRL_E_ROBOT_X] = v->cs.body_xyz[0];
  v->e[ILRL_E_ROBOT_Y] = v->cs.body_xyz[1];
  *reward = update_reward(v, action);
  inc_frame(v, v->skip > 0 ? v->skip : 2);
  check_target(v, rand_deg);
  v->terms[ILRL_T_LOWTARGET] = v->e[ILRL_E_LOW_TARGET_SCORE]; /* attribute as read after the step */
  ilrl_oracle_low_obs(v, obs70);
  int done = check_done(v);
  v->e[ILRL_E_T] += 1;
  if (v->e[ILRL_E_T] >= 3000) done = 1;
  return done;
}

/* HierarchicalHumanoidEnv.step with the high-level agent's action — REF hier_env.py:355-366, 538-571 */
void ilrl_oracle_high_step(orc_env* v, const double* action2, double* low_obs70) {
  double* e = v->e;
  e[ILRL_E_ROBOT_X] = v->cs.body_xyz[0]; e[ILRL_E_ROBOT_Y] = v->cs.body_xyz[1];
  double adeg = atan2(action2[1], action2[0]) * (180.0 / M_PI);
  double ndeg = adeg + v->cs.rpy[2] * (180.0 / M_PI);
  e[ILRL_E_HLDEG] = ndeg * (M_PI / 180.0);
  double ct = cos(e[ILRL_E_HLDEG]), st = sin(e[ILRL_E_HLDEG]);
  double wx = e[ILRL_E_ROBOT_X] + ct * 5, wy = e[ILRL_E_ROBOT_Y] + st * 5;
  e[ILRL_E_WALK_X] = wx; e[ILRL_E_WALK_Y] = wy;
  double vx = wx - e[ILRL_E_ROBOT_X], vy = wy - e[ILRL_E_ROBOT_Y], vz = 0;
  double dx = e[ILRL_E_SEP_X] - e[ILRL_E_ROBOT_X], dy = e[ILRL_E_SEP_Y] - e[ILRL_E_ROBOT_Y], dz = e[ILRL_E_SEP_Z];
  double len = sqrt(dx * dx + dy * dy + dz * dz), vn = sqrt(vx * vx + vy * vy + vz * vz);
  e[ILRL_E_SEP_X] = -vx / vn * len + e[ILRL_E_ROBOT_X];
  e[ILRL_E_SEP_Y] = -vy / vn * len + e[ILRL_E_ROBOT_Y];
  e[ILRL_E_SEP_Z] = -vz / vn * len + 0;
  e[ILRL_E_STEPS_REMAINING] = 5;
  e[ILRL_E_HIGH_PENDING] = 0;
  ilrl_oracle_low_obs(v, low_obs70);
}

/* HierarchicalHumanoidEnv.step with the low-level agent's action — REF hier_env.py:355-366, 583-642.
 * returns bit0 = done["__all__"], bit1 = high-level agent present in the returned dicts (reward/obs valid) */
int ilrl_oracle_hier_low_step(orc_env* v, const double* action, int rand_deg, int skip_physics, double* low_obs70,
                              double* low_reward, double* high_obs44, double* high_reward) {
  double* e = v->e;
  e[ILRL_E_ROBOT_X] = v->cs.body_xyz[0]; e[ILRL_E_ROBOT_Y] = v->cs.body_xyz[1]; /* stale by one step (Q13) */
  e[ILRL_E_STEPS_REMAINING] -= 1;
  double tau[NJ];
  ilrl_oracle_action_to_torque(action, tau);
  if (!skip_physics) ilrl_oracle_physics_step(v->phys, tau, 0);
  do_calc_state(v);
  *low_reward = update_reward(v, action);
  inc_frame(v, v->skip > 0 ? v->skip : 2);
  check_target(v, rand_deg);
  int done = check_done(v);
  e[ILRL_E_T] += 1;
  int ret = 0;
  *high_reward = 0;
  if (done || e[ILRL_E_T] >= 3000) {
    update_reward_high(v);
    *high_reward = v->terms[ILRL_T_DHIGHTARGET] * 0.3 + v->terms[ILRL_T_DRIFT] * 0.7;
    ilrl_oracle_high_obs(v, high_obs44);
    e[ILRL_E_CUM_ALIVE] = 0;
    ret = 3;
  } else if (e[ILRL_E_STEPS_REMAINING] <= 0) {
    update_reward_high(v);
    *high_reward = v->terms[ILRL_T_DHIGHTARGET] * 0.3 + v->terms[ILRL_T_DRIFT] * 0.7;
    ilrl_oracle_high_obs(v, high_obs44);
    e[ILRL_E_CUM_ALIVE] = 0;
    e[ILRL_E_HIGH_PENDING] = 1;
    ret = 2;
  }
  v->terms[ILRL_T_HIGHTARGET] = e[ILRL_E_HIGH_TARGET_SCORE];
  ilrl_oracle_low_obs(v, low_obs70);
  return ret;
}

/* calcEndPointScore — REF low_level_env.py:361-382 (not on any step path; exported term) */
double ilrl_oracle_endpoint_score(const orc_env* v) {
  orc_kin k;
  orc_fk(v->phys, &k);
  const double* ep = v->clip.ep + (int)v->e[ILRL_E_FRAME] * 27;
  /* link0_11 -> RightLeg (w1), right_foot -> RightFoot (w3), link0_18 -> LeftLeg (w1), left_foot -> LeftFoot (w3) */
  const double* part[4] = {k.rw[6], k.o[5], k.rw[10], k.o[8]};
  const int col[4] = {6, 9, 0, 3};
  const double wgt[4] = {1, 3, 1, 3};
  double s = 0;
  for (int i = 0; i < 4; i++) {
    double r[3], d[3];
    rotz(v->e[ILRL_E_HLDEG], ep + col[i], r);
    d[0] = v->e[ILRL_E_SEP_X] + r[0] - part[i][0];
    d[1] = v->e[ILRL_E_SEP_Y] + r[1] - part[i][1];
    d[2] = v->e[ILRL_E_SEP_Z] + r[2] - part[i][2];
    s += sqrt(dot3(d, d)) * wgt[i];
  }
  return exp(3 * (-s / ILRL_EP_W_SUM));
}

const double* ilrl_oracle_env_phys(orc_env* v) { return v->phys; }
double* ilrl_oracle_env_phys_mut(orc_env* v) { return v->phys; }
void ilrl_oracle_env_refresh(orc_env* v) { do_calc_state(v); }
void ilrl_oracle_env_terms(const orc_env* v, double* terms) { memcpy(terms, v->terms, sizeof v->terms); }

/* ------------------------------------------------------------------ CPU baseline driver (bench.py only)
 * Advance `n` low-level envs by `steps` env steps each with uniform random actions in [-1,1] and reset-on-done,
 * entirely in C (no Python in the loop).  Returns the number of env steps executed; *episodes gets the resets. */
static uint64_t sm64(uint64_t* s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static int g_action_mode = 0; /* 0: uniform in [-1,1] (bench workload); 1: N(0,1) clipped to [-1,1] (an untrained RLlib Gaussian policy) */
void ilrl_oracle_rollout_action_mode(int m) { g_action_mode = m; }
static double sm_gauss(uint64_t* st) {
  double u1 = ((double)(sm64(st) >> 11) + 1.0) * (1.0 / 9007199254740993.0), u2 = (double)(sm64(st) >> 11) * (1.0 / 9007199254740992.0);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
long ilrl_oracle_rollout(orc_env** envs, int n, int steps, uint64_t seed, long* episodes, double* reward_sum) {
  uint64_t st = seed * 0x2545F4914F6CDD1Dull + 1;
  long done_steps = 0, eps = 0;
  double rs = 0;
  for (int t = 0; t < steps; t++)
    for (int i = 0; i < n; i++) {
      orc_env* v = envs[i];
      double a[NJ], obs[70], rew;
      for (int k = 0; k < NJ; k++) {
        if (g_action_mode == 0) a[k] = (double)(sm64(&st) >> 11) * (2.0 / 9007199254740992.0) - 1.0;
        else { double g = sm_gauss(&st); a[k] = g < -1 ? -1 : (g > 1 ? 1 : g); }
      }
      int deg = (int)(sm64(&st) % 360) - 180;
      int done = ilrl_oracle_low_step(v, a, deg, 0, obs, &rew);
      rs += rew;
      done_steps++;
      if (done) {
        int sf = (int)(sm64(&st) % (uint64_t)(v->clip.max_frame - 5));
        int td = (int)(sm64(&st) % 360) - 180;
        ilrl_oracle_env_reset(v, sf, 0.0, td, obs);
        eps++;
      }
    }
  if (episodes) *episodes = eps;
  if (reward_sum) *reward_sum = rs;
  return done_steps;
}
