// ilrl_capi.cu — kernels + the C ABI declared in include/ilrl.h (libilrl_b200.so, sm_100a only, no CPU path).
//
// Data layout in HBM (per handle, N envs):
//   phys  [N][48] fp32  one 16-byte-aligned row per env (47 words used): the four lanes that step an env move its row, and
//   envf  [N][28] fp32  a row stays one contiguous block when envs are reached through the cost-grouping permutation (K8).
//                       Persistent state = 75 words + 1 rng counter = 304 B / env.  (Rounds 1-2 held it word-major,
//                       [47][N]: fine for neighbours in a warp, 8 sectors per access once the envs of a warp are scattered.)
//   rng   [N]     u32   Philox draw counter
//   clips: the 4 tables of every loaded motion clip, row-major fp32, read through the read-only path (L2-resident,
//          ~200 KB for all four clips).
// I/O with the caller (row-major [N,17] actions, [N,70] observations) is staged through shared memory so that global
// accesses are coalesced although every quad produces/consumes a whole row.
// The fused step kernel (K1) maps FOUR LANES to one env and is persistent (tiles of 16 envs handed out through an
// atomic counter); its physics lives in ilrl_chain.cuh, the env logic in ilrl_env.cuh.  The service kernels (reset,
// high-level step, harness) map one thread to one env.
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ilrl.h"
#include "ilrl_chain.cuh"

namespace ilrl {

constexpr int BLOCK = 64;  // threads per CTA of the thread-per-env service kernels (reset, high step, harness)
using chain::QE;
using chain::QT;
using SmemSmall = chain::SmemT<chain::LayoutSmall>;
using SmemLarge = chain::SmemT<chain::LayoutLarge>;
using SmemDense4 = chain::SmemT<chain::LayoutDense4>;
using SmemSelf = chain::SmemT<chain::LayoutSelf>;

struct StepArgs {
  int n;               // envs of the handle (stride of the structure-of-arrays state)
  int first, end;      // this launch steps envs [first, end)  (the whole batch, or one part of it: ilrl_step_host_async)
  int skip_frame;      // reference frames per env step (REF low_level_env.py:162)
  int id_base;         // global id of env 0 of this handle: the Philox counter of env i is id_base + i
  int skip_physics;
  int auto_reset;
  int max_timestep;
  float step_per_level;
  uint64_t seed;
  float* phys;         // [n][ILRL_PHYS_STRIDE]
  float* envf;         // [n][ILRL_ENV_STRIDE]
  uint32_t* rng;       // [n]
  const float* action; // [n,17]
  float* obs;          // [n,70]
  float* reward;       // [n]
  uint8_t* done;       // [n]
  float* terms;        // [n,12] or null
  float* high_obs;     // [n,44]   (hier)
  float* high_reward;  // [n]
  uint8_t* high_flags; // [n]
  const int32_t* forced_deg;  // [n] or null
  int forced_scalar;          // INT_MIN, or the heading every env uses at its next target re-sampling (ilrl_step_pull)
  double* stats;       // [16] or null: fp64 accumulators (counts stay exact past 2^24 env steps between two reads)
  float* jt;           // [n][ILRL_JT_STRIDE] jointTarget (MODE 2: hier_env_2.py:731)
  const float* forced_noise;  // [n,17] or null: joint noise an auto-reset uses instead of its own draws (MODE 2 harness)
  float* gscr;         // [n][GROWS][RW] overflow scratch for constraint rows beyond the shared-memory budget
  unsigned int* tile_counter;  // [2]: next tile to hand out, CTAs that have left (both zero between launches)
  int ntiles;
  const int* perm = nullptr;     // null, or [n]: the env that batch position p steps (envs grouped by cost, heaviest first: K8)
  uint8_t* cost = nullptr;       // null, or [n]: the key of that grouping, 0 = most expensive
  unsigned long long* ktime;   // null, or {first CTA start, last CTA end} of this launch in %globaltimer ns (ilrl_kernel_timing)
#ifdef ILRL_PROF
  long long* prof;     // [warps of the launch][PF_WORDS] phase cycles (measurement build only)
#endif
  chain::Terrain terr; // heightfield terrain (terr.h null: flat ground; only read by the TERR instantiations)
  chain::SelfC selfc;  // self-collision scratch (only read by the SELFC instantiations)
  ClipDesc clips[MAX_CLIPS];
};

__device__ __forceinline__ void load_state(const StepArgs& a, int i, Phys& s, EnvW& w) {
  const float* p = a.phys + (size_t)i * ILRL_PHYS_STRIDE;
#pragma unroll
  for (int k = 0; k < 3; k++) s.p[k] = p[0 + k];
#pragma unroll
  for (int k = 0; k < 4; k++) s.quat[k] = p[3 + k];
#pragma unroll
  for (int k = 0; k < 3; k++) s.v[k] = p[7 + k];
#pragma unroll
  for (int k = 0; k < 3; k++) s.w[k] = p[10 + k];
#pragma unroll
  for (int k = 0; k < NJ; k++) { s.q[k] = p[13 + k]; s.qd[k] = p[30 + k]; }
  const float* e = a.envf + (size_t)i * ILRL_ENV_STRIDE;
#pragma unroll
  for (int k = 0; k < ILRL_ENV_WORDS; k++) w.e[k] = e[k];
}
__device__ __forceinline__ void store_state(const StepArgs& a, int i, const Phys& s, const EnvW& w) {
  float* p = a.phys + (size_t)i * ILRL_PHYS_STRIDE;
#pragma unroll
  for (int k = 0; k < 3; k++) p[0 + k] = s.p[k];
#pragma unroll
  for (int k = 0; k < 4; k++) p[3 + k] = s.quat[k];
#pragma unroll
  for (int k = 0; k < 3; k++) p[7 + k] = s.v[k];
#pragma unroll
  for (int k = 0; k < 3; k++) p[10 + k] = s.w[k];
#pragma unroll
  for (int k = 0; k < NJ; k++) { p[13 + k] = s.q[k]; p[30 + k] = s.qd[k]; }
  float* e = a.envf + (size_t)i * ILRL_ENV_STRIDE;
#pragma unroll
  for (int k = 0; k < ILRL_ENV_WORDS; k++) e[k] = w.e[k];
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// shared-memory setup common to the quad kernels: model tables in (layouts that stage them)
template <class SM>
__device__ __forceinline__ void quad_smem_init(SM& sm) {
  if constexpr (SM::TSM) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&chain::kTables);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sm.T);
    // all loads of a thread in flight together (the copy is on every CTA's critical path)
    constexpr int PER = (chain::TABLE_WORDS + QT - 1) / QT;
    uint32_t v[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) { const int t = threadIdx.x + k * QT; v[k] = t < chain::TABLE_WORDS ? __ldg(src + t) : 0u; }
#pragma unroll
    for (int k = 0; k < PER; k++) { const int t = threadIdx.x + k * QT; if (t < chain::TABLE_WORDS) dst[t] = v[k]; }
  }
  if constexpr (SM::HSM) {   // the per-link part only (layouts that read the rest of the tables through L1)
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&chain::kTables.lc[0][0]);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sm.H.lc[0][0]);
    constexpr int PER = (chain::HOT_LC_WORDS + QT - 1) / QT;
    uint32_t v[PER];
#pragma unroll
    for (int k = 0; k < PER; k++) { const int t = threadIdx.x + k * QT; v[k] = t < chain::HOT_LC_WORDS ? __ldg(src + t) : 0u; }
#pragma unroll
    for (int k = 0; k < PER; k++) { const int t = threadIdx.x + k * QT; if (t < chain::HOT_LC_WORDS) dst[t] = v[k]; }
    if (threadIdx.x < 9) sm.H.Q[threadIdx.x] = chain::kTables.Q[threadIdx.x];
  }
  if constexpr (SM::SELF) {   // wide row 0 of every env must hold finite numbers (lanes past their count re-evaluate it)
    for (int k = 0; k < chain::WRW / 4; k++) sm.S.wr[threadIdx.x >> 2][(threadIdx.x & 3) * (chain::WRW / 4) + k] = 0.f;
    for (int p = threadIdx.x; p < chain::NSELF; p += QT) {   // the pair table, packed
      const float r = chain::kSelfReach[p];
      sm.S.pairs[p] = make_uint4((unsigned)chain::kSelfA0[p] | ((unsigned)chain::kSelfA1[p] << 8) |
                                     ((unsigned)chain::kSelfB0[p] << 16) | ((unsigned)chain::kSelfB1[p] << 24),
                                 __float_as_uint(4.f * r * r),
                                 __float_as_uint(kSphereR[chain::kSelfA0[p]] + kSphereR[chain::kSelfB0[p]] + (float)ILRL_CONTACT_BREAK), 0u);
    }
  }
}
// torques of this lane's chain into the link records: apply_action (REF humanoid.py:54-60): clip, gear x power,
// motor slot -> joint.  act: the env's action row in shared memory (null: joint-order torques in `torque`)
template <class SM>
__device__ __forceinline__ void set_torques(SM& sm, int e, int tid, int role, const float* act, const float* torque) {
#pragma unroll
  for (int c = 0; c < chain::NL; c++) {
    float* rec = chain::link_rec(sm, c, e, tid);
    const chain::LinkC& L = c < 3 ? chain::link_c(sm, 4, c) : chain::link_c(sm, role, c - 3);
    float t = 0.f;
    if (L.j >= 0) t = act ? L.gear * fminf(fmaxf(act[L.motor], -1.f), 1.f) : torque[L.j];
    rec[chain::W_TAU] = t;
  }
}

// ------------------------------------------------------------------------------------------------ K1: fused step
// MODE 0 = LowLevelHumanoidEnv.step (REF low_level_env.py:475-526), MODE 1 = HierarchicalHumanoidEnv low_level_step
// (REF hier_env.py:355-366, 583-642), MODE 2 = the hier_env_2.py variant (REF hier_env_2.py:408-419, 751-769: obs rows
// of 72 / 60 words, per-env jointTarget).  Four lanes = one env (ilrl_chain.cuh): the physics substeps run distributed
// over the quad; the env bookkeeping after them is computed redundantly by the four lanes (identical instruction
// stream, no divergence) and the outputs are dealt to the lanes for the stores.
// One tile (QE envs) of one env step: everything between a CTA picking its tile and moving on.  Shared by the
// launch-per-step kernel and the persistent serving kernel (K1s).  action: the [n,17] action array of this step.
template <int MODE, class SM, bool TERR, bool SELFC = false>
__device__ __forceinline__ void step_tile(const StepArgs& a, SM& sm, const int tile, const float* action, const int ko = 0) {
  const int tid = threadIdx.x, e = tid >> 2, role = tid & 3, qb = tid & ~3;
  // (ko: index of this step in a sequence launch, K1q - its outputs go to slab ko of [K, n, ...] arrays)
  float* const o_reward = a.reward + (size_t)ko * a.n;
  uint8_t* const o_done = a.done + (size_t)ko * a.n;
  float* const o_terms = a.terms ? a.terms + (size_t)ko * a.n * ILRL_TERM_WORDS : nullptr;
  const unsigned qm = 0xFu << ((tid & 31) & ~3);
  constexpr int OBSW = MODE == 2 ? ILRL_OBS_LOW2 : ILRL_OBS_LOW, HOBSW = MODE == 2 ? ILRL_OBS_HIGH2 : ILRL_OBS_HIGH;
  static_assert(SM::ES >= OBSW, "the obs row is staged in the env's scratch block");
  const int base = a.first + tile * QE;
  const bool valid = base + e < a.end;
  const int i = (valid && a.perm) ? a.perm[base + e] : base + e;
  chain::Prof pf;
#ifdef ILRL_PROF
  if (a.prof && (tid & 31) == 0) pf.p = a.prof + (size_t)(tile * (QT / 32) + (tid >> 5)) * chain::PF_WORDS;
  const long long t_tile = clock64();
#endif
  pf.start();
  // Head of a tile: every global load it needs (action row, waiting flag, base and chain state) is issued before any of
  // them is consumed - one memory round trip instead of three dependent ones (actions -> flag -> state).
  chain::Base b;
  float av[(NJ + 3) / 4], qv[chain::NL], qdv[chain::NL], pend_flag = 0.f;
#pragma unroll
  for (int m = 0; m < (NJ + 3) / 4; m++) av[m] = 0.f;
  if (valid) {  // (the 8 action rows of a warp are one contiguous 544-byte block)
    const float* arow = action + (size_t)i * NJ;
#pragma unroll
    for (int m = 0; m < (NJ + 3) / 4; m++) if (role + 4 * m < NJ) av[m] = arow[role + 4 * m];
    if (MODE >= 1) pend_flag = a.envf[(size_t)i * ILRL_ENV_STRIDE + ILRL_E_HIGH_PENDING];
    chain::load_base(a.phys, a.n, i, b);
    chain::load_links(a.phys, a.n, i, sm, role, qv, qdv);
    float* dst = sm.act(e);
#pragma unroll
    for (int m = 0; m < (NJ + 3) / 4; m++) if (role + 4 * m < NJ) dst[role + 4 * m] = av[m];
  }
  // row 0 of every env must hold finite numbers: lanes past their own row count evaluate it with a zero step
  reinterpret_cast<float4*>(&sm.rows[e][0])[role] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  bool write_obs = false;
  float st_ep = 0.f, st_ret = 0.f, st_len = 0.f, st_steps = 0.f, st_rew = 0.f, st_terms[11];
#pragma unroll
  for (int t = 0; t < 11; t++) st_terms[t] = 0.f;

  // skipped envs: hier envs waiting for a high-level action, and rows whose first action component is NaN (the
  // batched adapters' "no action for this env in this call"; the reference asserts finite actions, REF humanoid.py:55)
  const bool pending = valid && ((MODE >= 1 && pend_flag != 0.f) || isnan(sm.act(e)[0]));
  if (valid && pending) {
    if (role == 0) { o_reward[i] = 0.f; o_done[i] = 0; if (a.cost) a.cost[i] = 63; }
    if (o_terms)
      for (int t = role; t < ILRL_TERM_WORDS; t += 4) o_terms[(size_t)i * ILRL_TERM_WORDS + t] = 0.f;
  }
  // envs that step.  The physics substeps run between CTA barriers: the warps of a CTA (and, because tiles start
  // together, mostly the CTAs of an SM) then execute the same 70 KB of substep code at about the same time and share
  // the instruction cache — +4 % at 16384 envs, +5 % at 65536 (6 to 8 unsynchronised warps per SM otherwise thrash it).
  // Every lane runs the substeps (compile-time full shuffle masks, chain::substep): the lanes of an env that does
  // not step carry a benign dummy state that is never stored.
  const bool active = valid && !pending;
  float act[NJ];
  float stale_x = 0.f, stale_y = 0.f;
  float sumx = 0.f, sumy = 0.f, rfx = 0.f, rfy = 0.f;
  if (active) {
    chain::store_links(sm, e, tid, qv, qdv);
#pragma unroll
    for (int m = 0; m < NJ; m++) act[m] = sm.act(e)[m];
    set_torques(sm, e, tid, role, sm.act(e), nullptr);
    __syncwarp(qm);
    if (MODE >= 1) {
      // (Q13) robot_pos is refreshed from the PREVIOUS calc_state at the top of step()
      chain::pose_sums(b, sm, e, tid, role, qm, sumx, sumy, rfx, rfy);
      stale_x = (32.f * b.p[0] + sumx) * (1.f / 33.f);
      stale_y = (32.f * b.p[1] + sumy) * (1.f / 33.f);
    }
  } else {
    chain::dummy_state(sm, e, tid, b);
  }
  int rows_last = 0;
  if (!a.skip_physics) {
    float* gscr_tile = a.gscr + (size_t)min(base, a.n - 1) * chain::GROWS * chain::RW;
    chain::SelfC sct = a.selfc;

#pragma unroll 1
    for (int sub = 0; sub < ILRL_SUBSTEPS; sub++) {
      pf.mark(sub == 0 ? chain::PF_HEAD : chain::PF_INTEG);
#ifndef ILRL_NO_SUBSTEP_BAR
      __syncthreads();
#endif
      pf.mark(chain::PF_BARRIER);
      const int nrows = chain::substep<TERR, SELFC>(b, sm, gscr_tile, e, tid, role, active,
                                                    (float)(ILRL_FRAME_DT / ILRL_SUBSTEPS), pf, TERR ? &a.terr : nullptr,
                                                    SELFC ? &sct : nullptr);
      if (sub == ILRL_SUBSTEPS - 1) rows_last = nrows;
    }
  }
#ifndef ILRL_NO_TAIL_BAR
  __syncthreads();   // the once-per-step tail (4 k instructions) is entered together as well: +3 % at 65536 envs
#endif
  pf.mark(chain::PF_BARRIER);
  if (active) {
    chain::pose_sums(b, sm, e, tid, role, qm, sumx, sumy, rfx, rfy);
    Phys ps;
    chain::gather(b, sm, e, qb, qm, ps);
    pf.mark(chain::PF_T_POSE);
    EnvW w;
    {
      const float* ew = a.envf + (size_t)i * ILRL_ENV_STRIDE;
#pragma unroll
      for (int k = 0; k < ILRL_ENV_WORDS; k++) w.e[k] = ew[k];
    }
    if (MODE >= 1) {
      w.e[ILRL_E_ROBOT_X] = stale_x; w.e[ILRL_E_ROBOT_Y] = stale_y;
      w.e[ILRL_E_STEPS_REMAINING] -= 1.f;
    }
    const ClipDesc cl = a.clips[(int)w.e[ILRL_E_CLIP]];
    Calc c;
    float terms[ILRL_TERM_WORDS];
#pragma unroll
    for (int t = 0; t < ILRL_TERM_WORDS; t++) terms[t] = 0.f;
    calc_state(ps, sumx, sumy, w.e[ILRL_E_WALK_X], w.e[ILRL_E_WALK_Y], c);
    w.e[ILRL_E_OBS_SIN] = c.obs[1]; w.e[ILRL_E_OBS_COS] = c.obs[2];
    if (MODE == 0) { w.e[ILRL_E_ROBOT_X] = c.bx; w.e[ILRL_E_ROBOT_Y] = c.by; }
    float jt[MODE == 2 ? ILRL_JT_WORDS : 1];
    if (MODE == 2) {
#pragma unroll
      for (int k = 0; k < ILRL_JT_WORDS; k++) jt[k] = a.jt[(size_t)i * ILRL_JT_STRIDE + k];
    }
    float reward;
    if constexpr (MODE == 2) reward = update_reward2(c, w, jt, act, terms);   // (no frame advance in hier_env_2's low step)
    else { reward = update_reward<MODE>(ps, c, w, cl, act, terms); inc_frame(w, cl, a.skip_frame); }
    uint32_t ctr = a.rng[i];
    int deg;
    if (a.forced_deg && a.forced_deg[i] != INT_MIN) deg = a.forced_deg[i];
    else if (a.forced_scalar != INT_MIN) deg = a.forced_scalar;
    else {
      // the draw is consumed (counter advanced) only when the target actually switches
      uint32_t c2 = ctr;
      deg = rand_int(a.seed, (uint32_t)(i + a.id_base), c2, -180, 180);
      float dist = hyp(w.e[ILRL_E_ROBOT_X] - w.e[ILRL_E_TARGET_X], w.e[ILRL_E_ROBOT_Y] - w.e[ILRL_E_TARGET_Y]);
      if (dist <= (float)ILRL_TARGET_REACHED) ctr = c2;
    }
    check_target<MODE>(c, w, deg);
    if (MODE != 2) terms[ILRL_T_LOWTARGET] = w.e[ILRL_E_LOW_TARGET_SCORE];
    bool done = check_done<MODE>(w, terms[ILRL_T_ALIVE]);
    w.e[ILRL_E_T] += 1.f;
    if (w.e[ILRL_E_T] >= (float)a.max_timestep) done = true;
    w.e[ILRL_E_EP_RETURN] += reward;
    w.e[ILRL_E_EP_LEN] += 1.f;
    pf.mark(chain::PF_T_REWARD);
    float* so = &sm.scr[e][0];  // the env's scratch block is free after the substeps: stage the obs row there
    {
      float obs[OBSW];
      if constexpr (MODE == 2) write_low_obs2(c.obs, jt, obs); else write_low_obs(c.obs, w, cl, obs);
#pragma unroll
      for (int t = 0; t < OBSW; t++) if ((t & 3) == role) so[t] = obs[t];
    }
    write_obs = true;
    uint8_t hflags = 0;
    if (MODE >= 1) {
      if (done || w.e[ILRL_E_STEPS_REMAINING] <= 0.f) {
        float ho[HOBSW], hr;
        if constexpr (MODE == 2) {
          hr = update_reward_high2(ps, c, w, cl, terms, a.step_per_level);
          write_high_obs2(c, w, cl, ho);
        } else {
          update_reward_high(w, terms, a.step_per_level);
          hr = terms[ILRL_T_DHIGHTARGET] * 0.3f + terms[ILRL_T_DRIFT] * 0.7f;
          write_high_obs(c, w, ho);
          w.e[ILRL_E_CUM_ALIVE] = 0.f;
        }
        if (role == 0) a.high_reward[i] = hr;
#pragma unroll
        for (int t = 0; t < HOBSW; t++) if ((t & 3) == role) a.high_obs[(size_t)i * HOBSW + t] = ho[t];
        hflags = done ? 3 : 2;
        if (!done) { w.e[ILRL_E_HIGH_PENDING] = 1.f; hflags |= 4; }
      }
      terms[ILRL_T_HIGHTARGET] = w.e[ILRL_E_HIGH_TARGET_SCORE];
    }
    if (role == 0) { o_reward[i] = reward; o_done[i] = done ? 1 : 0; }
    if (o_terms) {
#pragma unroll
      for (int t = 0; t < ILRL_TERM_WORDS; t++) if ((t & 3) == role) o_terms[(size_t)i * ILRL_TERM_WORDS + t] = terms[t];
    }
    // (statistics: the four lanes of a quad hold the same values; lane `role` reduces slots = role mod 4 below)
    st_steps = 1.f; st_rew = reward;
#pragma unroll
    for (int t = 0; t < 11; t++) st_terms[t] = terms[t];
    if (done) {
      st_ep = 1.f; st_ret = w.e[ILRL_E_EP_RETURN]; st_len = w.e[ILRL_E_EP_LEN];
      if (a.auto_reset) {
        int sf = rand_int(a.seed, (uint32_t)(i + a.id_base), ctr, 0, cl.max_frame - 5);
        float yaw = MODE >= 1 ? (float)rand_int(a.seed, (uint32_t)(i + a.id_base), ctr, -180, 180) : 0.f;
        int tdeg = rand_int(a.seed, (uint32_t)(i + a.id_base), ctr, -180, 180);
        float noise[MODE == 2 ? NJ : 1];
        if (MODE == 2) {
          draw_reset_noise(a.seed, (uint32_t)(i + a.id_base), ctr, noise);
          if (a.forced_noise) {
#pragma unroll
            for (int j = 0; j < NJ; j++) noise[j] = a.forced_noise[(size_t)i * NJ + j];
          }
        }
        ResetCtx rx;
        reset_pose<MODE>(ps, w, cl, sf, yaw, tdeg, rx, nullptr, MODE == 2 ? noise : nullptr);
        chain::scatter(ps, sm, e, qb, role, qm, b);
        chain::pose_sums(b, sm, e, tid, role, qm, sumx, sumy, rfx, rfy);
        reset_finish<MODE>(ps, w, cl, rx, rfx, rfy, sumx, sumy, a.step_per_level, a.skip_frame, c);
        if constexpr (MODE == 0) {
          float obs[70];
          write_low_obs(c.obs, w, cl, obs);
#pragma unroll
          for (int t = 0; t < 70; t++) if ((t & 3) == role) so[t] = obs[t];
        } else {
          float ho[HOBSW];
          if constexpr (MODE == 2) write_high_obs2(c, w, cl, ho); else write_high_obs(c, w, ho);
#pragma unroll
          for (int t = 0; t < HOBSW; t++) if ((t & 3) == role) a.high_obs[(size_t)i * HOBSW + t] = ho[t];
          hflags |= 4;
        }
      }
    }
    pf.mark(chain::PF_T_OBS);
    if (role == 0) {
      if (MODE >= 1) a.high_flags[i] = hflags;
      a.rng[i] = ctr;
      // K8 key, 0 = most expensive: torso height (a fallen env tests and touches the ground with many spheres and is
      // about to end and auto-reset - the longest path of the tail) less 5 cm per constraint row of the last substep;
      // 1.6 cm buckets (tools/sort_experiment.py compares the candidates)
      if (a.cost) {
        const float ze = fminf(ps.p[2], 1.3f) - 0.05f * (float)((done && a.auto_reset) ? 0 : rows_last);
        a.cost[i] = (uint8_t)min(max((int)((ze - 0.3f) * 64.f), 0), 63);
      }
    }
    chain::store_phys(a.phys, a.n, i, role, ps);
    {
      float* ew = a.envf + (size_t)i * ILRL_ENV_STRIDE;
#pragma unroll
      for (int k = 0; k < ILRL_ENV_WORDS; k++) if ((k & 3) == role) ew[k] = w.e[k];
    }
    pf.mark(chain::PF_T_STORE);
  }
  // observations: only envs that stepped write their row (pending hier envs keep theirs).  Each warp stores the rows of
  // its own 8 envs (one contiguous 2240-byte block of the caller's array) from the staging area, 8 bytes per lane.
  {
    const unsigned okmask = __ballot_sync(0xffffffffu, write_obs);   // (includes the staging writes' __syncwarp)
    const int e0 = e & ~7, lane = tid & 31;
    float* const oall = a.obs + (size_t)ko * a.n * OBSW;
#pragma unroll 3   // (three rounds of loads in flight: +0.2 % at 4096 envs, +0.9 % at 16384 / 65536; 5 or 9: the same)
    for (int f0 = 0; f0 < 8 * (OBSW / 2); f0 += 32) {
      const int f = f0 + lane, r = min(f / (OBSW / 2), 7), c2 = f - r * (OBSW / 2);
      const int ir = __shfl_sync(0xffffffffu, i, 4 * r);   // (without grouping the 8 rows are one contiguous block)
      if (f < 8 * (OBSW / 2) && ((okmask >> (4 * r)) & 1u))
        *reinterpret_cast<float2*>(oall + (size_t)ir * OBSW + 2 * c2) = *reinterpret_cast<const float2*>(&sm.scr[e0 + r][2 * c2]);
    }
  }
  // K5: episode / reward statistics -> one atomicAdd per warp per slot
  if (a.stats) {
    float v[16] = {st_ep, st_ret, st_len, st_steps, st_rew};
#pragma unroll
    for (int t = 0; t < 11; t++) v[5 + t] = st_terms[t];
    // slot 4 m + role is summed over the warp's 8 envs by the lanes of that role (12 shuffles instead of 48; the same
    // butterfly over the same envs, so the sums are the ones lane 0 used to form alone)
#pragma unroll
    for (int m = 0; m < 4; m++) {
      float x = role == 0 ? v[4 * m] : role == 1 ? v[4 * m + 1] : role == 2 ? v[4 * m + 2] : v[4 * m + 3];
#pragma unroll
      for (int o = 16; o >= 4; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if ((tid & 31) < 4 && x != 0.f) atomicAdd(a.stats + 4 * m + role, (double)x);
    }
  }
  pf.mark(chain::PF_TAIL);
#ifdef ILRL_PROF
  if (pf.p) pf.p[chain::PF_TOTAL] += clock64() - t_tile;
#endif
  __syncthreads();  // every warp is done with this tile's shared memory (and with s_tile) before the next one
}


template <int MODE, class SM, bool TERR = false, bool SELFC = false>
__global__ void __launch_bounds__(QT) step_kernel(const StepArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  SM& sm = *reinterpret_cast<SM*>(smraw);
  const int tid = threadIdx.x;
  // Programmatic dependent launch: consecutive step kernels of a stream form a chain, and everything up to here depends
  // on nothing the previous step wrote.  The trigger lets the NEXT launch's CTAs become resident as this one's finish (a
  // step ends with its slowest warp: most SMs idle for the last ~20 % of it) and run their prologue - model tables into
  // shared memory - early; the wait below holds them until this grid has completed and flushed.  (No-ops when the launch
  // does not carry the attribute.)
  asm volatile("griddepcontrol.launch_dependents;");
  quad_smem_init(sm);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (a.ktime && tid == 0) atomicMin(a.ktime, globaltimer_ns());
  // Persistent CTAs: the grid is at most what is resident at once (SMs x CTAs per SM) and tiles of QE envs are handed
  // out through an atomic counter, so a large batch has no partially filled last wave and no per-tile table copy, and
  // CTAs that drew cheap tiles simply take more of them.  The last CTA to leave resets the counters (graph-safe).
  // (The first tile of a CTA is its own index: a batch that fits one wave never touches the counter.)
  __shared__ int s_tile;
  const bool one_wave = a.ntiles <= (int)gridDim.x;
  for (int it = 0;; it++) {
  if (it > 0) {
    if (one_wave) break;
    if (tid == 0) s_tile = (int)gridDim.x + (int)atomicAdd(a.tile_counter, 1u);
  } else if (tid == 0) s_tile = (int)blockIdx.x;
  __syncthreads();
  const int tile = s_tile;
  if (tile >= a.ntiles) break;
  step_tile<MODE, SM, TERR, SELFC>(a, sm, tile, a.action);
  }  // tile loop
  if (!one_wave && tid == 0 && atomicAdd(a.tile_counter + 1, 1u) == gridDim.x - 1) {  // last CTA out: ready for the next launch
    a.tile_counter[0] = 0u;
    a.tile_counter[1] = 0u;
  }
  if (a.ktime && tid == 0) atomicMax(a.ktime + 1, globaltimer_ns());
}

// ------------------------------------------------------------------------------------------------ K8: cost grouping
// A warp steps 8 envs in lock step and pays for the one with the most constraint rows (warp-uniform Gauss-Seidel loops
// over the warp's maximum, row items pooled per warp).  In batches of more than one wave, where throughput follows the
// MEAN warp time, envs of similar cost are therefore put into the same warps: every few steps this kernel sorts the env
// ids by a cost key the step kernel leaves per env (torso height and constraint rows of the last substep, see step_tile)
// into `perm`, most expensive first - the tile queue then also hands out the long tiles first.  Results do
// not depend on the grouping (an env's arithmetic never involves its warp neighbours): bit-identical, tests/test_gpu_api.py.
// One CTA, counting sort on 64 keys: warp-private histograms in shared memory (an atomic only ever conflicts with lanes of
// its own warp; __match_any_sync costs a round per distinct key and was 10x slower), scan in (key, warp) order, scatter.
// (The order of equal keys inside one 32-env group is whatever the hardware serialises; results do not depend on it.)
// cost: 32 chunks of `chunk` bytes (a multiple of 512; entries past n hold 255 = "no env").  Warp w owns chunk w and reads it
// 16 keys per lane and load, four loads in flight.
template <bool SCATTER>
__device__ __forceinline__ void cost_sort_pass(const uint8_t* cost, int* perm, const int chunk, unsigned (*hist)[64],
                                               const unsigned* start) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint4* src = reinterpret_cast<const uint4*>(cost + (size_t)w * chunk);
  const int nld = chunk / 512;
  for (int l0 = 0; l0 < nld; l0 += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) v[u] = l0 + u < nld ? src[(l0 + u) * 32 + lane] : make_uint4(~0u, ~0u, ~0u, ~0u);
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const unsigned wd[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int b = 0; b < 16; b++) {
        const int raw = (int)((wd[b >> 2] >> (8 * (b & 3))) & 255u);
        if (raw != 255) {   // (0 = most expensive: first)
          const unsigned at = atomicAdd(&hist[w][min(raw, 63)], 1u);   // warp-private counters: conflicts only inside this warp
          if (SCATTER) perm[start[min(raw, 63)] + at] = w * chunk + ((l0 + u) * 32 + lane) * 16 + b;
        }
      }
    }
  }
}
__global__ void __launch_bounds__(1024) cost_sort_kernel(const uint8_t* __restrict__ cost, int* __restrict__ perm, const int chunk) {
  __shared__ unsigned hist[32][64];
  __shared__ unsigned start[64];
  for (int k = threadIdx.x; k < 32 * 64; k += 1024) (&hist[0][0])[k] = 0u;
  __syncthreads();
  cost_sort_pass<false>(cost, perm, chunk, hist, start);
  __syncthreads();
  if (threadIdx.x < 64) {   // exclusive scan in (key, warp) order
    unsigned s = 0;
    for (int ww = 0; ww < 32; ww++) { const unsigned c = hist[ww][threadIdx.x]; hist[ww][threadIdx.x] = s; s += c; }
    start[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned s = 0;
    for (int k = 0; k < 64; k++) { const unsigned c = start[k]; start[k] = s; s += c; }
  }
  __syncthreads();
  cost_sort_pass<true>(cost, perm, chunk, hist, start);
}
__global__ void iota_kernel(int* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

// ------------------------------------------------------------------------------------------------ K1q: action sequences
// K consecutive env steps of the low-level env in ONE launch, for callers that hold the actions of all K steps before
// the first one runs (open-loop playback of recorded or scripted action sequences, random-action rollouts).  Envs never
// interact, so a CTA takes its tile through all K steps without waiting for any other CTA: the grid-wide barrier that a
// launch per step implies - every step ends with its slowest warp, ~25 % above the mean at 4096 envs -
// is paid once per K steps instead of once per step.  action [K,n,17]; obs [K,n,70], reward / done [K,n], terms [K,n,12].
template <class SM>
__global__ void __launch_bounds__(QT) step_seq_kernel(const StepArgs a, const int ksteps) {
  extern __shared__ __align__(16) unsigned char smraw[];
  SM& sm = *reinterpret_cast<SM*>(smraw);
  quad_smem_init(sm);
  if (a.ktime && threadIdx.x == 0) atomicMin(a.ktime, globaltimer_ns());
  __syncthreads();
  for (int tile = (int)blockIdx.x; tile < a.ntiles; tile += (int)gridDim.x)
    for (int k = 0; k < ksteps; k++)   // (step_tile ends with a CTA barrier: step k's stores are visible to step k+1's loads)
      step_tile<0, SM, false>(a, sm, tile, a.action + (size_t)k * a.n * NJ, k);
  if (a.ktime && threadIdx.x == 0) atomicMax(a.ktime + 1, globaltimer_ns());
}

// ------------------------------------------------------------------------------------------------ K1s: persistent serving
// The end-to-end path pays, per env step, a kernel launch and a stream synchronisation (~13 us of a 90 us step) although
// the kernel itself is latency-bound at ~75 us.  Between ilrl_serve_begin and ilrl_serve_end ONE launch of this kernel
// stays resident (the batch must fit one wave: every CTA owns one tile for the whole session) and the host drives it
// through a doorbell in mapped, page-locked host memory:
//   host:  actions already in their mapped buffer -> ctl.action = its device alias, ctl.seq = t (store order; x86 TSO)
//   CTA 0: polls ctl.seq over PCIe, republishes {action, seq} in device memory; the other CTAs poll that copy in L2
//   CTAs:  one step_tile each (reading the actions from / writing obs, reward, done to the mapped host buffers in place)
//   last CTA to finish (device counter): system-scope fence, then ack = t into mapped host memory; the host spins on it.
// A watchdog ends the kernel when no doorbell arrives for `timeout_ns` (a host that died must not leave the GPU busy).
struct ServeCtl {            // mapped host memory, written by the host
  volatile long long action; // device alias of this step's [n,17] actions
  volatile int seq;          // step number, > 0; -1 = quit
  int pad;
};
struct ServeDev {            // device memory
  volatile long long action;
  volatile int seq;
  unsigned int arrive;
};
template <int MODE, class SM>
__global__ void __launch_bounds__(QT) serve_kernel(const StepArgs a, const ServeCtl* ctl, ServeDev* dv, volatile int* ack,
                                                   unsigned long long timeout_ns) {
  extern __shared__ __align__(16) unsigned char smraw[];
  SM& sm = *reinterpret_cast<SM*>(smraw);
  const int tid = threadIdx.x;
  __shared__ int s_seq;
  __shared__ long long s_action;
  quad_smem_init(sm);
  int last = 0;
  for (;;) {
    if (tid == 0) {
      int q;
      long long act = 0;
      if (blockIdx.x == 0) {
        const unsigned long long t0 = globaltimer_ns();
        while ((q = ctl->seq) == last) {
          if (globaltimer_ns() - t0 > timeout_ns) { q = -2; break; }
          __nanosleep(200);
        }
        if (q > 0) act = ctl->action;   // (written before seq by the host)
        dv->action = act;
        __threadfence();
        dv->seq = q;
      } else {
        while ((q = dv->seq) == last) __nanosleep(100);
        __threadfence();
        act = dv->action;
      }
      s_seq = q; s_action = act;
    }
    __syncthreads();
    const int seq = s_seq;
    if (seq < 0) break;
    last = seq;
    if ((int)blockIdx.x < a.ntiles)
      step_tile<MODE, SM, false>(a, sm, (int)blockIdx.x, reinterpret_cast<const float*>(s_action));
    if (tid == 0) {
      __threadfence_system();   // this CTA's rows have left for host memory before it is counted
      if (atomicAdd(&dv->arrive, 1u) == gridDim.x - 1) {
        dv->arrive = 0u;
        __threadfence_system();
        *ack = seq;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ K2: reset
struct ResetArgs {
  int n;
  int skip_frame;
  int id_base;
  float step_per_level;
  uint64_t seed;
  float* phys; float* envf; uint32_t* rng;
  const uint8_t* mask; const int32_t* start_frame; const int32_t* target_deg; const float* yaw_deg; const float* target_xy;
  const float* noise;   // [n,17] or null (MODE 2: joint noise of robot_specific_reset; null = drawn)
  float* obs;
  uint8_t* high_flags;
  ClipDesc clips[MAX_CLIPS];
};
struct StateView { int n; float* phys; float* envf; uint32_t* rng; };

template <int MODE>
__global__ void __launch_bounds__(BLOCK) reset_kernel(const ResetArgs a) {
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= a.n) return;
  if (a.mask && !a.mask[i]) return;
  StepArgs v; v.n = a.n; v.phys = a.phys; v.envf = a.envf;
  Phys s; EnvW w;
  load_state(v, i, s, w);
  const ClipDesc cl = a.clips[(int)w.e[ILRL_E_CLIP]];
  uint32_t ctr = a.rng[i];
  int sf = a.start_frame ? a.start_frame[i] : rand_int(a.seed, (uint32_t)(i + a.id_base), ctr, 0, cl.max_frame - 5);
  float yaw = a.yaw_deg ? a.yaw_deg[i] : (MODE >= 1 ? (float)rand_int(a.seed, (uint32_t)(i + a.id_base), ctr, -180, 180) : 0.f);
  int tdeg = (a.target_deg || a.target_xy) ? (a.target_deg ? a.target_deg[i] : 0) : rand_int(a.seed, (uint32_t)(i + a.id_base), ctr, -180, 180);
  float txy[2] = {0.f, 0.f};
  if (a.target_xy) { txy[0] = a.target_xy[2 * i]; txy[1] = a.target_xy[2 * i + 1]; }
  Work k; Calc c;
  float noise[NJ];
  if (MODE == 2) {
    draw_reset_noise(a.seed, (uint32_t)(i + a.id_base), ctr, noise);
    if (a.noise) for (int j = 0; j < NJ; j++) noise[j] = a.noise[(size_t)i * NJ + j];
  }
  reset_env<MODE>(s, w, cl, sf, yaw, tdeg, a.step_per_level, a.skip_frame, k, c, a.target_xy ? txy : nullptr,
                  MODE == 2 ? noise : nullptr);
  a.rng[i] = ctr;
  store_state(v, i, s, w);
  if (a.obs) {
    if (MODE == 0) {
      float o[70];
      write_low_obs(c.obs, w, cl, o);
      for (int t = 0; t < 70; t++) a.obs[(size_t)i * 70 + t] = o[t];
    } else if (MODE == 1) {
      float o[44];
      write_high_obs(c, w, o);
      for (int t = 0; t < 44; t++) a.obs[(size_t)i * 44 + t] = o[t];
    } else {
      float o[ILRL_OBS_HIGH2];
      write_high_obs2(c, w, cl, o);
      for (int t = 0; t < ILRL_OBS_HIGH2; t++) a.obs[(size_t)i * ILRL_OBS_HIGH2 + t] = o[t];
    }
  }
  if (MODE >= 1 && a.high_flags) a.high_flags[i] = 4;
}

// ------------------------------------------------------------------------------------------------ K4: high-level step
struct HighArgs {
  int n; float step_per_level; int skip_frame;
  float* phys; float* envf; float* jt;
  const float* action2; float* low_obs;   // MODE 1: [n,2] / [n,70]; MODE 2: [n,36] / [n,72]
  ClipDesc clips[MAX_CLIPS];
};
template <int MODE>
__global__ void __launch_bounds__(BLOCK) high_step_kernel(const HighArgs a) {
  constexpr int AW = MODE == 2 ? ILRL_ACT_HIGH2 : ILRL_ACT_HIGH;
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= a.n) return;
  StepArgs v; v.n = a.n; v.phys = a.phys; v.envf = a.envf;
  Phys s; EnvW w;
  load_state(v, i, s, w);
  if (w.e[ILRL_E_HIGH_PENDING] == 0.f || isnan(a.action2[AW * i])) return;  // not waiting / no action in this call
  const ClipDesc cl = a.clips[(int)w.e[ILRL_E_CLIP]];
  Work k; Calc c;
  fk(s, k);
  calc_state(s, k.sumx, k.sumy, w.e[ILRL_E_WALK_X], w.e[ILRL_E_WALK_Y], c);
  c.obs[1] = w.e[ILRL_E_OBS_SIN]; c.obs[2] = w.e[ILRL_E_OBS_COS];  // cur_obs predates the walk-target change below
  w.e[ILRL_E_ROBOT_X] = c.bx; w.e[ILRL_E_ROBOT_Y] = c.by;
  if constexpr (MODE == 2) {
    // REF hier_env_2.py:699-741: the action's tail becomes the joint targets, the frame advances, nothing else changes
    float jt[ILRL_JT_WORDS];
    for (int t = 0; t < ILRL_JT_WORDS; t++) {
      jt[t] = a.action2[(size_t)AW * i + 2 + t];
      a.jt[(size_t)i * ILRL_JT_STRIDE + t] = jt[t];
    }
    w.e[ILRL_E_STEPS_REMAINING] = a.step_per_level;
    w.e[ILRL_E_HIGH_PENDING] = 0.f;
    inc_frame<2>(w, cl, a.skip_frame);
    float o[ILRL_OBS_LOW2];
    write_low_obs2(c.obs, jt, o);
    for (int t = 0; t < ILRL_OBS_LOW2; t++) a.low_obs[(size_t)i * ILRL_OBS_LOW2 + t] = o[t];
    float* e = a.envf + (size_t)i * ILRL_ENV_STRIDE;
#pragma unroll
    for (int kk = 0; kk < ILRL_ENV_WORDS; kk++) e[kk] = w.e[kk];
    return;
  }
  const float R2D = 57.29577951308232f, D2R = 0.017453292519943295f;
  float ndeg = atan2f(a.action2[2 * i + 1], a.action2[2 * i]) * R2D + c.yaw * R2D;
  float h = ndeg * D2R, sh, ch;
  w.e[ILRL_E_HLDEG] = h;
  sincosf(h, &sh, &ch);
  float wx = w.e[ILRL_E_ROBOT_X] + ch * 5.f, wy = w.e[ILRL_E_ROBOT_Y] + sh * 5.f;
  w.e[ILRL_E_WALK_X] = wx; w.e[ILRL_E_WALK_Y] = wy;
  float vx = wx - w.e[ILRL_E_ROBOT_X], vy = wy - w.e[ILRL_E_ROBOT_Y];
  float dx = w.e[ILRL_E_SEP_X] - w.e[ILRL_E_ROBOT_X], dy = w.e[ILRL_E_SEP_Y] - w.e[ILRL_E_ROBOT_Y], dz = w.e[ILRL_E_SEP_Z];
  float len = sqrtf(dx * dx + dy * dy + dz * dz), ivn = rsqrtf(vx * vx + vy * vy);
  w.e[ILRL_E_SEP_X] = -vx * ivn * len + w.e[ILRL_E_ROBOT_X];
  w.e[ILRL_E_SEP_Y] = -vy * ivn * len + w.e[ILRL_E_ROBOT_Y];
  w.e[ILRL_E_SEP_Z] = 0.f;
  w.e[ILRL_E_STEPS_REMAINING] = a.step_per_level;
  w.e[ILRL_E_HIGH_PENDING] = 0.f;
  float o[70];
  write_low_obs(c.obs, w, cl, o);
  for (int t = 0; t < 70; t++) a.low_obs[(size_t)i * 70 + t] = o[t];
  float* e = a.envf + (size_t)i * ILRL_ENV_STRIDE;
#pragma unroll
  for (int kk = 0; kk < ILRL_ENV_WORDS; kk++) e[kk] = w.e[kk];
}

// ------------------------------------------------------------------------------------------------ harness kernels
__global__ void state_get_kernel(StateView v, float* phys_aos, float* envf_aos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n) return;
  if (phys_aos) for (int k = 0; k < ILRL_PHYS_WORDS; k++) phys_aos[(size_t)i * ILRL_PHYS_WORDS + k] = v.phys[(size_t)i * ILRL_PHYS_STRIDE + k];
  if (envf_aos) for (int k = 0; k < ILRL_ENV_WORDS; k++) envf_aos[(size_t)i * ILRL_ENV_WORDS + k] = v.envf[(size_t)i * ILRL_ENV_STRIDE + k];
}
__global__ void state_set_kernel(StateView v, const float* phys_aos, const float* envf_aos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n) return;
  if (phys_aos) for (int k = 0; k < ILRL_PHYS_WORDS; k++) v.phys[(size_t)i * ILRL_PHYS_STRIDE + k] = phys_aos[(size_t)i * ILRL_PHYS_WORDS + k];
  if (envf_aos) for (int k = 0; k < ILRL_ENV_WORDS; k++) v.envf[(size_t)i * ILRL_ENV_STRIDE + k] = envf_aos[(size_t)i * ILRL_ENV_WORDS + k];
}
// resident jointTarget rows (stride ILRL_JT_STRIDE) <-> caller's [n,34] (exactly one of out / in is non-null)
__global__ void jt_copy_kernel(int n, float* jt, float* out, const float* in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < ILRL_JT_WORDS; k++) {
    if (out) out[(size_t)i * ILRL_JT_WORDS + k] = jt[(size_t)i * ILRL_JT_STRIDE + k];
    else jt[(size_t)i * ILRL_JT_STRIDE + k] = in[(size_t)i * ILRL_JT_WORDS + k];
  }
}
__global__ void clip_ids_kernel(StateView v, const int32_t* ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n) return;
  v.envf[(size_t)i * ILRL_ENV_STRIDE + ILRL_E_CLIP] = ids ? (float)ids[i] : 0.f;
}
template <class SM, bool TERR = false, bool SELFC = false>
__global__ void __launch_bounds__(QT) physics_only_kernel(StateView v, const float* torque, float* gscr_all, int nsub,
                                                          const chain::Terrain terr = chain::Terrain(),
                                                          const chain::SelfC selfc = chain::SelfC()) {
  extern __shared__ __align__(16) unsigned char smraw[];
  SM& sm = *reinterpret_cast<SM*>(smraw);
  const int tid = threadIdx.x, e = tid >> 2, role = tid & 3, qb = tid & ~3;
  const unsigned qm = 0xFu << ((tid & 31) & ~3);
  const int i = blockIdx.x * QE + e;
  quad_smem_init(sm);
  __syncthreads();
  const bool valid = i < v.n;
  chain::Base b;
  if (valid) {
    float qv[chain::NL], qdv[chain::NL];
    chain::load_base(v.phys, v.n, i, b);
    chain::load_links(v.phys, v.n, i, sm, role, qv, qdv);
    chain::store_links(sm, e, tid, qv, qdv);
    set_torques(sm, e, tid, role, nullptr, torque + (size_t)i * NJ);
  } else {
    chain::dummy_state(sm, e, tid, b);
  }
  // row 0 of every env must hold finite numbers (as in step_tile): lanes past their own row count evaluate it with a zero
  // step, and with self-collision an env may have rows of the wide kind only
  reinterpret_cast<float4*>(&sm.rows[e][0])[role] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
  float* gscr_tile = gscr_all + (size_t)blockIdx.x * QE * chain::GROWS * chain::RW;
  for (int sub = 0; sub < nsub; sub++)
  {
    chain::Prof pf;
    chain::SelfC sct = selfc;

    chain::substep<TERR, SELFC>(b, sm, gscr_tile, e, tid, role, valid, (float)(ILRL_FRAME_DT / ILRL_SUBSTEPS), pf,
                                TERR ? &terr : nullptr, SELFC ? &sct : nullptr);
  }
  if (!valid) return;
  Phys ps;
  chain::gather(b, sm, e, qb, qm, ps);
  chain::store_phys(v.phys, v.n, i, role, ps);
}
struct EpArgs { StateView v; float* score; ClipDesc clips[MAX_CLIPS]; };
__global__ void __launch_bounds__(BLOCK) endpoint_kernel(const EpArgs a) {
  const int i = blockIdx.x * BLOCK + threadIdx.x;
  if (i >= a.v.n) return;
  StepArgs sa; sa.n = a.v.n; sa.phys = a.v.phys; sa.envf = a.v.envf;
  Phys s; EnvW w;
  load_state(sa, i, s, w);
  Work k;
  fk(s, k);
  a.score[i] = endpoint_score(s, k, w, a.clips[(int)w.e[ILRL_E_CLIP]]);
}
// Generalised advantage estimation over a [T, N] rollout: one thread per env walks its column backwards
// (adjacent threads = adjacent envs: every access of a warp is one coalesced line).
__global__ void gae_kernel(const float* __restrict__ rew, const float* __restrict__ val, const uint8_t* __restrict__ done,
                           float gamma, float lam, float* __restrict__ adv, float* __restrict__ ret, int T, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
  float vnext = val[(size_t)T * n + i];
  for (int t = T - 1; t >= 0; t--) {
    const size_t k = (size_t)t * n + i;
    const float nd = done[k] ? 0.f : 1.f;
    const float v = val[k];
    const float delta = rew[k] + gamma * vnext * nd - v;
    a = delta + gamma * lam * nd * a;
    adv[k] = a;
    ret[k] = a + v;
    vnext = v;
  }
}
// RLlib's per-step bookkeeping columns of a [T, N] fragment from its done flags: `t` (step index inside the episode) and
// `eps_id` (unique per episode: global env id << 32 | episode counter).  One thread per env walks its column; the
// carries (t of the env's next step, its episode counter) continue from fragment to fragment.
__global__ void episode_columns_kernel(const uint8_t* __restrict__ done, int32_t* t_carry, int32_t* eps_carry,
                                       int32_t* __restrict__ t_out, long long* __restrict__ eps_out, int T, int n,
                                       long long id_base) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int t = t_carry[i], ep = eps_carry[i];
  const long long hi = (id_base + i) << 32;
  for (int k = 0; k < T; k++) {
    const size_t j = (size_t)k * n + i;
    t_out[j] = t;
    eps_out[j] = hi | (long long)(unsigned)ep;
    if (done[j]) { t = 0; ep++; } else t++;
  }
  t_carry[i] = t; eps_carry[i] = ep;
}
// GAE for the HIGH-level agent of the hierarchical env, whose decisions are irregular in time: per env, a decision is
// taken at every tick whose flag has bit2 (waiting) and its outcome — reward, end of episode, the next decision's value
// — is what ilrl_high_readout reports at the next tick with bit1 (REF hier_env.py:524-536, 613-631).  Rows 0..T of
// reward / flags / value are the readouts before tick t (row T: after the last step).  A decision whose outcome lies
// beyond the fragment is marked invalid (valid = 0) and contributes only its value as the bootstrap of the one before.
__global__ void gae_decisions_kernel(const float* __restrict__ rew, const uint8_t* __restrict__ flags,
                                     const float* __restrict__ val, float gamma, float lam, float* __restrict__ adv,
                                     float* __restrict__ ret, uint8_t* __restrict__ valid, int T, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int tau = -1;          // nearest later tick that reports an outcome
  float a_tau = 0.f;     // advantage of the decision taken at tau (0 if none / invalid / episode ended there)
  {
    const uint8_t f = flags[(size_t)T * n + i];
    if (f & 2) tau = T;
  }
  for (int t = T - 1; t >= 0; t--) {
    const size_t k = (size_t)t * n + i;
    const uint8_t f = flags[k];
    float a = 0.f;
    bool ok = false;
    if (f & 4) {
      if (tau >= 0) {
        const size_t kt = (size_t)tau * n + i;
        const bool ended = flags[kt] & 1;
        const float v = val[k];
        const float delta = rew[kt] + (ended ? 0.f : gamma * val[kt]) - v;
        a = delta + (ended ? 0.f : gamma * lam * a_tau);
        adv[k] = a;
        ret[k] = a + v;
        ok = true;
      } else {
        adv[k] = 0.f;
        ret[k] = val[k];
      }
    } else {
      adv[k] = 0.f;
      ret[k] = 0.f;
    }
    valid[k] = ok ? 1 : 0;
    if (f & 2) { tau = t; a_tau = ok ? a : 0.f; }
  }
}
// everything a reference-shaped env object mirrors after a call, one packed row per env (ILRL_PULL_* layout of ilrl.h):
// obs 72 | reward | done | terms 12 | envf 28 | phys 47 | high obs 60 | high reward | high flags | jointTarget 34
// (modes 0 / 1 fill 70 / 44 of the obs columns, the rest is 0)
struct PullArgs {
  int n, obs_w, hobs_w;
  const float* jt;
  const float *obs, *reward, *terms, *phys, *envf, *high_obs, *high_reward;
  const uint8_t *done, *high_flags;
  float* out;
};
__global__ void pull_kernel(const PullArgs a) {
  const int i = blockIdx.x;
  float* o = a.out + (size_t)i * ILRL_PULL_WORDS;
  for (int t = threadIdx.x; t < ILRL_PULL_WORDS; t += blockDim.x) {
    float v;
    if (t < 72) v = t < a.obs_w ? a.obs[(size_t)i * a.obs_w + t] : 0.f;
    else if (t == 72) v = a.reward[i];
    else if (t == 73) v = (float)a.done[i];
    else if (t < 86) v = a.terms[(size_t)i * ILRL_TERM_WORDS + (t - 74)];
    else if (t < 114) v = a.envf[(size_t)i * ILRL_ENV_STRIDE + (t - 86)];
    else if (t < 161) v = a.phys[(size_t)i * ILRL_PHYS_STRIDE + (t - 114)];
    else if (t < 221) v = (t - 161) < a.hobs_w ? a.high_obs[(size_t)i * a.hobs_w + (t - 161)] : 0.f;
    else if (t == 221) v = a.high_reward[i];
    else if (t == 222) v = (float)a.high_flags[i];
    else v = a.jt ? a.jt[(size_t)i * ILRL_JT_STRIDE + (t - 223)] : 0.f;
    o[t] = v;
  }
}
// the high-level agent's outputs kept in the handle -> caller buffers (any of them may be null): one launch
__global__ void high_readout_kernel(int n, int hobs_w, const float* __restrict__ obs, const float* __restrict__ rew,
                                    const uint8_t* __restrict__ flags, float* obs_out, float* rew_out, uint8_t* flags_out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (obs_out && t < hobs_w * n) obs_out[t] = obs[t];
  if (t < n) {
    if (rew_out) rew_out[t] = rew[t];
    if (flags_out) flags_out[t] = flags[t];
  }
}
__global__ void stats_fetch_kernel(double* acc, float* out) {
  int t = threadIdx.x;
  if (t < ILRL_STATS_WORDS) { out[t] = (float)acc[t]; acc[t] = 0.0; }
}

}  // namespace ilrl

// ================================================================================================ host side (C ABI)
using namespace ilrl;

constexpr int ILRL_MAX_PARTS = 8;
constexpr int KT_SLOTS = 8192;   // timed launches between two ilrl_kernel_timing calls
struct ilrl_env {
  ilrl_config cfg;
  int n;
  float* phys = nullptr;
  float* envf = nullptr;
  uint32_t* rng = nullptr;
  float* gscr = nullptr;
  float* jt = nullptr;                    // [n][ILRL_JT_STRIDE] jointTarget (mode 2 only)
  chain::Terrain terr = {nullptr, 0, 0, 0.f, 0.f, 1.f};   // ilrl_set_heightfield (mode 0); h = device copy owned by the handle
  float* terr_mem = nullptr;
  chain::SelfC selfc = {0};
  bool self_on = false;
  const float* forced_noise = nullptr;    // ilrl_set_forced_reset_noise
  int obs_w = ILRL_OBS_LOW, hobs_w = ILRL_OBS_HIGH, hact_w = ILRL_ACT_HIGH;   // row widths of the mode
  float* high_obs = nullptr;
  float* high_reward = nullptr;
  uint8_t* high_flags = nullptr;
  double* stats = nullptr;
  unsigned int* tile_counter = nullptr;   // [1 + ILRL_MAX_PARTS][2]: the whole batch, then one pair per part
  cudaStream_t part_stream[ILRL_MAX_PARTS] = {nullptr};   // ilrl_step_host_async: one stream per part
  bool part_busy[ILRL_MAX_PARTS] = {false};
  int32_t* clip_ids_dev = nullptr;        // staging for ilrl_set_clip_ids
  std::unordered_map<const void*, void*> alias;   // page-locked host buffer -> its device alias (host_alias)
#ifdef ILRL_PROF
  long long* prof = nullptr;
#endif
  int grid_small = 0, grid_large = 0, grid_dense4 = 0;  // resident CTAs of the step kernel in each layout
  float* clip_mem[MAX_CLIPS] = {nullptr};
  ClipDesc clips[MAX_CLIPS];
  bool clip_loaded[MAX_CLIPS] = {false};
  const int32_t* forced_deg = nullptr;
  // host-buffer path
  float *h_action = nullptr, *h_obs = nullptr, *h_reward = nullptr, *h_terms = nullptr;
  uint8_t* h_done = nullptr;
  float *d_action = nullptr, *d_obs = nullptr, *d_reward = nullptr, *d_terms = nullptr;
  uint8_t* d_done = nullptr;
  float *h_pull = nullptr, *h_pull_dev = nullptr;   // mapped staging of ilrl_step_pull / ilrl_pull (+ its action tail)
  int substeps = ILRL_SUBSTEPS;  // harness only (ilrl_debug_substeps)
  int* perm = nullptr;           // [n] env ids grouped by cost (K8); identity until the first sort
  uint8_t* cost = nullptr;       // [n] constraint rows of each env's last substep
  int group_every = 2;           // sort every this many whole-batch steps (ILRL_GROUP_EVERY; 0 = never group; measured 1/2/4/8)
  long long group_ctr = 0;
  int cost_chunk = 0;
  int layout = 0;                // shared-memory layout of the step kernel: 0 LayoutSmall, 1 LayoutLarge, 2 LayoutDense4
                                 // (chosen at create time from N)
  bool no_zero_copy = false;     // harness only (ilrl_debug_zero_copy): force the explicit-copy host path
  // ilrl_serve_*: the resident serving kernel and its doorbell
  bool serving = false;
  int serve_nparts = 0;
  ServeCtl* serve_ctl = nullptr;          // mapped host memory: one 256-byte slot per part (doorbell at +0, ack at +128)
  ServeDev* serve_dev = nullptr;          // [ILRL_MAX_PARTS]
  int serve_seq[8] = {0};
  int serve_posted[8] = {0};
  bool serve_live[8] = {false};           // the part has a resident kernel (non-empty part)
  int64_t launches = 0;
  bool timing = false;           // ilrl_kernel_timing: the step kernels stamp %globaltimer into ktime[slot]
  unsigned long long* ktime = nullptr;   // [KT_SLOTS][2] device
  int kt_used = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
};

static thread_local std::string g_create_err;

static int fail(ilrl_env* e, int code, const std::string& msg) {
  if (e) e->err = msg; else g_create_err = msg;
  return code;
}
#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t _r = (call);                                                                           \
    if (_r != cudaSuccess) return fail(env, ILRL_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_r)); \
  } while (0)

static inline int nblk(int n) { return (n + BLOCK - 1) / BLOCK; }

// Every entry point runs on the handle's device and leaves the caller's current device as it found it (a process
// that drives several GPUs, or torch with another current device, is not disturbed).
struct DeviceGuard {
  int prev = -1;
  cudaError_t status;
  explicit DeviceGuard(int dev) {
    status = cudaGetDevice(&prev);
    if (status == cudaSuccess && prev != dev) status = cudaSetDevice(dev); else if (status == cudaSuccess) prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_DEVICE_RAW(env) DeviceGuard _dg((env)->cfg.device); CK(_dg.status)
// (between ilrl_serve_begin and ilrl_serve_end the handle belongs to its resident kernel: only ilrl_serve_* may be called)
#define ON_DEVICE(env)                                                                                              \
  if ((env)->serving) return fail(env, ILRL_ERR_STATE, "the handle is serving (ilrl_serve_begin): call ilrl_serve_end first"); \
  ON_DEVICE_RAW(env)

// step-kernel launch, optionally with the programmatic-stream-serialization attribute (see the kernel's prologue).
// Only where it pays: with the 3-CTA-per-SM layout (and the terrain instantiation, which uses it) the early-resident CTAs
// of the next step crowd onto the SMs that finished first - three CTAs on some SMs, none on others - and the step gets
// SLOWER (6000 / 7000 envs -4 / -3 %, terrain 4096 envs -9 %); self-collision: no difference (tools/pdl_ab.py).
static bool g_pdl = [] { const char* e = getenv("ILRL_PDL"); return !(e && e[0] == '0'); }();
template <class K>
static void launch_step(K kernel, int grid, size_t smem, cudaStream_t st, const StepArgs& a, bool pdl = false) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(QT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (g_pdl && pdl) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, a);
}

extern "C" {

int ilrl_serve_end(ilrl_env* env);

const char* ilrl_last_error(const ilrl_env* env) { return env ? env->err.c_str() : g_create_err.c_str(); }

int ilrl_create(const ilrl_config* cfg, ilrl_env** out) {
  ilrl_env* env = nullptr;
  if (!cfg || !out) return fail(nullptr, ILRL_ERR_ARG, "ilrl_create: null argument");
  if (cfg->num_envs <= 0 || cfg->mode < 0 || cfg->mode > 2) return fail(nullptr, ILRL_ERR_ARG, "ilrl_create: bad num_envs/mode");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(nullptr, ILRL_ERR_CUDA, "ilrl_create: no CUDA device (this library has no CPU path)");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, ILRL_ERR_ARG, "ilrl_create: bad device ordinal");
  DeviceGuard dg(cfg->device);
  CK(dg.status);
  env = new (std::nothrow) ilrl_env();
  if (!env) return fail(nullptr, ILRL_ERR_ARG, "ilrl_create: out of host memory");
  env->cfg = *cfg;
  const bool m2 = cfg->mode == 2;   // hier_env_2.py:58-63, 160: step_per_level 20, skipFrame 5
  if (env->cfg.skip_frame <= 0) env->cfg.skip_frame = m2 ? 5 : 2;
  if (env->cfg.max_timestep <= 0) env->cfg.max_timestep = 3000;
  if (env->cfg.step_per_level <= 0) env->cfg.step_per_level = m2 ? 20 : 5;
  if (m2) { env->obs_w = ILRL_OBS_LOW2; env->hobs_w = ILRL_OBS_HIGH2; env->hact_w = ILRL_ACT_HIGH2; }
  const int n = env->n = cfg->num_envs;
  memset(env->clips, 0, sizeof env->clips);
#define CKC(call)                                                                              \
  do {                                                                                         \
    cudaError_t _r = (call);                                                                   \
    if (_r != cudaSuccess) {                                                                   \
      g_create_err = std::string(#call) + ": " + cudaGetErrorString(_r);                       \
      ilrl_destroy(env);                                                                       \
      return ILRL_ERR_CUDA;                                                                    \
    }                                                                                          \
  } while (0)
  CKC(cudaMalloc(&env->phys, sizeof(float) * ILRL_PHYS_STRIDE * n));
  CKC(cudaMalloc(&env->envf, sizeof(float) * ILRL_ENV_STRIDE * n));
  CKC(cudaMalloc(&env->rng, sizeof(uint32_t) * n));
  CKC(cudaMalloc(&env->gscr, sizeof(float) * (size_t)chain::GROWS * chain::RW * n));
  CKC(cudaMalloc(&env->perm, sizeof(int) * n));
  env->cost_chunk = (((n + 31) / 32) + 511) / 512 * 512;   // bytes per sorting warp; entries past n stay 255
  CKC(cudaMalloc(&env->cost, 32 * (size_t)env->cost_chunk));
  CKC(cudaMemset(env->cost, 255, 32 * (size_t)env->cost_chunk));
  CKC(cudaMemset(env->cost, 0, n));
  iota_kernel<<<nblk(n), BLOCK>>>(env->perm, n);
  if (const char* o = getenv("ILRL_GROUP_EVERY")) env->group_every = atoi(o);
  CKC(cudaFuncSetAttribute(step_kernel<0, SmemSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSmall)));
  CKC(cudaFuncSetAttribute(step_kernel<1, SmemSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSmall)));
  CKC(cudaFuncSetAttribute(step_seq_kernel<SmemSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSmall)));
  CKC(cudaFuncSetAttribute(step_seq_kernel<SmemLarge>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLarge)));
  CKC(cudaFuncSetAttribute(step_seq_kernel<SmemDense4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemDense4)));
  CKC(cudaFuncSetAttribute(step_kernel<0, SmemLarge, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLarge)));
  CKC(cudaFuncSetAttribute(step_kernel<0, SmemSmall, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSmall)));
  CKC(cudaFuncSetAttribute(physics_only_kernel<SmemLarge, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLarge)));
  CKC(cudaFuncSetAttribute(step_kernel<0, SmemSelf, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSelf)));
  CKC(cudaFuncSetAttribute(step_kernel<0, SmemSelf, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSelf)));
  CKC(cudaFuncSetAttribute(step_kernel<1, SmemSelf, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSelf)));
  CKC(cudaFuncSetAttribute(step_kernel<2, SmemSelf, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSelf)));
  CKC(cudaFuncSetAttribute(physics_only_kernel<SmemSelf, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSelf)));
  CKC(cudaFuncSetAttribute(physics_only_kernel<SmemSelf, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSelf)));
  CKC(cudaFuncSetAttribute(step_kernel<2, SmemSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSmall)));
  CKC(cudaFuncSetAttribute(step_kernel<2, SmemLarge>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLarge)));
  CKC(cudaFuncSetAttribute(step_kernel<2, SmemDense4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemDense4)));
  CKC(cudaFuncSetAttribute(physics_only_kernel<SmemSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSmall)));
  CKC(cudaFuncSetAttribute(step_kernel<0, SmemLarge>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLarge)));
  CKC(cudaFuncSetAttribute(step_kernel<1, SmemLarge>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLarge)));
  CKC(cudaFuncSetAttribute(physics_only_kernel<SmemLarge>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemLarge)));
  CKC(cudaFuncSetAttribute(step_kernel<0, SmemDense4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemDense4)));
  CKC(cudaFuncSetAttribute(step_kernel<1, SmemDense4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemDense4)));
  CKC(cudaFuncSetAttribute(physics_only_kernel<SmemDense4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemDense4)));
  {
    // Layout.  The step kernel is persistent: resident CTAs (SMs x CTAs per SM) pull 16-env tiles,
    // so a step takes x = tiles / resident CTAs "rounds", the last one partly filled.  Being latency-bound, a round costs
    // about the same however full it is; what differs between the layouts is how many envs a round holds (4736 / 7104 /
    // 9472 on 148 SMs) and how long it takes.  Measured on B200 (profiles/r2_v6_layout_sweep.txt), time of a step ~
    // floor(x) R + a + b frac(x) for x < 8 and x Rinf beyond (us, least squares): on-chip R 53.6 a 54.8 b 23.3 Rinf 57.5;
    // dense R 69.8 a 57.8 b 44.2 Rinf 73.2; dense4 R 76.8 a 67.2 b 48.1 Rinf 83.4.  Estimate all three, take the smallest.
    // Measured (M env-steps/s: on chip / dense / dense4; * = chosen by the estimate):
    //    2048: 28.8* / 25.5 / 22.2      4096: 56.3* / 49.6 / 42.9     5120: 48.4 / 56.4* / 51.6    6144: 55.0 / 64.6* / 61.4
    //    8192: 67.7 / 66.0 / 79.2*     10240: 63.1 / 74.4* / 74.5    12288: 71.4 / 79.1 / 82.2*   16384: 72.8 / 82.2 / 96.8*
    //   20480: 73.7 / 87.2 / 92.2*     24576: 74.9 / 85.8 / 101.9*   32768: 81.6 / 88.3 / 103.7*  65536: 82.8 / 96.0 / 112.5*
    //  131072: 83.3 / 98.2 / 113.6*
    // ILRL_LAYOUT=small|large|dense4 overrides (measurement aid).
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, cfg->device));
    int occ[3] = {0, 0, 0};
    if (cfg->mode == 0) {
      CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], step_kernel<0, SmemSmall>, QT, sizeof(SmemSmall)));
      CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], step_kernel<0, SmemLarge>, QT, sizeof(SmemLarge)));
      CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], step_kernel<0, SmemDense4>, QT, sizeof(SmemDense4)));
    } else if (cfg->mode == 1) {
      CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], step_kernel<1, SmemSmall>, QT, sizeof(SmemSmall)));
      CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], step_kernel<1, SmemLarge>, QT, sizeof(SmemLarge)));
      CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], step_kernel<1, SmemDense4>, QT, sizeof(SmemDense4)));
    } else {
      CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], step_kernel<2, SmemSmall>, QT, sizeof(SmemSmall)));
      CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], step_kernel<2, SmemLarge>, QT, sizeof(SmemLarge)));
      CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], step_kernel<2, SmemDense4>, QT, sizeof(SmemDense4)));
    }
    if (occ[0] < 1 || occ[1] < 1 || occ[2] < 1) { g_create_err = "step kernel does not fit on this device"; ilrl_destroy(env); return ILRL_ERR_CUDA; }
    env->grid_small = occ[0] * prop.multiProcessorCount;
    env->grid_large = occ[1] * prop.multiProcessorCount;
    env->grid_dense4 = occ[2] * prop.multiProcessorCount;
    if (const char* o = getenv("ILRL_LAYOUT")) {
      env->layout = (o[0] == 'l' || o[0] == 'L') ? 1 : (o[0] == 'd' || o[0] == 'D') ? 2 : 0;
    } else {
      const float tiles = (float)((n + QE - 1) / QE);
      const float R[3] = {53.6f, 69.8f, 76.8f}, A[3] = {54.8f, 57.8f, 67.2f}, B[3] = {23.3f, 44.2f, 48.1f},
                  Rinf[3] = {57.5f, 73.2f, 83.4f};
      const int grid[3] = {env->grid_small, env->grid_large, env->grid_dense4};
      float best = 0.f;
      for (int l = 0; l < 3; l++) {
        const float x = tiles / (float)grid[l], full = ceilf(x) - 1.f, frac = x - full;   // last round: 0 < frac <= 1
        // (from 8 rounds on the CTAs have drifted apart and the tile queue keeps every SM busy: time ~ x)
        const float est = x >= 8.f ? x * Rinf[l] : full * R[l] + A[l] + B[l] * frac;
        if (l == 0 || est < best) { best = est; env->layout = l; }
      }
    }
  }
  CKC(cudaMalloc(&env->high_obs, sizeof(float) * env->hobs_w * n));
  if (m2) {
    CKC(cudaMalloc(&env->jt, sizeof(float) * ILRL_JT_STRIDE * n));
    CKC(cudaMemset(env->jt, 0, sizeof(float) * ILRL_JT_STRIDE * n));   // self.jointTarget = [0] * 16 (hier_env_2.py:172)
  }
  CKC(cudaMalloc(&env->high_reward, sizeof(float) * n));
  CKC(cudaMalloc(&env->high_flags, n));
  CKC(cudaMalloc(&env->stats, sizeof(double) * ILRL_STATS_WORDS));
  CKC(cudaMalloc(&env->tile_counter, 2 * (1 + ILRL_MAX_PARTS) * sizeof(unsigned int)));
  CKC(cudaMemset(env->tile_counter, 0, 2 * (1 + ILRL_MAX_PARTS) * sizeof(unsigned int)));
  CKC(cudaMemset(env->phys, 0, sizeof(float) * ILRL_PHYS_STRIDE * n));
  CKC(cudaMemset(env->envf, 0, sizeof(float) * ILRL_ENV_STRIDE * n));
  CKC(cudaMemset(env->rng, 0, sizeof(uint32_t) * n));
  CKC(cudaMemset(env->high_obs, 0, sizeof(float) * env->hobs_w * n));
  CKC(cudaMemset(env->high_reward, 0, sizeof(float) * n));
  CKC(cudaMemset(env->high_flags, 0, n));
  CKC(cudaMemset(env->stats, 0, sizeof(double) * ILRL_STATS_WORDS));
#ifdef ILRL_PROF
  CKC(cudaMalloc(&env->prof, sizeof(long long) * chain::PF_WORDS * ((n + 7) / 8 + 2)));
  CKC(cudaMemset(env->prof, 0, sizeof(long long) * chain::PF_WORDS * ((n + 7) / 8 + 2)));
#endif
  CKC(cudaEventCreate(&env->ev0));
  CKC(cudaEventCreate(&env->ev1));
  CKC(cudaDeviceSynchronize());
  *out = env;
  return ILRL_OK;
}

void ilrl_destroy(ilrl_env* env) {
  if (!env) return;
  if (env->serving) ilrl_serve_end(env);
  DeviceGuard dg(env->cfg.device);
  cudaDeviceSynchronize();
  for (int p = 0; p < ILRL_MAX_PARTS; p++) {
    if (env->part_stream[p]) cudaStreamDestroy(env->part_stream[p]);
  }
  cudaFree(env->clip_ids_dev);
  cudaFree(env->terr_mem);
  cudaFreeHost(env->serve_ctl);
  cudaFree(env->serve_dev);
  cudaFreeHost(env->h_pull);
  cudaFree(env->ktime);
  cudaFree(env->phys); cudaFree(env->envf); cudaFree(env->rng); cudaFree(env->gscr); cudaFree(env->jt); cudaFree(env->perm); cudaFree(env->cost);
  cudaFree(env->high_obs); cudaFree(env->high_reward); cudaFree(env->high_flags); cudaFree(env->stats); cudaFree(env->tile_counter);
  for (int c = 0; c < MAX_CLIPS; c++) cudaFree(env->clip_mem[c]);
  cudaFreeHost(env->h_action); cudaFreeHost(env->h_obs); cudaFreeHost(env->h_reward); cudaFreeHost(env->h_terms);
  cudaFreeHost(env->h_done);
  cudaFree(env->d_action); cudaFree(env->d_obs); cudaFree(env->d_reward); cudaFree(env->d_terms); cudaFree(env->d_done);
  if (env->ev0) cudaEventDestroy(env->ev0);
  if (env->ev1) cudaEventDestroy(env->ev1);
  delete env;
}

int ilrl_load_clip(ilrl_env* env, int32_t clip, const float* pos, int32_t n_pos, const float* rel, int32_t n_rel,
                   const float* vel, int32_t n_vel, const float* ep, int32_t n_ep, int32_t max_frame) {
  if (!env) return ILRL_ERR_ARG;
  if (clip < 0 || clip >= MAX_CLIPS || !pos || !rel || !vel || !ep) return fail(env, ILRL_ERR_ARG, "ilrl_load_clip: bad argument");
  // every frame the step/reset path can index must exist in all four tables
  if (max_frame < 7 || max_frame > n_pos - 1 || max_frame > n_rel - 1 || max_frame > n_vel || max_frame > n_ep - 1)
    return fail(env, ILRL_ERR_ARG, "ilrl_load_clip: max_frame exceeds a table (clamp it; see DESIGN.md on motion13_13)");
  ON_DEVICE(env);
  size_t words = (size_t)14 * (n_pos + n_rel + n_vel) + (size_t)27 * n_ep;
  cudaFree(env->clip_mem[clip]);
  env->clip_mem[clip] = nullptr;
  CK(cudaMalloc(&env->clip_mem[clip], words * sizeof(float)));
  float* d = env->clip_mem[clip];
  ClipDesc& c = env->clips[clip];
  c.pos = d; CK(cudaMemcpy(d, pos, sizeof(float) * 14 * n_pos, cudaMemcpyHostToDevice)); d += 14 * n_pos;
  c.rel = d; CK(cudaMemcpy(d, rel, sizeof(float) * 14 * n_rel, cudaMemcpyHostToDevice)); d += 14 * n_rel;
  c.vel = d; CK(cudaMemcpy(d, vel, sizeof(float) * 14 * n_vel, cudaMemcpyHostToDevice)); d += 14 * n_vel;
  c.ep = d;  CK(cudaMemcpy(d, ep, sizeof(float) * 27 * n_ep, cudaMemcpyHostToDevice));
  c.n_pos = n_pos; c.n_vel = n_vel; c.max_frame = max_frame; c.pad = 0;
  env->clip_loaded[clip] = true;
  return ILRL_OK;
}

static StateView view(ilrl_env* env) { StateView v; v.n = env->n; v.phys = env->phys; v.envf = env->envf; v.rng = env->rng; return v; }

int ilrl_set_clip_ids(ilrl_env* env, const int32_t* ids_host) {
  if (!env) return ILRL_ERR_ARG;
  ON_DEVICE(env);
  int32_t* d = nullptr;
  if (ids_host) {
    for (int i = 0; i < env->n; i++)
      if (ids_host[i] < 0 || ids_host[i] >= MAX_CLIPS || !env->clip_loaded[ids_host[i]])
        return fail(env, ILRL_ERR_STATE, "ilrl_set_clip_ids: env refers to a clip that is not loaded");
    if (!env->clip_ids_dev) CK(cudaMalloc(&env->clip_ids_dev, sizeof(int32_t) * env->n));
    d = env->clip_ids_dev;
    CK(cudaMemcpy(d, ids_host, sizeof(int32_t) * env->n, cudaMemcpyHostToDevice));
  } else if (!env->clip_loaded[0]) {
    return fail(env, ILRL_ERR_STATE, "ilrl_set_clip_ids: clip 0 is not loaded");
  }
  clip_ids_kernel<<<(env->n + 255) / 256, 256>>>(view(env), d);
  env->launches++;
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  return ILRL_OK;
}

static int check_ready(ilrl_env* env) {
  if (!env->clip_loaded[0]) return fail(env, ILRL_ERR_STATE, "no motion clip loaded (ilrl_load_clip / ilrl_set_clip_ids first)");
  return ILRL_OK;
}

int ilrl_reset(ilrl_env* env, const uint8_t* mask, const int32_t* start_frame, const int32_t* target_deg,
               const float* yaw_deg, const float* target_xy, float* obs, void* stream) {
  if (!env) return ILRL_ERR_ARG;
  if (int r = check_ready(env)) return r;
  ON_DEVICE(env);
  ResetArgs a;
  a.n = env->n; a.skip_frame = env->cfg.skip_frame; a.id_base = env->cfg.env_id_base; a.step_per_level = (float)env->cfg.step_per_level; a.seed = env->cfg.seed;
  a.phys = env->phys; a.envf = env->envf; a.rng = env->rng;
  a.mask = mask; a.start_frame = start_frame; a.target_deg = target_deg; a.yaw_deg = yaw_deg; a.target_xy = target_xy; a.obs = obs;
  a.noise = env->forced_noise;
  a.high_flags = env->high_flags;
  memcpy(a.clips, env->clips, sizeof a.clips);
  cudaStream_t st = (cudaStream_t)stream;
  if (env->cfg.mode == 0) reset_kernel<0><<<nblk(env->n), BLOCK, 0, st>>>(a);
  else if (env->cfg.mode == 1) reset_kernel<1><<<nblk(env->n), BLOCK, 0, st>>>(a);
  else reset_kernel<2><<<nblk(env->n), BLOCK, 0, st>>>(a);
  env->launches++;
  CK(cudaGetLastError());
  return ILRL_OK;
}

static void fill_step_args(ilrl_env* env, StepArgs& a, const float* action, float* obs, float* reward, uint8_t* done,
                           float* terms, int skip_physics, int first, int count, int part, int forced_scalar) {
  a.n = env->n; a.first = first; a.end = first + count; a.skip_frame = env->cfg.skip_frame;
  a.id_base = env->cfg.env_id_base; a.skip_physics = skip_physics; a.auto_reset = env->cfg.auto_reset; a.max_timestep = env->cfg.max_timestep;
  a.step_per_level = (float)env->cfg.step_per_level; a.seed = env->cfg.seed;
  a.phys = env->phys; a.envf = env->envf; a.rng = env->rng;
  a.action = action; a.obs = obs; a.reward = reward; a.done = done; a.terms = terms;
  a.high_obs = env->high_obs; a.high_reward = env->high_reward; a.high_flags = env->high_flags;
  a.forced_deg = env->forced_deg; a.forced_scalar = forced_scalar; a.stats = env->stats; a.gscr = env->gscr;
  a.jt = env->jt; a.forced_noise = env->forced_noise; a.terr = env->terr; a.selfc = env->selfc;
  a.tile_counter = env->tile_counter + 2 * (part + 1);
#ifdef ILRL_PROF
  a.prof = env->prof;
#endif
  memcpy(a.clips, env->clips, sizeof a.clips);
  a.ktime = (env->timing && env->kt_used < KT_SLOTS) ? env->ktime + 2 * (size_t)env->kt_used++ : nullptr;
  a.ntiles = (count + QE - 1) / QE;
}

// part < 0: the whole batch; otherwise envs [first, first + count) on the part's own tile counters
static int do_step(ilrl_env* env, const float* action, float* obs, float* reward, uint8_t* done, float* terms,
                   cudaStream_t st, int skip_physics, int first = 0, int count = -1, int part = -1, int forced_scalar = INT_MIN) {
  if (!env) return ILRL_ERR_ARG;
  if (!action || !obs || !reward || !done) return fail(env, ILRL_ERR_ARG, "ilrl_step: null buffer");
  if (int r = check_ready(env)) return r;
  ON_DEVICE(env);
  if (count < 0) count = env->n;
  StepArgs a;
  fill_step_args(env, a, action, obs, reward, done, terms, skip_physics, first, count, part, forced_scalar);
  // (heightfield terrain / self-collision: their own instantiations - self-collision in the mid-size layout whatever
  // the batch size, the terrain alone on chip when the batch is small enough for that layout and mid-size otherwise)
  const bool special = env->terr.h || env->self_on;
  const int layout = !special ? env->layout : (!env->self_on && env->layout == 0) ? 0 : 1;
  const int qblk = min(a.ntiles, layout == 2 ? env->grid_dense4 : layout == 1 ? env->grid_large : env->grid_small);
  const int md = env->cfg.mode;
  // K8: whole-batch steps of more than one wave run on envs grouped by cost
  if (part < 0 && first == 0 && count == env->n && env->group_every > 0 && a.ntiles > qblk) {
    a.perm = env->perm; a.cost = env->cost;
    if (env->group_ctr++ % env->group_every == 0) {
      cost_sort_kernel<<<1, 1024, 0, st>>>(env->cost, env->perm, env->cost_chunk);
      env->launches++;
    }
  }
  if (special) {
    if (env->terr.h && env->self_on) launch_step(step_kernel<0, SmemSelf, true, true>, qblk, sizeof(SmemSelf), st, a);
    else if (env->self_on && md == 0) launch_step(step_kernel<0, SmemSelf, false, true>, qblk, sizeof(SmemSelf), st, a);
    else if (env->self_on && md == 1) launch_step(step_kernel<1, SmemSelf, false, true>, qblk, sizeof(SmemSelf), st, a);
    else if (env->self_on) launch_step(step_kernel<2, SmemSelf, false, true>, qblk, sizeof(SmemSelf), st, a);
    else if (layout == 0) launch_step(step_kernel<0, SmemSmall, true>, qblk, sizeof(SmemSmall), st, a, true);
    else launch_step(step_kernel<0, SmemLarge, true>, qblk, sizeof(SmemLarge), st, a);
  } else if (env->layout == 2) {
    if (md == 0) launch_step(step_kernel<0, SmemDense4>, qblk, sizeof(SmemDense4), st, a, true);
    else if (md == 1) launch_step(step_kernel<1, SmemDense4>, qblk, sizeof(SmemDense4), st, a, true);
    else launch_step(step_kernel<2, SmemDense4>, qblk, sizeof(SmemDense4), st, a, true);
  } else if (env->layout == 1) {
    if (md == 0) launch_step(step_kernel<0, SmemLarge>, qblk, sizeof(SmemLarge), st, a);
    else if (md == 1) launch_step(step_kernel<1, SmemLarge>, qblk, sizeof(SmemLarge), st, a);
    else launch_step(step_kernel<2, SmemLarge>, qblk, sizeof(SmemLarge), st, a);
  } else {
    if (md == 0) launch_step(step_kernel<0, SmemSmall>, qblk, sizeof(SmemSmall), st, a, true);
    else if (md == 1) launch_step(step_kernel<1, SmemSmall>, qblk, sizeof(SmemSmall), st, a, true);
    else launch_step(step_kernel<2, SmemSmall>, qblk, sizeof(SmemSmall), st, a, true);
  }
  env->launches++;
  CK(cudaGetLastError());
  return ILRL_OK;
}

int ilrl_step(ilrl_env* env, const float* action, float* obs, float* reward, uint8_t* done, float* terms, void* stream) {
  return do_step(env, action, obs, reward, done, terms, (cudaStream_t)stream, 0);
}

int ilrl_get_grouping(ilrl_env* env, uint8_t* cost_dev, int32_t* perm_dev, void* stream) {
  if (!env) return ILRL_ERR_ARG;
  ON_DEVICE(env);
  if (cost_dev) CK(cudaMemcpyAsync(cost_dev, env->cost, env->n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  if (perm_dev) CK(cudaMemcpyAsync(perm_dev, env->perm, sizeof(int) * env->n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return ILRL_OK;
}

int ilrl_step_sequence(ilrl_env* env, int32_t ksteps, const float* action, float* obs, float* reward, uint8_t* done,
                       float* terms, void* stream) {
  if (!env) return ILRL_ERR_ARG;
  if (!action || !obs || !reward || !done || ksteps < 1) return fail(env, ILRL_ERR_ARG, "ilrl_step_sequence: null buffer or ksteps < 1");
  if (env->cfg.mode != 0) return fail(env, ILRL_ERR_ARG, "ilrl_step_sequence: low-level env only (a hier env needs its high-level action between steps)");
  if (env->terr.h || env->self_on) return fail(env, ILRL_ERR_STATE, "ilrl_step_sequence: not available with a heightfield or self-collision");
  if (int r = check_ready(env)) return r;
  ON_DEVICE(env);
  cudaStream_t st = (cudaStream_t)stream;
  StepArgs a;
  fill_step_args(env, a, action, obs, reward, done, terms, 0, 0, env->n, -1, INT_MIN);
  const int qblk = min(a.ntiles, env->layout == 2 ? env->grid_dense4 : env->layout == 1 ? env->grid_large : env->grid_small);
  if (env->layout == 2) step_seq_kernel<SmemDense4><<<qblk, QT, sizeof(SmemDense4), st>>>(a, ksteps);
  else if (env->layout == 1) step_seq_kernel<SmemLarge><<<qblk, QT, sizeof(SmemLarge), st>>>(a, ksteps);
  else step_seq_kernel<SmemSmall><<<qblk, QT, sizeof(SmemSmall), st>>>(a, ksteps);
  env->launches++;
  CK(cudaGetLastError());
  return ILRL_OK;
}
int ilrl_step_no_physics(ilrl_env* env, const float* action, float* obs, float* reward, uint8_t* done, float* terms,
                         void* stream) {
  return do_step(env, action, obs, reward, done, terms, (cudaStream_t)stream, 1);
}

// device alias of a page-locked, mapped host buffer (null if it is pageable / not mappable).  The aliases are cached
// per handle (a rollout loop presents the same few buffers again and again; the two runtime queries cost ~2 us per
// buffer, i.e. more than the launch itself): a buffer must stay page-locked for as long as it is used with the handle.
static void* host_alias(ilrl_env* env, const void* p) {
  auto it = env->alias.find(p);
  if (it != env->alias.end()) return it->second;
  void* d = nullptr;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess || at.type != cudaMemoryTypeHost ||
      cudaHostGetDevicePointer(&d, const_cast<void*>(p), 0) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;   // (not cached: the caller may pin the buffer later)
  }
  if (env->alias.size() > 4096) env->alias.clear();
  env->alias[p] = d;
  return d;
}
static bool map_host_buffers(ilrl_env* env, const float* action_h, float* obs_h, float* reward_h, uint8_t* done_h, float* terms_h,
                             void** da, void** dobs, void** dr, void** dd, void** dt) {
  *dt = nullptr;
  return (*da = host_alias(env, action_h)) && (*dobs = host_alias(env, obs_h)) && (*dr = host_alias(env, reward_h)) &&
         (*dd = host_alias(env, done_h)) && (!terms_h || (*dt = host_alias(env, terms_h)));
}

static int ensure_io_buffers(ilrl_env* env);
int ilrl_step_host(ilrl_env* env, const float* action_h, float* obs_h, float* reward_h, uint8_t* done_h, float* terms_h,
                   void* stream) {
  if (!env) return ILRL_ERR_ARG;
  if (!action_h || !obs_h || !reward_h || !done_h) return fail(env, ILRL_ERR_ARG, "ilrl_step_host: null buffer");
  ON_DEVICE(env);
  const size_t n = env->n;
  cudaStream_t st = (cudaStream_t)stream;
  if (!env->no_zero_copy) {
    // Zero-copy: the step kernel reads the action tiles from, and writes obs / reward / done to, the caller's
    // page-locked buffers through their device mappings.  One launch + one synchronise; the output writes of CTAs
    // that finish early overlap the tail of the kernel instead of waiting for a separate D2H copy.
    void *da, *dobs, *dr, *dd, *dt;
    if (map_host_buffers(env, action_h, obs_h, reward_h, done_h, terms_h, &da, &dobs, &dr, &dd, &dt)) {
      int r = do_step(env, (const float*)da, (float*)dobs, (float*)dr, (uint8_t*)dd, (float*)dt, st, 0);
      if (r) return r;
      CK(cudaStreamSynchronize(st));
      return ILRL_OK;
    }
  }
  if (!env->h_action) {  // staging buffers: pinned host mirrors + device I/O, allocated on first use
    CK(cudaMallocHost(&env->h_action, sizeof(float) * 17 * n));
    CK(cudaMallocHost(&env->h_obs, sizeof(float) * env->obs_w * n));
    CK(cudaMallocHost(&env->h_reward, sizeof(float) * n));
    CK(cudaMallocHost(&env->h_terms, sizeof(float) * ILRL_TERM_WORDS * n));
    CK(cudaMallocHost(&env->h_done, n));
    if (int r = ensure_io_buffers(env)) return r;
  }
  // Page-locked caller buffers are DMA targets as they are; pageable ones go through the handle's pinned mirrors.
  auto pinned = [](const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
  };
  const bool pa = pinned(action_h), po = pinned(obs_h), pr = pinned(reward_h), pd = pinned(done_h),
             pt = terms_h && pinned(terms_h);
  if (!pa) memcpy(env->h_action, action_h, sizeof(float) * 17 * n);
  CK(cudaMemcpyAsync(env->d_action, pa ? action_h : env->h_action, sizeof(float) * 17 * n, cudaMemcpyHostToDevice, st));
  int r = do_step(env, env->d_action, env->d_obs, env->d_reward, env->d_done, terms_h ? env->d_terms : nullptr, st, 0);
  if (r) return r;
  CK(cudaMemcpyAsync(po ? obs_h : env->h_obs, env->d_obs, sizeof(float) * env->obs_w * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(pr ? reward_h : env->h_reward, env->d_reward, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(pd ? done_h : env->h_done, env->d_done, n, cudaMemcpyDeviceToHost, st));
  if (terms_h)
    CK(cudaMemcpyAsync(pt ? terms_h : env->h_terms, env->d_terms, sizeof(float) * ILRL_TERM_WORDS * n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (!po) memcpy(obs_h, env->h_obs, sizeof(float) * env->obs_w * n);
  if (!pr) memcpy(reward_h, env->h_reward, sizeof(float) * n);
  if (!pd) memcpy(done_h, env->h_done, n);
  if (terms_h && !pt) memcpy(terms_h, env->h_terms, sizeof(float) * ILRL_TERM_WORDS * n);
  return ILRL_OK;
}

// (A completion word in mapped host memory, set by the last CTA after a system-scope fence and polled by the host, was
// measured against the driver's stream synchronisation: no faster - 94 against 91 us per step - and it costs every CTA
// a fence and an atomic.  Not kept.)
static int wait_part(ilrl_env* env, int part) {
  ON_DEVICE(env);
  CK(cudaStreamSynchronize(env->part_stream[part]));
  env->part_busy[part] = false;
  return ILRL_OK;
}

// envs of part p of nparts: contiguous blocks of ceil(N / nparts) rounded up to whole 16-env tiles
static void part_range(const ilrl_env* env, int part, int nparts, int* first, int* count) {
  int per = (env->n + nparts - 1) / nparts;
  per = (per + QE - 1) / QE * QE;
  *first = min(part * per, env->n);
  *count = min(per, env->n - *first);
}

static int step_host_async(ilrl_env* env, int32_t part, int32_t nparts, const float* action_h, float* obs_h, float* reward_h,
                           uint8_t* done_h, float* terms_h, bool wait_first) {
  if (!env) return ILRL_ERR_ARG;
  if (nparts < 1 || nparts > ILRL_MAX_PARTS || part < 0 || part >= nparts)
    return fail(env, ILRL_ERR_ARG, "ilrl_step_host_async: part / nparts out of range (nparts <= 8)");
  if (!action_h || !obs_h || !reward_h || !done_h) return fail(env, ILRL_ERR_ARG, "ilrl_step_host_async: null buffer");
  ON_DEVICE(env);
  if (env->part_busy[part]) {
    if (!wait_first) return fail(env, ILRL_ERR_STATE, "ilrl_step_host_async: part is still in flight (ilrl_wait first)");
    if (int r = wait_part(env, part)) return r;
  }
  void *da, *dobs, *dr, *dd, *dt;
  if (!map_host_buffers(env, action_h, obs_h, reward_h, done_h, terms_h, &da, &dobs, &dr, &dd, &dt))
    return fail(env, ILRL_ERR_ARG, "ilrl_step_host_async: buffers must be page-locked and mapped (cudaHostAlloc / "
                                   "cudaHostRegister / torch pin_memory)");
  if (!env->part_stream[part]) CK(cudaStreamCreateWithFlags(&env->part_stream[part], cudaStreamNonBlocking));
  int first, count;
  part_range(env, part, nparts, &first, &count);
  if (count > 0) {
    int r = do_step(env, (const float*)da, (float*)dobs, (float*)dr, (uint8_t*)dd, (float*)dt, env->part_stream[part], 0,
                    first, count, part);
    if (r) return r;
  }
  env->part_busy[part] = true;
  return ILRL_OK;
}
int ilrl_step_host_async(ilrl_env* env, int32_t part, int32_t nparts, const float* action_h, float* obs_h, float* reward_h,
                         uint8_t* done_h, float* terms_h) {
  return step_host_async(env, part, nparts, action_h, obs_h, reward_h, done_h, terms_h, false);
}
int ilrl_wait_step_host_async(ilrl_env* env, int32_t part, int32_t nparts, const float* action_h, float* obs_h,
                              float* reward_h, uint8_t* done_h, float* terms_h) {
  return step_host_async(env, part, nparts, action_h, obs_h, reward_h, done_h, terms_h, true);
}

int ilrl_wait(ilrl_env* env, int32_t part) {
  if (!env) return ILRL_ERR_ARG;
  if (part < 0 || part >= ILRL_MAX_PARTS) return fail(env, ILRL_ERR_ARG, "ilrl_wait: part out of range");
  if (!env->part_busy[part]) return ILRL_OK;
  return wait_part(env, part);
}

static int ensure_io_buffers(ilrl_env* env) {
  const size_t n = env->n;
  if (!env->d_action) {
    CK(cudaMalloc(&env->d_action, sizeof(float) * 17 * n));
    CK(cudaMalloc(&env->d_obs, sizeof(float) * env->obs_w * n));
    CK(cudaMalloc(&env->d_reward, sizeof(float) * n));
    CK(cudaMalloc(&env->d_terms, sizeof(float) * ILRL_TERM_WORDS * n));
    CK(cudaMalloc(&env->d_done, n));
    CK(cudaMemset(env->d_obs, 0, sizeof(float) * env->obs_w * n));
    CK(cudaMemset(env->d_reward, 0, sizeof(float) * n));
    CK(cudaMemset(env->d_terms, 0, sizeof(float) * ILRL_TERM_WORDS * n));
    CK(cudaMemset(env->d_done, 0, n));
  }
  if (!env->h_pull) {
    CK(cudaHostAlloc(&env->h_pull, sizeof(float) * ILRL_PULL_WORDS * n + sizeof(float) * 17 * n, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void**)&env->h_pull_dev, env->h_pull, 0));
  }
  return ILRL_OK;
}
static int do_pull(ilrl_env* env, float* pull_host, const float* obs_dev, cudaStream_t st) {
  PullArgs a;
  a.n = env->n; a.obs = obs_dev ? obs_dev : env->d_obs; a.reward = env->d_reward; a.terms = env->d_terms; a.done = env->d_done;
  a.phys = env->phys; a.envf = env->envf; a.high_obs = env->high_obs; a.high_reward = env->high_reward;
  a.high_flags = env->high_flags; a.out = env->h_pull_dev;
  a.obs_w = env->obs_w; a.hobs_w = env->hobs_w; a.jt = env->jt;
  pull_kernel<<<env->n, 64, 0, st>>>(a);
  env->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  memcpy(pull_host, env->h_pull, sizeof(float) * ILRL_PULL_WORDS * env->n);
  return ILRL_OK;
}
int ilrl_step_pull(ilrl_env* env, const float* action_host, int32_t forced_target_deg, float* pull_host, void* stream) {
  if (!env) return ILRL_ERR_ARG;
  if (!action_host || !pull_host) return fail(env, ILRL_ERR_ARG, "ilrl_step_pull: null buffer");
  ON_DEVICE(env);
  if (int r = ensure_io_buffers(env)) return r;
  cudaStream_t st = (cudaStream_t)stream;
  // the actions travel through the tail of the mapped staging block: the step kernel reads them in place
  float* ah = env->h_pull + (size_t)ILRL_PULL_WORDS * env->n;
  memcpy(ah, action_host, sizeof(float) * 17 * env->n);
  int r = do_step(env, env->h_pull_dev + (size_t)ILRL_PULL_WORDS * env->n, env->d_obs, env->d_reward, env->d_done, env->d_terms,
                  st, 0, 0, -1, -1, forced_target_deg);
  if (r) return r;
  return do_pull(env, pull_host, nullptr, st);
}
int ilrl_pull(ilrl_env* env, const float* obs_dev, float* pull_host, void* stream) {
  if (!env) return ILRL_ERR_ARG;
  if (!pull_host) return fail(env, ILRL_ERR_ARG, "ilrl_pull: null buffer");
  ON_DEVICE(env);
  if (int r = ensure_io_buffers(env)) return r;
  return do_pull(env, pull_host, obs_dev, (cudaStream_t)stream);
}

// ---- persistent serving (K1s)
static const unsigned long long SERVE_WATCHDOG_NS = 2000000000ull;   // a serving kernel leaves after 2 s without a doorbell
static double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static inline ServeCtl* serve_ctl_of(ilrl_env* env, int part) {
  return reinterpret_cast<ServeCtl*>(reinterpret_cast<char*>(env->serve_ctl) + 256 * part);
}
static inline volatile int* serve_ack_of(ilrl_env* env, int part) {
  return reinterpret_cast<volatile int*>(reinterpret_cast<char*>(env->serve_ctl) + 256 * part + 128);   // its own cache line
}
int ilrl_serve_begin(ilrl_env* env, int32_t nparts, float* obs_h, float* reward_h, uint8_t* done_h, float* terms_h) {
  if (!env) return ILRL_ERR_ARG;
  if (!obs_h || !reward_h || !done_h) return fail(env, ILRL_ERR_ARG, "ilrl_serve_begin: null buffer");
  if (nparts < 1 || nparts > ILRL_MAX_PARTS) return fail(env, ILRL_ERR_ARG, "ilrl_serve_begin: nparts out of range (1..8)");
  if (int r = check_ready(env)) return r;
  ON_DEVICE(env);
  if (env->cfg.mode != 0) return fail(env, ILRL_ERR_ARG, "ilrl_serve_begin: low-level mode only (the hierarchical modes need their high-level call between steps)");
  if (env->terr.h || env->self_on) return fail(env, ILRL_ERR_ARG, "ilrl_serve_begin: not available with a heightfield terrain / self-collision");
  int total_tiles = 0;
  for (int p = 0; p < nparts; p++) {
    int first, count;
    part_range(env, p, nparts, &first, &count);
    total_tiles += (count + QE - 1) / QE;
  }
  if (total_tiles > env->grid_small)   // every serving kernel must be resident at once: a waiting one would never start
    return fail(env, ILRL_ERR_ARG, "ilrl_serve_begin: the batch must fit one wave of resident CTAs (N <= 16 x SMs x 2): use ilrl_step_host_async");
  for (int p = 0; p < ILRL_MAX_PARTS; p++) if (env->part_busy[p]) return fail(env, ILRL_ERR_STATE, "ilrl_serve_begin: a part is still in flight");
  void *dobs, *dr, *dd, *dt = nullptr;
  if (!(dobs = host_alias(env, obs_h)) || !(dr = host_alias(env, reward_h)) || !(dd = host_alias(env, done_h)) ||
      (terms_h && !(dt = host_alias(env, terms_h))))
    return fail(env, ILRL_ERR_ARG, "ilrl_serve_begin: buffers must be page-locked and mapped");
  if (!env->serve_ctl) {
    CK(cudaHostAlloc((void**)&env->serve_ctl, 256 * ILRL_MAX_PARTS, cudaHostAllocMapped));
    CK(cudaMalloc((void**)&env->serve_dev, sizeof(ServeDev) * ILRL_MAX_PARTS));
  }
  memset((void*)env->serve_ctl, 0, 256 * ILRL_MAX_PARTS);
  CK(cudaMemset(env->serve_dev, 0, sizeof(ServeDev) * ILRL_MAX_PARTS));
  CK(cudaDeviceSynchronize());
  void* ctl_dev = nullptr;
  CK(cudaHostGetDevicePointer(&ctl_dev, (void*)env->serve_ctl, 0));
  CK(cudaFuncSetAttribute(serve_kernel<0, SmemSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemSmall)));
  for (int p = 0; p < nparts; p++) {
    int first, count;
    part_range(env, p, nparts, &first, &count);
    env->serve_seq[p] = 0;
    env->serve_posted[p] = 0;
    env->serve_live[p] = false;
    if (count <= 0) continue;
    if (!env->part_stream[p]) CK(cudaStreamCreateWithFlags(&env->part_stream[p], cudaStreamNonBlocking));
    StepArgs a;
    a.n = env->n; a.first = first; a.end = first + count; a.skip_frame = env->cfg.skip_frame;
    a.id_base = env->cfg.env_id_base; a.skip_physics = 0; a.auto_reset = env->cfg.auto_reset; a.max_timestep = env->cfg.max_timestep;
    a.step_per_level = (float)env->cfg.step_per_level; a.seed = env->cfg.seed;
    a.phys = env->phys; a.envf = env->envf; a.rng = env->rng;
    a.action = nullptr; a.obs = (float*)dobs; a.reward = (float*)dr; a.done = (uint8_t*)dd; a.terms = (float*)dt;
    a.high_obs = env->high_obs; a.high_reward = env->high_reward; a.high_flags = env->high_flags;
    a.forced_deg = env->forced_deg; a.forced_scalar = INT_MIN; a.stats = env->stats; a.gscr = env->gscr;
    a.jt = env->jt; a.forced_noise = env->forced_noise; a.terr = env->terr; a.selfc = env->selfc;
    a.tile_counter = env->tile_counter; a.ntiles = (count + QE - 1) / QE; a.ktime = nullptr;
#ifdef ILRL_PROF
    a.prof = nullptr;
#endif
    memcpy(a.clips, env->clips, sizeof a.clips);
    serve_kernel<0, SmemSmall><<<a.ntiles, QT, sizeof(SmemSmall), env->part_stream[p]>>>(
        a, (const ServeCtl*)((char*)ctl_dev + 256 * p), env->serve_dev + p, (volatile int*)((char*)ctl_dev + 256 * p + 128),
        SERVE_WATCHDOG_NS);
    env->launches++;
    CK(cudaGetLastError());
    env->serve_live[p] = true;
  }
  env->serving = true;
  env->serve_nparts = nparts;
  return ILRL_OK;
}
static int serve_gone(ilrl_env* env, const char* who) {
  env->serving = false;
  for (int p = 0; p < ILRL_MAX_PARTS; p++) env->serve_posted[p] = 0;
  for (int p = 0; p < env->serve_nparts; p++)     // the other parts' kernels leave as well
    if (env->serve_live[p]) serve_ctl_of(env, p)->seq = -1;
  cudaGetLastError();
  return fail(env, ILRL_ERR_STATE, std::string(who) + ": a serving kernel is gone (watchdog after 2 s without a step, or a device error)");
}
int ilrl_serve_post(ilrl_env* env, int32_t part, const float* action_h) {
  if (!env) return ILRL_ERR_ARG;
  if (!env->serving) return fail(env, ILRL_ERR_STATE, "ilrl_serve_post: not serving (ilrl_serve_begin first)");
  if (part < 0 || part >= env->serve_nparts) return fail(env, ILRL_ERR_ARG, "ilrl_serve_post: part out of range");
  if (env->serve_posted[part]) return fail(env, ILRL_ERR_STATE, "ilrl_serve_post: the part's previous step has not been waited for");
  if (!env->serve_live[part]) return ILRL_OK;   // (an empty part)
  void* da = action_h ? host_alias(env, action_h) : nullptr;
  if (!da) return fail(env, ILRL_ERR_ARG, "ilrl_serve_post: actions must be in page-locked, mapped host memory");
  ServeCtl* c = serve_ctl_of(env, part);
  c->action = (long long)(uintptr_t)da;
  __atomic_thread_fence(__ATOMIC_RELEASE);
  c->seq = ++env->serve_seq[part];
  env->serve_posted[part] = 1;
  return ILRL_OK;
}
int ilrl_serve_wait(ilrl_env* env, int32_t part) {
  if (!env) return ILRL_ERR_ARG;
  if (!env->serving) return fail(env, ILRL_ERR_STATE, "ilrl_serve_wait: not serving");
  if (part < 0 || part >= env->serve_nparts) return fail(env, ILRL_ERR_ARG, "ilrl_serve_wait: part out of range");
  if (!env->serve_posted[part]) return ILRL_OK;
  const int want = env->serve_seq[part];
  volatile int* ack = serve_ack_of(env, part);
  double t0 = 0.0;
  for (unsigned spins = 0; *ack != want; spins++) {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
    if ((spins & 0xffffu) == 0xffffu) {   // rarely: has the kernel died (watchdog, device error)?
      if (t0 == 0.0) t0 = now_s();
      if (cudaStreamQuery(env->part_stream[part]) != cudaErrorNotReady || now_s() - t0 > 10.0)
        return serve_gone(env, "ilrl_serve_wait");
    }
  }
  __atomic_thread_fence(__ATOMIC_ACQUIRE);
  env->serve_posted[part] = 0;
  return ILRL_OK;
}
int ilrl_serve_step(ilrl_env* env, const float* action_h) {
  if (!env) return ILRL_ERR_ARG;
  for (int p = 0; p < env->serve_nparts; p++) if (int r = ilrl_serve_post(env, p, action_h)) return r;
  for (int p = 0; p < env->serve_nparts; p++) if (int r = ilrl_serve_wait(env, p)) return r;
  return env->serving ? ILRL_OK : fail(env, ILRL_ERR_STATE, "ilrl_serve_step: not serving (ilrl_serve_begin first)");
}
int ilrl_serve_end(ilrl_env* env) {
  if (!env) return ILRL_ERR_ARG;
  if (!env->serving) return ILRL_OK;
  ON_DEVICE_RAW(env);
  for (int p = 0; p < env->serve_nparts; p++) if (env->serve_posted[p] && env->serving) ilrl_serve_wait(env, p);
  cudaError_t r = cudaSuccess;
  for (int p = 0; p < env->serve_nparts; p++) {
    if (!env->serve_live[p]) continue;
    serve_ctl_of(env, p)->seq = -1;
    cudaError_t rp = cudaStreamSynchronize(env->part_stream[p]);
    if (rp != cudaSuccess) r = rp;
    env->serve_live[p] = false;
    env->serve_posted[p] = 0;
  }
  env->serving = false;
  if (r != cudaSuccess) return fail(env, ILRL_ERR_CUDA, std::string("ilrl_serve_end: ") + cudaGetErrorString(r));
  return ILRL_OK;
}

int ilrl_set_config(ilrl_env* env, int32_t max_timestep, int32_t step_per_level, int32_t skip_frame) {
  if (!env) return ILRL_ERR_ARG;
  if (max_timestep > 0) env->cfg.max_timestep = max_timestep;
  if (step_per_level > 0) env->cfg.step_per_level = step_per_level;
  if (skip_frame > 0) env->cfg.skip_frame = skip_frame;
  return ILRL_OK;
}

int ilrl_high_step(ilrl_env* env, const float* action2, float* low_obs, void* stream) {
  if (!env) return ILRL_ERR_ARG;
  if (env->cfg.mode < 1) return fail(env, ILRL_ERR_ARG, "ilrl_high_step: handle is not in hier mode");
  if (!action2 || !low_obs) return fail(env, ILRL_ERR_ARG, "ilrl_high_step: null buffer");
  if (int r = check_ready(env)) return r;
  ON_DEVICE(env);
  HighArgs a;
  a.n = env->n; a.step_per_level = (float)env->cfg.step_per_level; a.skip_frame = env->cfg.skip_frame;
  a.phys = env->phys; a.envf = env->envf; a.jt = env->jt;
  a.action2 = action2; a.low_obs = low_obs;
  memcpy(a.clips, env->clips, sizeof a.clips);
  if (env->cfg.mode == 1) high_step_kernel<1><<<nblk(env->n), BLOCK, 0, (cudaStream_t)stream>>>(a);
  else high_step_kernel<2><<<nblk(env->n), BLOCK, 0, (cudaStream_t)stream>>>(a);
  env->launches++;
  CK(cudaGetLastError());
  return ILRL_OK;
}

int ilrl_high_readout(ilrl_env* env, float* high_obs, float* high_reward, uint8_t* high_flags, void* stream) {
  if (!env) return ILRL_ERR_ARG;
  if (env->cfg.mode < 1) return fail(env, ILRL_ERR_ARG, "ilrl_high_readout: handle is not in hier mode");
  ON_DEVICE(env);
  const int words = env->hobs_w * env->n;
  high_readout_kernel<<<(words + 255) / 256, 256, 0, (cudaStream_t)stream>>>(env->n, env->hobs_w, env->high_obs, env->high_reward, env->high_flags,
                                                                           high_obs, high_reward, high_flags);
  env->launches++;
  CK(cudaGetLastError());
  return ILRL_OK;
}

int ilrl_get_state(ilrl_env* env, float* phys, float* envf, void* stream) {
  if (!env) return ILRL_ERR_ARG;
  ON_DEVICE(env);
  state_get_kernel<<<(env->n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(view(env), phys, envf);
  env->launches++;
  CK(cudaGetLastError());
  return ILRL_OK;
}
int ilrl_set_state(ilrl_env* env, const float* phys, const float* envf, void* stream) {
  if (!env) return ILRL_ERR_ARG;
  ON_DEVICE(env);
  state_set_kernel<<<(env->n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(view(env), phys, envf);
  env->launches++;
  CK(cudaGetLastError());
  return ILRL_OK;
}
int ilrl_set_forced_target_deg(ilrl_env* env, const int32_t* deg) {
  if (!env) return ILRL_ERR_ARG;
  env->forced_deg = deg;
  return ILRL_OK;
}
int ilrl_set_heightfield(ilrl_env* env, const float* heights, int32_t rows, int32_t cols, float zoff) {
  if (!env) return ILRL_ERR_ARG;
  ON_DEVICE(env);
  CK(cudaDeviceSynchronize());   // no step may be reading the old terrain
  if (!heights) {                // back to the flat ground plane
    cudaFree(env->terr_mem);
    env->terr_mem = nullptr;
    env->terr = chain::Terrain{nullptr, 0, 0, 0.f, 0.f, 1.f};
    return ILRL_OK;
  }
  if (env->cfg.mode != 0) return fail(env, ILRL_ERR_ARG, "ilrl_set_heightfield: only the low-level env has a terrain (REF low_level_env.py:43-47)");
  if (rows < 2 || cols < 2 || rows > 4096 || cols > 4096) return fail(env, ILRL_ERR_ARG, "ilrl_set_heightfield: bad grid size");
  float hmax = -1e30f, g2max = 0.f;
  for (int j = 0; j < cols; j++)
    for (int i = 0; i < rows; i++) {
      const float h = heights[i + (size_t)j * rows];
      if (!(h == h) || fabsf(h) > 1e6f) return fail(env, ILRL_ERR_ARG, "ilrl_set_heightfield: non-finite height");
      hmax = fmaxf(hmax, h);
      if (i + 1 < rows && j + 1 < cols) {   // both triangles of the cell
        const float h10 = heights[i + 1 + (size_t)j * rows], h01 = heights[i + (size_t)(j + 1) * rows],
                    h11 = heights[i + 1 + (size_t)(j + 1) * rows];
        g2max = fmaxf(g2max, (h10 - h) * (h10 - h) + (h01 - h) * (h01 - h));
        g2max = fmaxf(g2max, (h11 - h01) * (h11 - h01) + (h11 - h10) * (h11 - h10));
      }
    }
  if (g2max >= 1.f) return fail(env, ILRL_ERR_ARG, "ilrl_set_heightfield: a triangle is steeper than 45 degrees (the friction "
                                                   "frame of this path assumes |n.z| > 0.707)");
  if (env->terr_mem && (env->terr.rows != rows || env->terr.cols != cols)) { cudaFree(env->terr_mem); env->terr_mem = nullptr; }
  if (!env->terr_mem) CK(cudaMalloc(&env->terr_mem, sizeof(float) * (size_t)rows * cols));
  CK(cudaMemcpy(env->terr_mem, heights, sizeof(float) * (size_t)rows * cols, cudaMemcpyHostToDevice));
  env->terr = chain::Terrain{env->terr_mem, rows, cols, zoff, hmax + zoff, 1.f / sqrtf(1.f + g2max)};
  return ILRL_OK;
}
int ilrl_set_self_collision(ilrl_env* env, int32_t on) {
  if (!env) return ILRL_ERR_ARG;
  ON_DEVICE(env);
  CK(cudaDeviceSynchronize());
  if (!on) { env->self_on = false; return ILRL_OK; }
  env->self_on = true;
  return ILRL_OK;
}
int ilrl_set_forced_reset_noise(ilrl_env* env, const float* noise17) {
  if (!env) return ILRL_ERR_ARG;
  if (env->cfg.mode != 2) return fail(env, ILRL_ERR_ARG, "ilrl_set_forced_reset_noise: only the hier_env_2 mode keeps reset noise");
  env->forced_noise = noise17;
  return ILRL_OK;
}
int ilrl_get_joint_target(ilrl_env* env, float* jt, void* stream) {
  if (!env) return ILRL_ERR_ARG;
  if (env->cfg.mode != 2 || !jt) return fail(env, ILRL_ERR_ARG, "ilrl_get_joint_target: mode 2 handle and a buffer needed");
  ON_DEVICE(env);
  jt_copy_kernel<<<(env->n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(env->n, env->jt, jt, nullptr);
  env->launches++;
  CK(cudaGetLastError());
  return ILRL_OK;
}
int ilrl_set_joint_target(ilrl_env* env, const float* jt, void* stream) {
  if (!env) return ILRL_ERR_ARG;
  if (env->cfg.mode != 2 || !jt) return fail(env, ILRL_ERR_ARG, "ilrl_set_joint_target: mode 2 handle and a buffer needed");
  ON_DEVICE(env);
  jt_copy_kernel<<<(env->n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(env->n, env->jt, nullptr, jt);
  env->launches++;
  CK(cudaGetLastError());
  return ILRL_OK;
}
int ilrl_physics_only(ilrl_env* env, const float* torque, void* stream) {
  if (!env) return ILRL_ERR_ARG;
  if (!torque) return fail(env, ILRL_ERR_ARG, "ilrl_physics_only: null buffer");
  ON_DEVICE(env);
  if (env->terr.h && env->self_on)
    physics_only_kernel<SmemSelf, true, true><<<(env->n + QE - 1) / QE, QT, sizeof(SmemSelf), (cudaStream_t)stream>>>(view(env), torque, env->gscr, env->substeps, env->terr, env->selfc);
  else if (env->self_on)
    physics_only_kernel<SmemSelf, false, true><<<(env->n + QE - 1) / QE, QT, sizeof(SmemSelf), (cudaStream_t)stream>>>(view(env), torque, env->gscr, env->substeps, env->terr, env->selfc);
  else if (env->terr.h)
    physics_only_kernel<SmemLarge, true><<<(env->n + QE - 1) / QE, QT, sizeof(SmemLarge), (cudaStream_t)stream>>>(view(env), torque, env->gscr, env->substeps, env->terr);
  else if (env->layout == 2)
    physics_only_kernel<SmemDense4><<<(env->n + QE - 1) / QE, QT, sizeof(SmemDense4), (cudaStream_t)stream>>>(view(env), torque, env->gscr, env->substeps);
  else if (env->layout == 1)
    physics_only_kernel<SmemLarge><<<(env->n + QE - 1) / QE, QT, sizeof(SmemLarge), (cudaStream_t)stream>>>(view(env), torque, env->gscr, env->substeps);
  else
    physics_only_kernel<SmemSmall><<<(env->n + QE - 1) / QE, QT, sizeof(SmemSmall), (cudaStream_t)stream>>>(view(env), torque, env->gscr, env->substeps);
  env->launches++;
  CK(cudaGetLastError());
  return ILRL_OK;
}
int ilrl_endpoint_score(ilrl_env* env, float* score, void* stream) {
  if (!env) return ILRL_ERR_ARG;
  if (!score) return fail(env, ILRL_ERR_ARG, "ilrl_endpoint_score: null buffer");
  if (int r = check_ready(env)) return r;
  ON_DEVICE(env);
  EpArgs a; a.v = view(env); a.score = score;
  memcpy(a.clips, env->clips, sizeof a.clips);
  endpoint_kernel<<<nblk(env->n), BLOCK, 0, (cudaStream_t)stream>>>(a);
  env->launches++;
  CK(cudaGetLastError());
  return ILRL_OK;
}
int ilrl_stats(ilrl_env* env, float* stats16, void* stream) {
  if (!env) return ILRL_ERR_ARG;
  if (!stats16) return fail(env, ILRL_ERR_ARG, "ilrl_stats: null buffer");
  ON_DEVICE(env);
  stats_fetch_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(env->stats, stats16);
  env->launches++;
  CK(cudaGetLastError());
  return ILRL_OK;
}
int ilrl_gae(const float* reward, const float* value, const uint8_t* done, float gamma, float lam, float* advantage,
             float* value_target, int32_t T, int32_t n, void* stream) {
  if (!reward || !value || !done || !advantage || !value_target || T <= 0 || n <= 0) return ILRL_ERR_ARG;
  gae_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(reward, value, done, gamma, lam, advantage, value_target, T, n);
  return cudaGetLastError() == cudaSuccess ? ILRL_OK : ILRL_ERR_CUDA;
}
int ilrl_episode_columns(const uint8_t* done, int32_t* t_carry, int32_t* eps_carry, int32_t* t_out, int64_t* eps_id_out,
                          int32_t T, int32_t n, int64_t env_id_base, void* stream) {
  if (!done || !t_carry || !eps_carry || !t_out || !eps_id_out || T <= 0 || n <= 0) return ILRL_ERR_ARG;
  episode_columns_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(done, t_carry, eps_carry, t_out,
                                                                          (long long*)eps_id_out, T, n, (long long)env_id_base);
  return cudaGetLastError() == cudaSuccess ? ILRL_OK : ILRL_ERR_CUDA;
}
int ilrl_gae_decisions(const float* reward, const uint8_t* flags, const float* value, float gamma, float lam,
                       float* advantage, float* value_target, uint8_t* valid, int32_t T, int32_t n, void* stream) {
  if (!reward || !flags || !value || !advantage || !value_target || !valid || T <= 0 || n <= 0) return ILRL_ERR_ARG;
  gae_decisions_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(reward, flags, value, gamma, lam, advantage,
                                                                         value_target, valid, T, n);
  return cudaGetLastError() == cudaSuccess ? ILRL_OK : ILRL_ERR_CUDA;
}
/* harness only, not in ilrl.h: on = 0 forces ilrl_step_host onto explicit copies even for page-locked buffers */
int ilrl_debug_zero_copy(ilrl_env* env, int32_t on) { if (!env) return ILRL_ERR_ARG; env->no_zero_copy = on == 0; return ILRL_OK; }
#ifdef ILRL_PROF
/* measurement build only: copy out and clear the per-warp phase cycles ([ceil(N/8)][16] int64) */
int ilrl_debug_profile(ilrl_env* env, long long* out_host) {
  if (!env || !out_host) return ILRL_ERR_ARG;
  ON_DEVICE(env);
  const size_t bytes = sizeof(long long) * chain::PF_WORDS * ((env->n + 7) / 8);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out_host, env->prof, bytes, cudaMemcpyDeviceToHost));
  CK(cudaMemset(env->prof, 0, bytes));
  return ILRL_OK;
}
#endif
/* harness only, not in ilrl.h: number of substeps ilrl_physics_only runs */
int ilrl_debug_substeps(ilrl_env* env, int32_t n) { if (!env || n < 1) return ILRL_ERR_ARG; env->substeps = n; return ILRL_OK; }
int64_t ilrl_launch_count(const ilrl_env* env) { return env ? env->launches : 0; }
int ilrl_kernel_timing(ilrl_env* env, int32_t on, float* ms_out, int64_t* launches_out) {
  if (!env) return ILRL_ERR_ARG;
  ON_DEVICE(env);
  float ms = 0.f;
  int64_t cnt = 0;
  if (env->ktime && env->kt_used > 0) {   // durations of the launches stamped since the last call
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> h(2 * (size_t)env->kt_used);
    CK(cudaMemcpy(h.data(), env->ktime, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost));
    for (int k = 0; k < env->kt_used; k++)
      if (h[2 * k + 1] >= h[2 * k]) { ms += (float)((double)(h[2 * k + 1] - h[2 * k]) * 1e-6); cnt++; }
  }
  if (ms_out) *ms_out = ms;
  if (launches_out) *launches_out = cnt;
  env->timing = on != 0;
  env->kt_used = 0;
  if (env->timing) {
    if (!env->ktime) CK(cudaMalloc(&env->ktime, sizeof(unsigned long long) * 2 * KT_SLOTS));
    std::vector<unsigned long long> init(2 * (size_t)KT_SLOTS);
    for (int k = 0; k < KT_SLOTS; k++) { init[2 * k] = ~0ull; init[2 * k + 1] = 0ull; }
    CK(cudaMemcpy(env->ktime, init.data(), sizeof(unsigned long long) * init.size(), cudaMemcpyHostToDevice));
  }
  return ILRL_OK;
}

}  // extern "C"
