/* ilrl.h — C ABI of the B200-native batched humanoid-imitation env (libilrl_b200.so).
 *
 * Drop-in boundary for ONE hot path of AdityaPutraS/Imitation-Learning-RL: the env step / reset + imitation reward +
 * observation that the reference reaches through
 *     LowLevelHumanoidEnv.{reset,resetFromFrame,step}          REF low_level_env.py:224-305, 322-323, 475-526
 *     HierarchicalHumanoidEnv.{reset,resetFromFrame,step}      REF hier_env.py:235-319, 355-366, 538-642
 *     hier_env_2.HierarchicalHumanoidEnv.{reset,...,step}      REF hier_env_2.py:254-352, 408-419, 699-769  (mode 2)
 *     CustomHumanoidRobot.apply_action / calc_state            REF humanoid.py:54-60 (+ pybullet_envs WalkerBase)
 *     scene.global_step() -> pybullet.stepSimulation           REF low_level_env.py:479, hier_env.py:589
 * The reference has no FFI of its own (it is pure Python over pybullet's C extension); the entry points below are
 * what a ctypes binding inside those two env classes binds (shown in INTEGRATION.md).
 *
 * Conventions
 *   - plain C types only; every array argument named *_dev is a DEVICE pointer to caller-owned contiguous memory,
 *     every *_host argument a HOST pointer; `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - calls taking device pointers are asynchronous on `stream`; calls taking host pointers return after the data
 *     has arrived (they synchronise `stream`).
 *   - return 0 on success, a negative ilrl_status otherwise; ilrl_last_error() gives the text.
 *   - one handle = one GPU = N envs; a handle is not thread-safe.
 *   - there is NO CPU fallback: if no CUDA device is usable ilrl_create fails with ILRL_ERR_CUDA.
 * Layouts: observations [N,70] (low) / [N,44] (high) row-major fp32, actions [N,17] / [N,2] fp32, reward [N] fp32,
 * done [N] uint8, terms [N,12] fp32 (ILRL_T_* order, ilrl_constants.h), phys [N,47] and envf [N,28] fp32
 * (ILRL_PHYS_WORDS / ILRL_E_* order).
 * Mode 2 (hier_env_2.py: the high level hands the low level 17 joint (position, velocity) targets) has wider rows:
 * low obs [N,72], high obs [N,60], high action [N,36] (2 unused + jointTarget[34]); everywhere below "[N,70]",
 * "[N,44]" and "[N,2]" read as those widths for a mode-2 handle.  Its defaults are skip_frame 5, step_per_level 20.
 * The reference's hier_env_2 reads a data directory and a robot MJCF it does not ship; the declared substitutions
 * ("Joints CSV With Hand", humanoid_symmetric_2.xml) are stated in DESIGN.md section 4.
 */
#ifndef ILRL_H
#define ILRL_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ilrl_env ilrl_env; /* opaque */

typedef enum {
  ILRL_OK = 0,
  ILRL_ERR_ARG = -1,   /* bad argument (NULL, out of range, clip not loaded, wrong mode) */
  ILRL_ERR_CUDA = -2,  /* CUDA runtime error or no device */
  ILRL_ERR_STATE = -3  /* call order violated (e.g. step before every clip id has a loaded clip) */
} ilrl_status;

typedef struct {
  int32_t device;         /* CUDA ordinal */
  int32_t num_envs;       /* N */
  int32_t mode;           /* 0 = LowLevelHumanoidEnv, 1 = HierarchicalHumanoidEnv (hier_env.py), 2 = hier_env_2.py variant */
  int32_t auto_reset;     /* 1: an env that finishes is reset inside the same step kernel (obs = first obs of the
                             new episode, done = 1); 0: the caller resets, as the reference env does (Q19) */
  uint64_t seed;          /* Philox key for start frames, headings and re-sampled targets */
  int32_t skip_frame;     /* 2  REF low_level_env.py:162 (mode 2: 5, REF hier_env_2.py:160); <= 0 = the mode's default */
  int32_t max_timestep;   /* 3000 REF low_level_env.py:73 */
  int32_t step_per_level; /* 5  REF hier_env.py:58 (mode 2: 20, REF hier_env_2.py:58); <= 0 = the mode's default */
  int32_t env_id_base;    /* global id of env 0 of this handle (0 on a single GPU).  The random draws of env i are a
                             function of (seed, env_id_base + i, draw index) only, so sharding a batch over GPUs with
                             the same seed does not change any env's trajectory */
} ilrl_config;

int ilrl_create(const ilrl_config* cfg, ilrl_env** out);
void ilrl_destroy(ilrl_env* env);
const char* ilrl_last_error(const ilrl_env* env); /* env may be NULL: error of the last failed ilrl_create */

/* Stage one motion clip (the four CSV tables the reference reads at REF low_level_env.py:58-71) once in HBM.
 * Row-major host fp32; max_frame as the reference computes it (len(pos) - 1, REF low_level_env.py:80-82) — callers
 * clamp it for motion13_13 whose velocity table is short (declared divergence). clip in [0, 8). */
int ilrl_load_clip(ilrl_env* env, int32_t clip, const float* pos14_host, int32_t n_pos, const float* rel14_host,
                   int32_t n_rel, const float* vel14_host, int32_t n_vel, const float* ep27_host, int32_t n_ep,
                   int32_t max_frame);
/* Which clip each env imitates (`reference_name` / `selected_motion`). clip_of_env_host[N]; NULL = all envs clip 0. */
int ilrl_set_clip_ids(ilrl_env* env, const int32_t* clip_of_env_host);

/* reset()/resetFromFrame() for the envs whose mask byte is non-zero (mask_dev NULL = all).
 * start_frame_dev / target_deg_dev / reset_yaw_deg_dev: per-env overrides, NULL = drawn as the reference draws them
 * (start frame U{0..max_frame-6}, heading U{-180..179} deg; reset_yaw 0 in low mode, U{-180..179} in hier mode).
 * target_xy_dev [N,2]: explicit first target instead of 5 m along the drawn heading (the reference's
 * usePredefinedTarget path, REF low_level_env.py:253-255, hier_env.py:261-265); NULL = none.
 * obs_dev: [N,70] in low mode, [N,44] (high-level obs) in hier mode; rows of unmasked envs are left untouched. */
int ilrl_reset(ilrl_env* env, const uint8_t* mask_dev, const int32_t* start_frame_dev, const int32_t* target_deg_dev,
               const float* reset_yaw_deg_dev, const float* target_xy_dev, float* obs_dev, void* stream);

/* One low-level env step for every env: apply_action -> 4 physics substeps -> calc_state -> reward -> frame advance
 * -> target bookkeeping -> observation -> termination (-> reset if auto_reset).  terms_dev may be NULL.
 * In hier mode envs that are waiting for a high-level action are skipped (done 0, reward 0, obs row untouched) and
 * the high-level outputs are kept inside the handle until ilrl_high_readout.  In both modes an env whose action row
 * starts with NaN is skipped the same way ("no action for this env in this call"). */
int ilrl_step(ilrl_env* env, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
              float* terms_dev, void* stream);

/* K consecutive steps of the LOW-LEVEL env (mode 0, flat ground, no self-collision) in one launch, for callers that hold
 * the actions of all K steps up front (open-loop playback of recorded / scripted action sequences, random-action
 * rollouts): action_dev [K,N,17]; obs_dev [K,N,70], reward_dev / done_dev [K,N], terms_dev [K,N,12] or NULL.  Bit-identical
 * to K calls of ilrl_step (auto-reset included); envs do not wait for each other between steps. */
int ilrl_step_sequence(ilrl_env* env, int32_t ksteps, const float* action_dev, float* obs_dev, float* reward_dev,
                       uint8_t* done_dev, float* terms_dev, void* stream);

/* Same step through HOST buffers: copies the actions up, runs the step, copies obs/reward/done back and waits.
 * When every buffer is page-locked and mapped (cudaHostAlloc / cudaHostRegister / torch pin_memory) the kernel reads
 * and writes them in place (zero-copy: one launch, one synchronise); otherwise page-locked buffers are DMA endpoints
 * and pageable ones are staged through pinned mirrors owned by the handle.
 * This is the call the reference-facing Python env classes make (end-to-end path measured by bench.py "e2e"). */
int ilrl_step_host(ilrl_env* env, const float* action_host, float* obs_host, float* reward_host, uint8_t* done_host,
                   float* terms_host, void* stream);

/* Double-buffered host stepping: the batch is cut into `nparts` (<= 8) contiguous parts of whole 16-env tiles and each
 * part steps on its own stream, so that a rollout worker computes the actions of one part while another part steps
 * (the asynchronous vector-env pattern; what hides the launch, PCIe and synchronise latencies that ilrl_step_host
 * exposes once per step).  The arguments are the FULL [N, ...] arrays, page-locked and mapped (else ILRL_ERR_ARG):
 * the kernel reads rows [first, first + count) of the actions and writes the same rows of obs / reward / done
 * (/ terms) in place.  The call returns as soon as the step is enqueued; ilrl_wait(part) returns when the part's
 * outputs have landed in host memory.  One step per part may be in flight (ILRL_ERR_STATE otherwise).  Parts are
 * independent envs: stepping the parts separately gives bit-identical results to one ilrl_step_host of the batch.
 * Replaces the same reference call as ilrl_step_host (REF low_level_env.py:322-323 under a vectorised worker). */
int ilrl_step_host_async(ilrl_env* env, int32_t part, int32_t nparts, const float* action_host, float* obs_host,
                         float* reward_host, uint8_t* done_host, float* terms_host);
int ilrl_wait(ilrl_env* env, int32_t part);
/* ilrl_wait(part) followed by ilrl_step_host_async(part, ...) in one call (the steady state of a double-buffered loop:
 * the caller has already consumed the part's previous outputs when it presents the next actions... if it has not, it
 * calls ilrl_wait itself).  Saves one boundary crossing per part and step. */
int ilrl_wait_step_host_async(ilrl_env* env, int32_t part, int32_t nparts, const float* action_host, float* obs_host,
                              float* reward_host, uint8_t* done_host, float* terms_host);

/* Persistent serving: the lowest-latency form of the host path.  Between ilrl_serve_begin and ilrl_serve_end the step
 * kernel stays RESIDENT on the GPU — one launch per part (`nparts` <= 8 contiguous parts of whole 16-env tiles, as in
 * ilrl_step_host_async) — and every part is driven through its own doorbell in mapped host memory, so an env step costs
 * neither a kernel launch nor a stream synchronisation: ilrl_serve_post(part) publishes the step's actions (the FULL
 * [N,17] array, any page-locked, mapped buffer, already filled: the part reads its own rows), ilrl_serve_wait(part)
 * spins until obs / reward / done (/ terms) of the part's rows have landed in the buffers given to ilrl_serve_begin.
 * With several parts a worker overlaps its own work on one part with the steps of the others, and the parts'
 * observation bursts no longer hit PCIe at the same moment.  ilrl_serve_step = post all parts, wait for all.
 * Results are identical to ilrl_step_host.  Low-level mode, flat ground, and a batch that fits one wave of resident CTAs
 * (N <= 16 x 2 x SMs = 4736 on a B200; larger batches use ilrl_step_host_async).  While serving, every other call on
 * the handle fails with ILRL_ERR_STATE, and device-wide synchronisation (cudaDeviceSynchronize, torch.cuda.synchronize,
 * ilrl_create of another handle) must be avoided: it would wait for the resident kernels, which leave by themselves
 * after 2 s without a step (watchdog; the next ilrl_serve_* call then reports ILRL_ERR_STATE).  Replaces the same
 * reference call as ilrl_step_host (REF low_level_env.py:322-323). */
int ilrl_serve_begin(ilrl_env* env, int32_t nparts, float* obs_host, float* reward_host, uint8_t* done_host, float* terms_host);
int ilrl_serve_post(ilrl_env* env, int32_t part, const float* action_host);
int ilrl_serve_wait(ilrl_env* env, int32_t part);
int ilrl_serve_step(ilrl_env* env, const float* action_host);
int ilrl_serve_end(ilrl_env* env);

/* The reference-shaped single-env classes (LowLevelHumanoidEnv / HierarchicalHumanoidEnv, N = 1 views) mirror, after every
 * call, what the reference keeps as Python attributes: the observation, reward, done, the 12 reward terms, the env words
 * (frame, target, robot_pos, ...), the physics state and the high-level agent's outputs.  ilrl_step_pull does one
 * blocking step from HOST actions [N,17] and returns all of it as ONE packed host row per env (one launch for the step,
 * one for the packing, one synchronise; nothing else crosses the bus):
 *   [N][ILRL_PULL_WORDS] fp32 = obs 72 | reward | done | terms 12 | envf 28 | phys 47 | high obs 60 | high reward | flags
 *                               | jointTarget 34      (obs / high obs columns beyond the mode's width and, outside mode 2,
 *                               the jointTarget columns are 0)
 * forced_target_deg: INT32_MIN, or the heading (integer degrees) every env uses if it re-samples its target in this step
 * (the N = 1 views draw it from the env object's own generator, as the reference does, REF low_level_env.py:240-245).
 * ilrl_pull packs the same row without stepping (after a reset or a high-level step); obs_dev NULL = the observation
 * of the last ilrl_step_pull, else a device [N,70] buffer to take the observation columns from. */
#define ILRL_PULL_WORDS 257
int ilrl_step_pull(ilrl_env* env, const float* action_host, int32_t forced_target_deg, float* pull_host, void* stream);
int ilrl_pull(ilrl_env* env, const float* obs_dev, float* pull_host, void* stream);

/* The reference's drivers assign these attributes on a live env (REF env_vis_hier.py:52 `env.max_timestep = 100000`;
 * hier_env_2.py:58-63 uses step_per_level 20 and skipFrame 5): change them on the handle.  A non-positive argument
 * keeps the current value.  Takes effect from the next step / reset. */
int ilrl_set_config(ilrl_env* env, int32_t max_timestep, int32_t step_per_level, int32_t skip_frame);

/* hier mode: high-level agent's action (cos, sin of the heading) for every env that is waiting for one; the others
 * ignore their row, and so does a waiting env whose row starts with NaN.  low_obs_dev [N,70]: the low-level obs
 * the reference returns from high_level_step.  Mode 2: action [N,36] whose columns 2..35 become the env's jointTarget
 * (REF hier_env_2.py:731), low_obs_dev [N,72]; the frame advances by skip_frame here, not in the low-level step. */
int ilrl_high_step(ilrl_env* env, const float* action2_dev, float* low_obs_dev, void* stream);
/* hier mode: high-level obs [N,44], reward [N] and flags [N] (bit0 = episode ended this step, bit1 = high-level
 * agent present in the reference's returned dicts, bit2 = env is now waiting for a high-level action). */
int ilrl_high_readout(ilrl_env* env, float* high_obs_dev, float* high_reward_dev, uint8_t* high_flags_dev,
                      void* stream);

/* Parity harness: read / overwrite the full per-env state (AoS fp32 as documented above). */
int ilrl_get_state(ilrl_env* env, float* phys_dev, float* envf_dev, void* stream);
int ilrl_set_state(ilrl_env* env, const float* phys_dev, const float* envf_dev, void* stream);
/* Parity harness: the random heading the next target re-sampling uses, per env (INT32_MIN entry = draw normally);
 * NULL clears the override.  The pointer must stay valid until cleared. */
int ilrl_set_forced_target_deg(ilrl_env* env, const int32_t* deg_dev);
/* Heightfield terrain of the `useCustomEnv=True` low-level env (REF humanoid.py:68-144 CustomScene: 256 x 256 samples,
 * 1 m cells, the terrain body at z = 0.25; REF env_vis_low.py:155-171 replaces the data with a ramp).  Replaces
 * CustomScene.episode_restart's createCollisionShape(GEOM_HEIGHTFIELD, ...) / replaceHeightfieldData for every env of
 * the handle (one terrain per handle).  heights_host[i + j * rows] as the reference's heightfieldData; world height of a
 * sample = value + zoff (Bullet centres the shape on (min + max) / 2: zoff = 0.25 - (min + max) / 2 mirrors the
 * reference).  The step then runs the terrain instantiation of the kernel: ground contacts against the plane of the
 * triangle under each candidate sphere.  NULL = flat ground again.  Mode 0 only; slopes must stay below 45 degrees.
 * Synchronises the device. */
int ilrl_set_heightfield(ilrl_env* env, const float* heights_host, int32_t rows, int32_t cols, float zoff);

/* Self-collision of the robot (REF humanoid.py:13 `self_collision = True`; Bullet: URDF_USE_SELF_COLLISION |
 * URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS): on = 1 makes the step run the self-collision instantiation of the kernel
 * - every pair of the 14 limb / waist geoms whose bodies are not ancestor-related (66 capsule - capsule pairs), at most
 * 8 contacts per env, two-body rows (normal + friction, coefficient 2.0 x 2.0) after the ground contacts.  Off by
 * default (the path north_star scopes has ground contact only); every mode.  Synchronises the device. */
int ilrl_set_self_collision(ilrl_env* env, int32_t on);

/* Mode 2 parity harness.  hier_env_2's reset leaves WalkerBase.robot_specific_reset's joint noise (uniform(-0.1, 0.1))
 * in the six arm joints (its setJointsOrientation writes the abdomen and the legs only, REF hier_env_2.py:214-252).
 * noise17_dev [N,17] (ordered_joints order; only the arm entries matter): used by ilrl_reset and auto-resets instead
 * of the env's own draws; NULL = draw.  The pointer must stay valid until cleared. */
int ilrl_set_forced_reset_noise(ilrl_env* env, const float* noise17_dev);
/* Mode 2: read / overwrite the per-env jointTarget [N,34]. */
int ilrl_get_joint_target(ilrl_env* env, float* jt_dev, void* stream);
int ilrl_set_joint_target(ilrl_env* env, const float* jt_dev, void* stream);
/* Measurement aid: the cost key [N] u8 (0 = most expensive) each env left at its last grouped step and the current
 * env order [N] i32 of whole-batch steps (K8; identity until a batch of more than one wave has stepped).
 * Either pointer may be NULL.  Grouping never changes results; ILRL_GROUP_EVERY=0 in the environment switches it off. */
int ilrl_get_grouping(ilrl_env* env, uint8_t* cost_dev, int32_t* perm_dev, void* stream);
/* Parity harness (K3): everything ilrl_step does EXCEPT the physics, on the state currently held. */
int ilrl_step_no_physics(ilrl_env* env, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                         float* terms_dev, void* stream);
/* Parity harness: physics only (one env step = 4 substeps) with joint-order torques [N,17]; no env bookkeeping. */
int ilrl_physics_only(ilrl_env* env, const float* torque_dev, void* stream);
/* calcEndPointScore (REF low_level_env.py:361-382; not on any step path, used by param_check.py): score [N]. */
int ilrl_endpoint_score(ilrl_env* env, float* score_dev, void* stream);

/* Episode statistics accumulated on the device since the last call (then zeroed): stats16_dev[16] =
 * {episodes, sum return, sum length, steps, sum reward, sum of the 11 first terms}.  The caller all-reduces the 16
 * floats across ranks (NCCL sum) — the only collective of the path. */
int ilrl_stats(ilrl_env* env, float* stats16_dev, void* stream);

/* Rollout post-processing on the device (what RLlib's `compute_advantages` does on the host
 * for the reference's PPO configs, gamma / lambda of REF train_config.py:97-98): generalised advantage estimation
 * over a [T, N] rollout laid out step-major.  reward_dev [T,N], value_dev [T+1,N] (row T = value of the state after
 * the last step), done_dev [T,N] uint8 (a done step does not bootstrap), outputs advantage_dev / value_target_dev
 * [T,N].  No handle: pure function of its arguments, asynchronous on `stream`. */
int ilrl_gae(const float* reward_dev, const float* value_dev, const uint8_t* done_dev, float gamma, float lambda_,
             float* advantage_dev, float* value_target_dev, int32_t T, int32_t n, void* stream);

/* The two bookkeeping columns RLlib keeps next to every step of a sample batch, for a [T, N] fragment laid out step-major:
 * t_dev [T,N] int32 = step index inside the episode, eps_id_dev [T,N] int64 = (env_id_base + env) << 32 | episode counter.
 * t_carry_dev / eps_carry_dev [N] int32 (in / out): t of the env's next step and its episode counter, continued from
 * fragment to fragment (zero them once).  No handle; asynchronous on `stream`. */
int ilrl_episode_columns(const uint8_t* done_dev, int32_t* t_carry_dev, int32_t* eps_carry_dev, int32_t* t_dev,
                         int64_t* eps_id_dev, int32_t T, int32_t n, int64_t env_id_base, void* stream);

/* GAE for the high-level agent of the hierarchical env over a fragment of T ticks (rows 0..T of the ilrl_high_readout
 * outputs, taken before every tick and once after the last; value_dev [T+1,N] = value estimates of those high-level
 * observations).  A decision is taken where flags has bit2; its reward / termination / successor value are the ones
 * reported at the next row with bit1 (REF hier_env.py:524-536, 613-631).  Outputs [T,N]: advantage, value target and
 * valid (1 = a decision with its outcome inside the fragment; everything else is 0 / masked out). */
int ilrl_gae_decisions(const float* reward_dev, const uint8_t* flags_dev, const float* value_dev, float gamma,
                       float lambda_, float* advantage_dev, float* value_target_dev, uint8_t* valid_dev, int32_t T,
                       int32_t n, void* stream);

/* ---- fused policy / value forward for on-device rollout collection -------------------
 * Replaces the per-step model forward + action sampling RLlib does around env.step for the reference's policies
 * (REF train_config.py:91-113 low level 70-256-256-17, :262-286 high level 44-256-256-2: fcnet_hiddens [256, 256],
 * tanh, free log-std, separate value branch).  tcgen05 tensor-core kernel, bf16 operands, fp32 accumulation.
 * No handle: pure functions of their arguments, asynchronous on `stream`; every pointer is a device pointer.
 *
 * ilrl_policy_pack: fp32 parameters in torch.nn.Linear layout ([out, in] row major; w1 [256, obs_dim], w2 [256, 256],
 * w3_pi [act_dim, 256], w3_vf [1, 256], log_std [act_dim]) -> the packed blob (ilrl_policy_blob_bytes() bytes, 16-byte
 * aligned) the step reads.  obs_dim <= 80, act_dim <= 32.  Repack after every optimizer update.
 * ilrl_policy_step, for n envs: mean, value = model(obs); a = mean + exp(log_std) * noise (noise NULL: a = mean);
 * writes action [n, act_dim] (raw sample), action_clipped = clip(a, -1, 1) (what env.step consumes, RLlib
 * clip_actions=True), logp [n] (diagonal Gaussian log-density of a) and value [n].  Any output may be NULL; with all
 * three policy outputs NULL only the value net runs (bootstrap value of the last observation). */
int64_t ilrl_policy_blob_bytes(void);
int ilrl_policy_pack(const float* w1_pi, const float* b1_pi, const float* w2_pi, const float* b2_pi, const float* w3_pi,
                     const float* b3_pi, const float* w1_vf, const float* b1_vf, const float* w2_vf, const float* b2_vf,
                     const float* w3_vf, const float* b3_vf, const float* log_std, int32_t obs_dim, int32_t act_dim,
                     void* blob_dev, void* stream);
int ilrl_policy_step(const void* blob_dev, const float* obs_dev, const float* noise_dev, float* action_dev,
                     float* action_clipped_dev, float* logp_dev, float* value_dev, int32_t obs_dim, int32_t act_dim,
                     int32_t n, void* stream);

/* How many kernels of this library have been launched through the handle (bench.py "gpu_launches"). */
int64_t ilrl_launch_count(const ilrl_env* env);
/* Time of the step kernels only, measured ON THE DEVICE: while enabled (on = 1) every step kernel stamps %globaltimer
 * when its first CTA starts and when its last CTA ends (two atomics per CTA; no extra launch, no synchronisation, so
 * the launches still run back to back).  Returns the accumulated kernel milliseconds and the number of timed launches
 * since the previous call (it synchronises the device to read them; at most 8192 launches are kept), then clears both. */
int ilrl_kernel_timing(ilrl_env* env, int32_t on, float* ms_out, int64_t* launches_out);

#ifdef __cplusplus
}
#endif
#endif
