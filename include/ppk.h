/*
 * ppk.h -- C ABI of the B200-native humanoid ping-pong task hot path.
 *
 * Drop-in for the per-env task step of mjmj531/isaacgym's
 * tasks/humanoid_pingpong*.py (aliases as in SURVEY.md: BASE, A3, TILT, NES, A4,
 * ALIGN, ADOF).  The reference has no FFI of its own (it is Python calling ATen);
 * the boundary this library replaces is the set of VecTask methods of the task
 * classes and the free functions they call.  Each entry point below names the
 * reference interface it replaces (file:line under /root/reference/tasks/).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is DEVICE memory owned by the
 *    caller (PhysX state tensors, VecTask buffers); the library never allocates,
 *    frees or keeps state between calls (except the explicit host session below);
 *  - fp32 state/obs/reward, int64 reset/progress/env ids, int32 actor indices,
 *    bool flags as uint8 (0/1) -- the reference dtypes, never narrowed;
 *  - work is enqueued on `stream` (a cudaStream_t; NULL = legacy default stream)
 *    and no call synchronises with the host;
 *  - every function returns 0 (PPK_OK) or a negative PpkError; nothing throws or exits.
 */
#ifndef PPK_H_
#define PPK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PPK_API __attribute__((visibility("default")))
#else
#define PPK_API
#endif

#define PPK_ABI_VERSION 2
#define PPK_MAX_BODY_IDS 32   /* J: rows gathered by the observation functions        */
#define PPK_MAX_FLAGS 12      /* bool flag / counter tensors of one variant            */
#define PPK_STATS_SLOTS 64    /* stats[PPK_STATS_SLOTS][PPK_NUM_STATS] partial sums     */
#define PPK_NUM_STATS 8
#define PPK_MOMENT_SLOTS 64   /* obs_moments[PPK_MOMENT_SLOTS][2 * num_obs] partial column sums  */

typedef enum PpkError {
  PPK_OK = 0,
  PPK_ERR_NULL = -1,        /* a required pointer is NULL                               */
  PPK_ERR_SHAPE = -2,       /* sizes inconsistent with the variant                      */
  PPK_ERR_ALIGN = -3,       /* pointer not aligned to its element type                  */
  PPK_ERR_VARIANT = -4,     /* unknown variant / phase not supported by the variant     */
  PPK_ERR_LAUNCH = -5,      /* cudaGetLastError() after the launch                      */
  PPK_ERR_ABI = -6,         /* struct_size does not match this library                  */
  PPK_ERR_CUDA = -7         /* a CUDA runtime call failed (host session only)           */
} PpkError;

typedef enum PpkVariant {
  PPK_BASE = 0,   /* humanoid_pingpong.py                 HumanoidPingpong (5 actors)      */
  PPK_A3 = 1,     /* humanoid_interos_edit_pingpong_only_3_actor.py                       */
  PPK_TILT = 2,   /* humanoid_pingpong_3_actor_tilt.py                                     */
  PPK_NES = 3,    /* humanoid_pingpong_3_actor_tilt_no_earlystop.py                        */
  PPK_ALIGN = 4,  /* humanoid_pingpong_alignment.py (reward definition #1, ALIGN:1097)     */
  PPK_A4 = 5,     /* humanoid_pingpong_4_actor_tilt.py (two humanoids)                     */
  PPK_ADOF = 6,   /* humanoid_pingpong_3_actor_all_dof.py (27 DOF, imitation terms)        */
  PPK_ALIGN2 = 7  /* humanoid_pingpong_alignment.py reward definition #2 (ALIGN:1233-1351):
                     two humanoids on the A4 tensor layout, `last_hitter` state (8(f) rank 3) */
} PpkVariant;

/* Phases of post_physics_step (TILT:1022-1052); OR them together. */
typedef enum PpkPhase {
  PPK_PHASE_PROGRESS = 1,  /* progress_buf += 1                          TILT:1023        */
  PPK_PHASE_REWARD = 2,    /* compute_reward: rew_buf, reset_buf, flags  TILT:739-759     */
  PPK_PHASE_RESET = 4,     /* reset_idx(nonzero(reset_buf)), per env     TILT:1034-1036   */
  PPK_PHASE_OBS = 8,       /* compute_observations: obs_buf              TILT:770-799     */
  PPK_PHASE_STATS = 16,    /* accumulate the logged statistics           TILT:763-766     */
  PPK_PHASE_ALL = 31,
  /* not part of the task step proper (opt-in, needs PPK_PHASE_OBS): while the obs tile of
   * a CTA is still in shared memory, add its fp64 column sums and sums of squares to PpkBuffers.obs_moments, so that
   * the learner's RunningMeanStd update (normalize_input: True) never re-reads obs_buf: ppk_rms_fold_step_moments.
   * A3 / TILT / NES / ALIGN / A4 / ALIGN2. */
  PPK_PHASE_MOMENTS = 32
} PpkPhase;

/* Statistics accumulated per slot (sum over envs; divide by N on the host). */
typedef enum PpkStat {
  PPK_STAT_REWARD = 0,       /* sum rew_buf (A4: humanoid-1 reward)       TILT:764  */
  PPK_STAT_PROGRESS = 1,     /* sum progress_buf after +1, before reset   TILT:765  */
  PPK_STAT_RESETS = 2,       /* number of envs with reset_buf == 1                  */
  PPK_STAT_FALL_DOWN = 3,    /* ADOF:1164 sums of the five bool "count" tensors     */
  PPK_STAT_CLOSER = 4,
  PPK_STAT_HIT_PADDLE = 5,
  PPK_STAT_CROSS_NET = 6,
  PPK_STAT_HIT_TABLE = 7
} PpkStat;

/* Static description of one task variant: what the class constructor hard-codes or
 * reads from the YAML (TILT:63-107,125-127,168,183; ADOF:98-116; A4:125-127,157-172). */
typedef struct PpkTask {
  uint32_t struct_size;           /* = sizeof(PpkTask)                                   */
  int32_t variant;                /* PpkVariant                                          */
  int32_t num_actors;             /* A: root-state rows per env                          */
  int32_t num_bodies;             /* B: rigid-body rows per env                          */
  int32_t num_dofs;               /* D                                                   */
  int32_t humanoid_actor[2];      /* root row of humanoid 1 / 2 (x read by the rewards)  */
  int32_t ball_actor;             /* root row of the ball (BASE: ball1; ball2 = +1)      */
  int32_t paddle_body[2];         /* rigid-body row of paddle 1 / 2                      */
  int32_t pelvis_body;            /* ADOF:206                                            */
  int32_t num_body_ids;           /* J (bodyStatesId)                                    */
  int32_t body_ids[2][PPK_MAX_BODY_IDS];   /* [0] humanoid 1, [1] humanoid 2 (A4)        */
  int32_t num_balance_ids;        /* ADOF bodyStatesIdBalance                            */
  int32_t balance_ids[PPK_MAX_BODY_IDS];
  int64_t max_episode_length;     /* episodeLength                                       */
  float alpha;                    /* alphaVelocityReward                                 */
  float power_coefficient;        /* powerCoefficient                                    */
  float penalty;                  /* penalty                                             */
  float hit_table_reward;         /* hitTableReward                                      */
  float not_hit_table_penalty;    /* nothitTablePenalty                                  */
  float cross_net_reward;         /* crossNetRewardFloat                                 */
  float die_penalty;              /* diePenaltyFloat                                     */
  float hit_paddle_reward;        /* hitPaddleReward                                     */
  float miss_paddle_penalty_coefficient;
  int32_t is_train;               /* ADOF:98 (termination distance 0.32 vs 1e6)          */
  int32_t reset_dof;              /* 0 for NES (NES:871-918 leaves the DOF state alone)  */
  int32_t write_flags;            /* 1: flag updates are stored back (eager TILT/NES/ADOF);
                                     0: flags are read-only, as in the TorchScript ALIGN/A4
                                     functions where `flag |= x` is out-of-place (D16)   */
} PpkTask;

/* Caller-owned device tensors of one env shard.  Unused members may be NULL. */
typedef struct PpkBuffers {
  uint32_t struct_size;                 /* = sizeof(PpkBuffers)                              */
  int64_t num_envs;                     /* N of this shard                                   */
  /* PhysX state (gymtorch.wrap_tensor views, TILT:153-174,208-211) */
  const float* rigid_body_states;       /* [N,B,13] read                                     */
  float* root_states;                   /* [N,A,13] read; rows rewritten on reset            */
  float* dof_states;                    /* [N,D,2]  read; rows rewritten on reset            */
  const float* dof_forces;              /* [N,D]    read                                     */
  /* saved pre-step ball state (TILT:1020): row n at pre_ball_states + n*pre_ball_stride,
   * vx at [pre_vx_offset], vz at [pre_vz_offset] (a full 13-float clone uses 13/7/9)        */
  float* pre_ball_states;
  int32_t pre_ball_stride, pre_vx_offset, pre_vz_offset;
  /* reset sources */
  const float* initial_root_states;     /* [N,A,13] TILT:186                                 */
  const float* initial_dof_states;      /* [N,D,2]  TILT:214                                 */
  const float* initial_body_states;     /* [N,B,13] ADOF:200 (imitation reference pose)      */
  const float* reset_ball_vel;          /* [N,3] launch velocity env n takes when it resets  */
  const float* reset_ball_pos_yz;       /* [N,2] ADOF:976-979                                */
  /* VecTask buffers (upstream allocate_buffers dtypes) */
  float* obs_buf;                       /* [N,obs_rows,num_obs] written                      */
  float* rew_buf;                       /* [N] (A4: [N,2]) written                           */
  int64_t* reset_buf;                   /* [N] written                                       */
  int64_t* progress_buf;                /* [N] read + written                                */
  /* bool flag tensors, uint8 0/1, order per variant:
   *  TILT  0 condition_calculated 1 reward_calculated 2 no_bounce_before_half_mask (TILT:241-243)
   *  A4    0..2 humanoid 1 as TILT, 3..5 humanoid 2
   *  NES   0 paddle_condition_calculated 1 missed_ball_calculated (NES:759-760)
   *  ALIGN, ALIGN2 0 reward_calculated (ALIGN:239)
   *  ADOF  0 paddle_condition_calculated 1 hit_table_calculated 2 die_penalty_calculated
   *        3 humanoid_die_calculated 4 closer_to_paddle_count 5 hit_paddle_count
   *        6 cross_net_count 7 hit_table_count 8 fall_down_count (ADOF:279-293)             */
  uint8_t* flags[PPK_MAX_FLAGS];
  /* pre_physics_step */
  float* actions;                       /* [N,D] (clamped in place when clip_actions > 0)    */
  const float* pd_action_offset;        /* [D] TILT:666                                      */
  const float* pd_action_scale;         /* [D] TILT:667                                      */
  float* pd_targets;                    /* [N,D] written                                     */
  /* statistics: PPK_STATS_SLOTS x PPK_NUM_STATS doubles, accumulated with atomics */
  double* stats;
  /* >= 64 bytes of device scratch, zeroed once by the caller (ADOF any-env-reset flag)    */
  uint32_t* scratch;
  /* optional: where the predicated reset of ppk_post_physics_step writes the root / DOF rows of
   * resetting envs (same shapes as root_states / dof_states).  NULL = in place.  The host session
   * points them at the caller's pinned host tensors so only reset rows travel back over PCIe. */
  float* root_states_out;
  float* dof_states_out;
  /* ---- VecTask.step envelope, all optional ---------------------------------
   * clip_actions > 0: ppk_pre_physics_step clamps `actions` to +-clip_actions in place before
   *   scaling them (upstream VecTask.step clamps before calling pre_physics_step; clipActions
   *   cfg/task/HumanoidPingpongTiltG1.yaml:26).
   * timeout_buf [N] int64: 1 where the env's episode ran out this step (progress >= L-1 before
   *   the reset cleared it) -- what RL libraries need to bootstrap truncated episodes.
   * reset_count / reset_actor_indices / reset_dof_indices: the fused step appends, for every env
   *   it resets, the rows `actor_indices.view(N,A)[env]` and `dof_indices.view(N,dof_per_env)[env]`
   *   as int32 (TILT:876-877) -- the lists gym.set_actor_root_state_tensor_indexed /
   *   set_dof_state_tensor_indexed take (TILT:881-888) -- and counts the envs in *reset_count, which
   *   ppk_pre_physics_step zeroes (callers that skip the pre-step zero it themselves).  Env order
   *   within the lists is unspecified. */
  float clip_actions;
  int32_t dof_indices_per_env;
  int64_t* timeout_buf;
  const int64_t* actor_indices;         /* [N*A]   TILT:645                                   */
  const int64_t* dof_indices;           /* [N*dof_indices_per_env] TILT:646, A4:889           */
  int32_t* reset_count;                 /* [1] device                                         */
  int32_t* reset_actor_indices;         /* [N*A] capacity                                     */
  int32_t* reset_dof_indices;           /* [N*dof_indices_per_env] capacity                   */
  /* ALIGN2 only: who hit the ball last, 1 or 2 (ALIGN:1253, returned by the reward ALIGN:1349-1351;
   * read + written; a reset puts it back to its initial value 2) */
  int64_t* last_hitter;
  /* ADOF, optional: the imitation reference pose repacked once at init as [N, n_balance, 6] =
   * initial_body_states[:, balance_ids][..., (0,1,2,7,8,9)] (what ADOF:1345-1349, 1908-1909 read of it).
   * When non-NULL the step reads this instead of initial_body_states (552 B per env instead of 28
   * rigid-body rows); the caller keeps it in sync if it ever rewrites the reference pose. */
  const float* initial_balance_states;
  /* VecTask.step envelope: > 0 clamps every observation to +-clip_observations where the step kernels
   * produce it (upstream VecTask.step: obs = clamp(obs_buf, -clip_obs, clip_obs); the upstream default
   * is inf = no clamp, and no YAML of the reference sets clipObservations).  <= 0: off. */
  float clip_observations;
  /* PPK_PHASE_MOMENTS: PPK_MOMENT_SLOTS x (2 * num_obs) doubles, zero-initialised by the caller once; slot s holds
   * [sum_r obs[r, :], sum_r obs[r, :]^2] of the CTAs that hash to it (rows = envs, or (env, humanoid) pairs for A4). */
  double* obs_moments;
} PpkBuffers;

PPK_API int ppk_abi_version(void);
PPK_API const char* ppk_strerror(int code);

/* The fused task step: any OR of PpkPhase, executed per env in the reference's order
 * progress+=1 -> reward/reset/flags -> predicated reset -> observations
 * (post_physics_step TILT:1022-1052, A3:997-1014, NES, ALIGN:1022-1049, A4:1028-1060,
 * ADOF:1149-1192 incl. the clear of the five counters when any env of the shard resets).
 * For BASE the order is BASE:587-596: progress+=1 -> reset_idx(prev reset_buf) -> obs -> reward. */
PPK_API int ppk_post_physics_step(const PpkTask* task, const PpkBuffers* buf, uint32_t phases, void* stream);

/* compute_reward(actions): TILT:739, A3:720, NES:745, ALIGN:736, A4:743 (h1+h2, defect D4),
 * ADOF:802, BASE:463.  = ppk_post_physics_step(PPK_PHASE_REWARD). */
PPK_API int ppk_compute_reward(const PpkTask* task, const PpkBuffers* buf, void* stream);

/* compute_observations(): TILT:770, A4:773-803, ADOF:867, BASE:493.  = PPK_PHASE_OBS. */
PPK_API int ppk_compute_observations(const PpkTask* task, const PpkBuffers* buf, void* stream);

/* reset_idx(env_ids) / _reset_idx: TILT:809/847-906, A3:826-872, NES:871-918, A4:853-912,
 * ALIGN:845-898, ADOF:965-1028, BASE:530-579.  `ball_vel` [k,3] (BASE: [2,3], one pair for all
 * envs) and `ball_pos_yz` [k,2] (ADOF) are the host-sampled launch values for env_ids[i];
 * either may be NULL to take rows env_ids[i] of buf->reset_ball_vel / reset_ball_pos_yz.
 * `actor_indices_out` [k*A] and `dof_indices_out` [k*dof_per_env] (int32, may be NULL) receive
 * the gather TILT:876-877 of `actor_indices` [N*A] / `dof_indices` [N*dof_per_env] (int64). */
PPK_API int ppk_reset_idx(const PpkTask* task, const PpkBuffers* buf, const int64_t* env_ids, int64_t num_ids,
                  const float* ball_vel, const float* ball_pos_yz, const int64_t* actor_indices,
                  const int64_t* dof_indices, int32_t dof_indices_per_env, int32_t* actor_indices_out,
                  int32_t* dof_indices_out, void* stream);

/* pre_physics_step(actions): TILT:1002-1020 (A4 per defect D6 takes [2*7] offset/scale):
 * pd_targets = offset + scale*actions; save the ball's pre-step velocity. */
PPK_API int ppk_pre_physics_step(const PpkTask* task, const PpkBuffers* buf, void* stream);

/* Device-side form of generate_random_speed_for_ball (TILT:307-318, NES:312-323, ADOF:357-367,
 * A3:300-302) + the per-env host loop that calls it (TILT:857-862, ADOF:975-988): fills
 * buf->reset_ball_vel [N,3] (and reset_ball_pos_yz [N,2] for ADOF) from a counter-based
 * Philox4x32-10 stream keyed by `seed`, counter = (env_offset + n, epoch).  With
 * refresh_consumed_only != 0 only the rows of envs whose reset_buf is set are redrawn (call it right
 * after ppk_post_physics_step so every reset consumes a fresh draw).  Same ranges and formulas as the
 * reference; parity with its Mersenne-Twister stream is statistical, not bitwise.  Not for BASE
 * (one host-drawn velocity pair per reset_idx call, BASE:542). */
PPK_API int ppk_sample_ball_launch(const PpkTask* task, const PpkBuffers* buf, uint64_t seed, uint64_t epoch,
                                   int64_t env_offset, int32_t refresh_consumed_only, void* stream);

/* Fold the stats slots into out[PPK_NUM_STATS] (device) and zero the slots. */
PPK_API int ppk_stats_reduce(double* stats, double* out, void* stream);

/* ---- host-buffer session: the CPU-pipeline form of the same step ------------------------
 * State tensors live in HOST memory (isaacgym `use_gpu_pipeline: False`); one call copies
 * only the rows the step consumes to the device, runs the fused step and copies
 * obs/rew/reset/progress/flags (and reset rows) back.  The session owns its device staging
 * buffers and streams; `buf` members are HOST pointers (pinned for full speed). */
typedef struct PpkHostSession PpkHostSession;
PPK_API int ppk_host_session_create(const PpkTask* task, int64_t max_envs, int32_t num_chunks, PpkHostSession** out);
PPK_API int ppk_host_session_destroy(PpkHostSession* s);
/* `host_buf->reset_ball_vel` / `reset_ball_pos_yz` are per-step inputs: they are read (in place when pinned, else
 * copied per chunk) by every call that includes PPK_PHASE_RESET, so refilling them in place between calls gives every
 * reset a fresh draw (TILT:857-862).  The initial_* tensors are uploaded when their pointers or num_envs change.
 * With PPK_PHASE_STATS `host_buf->stats` is a HOST array of PPK_NUM_STATS doubles that receives the sums of this
 * call (PpkStat order; divide by N); the session's device slots are cleared by it. */
PPK_API int ppk_host_post_physics_step(PpkHostSession* s, const PpkBuffers* host_buf, uint32_t phases);
/* bytes moved by the last call */
PPK_API int ppk_host_session_traffic(const PpkHostSession* s, int64_t* h2d_bytes, int64_t* d2h_bytes);

/* ---- learner side of the path: input normalisation + first MLP layer ----
 * rl_games `normalize_input: True` (cfg/train/HumanoidPingpongTiltG1PPO.yaml:51) = RunningMeanStd
 * (rl_games/algos_torch/running_mean_std.py; un-vendored, restated in oracle/policy_oracle.py):
 * fp64 running_mean / running_var [width] and count [1], epsilon 1e-5, output clamped to +-5.
 * `clip_obs` is VecTask.step's clamp of obs_buf (clipObservations; <= 0 = none = upstream default inf).
 * `moments` is a caller-owned, zero-initialised fp64 scratch of ppk_rms_scratch_doubles(width) entries: the
 * first 2*width are the column sums and sums of squares of the batches not yet merged (what data-parallel
 * ranks all-reduce), the rest is a ticket word and the copies of the accumulators the CTAs add into. */
typedef struct PpkRunningMeanStd {
  uint32_t struct_size;
  int32_t width;
  float epsilon;
  float clip_obs;
  double* running_mean;
  double* running_var;
  double* count;
  double* moments;
} PpkRunningMeanStd;

PPK_API size_t ppk_rms_scratch_doubles(int32_t width);
/* moments += (sum_r x, sum_r x^2) over obs [rows,width].  Data-parallel ranks all-reduce (SUM) `moments`
 * and the row count before merging, so that every rank holds the statistics of the global batch. */
PPK_API int ppk_rms_accumulate(const PpkRunningMeanStd* rms, const float* obs, int64_t rows, void* stream);
/* RunningMeanStd._update_mean_var_count_from_moments with the batch mean / unbiased variance of the
 * `batch_rows` rows accumulated in `moments`; clears `moments`. */
PPK_API int ppk_rms_merge(const PpkRunningMeanStd* rms, double batch_rows, void* stream);
/* The moments a step kernel accumulated with PPK_PHASE_MOMENTS: rms->moments[0..2W) += sum over the slots of
 * `obs_moments` (which are cleared), then -- merge != 0 -- ppk_rms_merge with `batch_rows` (merge == 0: data-parallel
 * ranks all-reduce rms->moments and the row count first, then call ppk_rms_merge). */
PPK_API int ppk_rms_fold_step_moments(const PpkRunningMeanStd* rms, double* obs_moments, double batch_rows, int32_t merge,
                                      void* stream);
/* RunningMeanStd.forward in training mode, single rank: accumulate + merge. */
PPK_API int ppk_rms_update(const PpkRunningMeanStd* rms, const float* obs, int64_t rows, void* stream);
/* RunningMeanStd.forward output: out = clamp((clamp(obs) - float(mean)) / sqrt(float(var) + eps), -5, 5), fp32. */
PPK_API int ppk_rms_normalize(const PpkRunningMeanStd* rms, const float* obs, int64_t rows, float* out, void* stream);

typedef enum PpkActivation { PPK_ACT_NONE = 0, PPK_ACT_ELU = 1 } PpkActivation;

/* The first nn.Linear(width -> units) of the actor (and, with `separate: True`, critic) MLP, packed
 * once at init into fp16 tensor-core operand tiles (the bias rides in the K padding).  units % 256 == 0;
 * width <= 31, 64..95 or 304..319 (the observation widths of the task variants: 24, 80, 94, 313). */
PPK_API size_t ppk_linear_packed_bytes(int32_t units, int32_t width);
PPK_API int ppk_linear_pack(const float* weight /*[units,width]*/, const float* bias /*[units] or NULL*/, int32_t units,
                            int32_t width, void* packed, size_t packed_bytes, void* stream);
/* out[rows,units] fp16 = act(fp16(fp16(norm(obs)) @ fp16(W)^T + fp16(b))): what
 * `torch.autocast(fp16)` computes for `act(linear(running_mean_std(obs)))` (`mixed_precision: True`,
 * yaml:50; fp32 accumulation).  `rms` may be NULL (normalize_input False). */
PPK_API int ppk_policy_first_layer(const PpkRunningMeanStd* rms, const float* obs, int64_t rows, int32_t width,
                                   const void* packed, int32_t units, int32_t activation, void* out_f16, void* stream);

/* The same layer as the ROLLOUT forward computes it (rl_games get_action_values: no autocast, fp32):
 * out[rows,units] fp32 = act(norm(obs) @ W^T + b) with every product formed from TF32 hi/lo splits of both operands
 * (3 tensor-core MMAs, fp32 accumulation): within ~1e-6 * sum_k |x_k w_k| of an fp32 FMA chain.  Weights are packed
 * once into hi/lo TF32 operand tiles (8 bytes per element).  units % 256 == 0; width <= 95 (24, 80, 94). */
PPK_API size_t ppk_linear_packed_bytes_f32(int32_t units, int32_t width);
PPK_API int ppk_linear_pack_f32(const float* weight /*[units,width]*/, const float* bias /*[units] or NULL*/, int32_t units,
                                int32_t width, void* packed, size_t packed_bytes, void* stream);
PPK_API int ppk_policy_first_layer_f32(const PpkRunningMeanStd* rms, const float* obs, int64_t rows, int32_t width,
                                       const void* packed, int32_t units, int32_t activation, float* out_f32, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PPK_H_ */
