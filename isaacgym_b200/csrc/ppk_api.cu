// C ABI (include/ppk.h) over the sm_100a kernels.  No torch types, no allocation, no host sync.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ppk.h"
#include "ppk_device.cuh"
#include "ppk_family.cuh"
#include "ppk_misc.cuh"
#include "ppk_adof.cuh"
#include "ppk_adof2.cuh"
#include "ppk_policy.cuh"
#include "ppk_policy_f32.cuh"

using namespace ppk;

namespace {

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int fill_args(const PpkTask* t, const PpkBuffers* b, uint32_t phases, KArgs* k) {
  if (t == nullptr || b == nullptr) return PPK_ERR_NULL;
  if (t->struct_size != sizeof(PpkTask) || b->struct_size != sizeof(PpkBuffers)) return PPK_ERR_ABI;
  if (t->variant < PPK_BASE || t->variant > PPK_ALIGN2) return PPK_ERR_VARIANT;
  if (b->num_envs < 0 || t->num_actors <= 0 || t->num_bodies <= 0 || t->num_dofs <= 0) return PPK_ERR_SHAPE;
  if (t->num_body_ids < 0 || t->num_body_ids > PPK_MAX_BODY_IDS || t->num_balance_ids < 0 ||
      t->num_balance_ids > PPK_MAX_BODY_IDS)
    return PPK_ERR_SHAPE;
  memset(k, 0, sizeof(*k));
  k->rb = b->rigid_body_states; k->root = b->root_states; k->dof = b->dof_states; k->force = b->dof_forces;
  k->last_hitter = reinterpret_cast<long long*>(b->last_hitter);
  k->init_bal = b->initial_balance_states;
  k->timeout = reinterpret_cast<long long*>(b->timeout_buf);
  if (b->reset_count != nullptr) {
    if (!b->actor_indices || !b->reset_actor_indices || b->dof_indices_per_env < 0 ||
        (b->dof_indices_per_env > 0 && (!b->dof_indices || !b->reset_dof_indices)))
      return PPK_ERR_NULL;
    k->actor_idx = reinterpret_cast<const long long*>(b->actor_indices);
    k->dof_idx = reinterpret_cast<const long long*>(b->dof_indices);
    k->dof_per_env = b->dof_indices_per_env;
    k->reset_count = b->reset_count;
    k->reset_actor_out = b->reset_actor_indices;
    k->reset_dof_out = b->reset_dof_indices;
  }
  k->root_out = b->root_states_out ? b->root_states_out : b->root_states;
  k->dof_out = b->dof_states_out ? b->dof_states_out : b->dof_states;
  k->pre = b->pre_ball_states; k->init_root = b->initial_root_states; k->init_dof = b->initial_dof_states;
  k->init_rb = b->initial_body_states; k->reset_vel = b->reset_ball_vel; k->reset_yz = b->reset_ball_pos_yz;
  k->obs = b->obs_buf; k->rew = b->rew_buf;
  k->reset = reinterpret_cast<long long*>(b->reset_buf);
  k->progress = reinterpret_cast<long long*>(b->progress_buf);
  for (int i = 0; i < PPK_MAX_FLAGS; ++i) k->flags[i] = b->flags[i];
  k->stats = b->stats; k->scratch = b->scratch;
  k->n = b->num_envs; k->max_len = t->max_episode_length;
  k->pre_stride = b->pre_ball_stride; k->pre_vx = b->pre_vx_offset; k->pre_vz = b->pre_vz_offset;
  k->A = t->num_actors; k->B = t->num_bodies; k->D = t->num_dofs;
  for (int h = 0; h < 2; ++h) {
    k->hum[h] = t->humanoid_actor[h];
    k->paddle_body[h] = t->paddle_body[h];
    k->paddle_j[h] = -1;
    for (int j = 0; j < t->num_body_ids; ++j) {
      int id = t->body_ids[h][j];
      if (h == 0 || t->variant == PPK_A4 || t->variant == PPK_ALIGN2) {
        if (id < 0 || id >= t->num_bodies) return PPK_ERR_SHAPE;
      }
      k->ids[h][j] = id;
      if (id == t->paddle_body[h]) k->paddle_j[h] = j;
    }
    if (t->paddle_body[h] < 0 || t->paddle_body[h] >= t->num_bodies) return PPK_ERR_SHAPE;
    if (t->humanoid_actor[h] < 0 || t->humanoid_actor[h] >= t->num_actors) return PPK_ERR_SHAPE;
  }
  for (int j = 0; j < t->num_balance_ids; ++j) {
    if (t->balance_ids[j] < 0 || t->balance_ids[j] >= t->num_bodies) return PPK_ERR_SHAPE;
    k->bal_ids[j] = t->balance_ids[j];
  }
  k->ball = t->ball_actor; k->pelvis_body = t->pelvis_body;
  if (t->ball_actor < 0 || t->ball_actor >= t->num_actors) return PPK_ERR_SHAPE;
  if (t->pelvis_body < 0 || t->pelvis_body >= t->num_bodies) return PPK_ERR_SHAPE;
  k->alpha = t->alpha; k->power_coef = t->power_coefficient; k->penalty = t->penalty;
  k->hit_table = t->hit_table_reward; k->not_hit = t->not_hit_table_penalty; k->cross_net = t->cross_net_reward;
  k->die_penalty = t->die_penalty; k->hit_paddle = t->hit_paddle_reward; k->miss_coef = t->miss_paddle_penalty_coefficient;
  k->term_dist = t->is_train ? 0.32f : 1e6f;           // ADOF:1404-1410, is_g1 branch
  k->phases = (int)phases; k->write_flags = t->write_flags; k->reset_dof = t->reset_dof;
  k->clip_obs = b->clip_observations;
  k->moments = b->obs_moments;
  // Tensor-map staging (family kernel): the rigid-body tensor is read as [N/2 env pairs, 2*B*13 floats], legal when
  // the tensors start 16-byte aligned, an env pair is a multiple of 16 bytes (B even), ids[1..J) are consecutive rows
  // (one run per env).  The boxes over-read <= 12 bytes on either side of a run; what falls outside a pair's row of
  // the tensor is zero-filled by the engine, never fetched.  Bit 0: loads, bit 1: the obs tile may be written back
  // with one bulk copy.
  bool bulk = t->num_body_ids >= 2 && b->rigid_body_states && b->root_states && b->dof_states && b->dof_forces &&
              (t->num_bodies % 2 == 0) && b->num_envs >= 2;
  const int nh = (t->variant == PPK_A4 || t->variant == PPK_ALIGN2) ? 2 : 1;
  if (bulk) {
    const void* al[] = {b->rigid_body_states, b->root_states, b->dof_states, b->dof_forces};
    for (const void* p : al) bulk = bulk && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0);
    for (int h = 0; h < nh && bulk; ++h) {
      for (int j = 2; j < t->num_body_ids; ++j) bulk = bulk && (t->body_ids[h][j] == t->body_ids[h][1] + j - 1);
    }
  }
  k->bulk_ok = bulk ? 1 : 0;
  if (bulk && b->obs_buf && (reinterpret_cast<uintptr_t>(b->obs_buf) & 15u) == 0) k->bulk_ok |= 2;
  k->row0_box = kRow0Max;
  if (bulk) {
    int max_off = 0;
    for (int h = 0; h < nh; ++h)
      for (int q = 0; q < 2; ++q) {
        const int s0 = q * t->num_bodies * kRow + t->body_ids[h][1] * kRow;
        const int r0 = q * t->num_bodies * kRow + t->body_ids[h][0] * kRow;
        k->span_c[h][q] = s0 & ~3; k->span_off[h][q] = s0 & 3;
        k->row0_c[h][q] = r0 & ~3; k->row0_off[h][q] = r0 & 3;
        if ((r0 & 3) > max_off) max_off = r0 & 3;
      }
    k->row0_box = (max_off + 10 <= 12) ? 12 : 16;
  }
  return PPK_OK;
}

inline bool misaligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) != 0; }

int num_flags_of(int variant) {
  switch (variant) {
    case PPK_TILT: return 3;
    case PPK_A4: return 6;
    case PPK_NES: return 2;
    case PPK_ALIGN: return 1;
    case PPK_ALIGN2: return 1;
    case PPK_ADOF: return 9;
    default: return 0;
  }
}

// Pointers a given phase set dereferences; NULL or misaligned ones are rejected before the launch.
int check_step_pointers(const PpkTask* t, const PpkBuffers* b, uint32_t phases) {
  const int v = t->variant;
  const bool rew = phases & PPK_PHASE_REWARD, rst = phases & PPK_PHASE_RESET, obs = phases & PPK_PHASE_OBS;
  if (!b->rigid_body_states || !b->root_states || !b->dof_states || !b->progress_buf || !b->reset_buf) return PPK_ERR_NULL;
  if (v != PPK_BASE && !b->dof_forces) return PPK_ERR_NULL;
  if (rew && !b->rew_buf) return PPK_ERR_NULL;
  if (rew && v != PPK_BASE && !b->pre_ball_states) return PPK_ERR_NULL;
  if (obs && !b->obs_buf) return PPK_ERR_NULL;
  if (rst && (!b->initial_root_states || !b->reset_ball_vel)) return PPK_ERR_NULL;
  if (rst && t->reset_dof && !b->initial_dof_states) return PPK_ERR_NULL;
  if (v == PPK_ADOF) {
    if ((!b->initial_body_states && !b->initial_balance_states) || !b->initial_dof_states) return PPK_ERR_NULL;
    if (rst && (!b->reset_ball_pos_yz || !b->scratch)) return PPK_ERR_NULL;
  }
  if (v == PPK_ALIGN2 && (rew || rst) && !b->last_hitter) return PPK_ERR_NULL;
  if ((phases & PPK_PHASE_STATS) && !b->stats) return PPK_ERR_NULL;
  if (rew || rst)
    for (int i = 0; i < num_flags_of(v); ++i)
      if (!b->flags[i]) return PPK_ERR_NULL;
  const void* f32[] = {b->rigid_body_states, b->root_states, b->dof_states, b->dof_forces, b->pre_ball_states,
                       b->initial_root_states, b->initial_dof_states, b->initial_body_states, b->reset_ball_vel,
                       b->reset_ball_pos_yz, b->obs_buf, b->rew_buf};
  for (const void* p : f32)
    if (p && misaligned(p, 4)) return PPK_ERR_ALIGN;
  if (misaligned(b->progress_buf, 8) || misaligned(b->reset_buf, 8)) return PPK_ERR_ALIGN;
  if (b->stats && misaligned(b->stats, 8)) return PPK_ERR_ALIGN;
  if (rew && v != PPK_BASE && (b->pre_ball_stride <= 0 || b->pre_vx_offset < 0 || b->pre_vx_offset >= b->pre_ball_stride))
    return PPK_ERR_SHAPE;
  if (rew && (v == PPK_ALIGN || v == PPK_ALIGN2) && (b->pre_vz_offset < 0 || b->pre_vz_offset >= b->pre_ball_stride)) return PPK_ERR_SHAPE;
  return PPK_OK;
}

// Soft-start delay per first-wave slot, in SM cycles (profiles/r2_staging_probe.md): the time DRAM needs to serve one
// tile's bytes to one SM (~1 cycle per 26 bytes of a CTA's shared memory: 1400 cycles for the 36 KB tiles; the 18 KB
// tiles, twice as many slots per SM, measured best at 400); PPK_STAGGER overrides it (A/B runs)
int stagger_cycles(size_t smem_bytes) {
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("PPK_STAGGER");
    forced = e ? atoi(e) : -1;
  }
  if (forced >= 0) return forced;
  return smem_bytes > 24 * 1024 ? (int)(smem_bytes / 26) : (int)(smem_bytes / 46);
}

// the rigid-body tensor as [N/2 env pairs, 2*B*13 floats]; box = `box_floats` x `pairs` pairs
int make_rb_map(CUtensorMap* m, const float* rb, long long n, int num_bodies, int box_floats, int pairs) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return PPK_ERR_CUDA;
  const cuuint64_t dims[2] = {(cuuint64_t)2 * num_bodies * kRow, (cuuint64_t)(n / 2)};
  const cuuint64_t strides[1] = {(cuuint64_t)2 * num_bodies * kRow * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_floats, (cuuint32_t)pairs};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(rb), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_64B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? PPK_OK
             : PPK_ERR_CUDA;
}

// Up to ~2.5 waves of 32-env tiles (6 CTAs per SM) the step is latency- and tail-bound: 16-env tiles win (A3 16 384 envs
// 7.3 -> 7.1 us, TILT 65 536 envs 15.55 -> 15.3 us); beyond, the full-lane 32-env tiles do (1 M envs: 186 vs 193 us).
// PPK_SMALL_TILES=0/1 forces the choice (A/B runs, tests)
bool small_batch(long long n) {
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("PPK_SMALL_TILES");
    forced = e ? atoi(e) : -1;
  }
  if (forced >= 0) return forced != 0;
  return (n + 31) / 32 <= 21LL * sm_count();     // crossover measured between 98 304 (16-env tiles 1 % ahead) and 131 072 envs (32-env tiles 1 % ahead)
}

template <int V, int H, int J, int D, int A, int TILE>
int launch_family(const KArgs& k0, cudaStream_t s) {
  using L = FamilyLayout<H, J, D, A, TILE>;
  auto kern = family_step_kernel<V, H, J, D, A, TILE>;
  constexpr size_t smem = (size_t)L::kFloats * sizeof(float);
  static SmemOptIn opt;
  if (!opt.ensure(kern, smem)) return PPK_ERR_LAUNCH;
  KArgs k = k0;
  CUtensorMap m_span, m_row0;
  memset(&m_span, 0, sizeof(m_span));
  memset(&m_row0, 0, sizeof(m_row0));
  const long long tiles = (k.n + TILE - 1) / TILE;      // one CTA per tile
  if (k.bulk_ok & 1) {
    if (make_rb_map(&m_span, k.rb, k.n, k.B, L::kSpanF, L::kPairs) != PPK_OK ||
        make_rb_map(&m_row0, k.rb, k.n, k.B, k.row0_box, L::kPairs) != PPK_OK)
      k.bulk_ok = 0;                                     // no driver entry point: the LDG path still works
  }
  // soft start: only when the grid does not fit the GPU at once
  static int occ = 0, sms = 0;
  if (occ == 0) {
    int dev = 0, o = 0, n_sm = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kFamilyThreads, smem) == cudaSuccess &&
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && o > 0 && n_sm > 0) {
      sms = n_sm;
      occ = o;
    } else {
      cudaGetLastError();
    }
  }
  k.num_sms = sms > 0 ? sms : 1;
  k.first_wave = occ * sms;
  k.stagger = (occ > 0 && tiles > (long long)occ * sms) ? stagger_cycles(smem) : 0;
  if (launch_pdl(kern, (unsigned)tiles, kFamilyThreads, smem, s, k, m_span, m_row0) != cudaSuccess) {
    cudaGetLastError();
    return PPK_ERR_LAUNCH;
  }
  return PPK_OK;
}

}  // namespace

extern "C" {

int ppk_abi_version(void) { return PPK_ABI_VERSION; }

const char* ppk_strerror(int code) {
  switch (code) {
    case PPK_OK: return "ok";
    case PPK_ERR_NULL: return "a required device pointer is NULL";
    case PPK_ERR_SHAPE: return "sizes are inconsistent with the task variant";
    case PPK_ERR_ALIGN: return "a pointer is not aligned to its element type";
    case PPK_ERR_VARIANT: return "unknown task variant or unsupported phase";
    case PPK_ERR_LAUNCH: return "CUDA kernel launch failed";
    case PPK_ERR_ABI: return "struct_size mismatch: header and library disagree";
    case PPK_ERR_CUDA: return "a CUDA runtime call failed";
    default: return "unknown ppk error";
  }
}

int ppk_post_physics_step(const PpkTask* t, const PpkBuffers* b, uint32_t phases, void* stream) {
  KArgs k;
  int rc = fill_args(t, b, phases, &k);
  if (rc != PPK_OK) return rc;
  if ((phases & ~((uint32_t)PPK_PHASE_ALL | PPK_PHASE_MOMENTS | kPhaseDeferCounterClear)) != 0 || (phases & PPK_PHASE_ALL) == 0)
    return PPK_ERR_VARIANT;
  if (phases & PPK_PHASE_MOMENTS) {
    if (t->variant == PPK_BASE || t->variant == PPK_ADOF || !(phases & PPK_PHASE_OBS)) return PPK_ERR_VARIANT;
    if (!b->obs_moments) return PPK_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(b->obs_moments) & 7u) return PPK_ERR_ALIGN;
  }
  if (b->num_envs == 0) return PPK_OK;      // empty shard: nothing to dereference
  rc = check_step_pointers(t, b, phases);
  if (rc != PPK_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (t->variant) {
    case PPK_BASE: {
      if (t->num_actors < 5 || t->ball_actor + 1 >= t->num_actors) return PPK_ERR_SHAPE;
      const long long blocks = ((b->num_envs + 31) / 32 + kBaseWarps - 1) / kBaseWarps;
      base_step_kernel<<<(unsigned)blocks, kBaseWarps * 32, 0, s>>>(k);
      return cudaGetLastError() == cudaSuccess ? PPK_OK : PPK_ERR_LAUNCH;
    }
    case PPK_A3:
    case PPK_TILT:
    case PPK_NES:
    case PPK_ALIGN:
      if (t->num_actors != 3 || t->num_dofs != 7 || t->num_body_ids != 10) return PPK_ERR_SHAPE;
      // 32-env tiles use every lane of the lane = env warps; a batch that does not even fill one wave of them is
      // latency-bound instead: 16-env tiles put twice as many CTAs to work and halve the rotation passes per CTA
      if (small_batch(b->num_envs)) {
        if (t->variant == PPK_A3) return launch_family<PPK_A3, 1, 10, 7, 3, 16>(k, s);
        if (t->variant == PPK_TILT) return launch_family<PPK_TILT, 1, 10, 7, 3, 16>(k, s);
        if (t->variant == PPK_NES) return launch_family<PPK_NES, 1, 10, 7, 3, 16>(k, s);
        return launch_family<PPK_ALIGN, 1, 10, 7, 3, 16>(k, s);
      }
      if (t->variant == PPK_A3) return launch_family<PPK_A3, 1, 10, 7, 3, 32>(k, s);
      if (t->variant == PPK_TILT) return launch_family<PPK_TILT, 1, 10, 7, 3, 32>(k, s);
      if (t->variant == PPK_NES) return launch_family<PPK_NES, 1, 10, 7, 3, 32>(k, s);
      return launch_family<PPK_ALIGN, 1, 10, 7, 3, 32>(k, s);
    case PPK_A4:
      if (t->num_actors != 4 || t->num_dofs != 14 || t->num_body_ids != 10) return PPK_ERR_SHAPE;
      return launch_family<PPK_A4, 2, 10, 14, 4, 16>(k, s);
    case PPK_ALIGN2:
      if (t->num_actors != 4 || t->num_dofs != 14 || t->num_body_ids != 10) return PPK_ERR_SHAPE;
      return launch_family<PPK_ALIGN2, 2, 10, 14, 4, 16>(k, s);
    case PPK_ADOF:
      if (t->num_actors != 3 || t->num_dofs != 27 || t->num_body_ids != 10 || t->num_balance_ids != 23) return PPK_ERR_SHAPE;
      return launch_adof(k, s);
    default:
      return PPK_ERR_VARIANT;
  }
}

int ppk_compute_reward(const PpkTask* t, const PpkBuffers* b, void* stream) {
  return ppk_post_physics_step(t, b, PPK_PHASE_REWARD, stream);
}

int ppk_compute_observations(const PpkTask* t, const PpkBuffers* b, void* stream) {
  return ppk_post_physics_step(t, b, PPK_PHASE_OBS, stream);
}

int ppk_reset_idx(const PpkTask* t, const PpkBuffers* b, const int64_t* env_ids, int64_t num_ids, const float* ball_vel,
                  const float* ball_pos_yz, const int64_t* actor_indices, const int64_t* dof_indices,
                  int32_t dof_indices_per_env, int32_t* actor_indices_out, int32_t* dof_indices_out, void* stream) {
  KArgs k;
  int rc = fill_args(t, b, PPK_PHASE_RESET, &k);
  if (rc != PPK_OK) return rc;
  if (num_ids < 0) return PPK_ERR_SHAPE;
  if (num_ids == 0) return PPK_OK;
  if (!env_ids || !b->root_states || !b->initial_root_states || !b->progress_buf) return PPK_ERR_NULL;
  if (t->reset_dof && (!b->dof_states || !b->initial_dof_states)) return PPK_ERR_NULL;
  if (!ball_vel && !b->reset_ball_vel) return PPK_ERR_NULL;
  if (t->variant == PPK_ADOF && !ball_pos_yz && !b->reset_ball_pos_yz) return PPK_ERR_NULL;
  if (t->variant == PPK_BASE && (!ball_vel || !b->reset_buf)) return PPK_ERR_NULL;
  if (actor_indices_out && !actor_indices) return PPK_ERR_NULL;
  if (dof_indices_out && (!dof_indices || dof_indices_per_env <= 0 || dof_indices_per_env > 32)) return PPK_ERR_SHAPE;
  if (t->num_actors > 32) return PPK_ERR_SHAPE;
  if (misaligned(env_ids, 8)) return PPK_ERR_ALIGN;
  ResetArgs r;
  memset(&r, 0, sizeof(r));
  r.env_ids = reinterpret_cast<const long long*>(env_ids); r.num_ids = num_ids;
  r.ball_vel = ball_vel; r.ball_yz = ball_pos_yz;
  r.actor_indices = reinterpret_cast<const long long*>(actor_indices);
  r.dof_indices = reinterpret_cast<const long long*>(dof_indices);
  r.dof_per_env = dof_indices_out ? dof_indices_per_env : 0;
  r.actor_out = actor_indices_out; r.dof_out = dof_indices_out;
  r.variant = t->variant;
  r.last_hitter = reinterpret_cast<long long*>(b->last_hitter);
  if (t->variant == PPK_ALIGN2 && !b->last_hitter) return PPK_ERR_NULL;
  // flags written by _reset_idx: TILT:902-905, NES:913-917, ALIGN:897, A4:908-911, ADOF:1023-1026
  // (ADOF resets only its four *_calculated flags; the five counters are cleared elsewhere)
  int nf = num_flags_of(t->variant);
  if (t->variant == PPK_ADOF) nf = 4;
  r.num_flags = nf;
  for (int i = 0; i < nf; ++i) {
    if (!b->flags[i]) return PPK_ERR_NULL;
    r.flag_reset_value[i] = ((t->variant == PPK_TILT || t->variant == PPK_A4) && (i % 3) == 2) ? 1 : 0;
  }
  const int warps = 4;
  const long long blocks = (num_ids + warps - 1) / warps;
  reset_idx_kernel<<<(unsigned)blocks, warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(k, r);
  return cudaGetLastError() == cudaSuccess ? PPK_OK : PPK_ERR_LAUNCH;
}

int ppk_pre_physics_step(const PpkTask* t, const PpkBuffers* b, void* stream) {
  if (t == nullptr || b == nullptr) return PPK_ERR_NULL;
  if (t->struct_size != sizeof(PpkTask) || b->struct_size != sizeof(PpkBuffers)) return PPK_ERR_ABI;
  if (!b->actions || !b->pd_action_offset || !b->pd_action_scale || !b->pd_targets) return PPK_ERR_NULL;
  const bool save_ball = t->variant != PPK_BASE;          // BASE:581-585 has no ball clone
  if (save_ball && (!b->pre_ball_states || !b->root_states)) return PPK_ERR_NULL;
  if (save_ball && (b->pre_ball_stride <= 0 || b->pre_vx_offset < 0 || b->pre_vx_offset >= b->pre_ball_stride ||
                    b->pre_vz_offset >= b->pre_ball_stride))
    return PPK_ERR_SHAPE;
  if (b->num_envs < 0 || t->num_dofs <= 0) return PPK_ERR_SHAPE;
  if (b->num_envs == 0) return PPK_OK;
  const long long total = b->num_envs * t->num_dofs;
  const int vec = ((reinterpret_cast<uintptr_t>(b->actions) | reinterpret_cast<uintptr_t>(b->pd_targets)) & 15u) == 0;
  const long long per_block = vec ? 1024 : 256;         // elements of the [N, D] tensors per CTA and pass
  long long action_blocks = (total + per_block - 1) / per_block;
  if (action_blocks > sm_count() * 32) action_blocks = sm_count() * 32;
  long long ball_blocks = save_ball ? (b->num_envs + 255) / 256 : 0;       // one thread per env: every scattered read in flight at once
  if (ball_blocks > sm_count() * 64) ball_blocks = sm_count() * 64;
  pre_step_kernel<<<(unsigned)(action_blocks + ball_blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      b->actions, b->clip_actions, b->pd_action_offset, b->pd_action_scale, b->pd_targets, b->num_envs, t->num_dofs, b->root_states,
      t->num_actors * kRow, t->ball_actor, save_ball ? b->pre_ball_states : nullptr, b->pre_ball_stride,
      b->pre_vx_offset, b->pre_vz_offset, b->reset_count, (unsigned)action_blocks, vec);
  return cudaGetLastError() == cudaSuccess ? PPK_OK : PPK_ERR_LAUNCH;
}

int ppk_sample_ball_launch(const PpkTask* t, const PpkBuffers* b, uint64_t seed, uint64_t epoch, int64_t env_offset,
                           int32_t refresh_consumed_only, void* stream) {
  if (t == nullptr || b == nullptr) return PPK_ERR_NULL;
  if (t->struct_size != sizeof(PpkTask) || b->struct_size != sizeof(PpkBuffers)) return PPK_ERR_ABI;
  if (t->variant <= PPK_BASE || t->variant > PPK_ALIGN2) return PPK_ERR_VARIANT;
  if (b->num_envs < 0 || env_offset < 0) return PPK_ERR_SHAPE;
  if (b->num_envs == 0) return PPK_OK;
  if (!b->reset_ball_vel || (t->variant == PPK_ADOF && !b->reset_ball_pos_yz)) return PPK_ERR_NULL;
  if (refresh_consumed_only && !b->reset_buf) return PPK_ERR_NULL;
  const long long blocks = (b->num_envs + 255) / 256;
  // a programmatic dependent of the step kernel it usually follows: resident when that grid retires (it waits for it
  // before its first access: the step reads the rows this kernel rewrites)
  if (launch_pdl(sample_launch_kernel, (unsigned)blocks, 256u, 0, static_cast<cudaStream_t>(stream),
                 const_cast<float*>(b->reset_ball_vel), const_cast<float*>(b->reset_ball_pos_yz),
                 refresh_consumed_only ? reinterpret_cast<const long long*>(b->reset_buf) : nullptr, (long long)b->num_envs,
                 (int)t->variant, (unsigned long long)seed, (unsigned long long)epoch, (long long)env_offset) != cudaSuccess) {
    cudaGetLastError();
    return PPK_ERR_LAUNCH;
  }
  return PPK_OK;
}

// Used by the host session: the shard-wide ADOF counter clear after all chunks ran (ADOF:1162-1175).
int ppk_internal_adof_clear(const PpkBuffers* b, void* stream) {
  if (!b || !b->scratch) return PPK_ERR_NULL;
  for (int i = 4; i < 9; ++i)
    if (!b->flags[i]) return PPK_ERR_NULL;
  return launch_adof_clear(b->scratch, b->flags, b->num_envs, static_cast<cudaStream_t>(stream));
}

#ifdef PPK_TRACE
// debug builds only (-DPPK_TRACE): device buffer of >= blocks*8*4 uint64 receiving timeline stamps
PPK_API int ppk_debug_set_trace(unsigned long long* buf) {
  return cudaMemcpyToSymbol(ppk::g_trace, &buf, sizeof(buf)) == cudaSuccess ? PPK_OK : PPK_ERR_CUDA;
}
#endif

namespace {
int rms_args(const PpkRunningMeanStd* r, RmsArgs* a) {
  if (r == nullptr) return PPK_ERR_NULL;
  if (r->struct_size != sizeof(PpkRunningMeanStd)) return PPK_ERR_ABI;
  if (r->width <= 0 || r->width > 512) return PPK_ERR_SHAPE;
  if (!r->running_mean || !r->running_var || !r->count) return PPK_ERR_NULL;
  a->width = r->width; a->eps = r->epsilon; a->clip = r->clip_obs;
  a->mean = r->running_mean; a->var = r->running_var; a->count = r->count;
  return PPK_OK;
}
}  // namespace

namespace {
// thread layout of the rms kernels: x = column group (4 columns when the rows are 16-byte aligned), y = rows
int rms_launch_moments(const PpkRunningMeanStd* rms, const RmsArgs& a, const float* obs, int64_t rows, int merge,
                       cudaStream_t s) {
  const bool vec = (a.width % 4 == 0) && ((reinterpret_cast<uintptr_t>(obs) & 15u) == 0);
  const int wx = vec ? a.width / 4 : a.width;
  int R = 512 / wx;
  if (R < 1) R = 1;
  if (R < (vec ? 8 : 2)) R = (wx * (vec ? 8 : 2) <= 1024) ? (vec ? 8 : 2) : R;     // the fold needs 2*VEC row slots
  long long blocks = (rows + (long long)R * 4 - 1) / ((long long)R * 4);
  if (blocks > 2 * sm_count()) blocks = 2 * sm_count();   // two CTAs per SM: the whole batch is in flight in one pass
  if (blocks > kRmsMaxCtas) blocks = kRmsMaxCtas;
  if (blocks < 1) blocks = 1;
  const size_t smem = sizeof(double) * 2 * (vec ? 4 : 1) * R * wx;
  if (vec) {
    if (R < 8) return PPK_ERR_SHAPE;
    static SmemOptIn opt;
    if (!opt.ensure(rms_moments_kernel<4>, 96 * 1024)) return PPK_ERR_LAUNCH;
    rms_moments_kernel<4><<<(unsigned)blocks, dim3(wx, R), smem, s>>>(a, obs, rows, rms->moments, merge, (double)rows);
  } else {
    if (R < 2) return PPK_ERR_SHAPE;
    rms_moments_kernel<1><<<(unsigned)blocks, dim3(wx, R), smem, s>>>(a, obs, rows, rms->moments, merge, (double)rows);
  }
  return cudaGetLastError() == cudaSuccess ? PPK_OK : PPK_ERR_LAUNCH;
}
}  // namespace

int ppk_rms_accumulate(const PpkRunningMeanStd* rms, const float* obs, int64_t rows, void* stream) {
  RmsArgs a;
  int rc = rms_args(rms, &a);
  if (rc != PPK_OK) return rc;
  if (rows < 0) return PPK_ERR_SHAPE;
  if (rows == 0) return PPK_OK;
  if (!obs || !rms->moments) return PPK_ERR_NULL;
  return rms_launch_moments(rms, a, obs, rows, 0, static_cast<cudaStream_t>(stream));
}

int ppk_rms_merge(const PpkRunningMeanStd* rms, double batch_rows, void* stream) {
  RmsArgs a;
  int rc = rms_args(rms, &a);
  if (rc != PPK_OK) return rc;
  if (!rms->moments) return PPK_ERR_NULL;
  if (!(batch_rows > 0.0)) return PPK_ERR_SHAPE;
  rms_merge_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, rms->moments, batch_rows);
  return cudaGetLastError() == cudaSuccess ? PPK_OK : PPK_ERR_LAUNCH;
}

int ppk_rms_fold_step_moments(const PpkRunningMeanStd* rms, double* obs_moments, double batch_rows, int32_t merge, void* stream) {
  RmsArgs a;
  int rc = rms_args(rms, &a);
  if (rc != PPK_OK) return rc;
  if (!rms->moments || !obs_moments) return PPK_ERR_NULL;
  if (merge && !(batch_rows > 0.0)) return PPK_ERR_SHAPE;
  if (launch_pdl(rms_fold_step_kernel, 1u, 1024u, 0, static_cast<cudaStream_t>(stream), a, rms->moments, obs_moments, batch_rows,
                 (int)merge) != cudaSuccess) {
    cudaGetLastError();
    return PPK_ERR_LAUNCH;
  }
  return PPK_OK;
}

int ppk_rms_update(const PpkRunningMeanStd* rms, const float* obs, int64_t rows, void* stream) {
  RmsArgs a;
  int rc = rms_args(rms, &a);
  if (rc != PPK_OK) return rc;
  if (rows < 0) return PPK_ERR_SHAPE;
  if (rows == 0) return PPK_OK;
  if (!obs || !rms->moments) return PPK_ERR_NULL;
  return rms_launch_moments(rms, a, obs, rows, 1, static_cast<cudaStream_t>(stream));     // moments + merge, one launch
}

int ppk_rms_normalize(const PpkRunningMeanStd* rms, const float* obs, int64_t rows, float* out, void* stream) {
  RmsArgs a;
  int rc = rms_args(rms, &a);
  if (rc != PPK_OK) return rc;
  if (rows < 0) return PPK_ERR_SHAPE;
  if (rows == 0) return PPK_OK;
  if (!obs || !out) return PPK_ERR_NULL;
  const bool vec = (a.width % 4 == 0) && (((reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0);
  const int wx = vec ? a.width / 4 : a.width;
  int R = 256 / wx;
  if (R < 1) R = 1;
  long long blocks = (rows + R - 1) / R;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (vec) rms_apply_kernel<4><<<(unsigned)blocks, dim3(wx, R), 0, s>>>(a, obs, rows, out);
  else rms_apply_kernel<1><<<(unsigned)blocks, dim3(wx, R), 0, s>>>(a, obs, rows, out);
  return cudaGetLastError() == cudaSuccess ? PPK_OK : PPK_ERR_LAUNCH;
}

size_t ppk_rms_scratch_doubles(int32_t width) {
  if (width <= 0 || width > 512) return 0;
  return (size_t)2 * width * (1 + kRmsSlots) + 1;
}

size_t ppk_linear_packed_bytes(int32_t units, int32_t width) {
  if (units <= 0 || width <= 0 || units % kFlN != 0) return 0;
  return (size_t)units * fl_kpad(width) * sizeof(__half);
}

int ppk_linear_pack(const float* weight, const float* bias, int32_t units, int32_t width, void* packed, size_t packed_bytes,
                    void* stream) {
  if (!weight || !packed) return PPK_ERR_NULL;
  const size_t need = ppk_linear_packed_bytes(units, width);
  if (need == 0 || packed_bytes < need) return PPK_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(packed) & 15u) return PPK_ERR_ALIGN;
  const int kp = fl_kpad(width);
  linear_pack_kernel<<<sm_count() * 2, 256, 0, static_cast<cudaStream_t>(stream)>>>(weight, bias, units, width, kp,
                                                                              static_cast<__half*>(packed));
  return cudaGetLastError() == cudaSuccess ? PPK_OK : PPK_ERR_LAUNCH;
}

extern "C++" {
namespace {
// out [rows, units] fp16 as a 2-D tensor; box = one epilogue warp's [32 rows x 64 units] tile, 128B swizzle
int make_out_map(CUtensorMap* m, void* out, long long rows, int units) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return PPK_ERR_CUDA;
  const cuuint64_t dims[2] = {(cuuint64_t)units, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)units * sizeof(__half)};
  const cuuint32_t box[2] = {(cuuint32_t)kFlEpiCols, 32};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? PPK_OK
             : PPK_ERR_CUDA;
}

template <int KP, int KBLK>
int launch_first_layer(const FlArgs& k, cudaStream_t s) {
  using L = FlLayout<KP, KBLK>;
  CUtensorMap out_map;
  int rc = make_out_map(&out_map, k.out, k.rows, k.units);
  if (rc != PPK_OK) return rc;
  static SmemOptIn opt0, opt1;
  if (!opt0.ensure(first_layer_kernel<KP, KBLK, 0>, L::kBytes) || !opt1.ensure(first_layer_kernel<KP, KBLK, 1>, L::kBytes))
    return PPK_ERR_LAUNCH;
  const long long units = ((k.rows + kFlM - 1) / kFlM) * (k.units / kFlN);
  const unsigned grid = (unsigned)(units < sm_count() ? units : sm_count());      // persistent: one CTA per SM
  if (k.activation == PPK_ACT_ELU) first_layer_kernel<KP, KBLK, 1><<<grid, kFlThreads, L::kBytes, s>>>(k, out_map);
  else first_layer_kernel<KP, KBLK, 0><<<grid, kFlThreads, L::kBytes, s>>>(k, out_map);
  return cudaGetLastError() == cudaSuccess ? PPK_OK : PPK_ERR_LAUNCH;
}
}  // namespace
}  // extern "C++"

int ppk_policy_first_layer(const PpkRunningMeanStd* rms, const float* obs, int64_t rows, int32_t width, const void* packed,
                           int32_t units, int32_t activation, void* out_f16, void* stream) {
  FlArgs k;
  memset(&k, 0, sizeof(k));
  if (rms != nullptr) {
    int rc = rms_args(rms, &k.rms);
    if (rc != PPK_OK) return rc;
    if (rms->width != width) return PPK_ERR_SHAPE;
  }
  if (rows < 0 || width <= 0 || units <= 0 || units % kFlN != 0) return PPK_ERR_SHAPE;
  if (activation != PPK_ACT_NONE && activation != PPK_ACT_ELU) return PPK_ERR_VARIANT;
  if (rows == 0) return PPK_OK;
  if (!obs || !packed || !out_f16) return PPK_ERR_NULL;
  if ((reinterpret_cast<uintptr_t>(packed) & 15u) || (reinterpret_cast<uintptr_t>(out_f16) & 15u) ||
      (reinterpret_cast<uintptr_t>(obs) & 3u))
    return PPK_ERR_ALIGN;
  const int kp = fl_kpad(width);
  k.obs = obs; k.rows = rows; k.width = width; k.units = units; k.activation = activation;
  k.packed = static_cast<const unsigned char*>(packed);
  k.out = static_cast<__half*>(out_f16);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (kp) {
    case 32: return launch_first_layer<32, 32>(k, s);
    case 80: return launch_first_layer<80, 80>(k, s);
    case 96: return launch_first_layer<96, 96>(k, s);       // TILT / NES / A3 / ALIGN (80) and A4 (94)
    case 320: return launch_first_layer<320, 64>(k, s);     // ADOF (313): weights streamed in K blocks of 64
    default: return PPK_ERR_SHAPE;
  }
}

// ---- fp32 variant (rollout forward): 3 x TF32 split products, fp32 out ------------------------------------------
extern "C++" {
namespace {
// default: single CTAs (202 us at 65 536 x 80 -> 2048, ELU).  PPK_FL32_CLUSTER=2: CTA pairs (cta_group::2, each CTA holds
// half of the B operand): parity-green and half the weight bytes per SM, but the pair's MMAs run at half the single-CTA
// rate with this operand layout (321 us), kept for A/B runs.  Process-wide: the packed weight layout depends on it.
int f32_cluster_size() {
  static const int v = [] { const char* e = getenv("PPK_FL32_CLUSTER"); return (e != nullptr && atoi(e) == 2) ? 2 : 1; }();
  return v;
}
}  // namespace
}  // extern "C++"

size_t ppk_linear_packed_bytes_f32(int32_t units, int32_t width) {
  if (units <= 0 || width <= 0 || units % kFlN != 0) return 0;
  return (size_t)units * f32_kpad(width) * 2 * sizeof(float);
}

int ppk_linear_pack_f32(const float* weight, const float* bias, int32_t units, int32_t width, void* packed, size_t packed_bytes,
                        void* stream) {
  if (!weight || !packed) return PPK_ERR_NULL;
  const size_t need = ppk_linear_packed_bytes_f32(units, width);
  if (need == 0 || packed_bytes < need) return PPK_ERR_SHAPE;
  if (reinterpret_cast<uintptr_t>(packed) & 15u) return PPK_ERR_ALIGN;
  linear_pack_f32_kernel<<<sm_count() * 2, 256, 0, static_cast<cudaStream_t>(stream)>>>(weight, bias, units, width, f32_kpad(width),
                                                                                  f32_cluster_size(), static_cast<float*>(packed));
  return cudaGetLastError() == cudaSuccess ? PPK_OK : PPK_ERR_LAUNCH;
}

extern "C++" {
namespace {
// out [rows, units] fp32 as a 2-D tensor; box = one epilogue warp's [32 rows x 32 units] tile, 128B swizzle
int make_out_map_f32(CUtensorMap* m, void* out, long long rows, int units) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return PPK_ERR_CUDA;
  const cuuint64_t dims[2] = {(cuuint64_t)units, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)units * sizeof(float)};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? PPK_OK
             : PPK_ERR_CUDA;
}

template <int KP, int ACT, int CL>
int launch_first_layer_f32_as(const F32Args& k, const CUtensorMap& out_map, cudaStream_t s) {
  using L = F32Layout<KP>;
  auto kern = first_layer_f32_kernel<KP, ACT, CL>;
  static SmemOptIn opt;
  if (!opt.ensure(kern, L::kBytes)) return PPK_ERR_LAUNCH;
  const long long groups = ((k.rows + kFlM - 1) / kFlM + CL - 1) / CL;
  const long long units = groups * (k.units / kFlN);
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kF32Threads);
  cfg.dynamicSmemBytes = L::kBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  // persistent: as many clusters as can be resident at once (one CTA per SM; GPCs with an odd SM count leave one out)
  static int max_clusters = 0;
  if (max_clusters == 0) {
    int n = 0;
    cfg.gridDim = dim3((unsigned)(sm_count() / CL * CL));
    if (CL > 1) { if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); return PPK_ERR_LAUNCH; } }
    else n = sm_count();
    max_clusters = n;
  }
  const long long clusters = units < max_clusters ? units : max_clusters;
  cfg.gridDim = dim3((unsigned)(clusters * CL));
  if (cudaLaunchKernelEx(&cfg, kern, k, out_map) != cudaSuccess) { cudaGetLastError(); return PPK_ERR_LAUNCH; }
  return PPK_OK;
}

template <int KP>
int launch_first_layer_f32(const F32Args& k, int activation, cudaStream_t s) {
  CUtensorMap out_map;
  int rc = make_out_map_f32(&out_map, k.out, k.rows, k.units);
  if (rc != PPK_OK) return rc;
  const bool pair = f32_cluster_size() == 2;
  if (activation == PPK_ACT_ELU) return pair ? launch_first_layer_f32_as<KP, 1, 2>(k, out_map, s) : launch_first_layer_f32_as<KP, 1, 1>(k, out_map, s);
  return pair ? launch_first_layer_f32_as<KP, 0, 2>(k, out_map, s) : launch_first_layer_f32_as<KP, 0, 1>(k, out_map, s);
}
}  // namespace
}  // extern "C++"

int ppk_policy_first_layer_f32(const PpkRunningMeanStd* rms, const float* obs, int64_t rows, int32_t width, const void* packed,
                               int32_t units, int32_t activation, float* out_f32, void* stream) {
  F32Args k;
  memset(&k, 0, sizeof(k));
  if (rms != nullptr) {
    int rc = rms_args(rms, &k.rms);
    if (rc != PPK_OK) return rc;
    if (rms->width != width) return PPK_ERR_SHAPE;
  }
  if (rows < 0 || width <= 0 || units <= 0 || units % kFlN != 0) return PPK_ERR_SHAPE;
  if (activation != PPK_ACT_NONE && activation != PPK_ACT_ELU) return PPK_ERR_VARIANT;
  if (rows == 0) return PPK_OK;
  if (!obs || !packed || !out_f32) return PPK_ERR_NULL;
  if ((reinterpret_cast<uintptr_t>(packed) & 15u) || (reinterpret_cast<uintptr_t>(out_f32) & 15u) ||
      (reinterpret_cast<uintptr_t>(obs) & 3u))
    return PPK_ERR_ALIGN;
  k.obs = obs; k.rows = rows; k.width = width; k.units = units;
  k.packed = static_cast<const unsigned char*>(packed);
  k.out = out_f32;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (f32_kpad(width)) {
    case 32: return launch_first_layer_f32<32>(k, activation, s);       // BASE (24)
    case 88: return launch_first_layer_f32<88>(k, activation, s);       // TILT / NES / A3 / ALIGN (80)
    case 96: return launch_first_layer_f32<96>(k, activation, s);       // A4 (94)
    default: return PPK_ERR_SHAPE;                                      // ADOF (313): the row tile does not fit
  }
}

int ppk_stats_reduce(double* stats, double* out, void* stream) {
  if (!stats || !out) return PPK_ERR_NULL;
  stats_reduce_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(stats, out);
  return cudaGetLastError() == cudaSuccess ? PPK_OK : PPK_ERR_LAUNCH;
}

}  // extern "C"
