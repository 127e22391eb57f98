// Small kernels around the fused step: BASE variant, pre_physics_step, reset_idx, statistics.
#pragma once
#include "ppk_async.cuh"
#include "ppk_device.cuh"

namespace ppk {

// ---- BASE (humanoid_pingpong.py) -----------------------------------------------------------------
// post_physics_step order BASE:587-596: progress += 1 -> reset_idx(envs whose reset_buf is set)
// -> compute_observations -> compute_reward.  No rotation: the obs row is a gather of pos/vel of
// paddle1, paddle2, ball1, ball2 (BASE:776-813).
constexpr int kBaseWarps = 1;          // 4096 envs = 128 warps: one per CTA spreads them over the SMs
constexpr int kBaseObs = 24;

__global__ void __launch_bounds__(kBaseWarps * 32)
base_step_kernel(const __grid_constant__ KArgs k) {
  __shared__ float obs_s[kBaseWarps][32 * (kBaseObs + 1)];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long env0 = ((long long)blockIdx.x * kBaseWarps + warp) * 32;
  if (env0 >= k.n) return;
  const int nvalid = (int)min(32LL, k.n - env0);
  const bool on = lane < nvalid;
  const long long env = env0 + (on ? lane : 0);
  const int phases = k.phases;
  const int rootN = k.A * kRow;

  long long prog = k.progress[env] + ((phases & PPK_PHASE_PROGRESS) ? 1 : 0);
  float* gr = k.root + (size_t)env * rootN;
  // every load the observation needs is requested up front, in one DRAM round trip with progress / reset;
  // lanes whose env is reset below re-read their two ball rows afterwards (cache hits)
  const float* p1 = k.rb + ((size_t)env * k.B + k.paddle_body[0]) * kRow;
  const float* p2 = k.rb + ((size_t)env * k.B + k.paddle_body[1]) * kRow;
  const float* b1 = gr + k.ball * kRow;
  const float* b2 = gr + (k.ball + 1) * kRow;
  float o[kBaseObs];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[c] = p1[c]; o[3 + c] = p1[7 + c];
    o[6 + c] = p2[c]; o[9 + c] = p2[7 + c];
    o[12 + c] = b1[c]; o[15 + c] = b1[7 + c];
    o[18 + c] = b2[c]; o[21 + c] = b2[7 + c];
  }
  // reset_idx of the envs flagged by the PREVIOUS step's reward (BASE:589-591)
  const bool base_resets = (phases & PPK_PHASE_RESET) && on && k.reset[env] != 0;
  if (phases & PPK_PHASE_RESET) append_reset_indices(k, base_resets, env, lane);
  // the warp resets its flagged envs cooperatively, all lanes copying (coalesced rows instead of one lane walking
  // 35 + 2*52 floats), four envs per pass with every load of the pass requested before the first store: with a tenth of
  // the envs flagged (the synthetic states) a pass per env was three dependent DRAM round trips per warp
  for (unsigned pending = __ballot_sync(0xffffffffu, base_resets); pending != 0;) {
    constexpr int G = 4;
    long long e[G];
    bool live[G];
#pragma unroll
    for (int u = 0; u < G; ++u) {
      live[u] = pending != 0;
      e[u] = env0 + (live[u] ? __ffs(pending) - 1 : 0);
      pending &= pending - 1;               // 0 stays 0
    }
    const int nroot = k.A * 7;              // pos + rot; velocities are NOT zeroed (BASE:533-534)
    for (int f0 = 0; f0 < nroot; f0 += 32) {
      const int f = f0 + lane;
      const int a = f / 7, c = f - a * 7;
      float v[G];
#pragma unroll
      for (int u = 0; u < G; ++u) v[u] = (live[u] && f < nroot) ? k.init_root[(size_t)e[u] * rootN + a * kRow + c] : 0.0f;
#pragma unroll
      for (int u = 0; u < G; ++u)
        if (live[u] && f < nroot) k.root_out[(size_t)e[u] * rootN + a * kRow + c] = v[u];   // = the input rows unless the caller splits the reset output off
    }
    if (lane < 6) {                         // ball1 <- velocity_1, ball2 <- velocity_2 (BASE:549-550)
      const float rv = k.reset_vel[lane];
#pragma unroll
      for (int u = 0; u < G; ++u)
        if (live[u]) k.root_out[(size_t)e[u] * rootN + (k.ball + lane / 3) * kRow + 7 + lane % 3] = rv;
    }
    if (k.reset_dof) {                      // BASE:552 (PpkTask.reset_dof; initial_dof_states required with it)
      for (int i0 = 0; i0 < 2 * k.D; i0 += 32) {
        const int i = i0 + lane;
        float v[G];
#pragma unroll
        for (int u = 0; u < G; ++u) v[u] = (live[u] && i < 2 * k.D) ? k.init_dof[(size_t)e[u] * 2 * k.D + i] : 0.0f;
#pragma unroll
        for (int u = 0; u < G; ++u)
          if (live[u] && i < 2 * k.D) k.dof_out[(size_t)e[u] * 2 * k.D + i] = v[u];
      }
    }
  }
  __syncwarp();                                             // the observations below read the rows just written
  if (base_resets) prog = 0;
  // time-out as BASE's reward sees it: from the progress AFTER the reset cleared it (BASE:587-596, 665)
  if (k.timeout != nullptr && on) k.timeout[env] = (prog >= k.max_len - 1) ? 1 : 0;
  if (on && (phases & (PPK_PHASE_PROGRESS | PPK_PHASE_RESET))) k.progress[env] = prog;

  if (base_resets) {
    // velocities 10:13 of the ball rows are not part of the obs, 7:10 were just rewritten: read the rows back
    const float* r1 = k.root_out + (size_t)env * rootN + k.ball * kRow;
    const float* r2 = r1 + kRow;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[12 + c] = r1[c]; o[15 + c] = r1[7 + c];
      o[18 + c] = r2[c]; o[21 + c] = r2[7 + c];
    }
  }
  if (phases & PPK_PHASE_OBS) {
#pragma unroll
    for (int c = 0; c < kBaseObs; ++c) obs_s[warp][lane * (kBaseObs + 1) + c] = o[c];
    __syncwarp();
    float* g = k.obs + (size_t)env0 * kBaseObs;
#pragma unroll
    for (int i = 0; i < kBaseObs; ++i) {
      int f = i * 32 + lane, e = f / kBaseObs, c = f - e * kBaseObs;
      if (e < nvalid) st_stream(g + f, clip_opt(obs_s[warp][e * (kBaseObs + 1) + c], k.clip_obs));
    }
  }
  if (phases & PPK_PHASE_REWARD) {
    // BASE:622-667: d1 = paddle1 <-> ball2, d2 = paddle2 <-> ball1
    float dx = o[0] - o[18], dy = o[1] - o[19], dz = o[2] - o[20];
    float d1 = sqrtf(dx * dx + dy * dy + dz * dz);
    dx = o[6] - o[12]; dy = o[7] - o[13]; dz = o[8] - o[14];
    float d2 = sqrtf(dx * dx + dy * dy + dz * dz);
    float r = 1.0f / (1.0f + d1 * d1) + 1.0f / (1.0f + d2 * d2);
    bool die = (o[14] < 0.1f) && (o[14] < 0.1f);          // ball1 tested twice (BASE:662)
    bool rst = (prog >= k.max_len - 1) || die;
    if (on) { k.rew[env] = r; k.reset[env] = rst ? 1 : 0; }
    if (phases & PPK_PHASE_STATS) {
      double s0 = warp_sum(on ? (double)r : 0.0), s1 = warp_sum(on ? (double)prog : 0.0), s2 = warp_sum((on && rst) ? 1.0 : 0.0);
      if (lane == 0) {
        double* slot = k.stats + (size_t)((env0 / 32) % PPK_STATS_SLOTS) * PPK_NUM_STATS;
        atomicAdd(slot + PPK_STAT_REWARD, s0); atomicAdd(slot + PPK_STAT_PROGRESS, s1); atomicAdd(slot + PPK_STAT_RESETS, s2);
      }
    }
  }
}

// ---- pre_physics_step (TILT:1002-1020) ---------------------------------------------------------------
// pd_tar[n,d] = offset[d] + scale[d]*actions[n,d]; save the ball's vx (and vz) for the next reward.
// CTAs [0, action_blocks) scale the actions, 16 bytes per thread and access when both tensors are 16-byte aligned (the
// [N, D] tensors are flat arrays here); the remaining CTAs save the ball velocities (two 4-byte reads at a 156-byte stride
// per env: a 64-byte granule fetched for 8 bytes, so this part is the larger half of the DRAM traffic) -- the two parts run
// side by side instead of one after the other in every thread (131 072 envs: 9.1 -> 7.0 us; 1 M envs: 39 us for ~165 MB of
// DRAM traffic, 67 MB of it the granules around the ball velocities).
__global__ void __launch_bounds__(256)
pre_step_kernel(float* __restrict__ actions, float clip, const float* __restrict__ offset, const float* __restrict__ scale,
                float* __restrict__ pd, long long n, int D, const float* __restrict__ root, int rootN, int ball,
                float* __restrict__ pre, int pre_stride, int pre_vx, int pre_vz, int* __restrict__ reset_count,
                unsigned action_blocks, int vec) {
  // (launched plainly: as a programmatic dependent this kernel and the step kernel behind it measured 1.5 % slower)
  // the compacted reset lists of the coming post_physics_step start empty (no extra memset node)
  if (reset_count != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *reset_count = 0;
  const long long total = n * D;
  if (blockIdx.x < action_blocks) {
    const long long stride = (long long)action_blocks * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long done = 0;                    // elements [0, done) are handled four at a time
    if (vec) {
      const long long n4 = total >> 2;
      done = n4 << 2;
      for (long long i4 = t0; i4 < n4; i4 += stride) {
        float4 a4 = __ldcs(reinterpret_cast<const float4*>(actions) + i4);
        float a[4] = {a4.x, a4.y, a4.z, a4.w}, o[4];
        int d = (int)((i4 << 2) % D);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (clip > 0.0f) a[j] = fminf(fmaxf(a[j], -clip), clip);      // VecTask.step: torch.clamp(actions, -clip, clip), kept in self.actions
          o[j] = __ldg(offset + d) + __ldg(scale + d) * a[j];
          if (++d == D) d = 0;
        }
        if (clip > 0.0f) reinterpret_cast<float4*>(actions)[i4] = make_float4(a[0], a[1], a[2], a[3]);
        __stcs(reinterpret_cast<float4*>(pd) + i4, make_float4(o[0], o[1], o[2], o[3]));
      }
    }
    for (long long i = done + t0; i < total; i += stride) {
      const int d = (int)(i % D);
      float a = ld_stream(actions + i);
      if (clip > 0.0f) {
        a = fminf(fmaxf(a, -clip), clip);
        actions[i] = a;
      }
      st_stream(pd + i, offset[d] + scale[d] * a);
    }
  } else if (pre != nullptr) {
    const long long stride = (long long)(gridDim.x - action_blocks) * blockDim.x;
    for (long long e = (long long)(blockIdx.x - action_blocks) * blockDim.x + threadIdx.x; e < n; e += stride) {
      const float* b = root + (size_t)e * rootN + ball * kRow;
      if (pre_stride == kRow) {          // the reference's full-row clone (TILT:1020)
        for (int c = 0; c < kRow; ++c) pre[(size_t)e * kRow + c] = b[c];
      } else {
        const float vx = __ldg(b + 7), vz = __ldg(b + 9);
        pre[(size_t)e * pre_stride + pre_vx] = vx;
        if (pre_vz >= 0 && pre_vz != pre_vx) pre[(size_t)e * pre_stride + pre_vz] = vz;
      }
    }
  }
}

// ---- reset_idx(env_ids) (TILT:847-906 and variants) ------------------------------------------------
// One warp per listed env: rewrite its root rows / DOF rows from the initial tensors, launch the
// ball, clear progress and flags, and gather the int32 actor / dof indices for the gym setters.
struct ResetArgs {
  const long long* env_ids;
  long long num_ids;
  const float* ball_vel;       // [k,3] (BASE [2,3]) or nullptr -> rows env_id of reset_vel
  const float* ball_yz;        // [k,2] ADOF or nullptr -> rows env_id of reset_yz
  const long long* actor_indices;
  const long long* dof_indices;
  int dof_per_env;
  int* actor_out;
  int* dof_out;
  int num_flags;               // flags written at reset
  int flag_reset_value[PPK_MAX_FLAGS];
  int variant;
  long long* last_hitter;      // ALIGN2: back to its initial value 2
};

__global__ void __launch_bounds__(128)
reset_idx_kernel(const __grid_constant__ KArgs k, const __grid_constant__ ResetArgs r) {
  const int lane = threadIdx.x & 31;
  const long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= r.num_ids) return;
  const long long env = r.env_ids[i];
  if (env < 0 || env >= k.n) return;
  const int rootN = k.A * kRow;
  const bool base = r.variant == PPK_BASE;
  const float* ir = k.init_root + (size_t)env * rootN;
  float* gr = k.root + (size_t)env * rootN;
  for (int f = lane; f < rootN; f += 32) {
    int c = f % kRow;
    if (c < 7) gr[f] = ir[f];
    else if (!base) gr[f] = 0.0f;
  }
  __syncwarp();
  if (lane < 3) {
    if (base) {
      gr[k.ball * kRow + 7 + lane] = r.ball_vel[lane];
      gr[(k.ball + 1) * kRow + 7 + lane] = r.ball_vel[3 + lane];
    } else {
      const float* bv = r.ball_vel ? r.ball_vel + (size_t)i * 3 : k.reset_vel + (size_t)env * 3;
      gr[k.ball * kRow + 7 + lane] = bv[lane];
    }
  }
  if (r.variant == PPK_ADOF && lane < 2) {
    const float* yz = r.ball_yz ? r.ball_yz + (size_t)i * 2 : k.reset_yz + (size_t)env * 2;
    gr[k.ball * kRow + 1 + lane] = yz[lane];
  }
  if (k.reset_dof) {
    const float* id = k.init_dof + (size_t)env * 2 * k.D;
    float* gd = k.dof + (size_t)env * 2 * k.D;
    for (int f = lane; f < 2 * k.D; f += 32) gd[f] = id[f];
  }
  if (lane == 0) {
    k.progress[env] = 0;
    if (base) k.reset[env] = 0;                      // BASE:579
    if (r.last_hitter != nullptr) r.last_hitter[env] = 2;
  }
  if (lane < r.num_flags) k.flags[lane][env] = (unsigned char)r.flag_reset_value[lane];
  if (r.actor_out != nullptr && lane < k.A) r.actor_out[i * k.A + lane] = (int)r.actor_indices[env * k.A + lane];
  if (r.dof_out != nullptr && lane < r.dof_per_env)
    r.dof_out[i * r.dof_per_env + lane] = (int)r.dof_indices[env * r.dof_per_env + lane];
}

// ---- statistics ----------------------------------------------------------------------------------
__global__ void stats_reduce_kernel(double* stats, double* out) {
  const int s = threadIdx.x;
  if (s >= PPK_NUM_STATS) return;
  double acc = 0.0;
  for (int slot = 0; slot < PPK_STATS_SLOTS; ++slot) {
    acc += stats[slot * PPK_NUM_STATS + s];
    stats[slot * PPK_NUM_STATS + s] = 0.0;
  }
  out[s] = acc;
}

// ---- device-side ball launch sampler ------------------------------------------
// The reference draws the launch velocity of every resetting env on the host with Python
// `random.uniform` inside a per-env loop (TILT:857-862, generate_random_speed_for_ball TILT:307-318,
// NES:312-323, ADOF:357-367, A3:300-302).  Here a counter-based Philox4x32-10 stream keyed by
// (seed, env, epoch) fills the per-env launch table consumed by the predicated reset.  Same ranges
// and formulas; bit parity with a Mersenne-Twister stream is impossible, so parity is statistical.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

__device__ __forceinline__ float u01(unsigned int x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }   // [0,1)

__global__ void __launch_bounds__(256)
sample_launch_kernel(float* __restrict__ vel, float* __restrict__ pos_yz, const long long* __restrict__ only_if_reset,
                     long long n, int variant, unsigned long long seed, unsigned long long epoch, long long env_offset) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  gdc_wait();                                                         // the step kernel ahead reads these rows and writes reset_buf
  if (e >= n) return;
  if (only_if_reset != nullptr && only_if_reset[e] == 0) return;      // refresh only the rows just consumed
  const unsigned long long g = (unsigned long long)(e + env_offset);  // global env id: shards draw disjoint streams
  const uint4 r = philox4x32_10(make_uint4((unsigned)g, (unsigned)(g >> 32), (unsigned)epoch, (unsigned)(epoch >> 32)),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  const float rad = 0.017453292519943295f;
  float s, a, z = 0.0f, vx, vy, vz;
  if (variant == PPK_TILT || variant == PPK_A4 || variant == PPK_ALIGN || variant == PPK_ALIGN2) {
    // s = -U(8.0, 8.6 | 8.8), a = U(-5,5) deg, z = U(2,10) deg;  v = s*(cos a cos z, sin a sin z, sin a)
    s = -(8.0f + u01(r.x) * ((variant == PPK_ALIGN || variant == PPK_ALIGN2) ? 0.8f : 0.6f));
    a = (-5.0f + u01(r.y) * 10.0f) * rad;
    z = (2.0f + u01(r.z) * 8.0f) * rad;
    vx = s * cosf(a) * cosf(z); vy = s * sinf(a) * sinf(z); vz = s * sinf(a);
  } else if (variant == PPK_NES || variant == PPK_ADOF) {
    // v = (-s cos a cos z, s sin a cos z, s sin z)
    const bool nes = variant == PPK_NES;
    s = nes ? (5.4f + u01(r.x) * 0.5f) : (5.0f + u01(r.x) * 0.4f);
    a = (nes ? (-5.0f + u01(r.y) * 10.0f) : (-8.0f + u01(r.y) * 11.0f)) * rad;
    z = (nes ? (10.0f + u01(r.z) * 7.0f) : (14.0f + u01(r.z) * 10.0f)) * rad;
    vx = -s * cosf(a) * cosf(z); vy = s * sinf(a) * cosf(z); vz = s * sinf(z);
  } else {
    // A3 / BASE: v = s*(cos a, sin a, 0), s = -U(6.5, 7.5)
    s = -(6.5f + u01(r.x) * 1.0f);
    a = (-5.0f + u01(r.y) * 10.0f) * rad;
    vx = s * cosf(a); vy = s * sinf(a); vz = 0.0f;
  }
  vel[e * 3 + 0] = vx; vel[e * 3 + 1] = vy; vel[e * 3 + 2] = vz;
  if (variant == PPK_ADOF && pos_yz != nullptr) {
    // ADOF:127-128,976-979: ball y ~ U(-0.5, 0.1), z ~ U(0.96, 1.05); second Philox block
    const uint4 q = philox4x32_10(make_uint4((unsigned)g, (unsigned)(g >> 32), (unsigned)epoch, (unsigned)(epoch >> 32) ^ 0x80000000u),
                                  make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    pos_yz[e * 2 + 0] = -0.5f + u01(q.x) * 0.6f;
    pos_yz[e * 2 + 1] = 0.96f + u01(q.y) * 0.09f;
  }
}

// ADOF:1162-1175: when any env of the shard reset this step, all five counters are cleared.
__global__ void __launch_bounds__(256)
adof_clear_counters_kernel(unsigned int* any_reset, unsigned char* c0, unsigned char* c1, unsigned char* c2,
                           unsigned char* c3, unsigned char* c4, long long n) {
  // scratch[0] = "some env reset" flag set by the step kernel, scratch[1] = ticket.  Every block
  // reads the flag, then takes a ticket; the last one re-arms both words for the next step, so no
  // memset node is needed in front of the step kernel.
  __shared__ unsigned int flag_s;
  gdc_wait();
  gdc_launch_dependents();
  if (threadIdx.x == 0) flag_s = any_reset[0];
  __syncthreads();
  const bool clear = flag_s != 0u;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(any_reset + 1, 1u);
    if (t == gridDim.x - 1) { any_reset[0] = 0u; any_reset[1] = 0u; }
  }
  if (!clear) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const uintptr_t all = reinterpret_cast<uintptr_t>(c0) | reinterpret_cast<uintptr_t>(c1) | reinterpret_cast<uintptr_t>(c2) |
                        reinterpret_cast<uintptr_t>(c3) | reinterpret_cast<uintptr_t>(c4);
  if (all & 3u) {     // a shard view at an odd env offset: clear bytewise
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
      c0[j] = 0; c1[j] = 0; c2[j] = 0; c3[j] = 0; c4[j] = 0;
    }
    return;
  }
  const long long words = (n + 3) / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += stride) {
    long long b = i * 4;
    if (b + 4 <= n) {
      *reinterpret_cast<unsigned int*>(c0 + b) = 0u; *reinterpret_cast<unsigned int*>(c1 + b) = 0u;
      *reinterpret_cast<unsigned int*>(c2 + b) = 0u; *reinterpret_cast<unsigned int*>(c3 + b) = 0u;
      *reinterpret_cast<unsigned int*>(c4 + b) = 0u;
    } else {
      for (long long j = b; j < n; ++j) { c0[j] = 0; c1[j] = 0; c2[j] = 0; c3[j] = 0; c4[j] = 0; }
    }
  }
}

}  // namespace ppk
