// Fused task step for the 3-/4-actor variants (A3, TILT, NES, ALIGN, A4, ALIGN2).
//
// One CTA of four warps owns one tile of 32 (env, humanoid) units: 32 envs (one humanoid) or 16 envs (two).
//   stage    The PhysX tensors are AoS with 52-byte rows: nothing is 16-byte aligned per env.  But an env PAIR is
//            (2 x B x 52 bytes, B even), so the rigid-body tensor is handed to the TMA engine as a 2-D tensor
//            [N/2 pairs, 2*B*13 floats]: the rows ids[1..J) of all even (odd) envs of the tile are ONE
//            cp.async.bulk.tensor.2d box (480 B x TILE/2 pairs), row ids[0] another (48 B x TILE/2), the inner
//            coordinate rounded down to the 16-byte boundary TMA needs and the remainder (0..3 floats) applied
//            when the staged rows are read.  L2 promotion 64 B: the smallest DRAM fetch granularity the engine
//            offers (58 MB instead of 66 MB per 65536-env launch, profiles/r2_staging_probe.md).  The tile's root,
//            DOF-state and DOF-force slices are contiguous: one 1-D bulk copy each.  Seven async copies per tile,
//            one mbarrier; no registers or LSU slots are spent on staging.  Tail tiles, misaligned tensors and
//            non-consecutive id lists take an LDG path into the same layout.
//   soft start  All CTAs of the first wave would issue their copies at once and (the memory system returns the
//            sectors in no particular order) complete together ~4 us later, leaving DRAM idle while they store and
//            the next wave starts.  First-wave CTA b delays its copies by (b / #SMs) x `stagger` cycles.
//   warp 0   lane = unit: heading frame (atan2f / sinf / cosf, restated op by op) -> table in shared memory
//   warp 1   lane = env: progress+1, reward, die / time-out mask, flag updates, statistics, the predicated reset
//            (root / DOF rows rewritten from the initial tensors), then the obs tail (dof_pos, 0.1*dof_vel, ball
//            in the heading frame)
//   warps 0, 2, 3  lane = (unit, body): both rotated vectors of one body -> the obs tile in shared memory
//   store    the tile's obs rows are contiguous in global memory: ONE bulk shared->global copy (full sectors only)
#pragma once
#include <cuda.h>

#include "ppk_async.cuh"
#include "ppk_device.cuh"

namespace ppk {

#ifdef PPK_TRACE
// Debug build only: per-warp timeline stamps (globaltimer ns) -> g_trace[(block*8 + warp)*4 + slot]
__device__ unsigned long long* g_trace = nullptr;
__device__ __forceinline__ void trace_stamp(int warp, int slot) {
  if (g_trace != nullptr && (threadIdx.x & 31) == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace[((size_t)blockIdx.x * 8 + warp) * 4 + slot] = t;
  }
}
#define PPK_STAMP(slot) trace_stamp(warp, slot)
#else
#define PPK_STAMP(slot)
#endif

constexpr int kFamilyWarps = 4;
constexpr int kFamilyThreads = 32 * kFamilyWarps;
constexpr int kRow0Max = 16;    // floats reserved per staged ids[0] row

template <int H, int J, int D, int A, int TILE>
struct FamilyLayout {
  static constexpr int kUnits = TILE * H;                               // (env, humanoid) pairs
  static constexpr int kPairs = TILE / 2;                               // env pairs = rows of a tensor box
  static constexpr int kSpanF = (((J - 1) * kRow + 3 + 3) / 4) * 4;     // 120: rows ids[1..J) + alignment slack
  static constexpr int kRootEnv = A * kRow;
  static constexpr int kTail = 2 * D + 6;              // dof_pos, 0.1*dof_vel, ball local pos, vel
  static constexpr int kObs = 6 * J + kTail;           // 80 (D=7) / 94 (D=14)
  static constexpr int kHd = 8;                        // a0, sz, cw, span offset | root pos xyz, row-0 offset
  // float offsets inside the CTA's shared memory
  static constexpr int kOffSpan = 0;                                   // [H][2][kPairs][kSpanF]
  static constexpr int kOffRow0 = kOffSpan + kUnits * kSpanF;          // [H][2][kPairs][row0_box]
  static constexpr int kOffRoot = kOffRow0 + kUnits * kRow0Max;        // [TILE][A*13]
  static constexpr int kOffDof = kOffRoot + TILE * kRootEnv;           // [TILE][2D]
  static constexpr int kOffForce = kOffDof + TILE * 2 * D;             // [TILE][D]
  static constexpr int kOffHd = kOffForce + TILE * D;                  // [kUnits][8]
  static constexpr int kOffObs = kOffHd + kUnits * kHd;                // [kUnits][kObs]
  static constexpr int kOffBar = kOffObs + kUnits * kObs;              // two 8-byte mbarriers
  static constexpr int kFloats = kOffBar + 4;
  static constexpr uint32_t kLinearTx = 4u * TILE * (kRootEnv + 3 * D);
  static_assert(kUnits <= 32, "one heading per lane of warp 0");
  static_assert(TILE % 4 == 0 && TILE <= 32, "tile");
  static_assert((kPairs * kSpanF * 4) % 128 == 0 && (kPairs * 12 * 4) % 128 == 0, "128-byte tensor box destinations");
  static_assert((kOffRow0 * 4) % 128 == 0, "128-byte tensor box destinations");
  static_assert(kOffRoot % 4 == 0 && kOffDof % 4 == 0 && kOffForce % 4 == 0 && kOffHd % 4 == 0 && kOffObs % 4 == 0,
                "16-byte bulk copies");
  static_assert((TILE * kRootEnv) % 4 == 0 && (TILE * D) % 4 == 0 && (kUnits * kObs) % 4 == 0, "16-byte bulk sizes");
  static_assert(kOffBar % 2 == 0, "mbarrier alignment");
};

// Named barriers (barrier 0 = __syncthreads):
//   1  heading table ready: warp 0 arrives, warp 3 waits        } one barrier per waiting warp: a waiter must never
//   3  heading table ready: warp 0 arrives, warp 2 waits        } wait for ANOTHER waiter's arrival
//   2  pre-reset state read: warp 1 arrives, warp 2 waits before it overwrites the staged rows of resetting envs

constexpr int kResetBatch = 4;      // resetting envs of a tile whose source rows are in flight at once

template <int V, int H, int J, int D, int A, int TILE>
__global__ void __launch_bounds__(kFamilyThreads, 6)
family_step_kernel(const __grid_constant__ KArgs k, const __grid_constant__ CUtensorMap m_span,
                   const __grid_constant__ CUtensorMap m_row0) {
  using L = FamilyLayout<H, J, D, A, TILE>;
  extern __shared__ __align__(128) float smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const long long env0 = (long long)blockIdx.x * TILE;
  const int nvalid = (int)min((long long)TILE, k.n - env0);

  float* span_s = smem + L::kOffSpan;
  float* row0_s = smem + L::kOffRow0;
  float* root_s = smem + L::kOffRoot;
  float* dof_s = smem + L::kOffDof;
  float* force_s = smem + L::kOffForce;
  float4* hd_s = reinterpret_cast<float4*>(smem + L::kOffHd);
  float* obs_s = smem + L::kOffObs;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::kOffBar);     // [0] row ids[0] boxes, [1] everything else

  const int phases = k.phases;
  const bool want_obs = (phases & PPK_PHASE_OBS) != 0;
  const bool fast = (k.bulk_ok & 1) && (nvalid == TILE);
  const int r0s = k.row0_box;
  // where unit (e, h)'s staged rows start (floats from the start of shared memory)
  auto span_at = [&](int e, int h) -> int {
    return L::kOffSpan + ((h * 2 + (e & 1)) * L::kPairs + (e >> 1)) * L::kSpanF + k.span_off[h][e & 1];
  };
  auto row0_at = [&](int e, int h) -> int {
    return L::kOffRow0 + ((h * 2 + (e & 1)) * L::kPairs + (e >> 1)) * r0s + k.row0_off[h][e & 1];
  };

  // ---- stage ---------------------------------------------------------------------------------------
  PPK_STAMP(0);
  // Programmatic dependent launch: up to here (parameters, address arithmetic) the CTA may run while the previous kernel
  // of the stream is still draining; nothing of global memory has been touched yet.
  gdc_wait();
  if (fast) {
    if (warp == 0) {
      // one elected lane of the converged warp 0 sets the barriers up and issues the seven copies (warp-uniform
      // operands: no per-lane serialisation); the other warps meet it at the __syncthreads below
      if (elect_one()) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        mbar_fence_init();
        if (k.stagger > 0 && (int)blockIdx.x < k.first_wave && (int)blockIdx.x >= k.num_sms) {
          const long long wait = (long long)((int)blockIdx.x / k.num_sms) * k.stagger;
          const long long t0 = clock64();
          while (clock64() - t0 < wait) __nanosleep(64);
        }
        const int p0 = (int)(env0 >> 1);
        // the heading frames need only row ids[0]: requested first, on their own barrier, so that warp 0 computes
        // them (atan2f / sinf / cosf: ~0.6 us) while the rest of the tile is still in flight
        mbar_arrive_expect_tx(bar, 4u * L::kUnits * r0s);
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
          for (int q = 0; q < 2; ++q) tma_load_2d_in(row0_s + (h * 2 + q) * L::kPairs * r0s, &m_row0, k.row0_c[h][q], p0, bar);
        mbar_arrive_expect_tx(bar + 1, 4u * L::kUnits * L::kSpanF + L::kLinearTx);
        bulk_g2s_in(root_s, k.root + (size_t)env0 * L::kRootEnv, 4u * TILE * L::kRootEnv, bar + 1);
        bulk_g2s_in(dof_s, k.dof + (size_t)env0 * 2 * D, 4u * TILE * 2 * D, bar + 1);
        bulk_g2s_in(force_s, k.force + (size_t)env0 * D, 4u * TILE * D, bar + 1);
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
          for (int q = 0; q < 2; ++q)
            tma_load_2d_in(span_s + (h * 2 + q) * L::kPairs * L::kSpanF, &m_span, k.span_c[h][q], p0, bar + 1);
      }
      __syncwarp();
    }
    __syncthreads();        // the initialised barriers are visible to every waiter
  } else {
    // generic path (tail tile, unaligned tensors, non-consecutive ids): plain loads, same layout
    const int env_stride = k.B * kRow;
    const float* g_rb = k.rb + (size_t)env0 * env_stride;
    for (int f = tid; f < L::kUnits * J * kRow; f += kFamilyThreads) {
      const int u = f / (J * kRow), r = f - u * (J * kRow);
      const int e = u / H, h = u - e * H;
      const int j = r / kRow, c = r - j * kRow;
      const float v = (e < nvalid) ? g_rb[(size_t)e * env_stride + k.ids[h][j] * kRow + c] : 0.0f;
      if (j == 0) { if (c < 10) smem[row0_at(e, h) + c] = v; }
      else smem[span_at(e, h) + (j - 1) * kRow + c] = v;
    }
    for (int f = tid; f < TILE * L::kRootEnv; f += kFamilyThreads)
      root_s[f] = (f < nvalid * L::kRootEnv) ? k.root[(size_t)env0 * L::kRootEnv + f] : 0.0f;
    for (int f = tid; f < TILE * 2 * D; f += kFamilyThreads)
      dof_s[f] = (f < nvalid * 2 * D) ? k.dof[(size_t)env0 * 2 * D + f] : 0.0f;
    for (int f = tid; f < TILE * D; f += kFamilyThreads)
      force_s[f] = (f < nvalid * D) ? k.force[(size_t)env0 * D + f] : 0.0f;
    __syncthreads();
  }
  PPK_STAMP(1);
  // Every CTA of this grid has been scheduled once the last one gets here: from then on the next kernel of the stream may
  // fill the slots the tail of this grid leaves empty and run its prologue (it blocks in gdc_wait until this grid is done).
  gdc_launch_dependents();

  const float clip = k.clip_obs;
  const bool clip_on = clip > 0.0f;
  const bool lane_env = lane < nvalid;
  const long long env = env0 + (lane_env ? lane : 0);
  const int le = (lane < TILE) ? lane : 0;     // staged row a lane = env warp reads (idle lanes read row 0)

  // lane = (unit, body): my_quat_rotate((0,0,sz,cw), v) of the body's position (relative to the root body) and
  // velocity, component by component in the operation order of the restated helper -> obs tile
  auto rotate_bodies = [&](int it0, int it1) {
    constexpr int kBodies = L::kUnits * J;
#pragma unroll 2
    for (int it = it0; it < it1; ++it) {
      const int p = it * 32 + lane;
      const bool on = p < kBodies;
      const int pp = on ? p : 0;
      const int u = pp / J, j = pp - u * J;
      const float4 fa = hd_s[u * 2], fb = hd_s[u * 2 + 1];
      const float a0 = fa.x, sz = fa.y, cw = fa.z;
      const float* row = smem + ((j == 0) ? __float_as_int(fb.w) : (__float_as_int(fa.w) + (j - 1) * kRow));
      const float px = row[0] - fb.x, py = row[1] - fb.y, pz = row[2] - fb.z;
      const float vx = row[7], vy = row[8], vz = row[9];
      float o0 = px * a0 + ((-(sz * py)) * cw) * 2.0f;
      float o1 = py * a0 + ((sz * px) * cw) * 2.0f;
      float o2 = pz * a0 + (sz * (sz * pz)) * 2.0f;
      float o3 = vx * a0 + ((-(sz * vy)) * cw) * 2.0f;
      float o4 = vy * a0 + ((sz * vx) * cw) * 2.0f;
      float o5 = vz * a0 + (sz * (sz * vz)) * 2.0f;
      if (clip_on) {
        o0 = clip_to(o0, clip); o1 = clip_to(o1, clip); o2 = clip_to(o2, clip);
        o3 = clip_to(o3, clip); o4 = clip_to(o4, clip); o5 = clip_to(o5, clip);
      }
      if (on) {
        float* o = obs_s + u * L::kObs + 3 * j;
        o[0] = o0; o[1] = o1; o[2] = o2;
        o[3 * J] = o3; o[3 * J + 1] = o4; o[3 * J + 2] = o5;
      }
    }
  };
  constexpr int kIters = (L::kUnits * J + 31) / 32;      // 10
#ifndef PPK_ROT_W0
#define PPK_ROT_W0 ((kIters * 2 + 2) / 5)
#endif
  constexpr int kItW0 = PPK_ROT_W0;                      // warp 0 (its frames are ready before the data)
  constexpr int kItW3 = PPK_ROT_W0;                      // warp 3; warp 2 takes the rest after the reset / tail
  static_assert(kItW0 + kItW3 <= kIters, "rotation split");   // warp 2 takes the rest after the reset / tail

  if (warp == 0) {
    // ================= warp 0: heading frames (lane = unit), then rotations ==============================
    if (!want_obs) return;
    if (fast) mbar_wait(bar, 0);
    if (lane < L::kUnits) {
      const int u = lane, e = u / H, h = u - e * H;
      const int so = span_at(e, h), ro = row0_at(e, h);
      const float* r0 = smem + ro;
      const Heading hq = heading_quat_inv(r0[3], r0[4], r0[5], r0[6]);
      hd_s[u * 2] = make_float4(2.0f * (hq.cw * hq.cw) - 1.0f, hq.sz, hq.cw, __int_as_float(so));
      hd_s[u * 2 + 1] = make_float4(r0[0], r0[1], r0[2], __int_as_float(ro));
    }
    bar_arrive(1, 64);
    bar_arrive(3, 64);
    __syncwarp();
    if (fast) mbar_wait(bar + 1, 0);
    PPK_STAMP(2);
    rotate_bodies(0, kItW0);
    PPK_STAMP(3);
  } else if (warp == 3) {
    if (!want_obs) return;
    if (fast) mbar_wait(bar + 1, 0);
    PPK_STAMP(2);
    bar_wait(1, 64);
    rotate_bodies(kItW0, kItW0 + kItW3);
    PPK_STAMP(3);
  } else if (warp == 1) {
    // ================= warp 1: reward, flags, statistics (lane = env) ===================================
    long long prog = 0, reset_prev = 0;
    float pre_vx = 0.0f, pre_vz = 0.0f;
    constexpr int NF = (V == PPK_TILT) ? 3 : (V == PPK_A4) ? 6 : (V == PPK_NES) ? 2 : (V == PPK_ALIGN || V == PPK_ALIGN2) ? 1 : 0;
    long long last_hitter = 2;
    bool flag[NF > 0 ? NF : 1];
    // per-env scalars come straight from global (coalesced, lane = env) and overlap the async copies
    if (lane_env) {
      prog = k.progress[env];
      if (!(phases & PPK_PHASE_REWARD)) reset_prev = k.reset[env];
      if (phases & PPK_PHASE_REWARD) {
        const float* p = k.pre + (size_t)env * k.pre_stride;
        pre_vx = ld_stream(p + k.pre_vx);
        if (V == PPK_ALIGN || V == PPK_ALIGN2) pre_vz = ld_stream(p + k.pre_vz);
        if (V == PPK_ALIGN2) last_hitter = k.last_hitter[env];
#pragma unroll
        for (int i = 0; i < NF; ++i) flag[i] = k.flags[i][env] != 0;
      }
    }
    if (fast) mbar_wait(bar + 1, 0);
    PPK_STAMP(2);

    // everything the reward reads of the (pre-reset) staged state, then warp 2 may overwrite the rows of resetting envs
    const float* my_root = root_s + le * L::kRootEnv;
    const float* my_ball = my_root + k.ball * kRow;
    const float bx = my_ball[0], by = my_ball[1], bz = my_ball[2];
    const float bvx = my_ball[7], bvz = my_ball[9];
    float hx[H];
#pragma unroll
    for (int h = 0; h < H; ++h) hx[h] = my_root[k.hum[h] * kRow];
    float power = 0.0f;
    if (phases & PPK_PHASE_REWARD) {
#pragma unroll
      for (int d = 0; d < D; ++d) power += fabsf(force_s[le * D + d] * dof_s[le * 2 * D + 2 * d + 1]);
    }
    bar_arrive(2, 64);

    long long p_new = prog + ((phases & PPK_PHASE_PROGRESS) ? 1 : 0);
    bool is_reset = reset_prev != 0;
    float rew[H];
#pragma unroll
    for (int h = 0; h < H; ++h) rew[h] = 0.0f;

    if (phases & PPK_PHASE_REWARD) {
      bool die = false;
      Scene sc[H];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        Scene& s = sc[h];
        s.bx = bx; s.by = by; s.bz = bz; s.vx = bvx; s.vz = bvz;
        s.pre_vx = pre_vx; s.pre_vz = pre_vz;
        const int pj = k.paddle_j[h];
        const float* pd;
        if (pj > 0) pd = smem + span_at(le, h) + (pj - 1) * kRow;
        else if (pj == 0) pd = smem + row0_at(le, h);
        else pd = k.rb + ((size_t)env * k.B + k.paddle_body[h]) * kRow;   // not among the observation bodies: in place
        s.px = pd[0]; s.py = pd[1]; s.pz = pd[2];
        s.hx = hx[h];
        s.power_reward = (-k.power_coef) * power;
        s.progress = p_new;
      }
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const Scene& s = sc[h];
        bool d = false;
        if (V == PPK_A3) rew[h] = reward_a3(s, k, d);
        if (V == PPK_TILT) rew[h] = reward_tilt<false>(s, k, flag[0], flag[1], flag[2], d);
        if (V == PPK_NES) rew[h] = reward_nes(s, k, flag[0], flag[1], d);
        if (V == PPK_ALIGN) rew[h] = reward_align(s, k, flag[0], d);
        if (V == PPK_A4) {
          if (h == 0) rew[h] = reward_tilt<false>(s, k, flag[0], flag[1], flag[2], d);
          else rew[h] = reward_tilt<true>(s, k, flag[NF > 3 ? 3 : 0], flag[NF > 4 ? 4 : 0], flag[NF > 5 ? 5 : 0], d);
        }
        die = die || d;
      }
      if (V == PPK_ALIGN2) reward_align2(sc[0], sc[H - 1], k, flag[0], last_hitter, rew[0], rew[H - 1], die);
      is_reset = (p_new >= k.max_len - 1) || die;
      if (lane_env) {
#pragma unroll
        for (int h = 0; h < H; ++h) k.rew[(size_t)env * H + h] = rew[h];
        k.reset[env] = is_reset ? 1 : 0;
      }
    }

    if (phases & PPK_PHASE_STATS) {
      double s_rew = warp_sum(lane_env ? (double)rew[0] : 0.0);
      double s_prog = warp_sum(lane_env ? (double)p_new : 0.0);
      double s_rst = warp_sum((lane_env && is_reset) ? 1.0 : 0.0);
      if (lane == 0) {
        double* slot = k.stats + (size_t)(blockIdx.x % PPK_STATS_SLOTS) * PPK_NUM_STATS;
        atomicAdd(slot + PPK_STAT_REWARD, s_rew);
        atomicAdd(slot + PPK_STAT_PROGRESS, s_prog);
        atomicAdd(slot + PPK_STAT_RESETS, s_rst);
      }
    }

    // per-env bookkeeping of the predicated reset (TILT:847-906); warp 2 rewrites the root / DOF rows
    const bool do_reset = (phases & PPK_PHASE_RESET) && is_reset && lane_env;
    if (phases & PPK_PHASE_RESET) append_reset_indices(k, do_reset, env, lane);
    if (k.timeout != nullptr && lane_env && (phases & (PPK_PHASE_REWARD | PPK_PHASE_PROGRESS)))
      k.timeout[env] = (p_new >= k.max_len - 1) ? 1 : 0;
    if (do_reset) p_new = 0;
    if (V == PPK_ALIGN2 && lane_env && (phases & (PPK_PHASE_REWARD | PPK_PHASE_RESET))) {
      if (do_reset) k.last_hitter[env] = 2;                       // a fresh rally starts with its initial value (ALIGN:1253)
      else if (phases & PPK_PHASE_REWARD) k.last_hitter[env] = last_hitter;
    }
    if (lane_env) {
      if (phases & (PPK_PHASE_PROGRESS | PPK_PHASE_RESET)) k.progress[env] = p_new;
      if (NF > 0 && (phases & (PPK_PHASE_REWARD | PPK_PHASE_RESET))) {
#pragma unroll
        for (int i = 0; i < NF; ++i) {
          // reset values: *_calculated -> False, no_bounce_before_half_mask -> True (TILT:902-905)
          const bool reset_val = (V == PPK_TILT || V == PPK_A4) ? ((i % 3) == 2) : false;
          if (do_reset) k.flags[i][env] = reset_val ? 1 : 0;
          else if ((phases & PPK_PHASE_REWARD) && k.write_flags) k.flags[i][env] = flag[i] ? 1 : 0;
        }
      }
    }
    PPK_STAMP(3);
    if (!want_obs) return;
  } else {
    // ================= warp 2: predicated reset, obs tail, remaining rotations ============================
    // Which envs reset is decidable from a few scalars (progress, ball height) read straight from global, long before
    // the tile's bulk data lands: the source rows of up to kResetBatch resetting envs are requested right away (all
    // lanes, coalesced) and are in registers when the data arrives.  A3's third condition (missed_ball: ball behind the
    // paddle) needs the paddle row: it is taken from the staged tile once that has landed and its envs join the loop
    // then -- a 4-byte read of that row straight from global ahead of the tensor-map copy of the same lines cost the
    // whole batch 10 % at 65 536 envs and 18 % at 1 M.
    const bool rst_phase = (phases & PPK_PHASE_RESET) != 0;
    const bool a3_late = V == PPK_A3 && rst_phase && (phases & PPK_PHASE_REWARD) && fast && k.paddle_j[0] >= 0;
    unsigned pending = 0u;
    bool is_reset = false;
    if (rst_phase) {
      if (lane_env) {
        const float* g_root = k.root + (size_t)env * L::kRootEnv;
        if (phases & PPK_PHASE_REWARD) {
          const long long p_new = k.progress[env] + ((phases & PPK_PHASE_PROGRESS) ? 1 : 0);
          bool die = false;
          if (V != PPK_NES) die = g_root[k.ball * kRow + 2] < 0.1f;
          if (V == PPK_A3 && !a3_late) {      // missed_ball (A3:1149): ball behind the paddle
            const float px = k.rb[((size_t)env * k.B + k.paddle_body[0]) * kRow];
            die = die || (g_root[k.ball * kRow] < px - 1e-3f);
          }
          is_reset = (p_new >= k.max_len - 1) || die;
        } else {
          is_reset = k.reset[env] != 0;
        }
      }
      pending = __ballot_sync(0xffffffffu, is_reset && lane_env);
    }
    constexpr int RE = L::kRootEnv;                 // <= 64 floats: two per lane
    static_assert(RE <= 64 && 2 * D <= 32, "one reset env per warp pass");
    float r_root[kResetBatch][2], r_dof[kResetBatch];
    int r_env[kResetBatch];
    bool waited = false;
    do {
      // ---- request the source rows of the next batch of resetting envs
#pragma unroll
      for (int i = 0; i < kResetBatch; ++i) {
        r_env[i] = -1;
        if (pending != 0u) {
          const int e = __ffs(pending) - 1;
          pending &= pending - 1u;
          r_env[i] = e;
          const long long ge = env0 + e;
          const float* ir = k.init_root + (size_t)ge * RE;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int f = lane + 32 * half;
            float v = 0.0f;
            if (f < RE) {
              const int a = f / kRow, c = f - a * kRow;
              if (c < 7) v = __ldg(ir + f);
              else if (a == k.ball && c < 10) v = __ldg(k.reset_vel + (size_t)ge * 3 + (c - 7));
            }
            r_root[i][half] = v;
          }
          r_dof[i] = (k.reset_dof && lane < 2 * D) ? __ldg(k.init_dof + (size_t)ge * 2 * D + lane) : 0.0f;
        }
      }
      if (!waited) {
        if (fast) mbar_wait(bar + 1, 0);
        PPK_STAMP(2);
        bar_wait(2, 64);            // warp 1 has read the pre-reset state of the tile
        waited = true;
        if (a3_late) {              // missed_ball from the staged rows (no env of this warp's batch has been rewritten yet)
          bool late = false;
          if (lane_env && !is_reset) {
            const int pj = k.paddle_j[0];
            const float px = pj > 0 ? smem[span_at(le, 0) + (pj - 1) * kRow] : smem[row0_at(le, 0)];
            late = root_s[le * L::kRootEnv + k.ball * kRow] < px - 1e-3f;
          }
          pending |= __ballot_sync(0xffffffffu, late);
        }
      }
      // ---- rewrite the rows: global tensors (what PhysX continues from) and the staged copy (what the obs see)
#pragma unroll
      for (int i = 0; i < kResetBatch; ++i) {
        const int e = r_env[i];
        if (e >= 0) {
          const long long ge = env0 + e;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int f = lane + 32 * half;
            if (f < RE) {
              k.root_out[(size_t)ge * RE + f] = r_root[i][half];
              root_s[e * RE + f] = r_root[i][half];
            }
          }
          if (k.reset_dof && lane < 2 * D) {
            k.dof_out[(size_t)ge * 2 * D + lane] = r_dof[i];
            dof_s[e * 2 * D + lane] = r_dof[i];
          }
        }
      }
    } while (pending != 0u);
    if (!want_obs) return;

    // ---- obs tail: dof_pos (D), 0.1*dof_vel (D) of every unit, lane = (unit, dof) --------------------
    __syncwarp();
    constexpr int kDofs = L::kUnits * D;
#pragma unroll 2
    for (int p = lane; p < kDofs; p += 32) {
      const int u = p / D, d = p - u * D;
      const int e = u / H;
      float q = dof_s[e * 2 * D + 2 * d];
      float qd = dof_s[e * 2 * D + 2 * d + 1] * 0.1f;
      if (clip_on) { q = clip_to(q, clip); qd = clip_to(qd, clip); }
      float* o = obs_s + u * L::kObs + 6 * J;
      o[d] = q;
      o[D + d] = qd;
    }
    // ---- ball in the heading frame (frames come from warp 0), lane = env --------------------------------
    bar_wait(3, 64);
    if (lane < TILE) {
      const float* my_ball = root_s + le * L::kRootEnv + k.ball * kRow;
      const float bx = my_ball[0], by = my_ball[1], bz = my_ball[2];
      const float bvx = my_ball[7], bvy = my_ball[8], bvz = my_ball[9];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const int u = le * H + h;
        const float4 fa = hd_s[u * 2], fb = hd_s[u * 2 + 1];
        const float a0 = fa.x, sz = fa.y, cw = fa.z;
        const float rx = bx - fb.x, ry = by - fb.y, rz = bz - fb.z;
        float o0 = rx * a0 + ((-(sz * ry)) * cw) * 2.0f;
        float o1 = ry * a0 + ((sz * rx) * cw) * 2.0f;
        float o2 = rz * a0 + (sz * (sz * rz)) * 2.0f;
        float o3 = bvx * a0 + ((-(sz * bvy)) * cw) * 2.0f;
        float o4 = bvy * a0 + ((sz * bvx) * cw) * 2.0f;
        float o5 = bvz * a0 + (sz * (sz * bvz)) * 2.0f;
        if (clip_on) {
          o0 = clip_to(o0, clip); o1 = clip_to(o1, clip); o2 = clip_to(o2, clip);
          o3 = clip_to(o3, clip); o4 = clip_to(o4, clip); o5 = clip_to(o5, clip);
        }
        float* o = obs_s + u * L::kObs + 6 * J + 2 * D;
        o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3; o[4] = o4; o[5] = o5;
      }
    }
    rotate_bodies(kItW0 + kItW3, kIters);
    PPK_STAMP(3);
  }

  // ---- store the tile's obs rows ----------------------------------------------------------------------
  float* g_obs = k.obs + (size_t)env0 * H * L::kObs;
  const bool bulk_store = fast && (k.bulk_ok & 2);
  if (bulk_store) {
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
      bulk_s2g_out(g_obs, obs_s, 4u * L::kUnits * L::kObs);
      bulk_commit();
    }
  } else {
    __syncthreads();
    const int nfl = nvalid * H * L::kObs;
    for (int f = tid; f < nfl; f += kFamilyThreads) st_stream(g_obs + f, obs_s[f]);
  }
  // ---- learner side, opt-in: column moments of the tile for RunningMeanStd, thread = column,
  // while the rows are still on chip -- the normaliser's update never re-reads obs_buf
  if ((phases & PPK_PHASE_MOMENTS) && tid < L::kObs) {
    static_assert(L::kObs <= kFamilyThreads, "one thread per obs column");
    double s = 0.0, ss = 0.0;
    const int rows = nvalid * H;
#pragma unroll 4
    for (int u = 0; u < rows; ++u) {
      const double v = (double)obs_s[u * L::kObs + tid];
      s += v;
      ss += v * v;
    }
    double* slot = k.moments + (size_t)(blockIdx.x % PPK_MOMENT_SLOTS) * 2 * L::kObs;
    atomicAdd(slot + tid, s);
    atomicAdd(slot + L::kObs + tid, ss);
  }
  if (bulk_store && tid == 0) bulk_wait_read();
}

}  // namespace ppk
