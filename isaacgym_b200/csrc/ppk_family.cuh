// Fused task step for the 3-/4-actor variants (A3, TILT, NES, ALIGN, A4).
//
// One CTA of three warps owns one tile of TILE consecutive envs.
//   stage    The PhysX tensors are AoS with 52-byte rows, so per-env vector loads are impossible.
//            Instead every env's rigid-body rows ids[1..J) (one contiguous 468-byte run) and row
//            ids[0] are fetched with 1-D bulk async copies (cp.async.bulk, the TMA engine) of the
//            enclosing 16-byte-aligned windows (480 B / 64 B), and the tile's slices of the root,
//            DOF and DOF-force tensors (contiguous across envs) with one bulk copy each.  All
//            copies complete on one mbarrier; no registers or LSU issue slots are spent on staging
//            and each 32-byte sector is fetched once.  Tail tiles / misaligned tensors use a plain
//            LDG path into the same layout.
//   warp 0   lane = env: progress+1, reward, die/time-out mask, flag updates, statistics, the
//            predicated reset (root/DOF rows rewritten from the initial tensors), ball in the
//            heading frame, and the obs "tail" (dof_pos, 0.1*dof_vel, ball).
//   warps 1,2  lane = (env, body): rotate pos/vel of the J bodies into the heading frame,
//            transpose inside the warp with shuffles and store obs row segments contiguously.
#pragma once
#include "ppk_async.cuh"
#include "ppk_device.cuh"

namespace ppk {

constexpr int kFamilyThreads = 96;

template <int H, int J, int D, int A, int TILE>
struct FamilyLayout {
  static constexpr int kSpanRows = J - 1;                               // rows ids[1..J)
  static constexpr int kSpanFloats = ((kSpanRows * kRow + 3 + 3) / 4) * 4;   // 120: run + alignment slack
  static constexpr int kRow0Floats = 16;                                // 10 used floats + slack
  static constexpr int kRootEnv = A * kRow;
  static constexpr int kTail = 2 * D + 6;              // dof_pos, 0.1*dof_vel, ball local pos, vel
  static constexpr int kSTail = kTail | 1;
  static constexpr int kObs = 6 * J + kTail;           // 80 (D=7) / 94 (D=14)
  static constexpr int kHdr = 5;                       // root pos (3) + heading quat (sz, cw)
  // float offsets inside the CTA's shared memory
  static constexpr int kOffRow0 = TILE * H * kSpanFloats;
  static constexpr int kOffRoot = kOffRow0 + TILE * H * kRow0Floats;
  static constexpr int kOffDof = kOffRoot + TILE * kRootEnv;
  static constexpr int kOffForce = kOffDof + TILE * 2 * D;
  static constexpr int kOffBar = kOffForce + TILE * D;            // 8-byte mbarrier
  static constexpr int kOffHdr = kOffBar + 4;
  static constexpr int kFloats = kOffHdr + 2 * TILE * H * kHdr;   // one heading table per obs warp
  static constexpr uint32_t kTxBytes =
      4u * (TILE * H * (kSpanFloats + kRow0Floats) + TILE * kRootEnv + TILE * 2 * D + TILE * D);
  static_assert((TILE * kRootEnv) % 4 == 0 && (TILE * 2 * D) % 4 == 0 && (TILE * D) % 4 == 0, "16-byte bulk sizes");
  static_assert(kOffBar % 2 == 0, "mbarrier alignment");
  static_assert(TILE * H * kSTail <= kOffBar - kOffRoot, "tail does not fit its alias region");
};

template <int V, int H, int J, int D, int A, int TILE>
__global__ void __launch_bounds__(kFamilyThreads, 8)
family_step_kernel(const __grid_constant__ KArgs k) {
  using L = FamilyLayout<H, J, D, A, TILE>;
  extern __shared__ __align__(128) float smem[];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long env0 = (long long)blockIdx.x * TILE;
  const int nvalid = (int)min((long long)TILE, k.n - env0);

  float* span_s = smem;
  float* row0_s = smem + L::kOffRow0;
  float* root_s = smem + L::kOffRoot;
  float* dof_s = smem + L::kOffDof;
  float* force_s = smem + L::kOffForce;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  float* tail_s = root_s;

  const int phases = k.phases;
  const bool bulk = k.bulk_ok && (nvalid == TILE);
  const int env_stride = k.B * kRow;
  const float* g_rb = k.rb + (size_t)env0 * env_stride;

  // position of an env's run / row inside its 16-byte aligned staging window
  auto span_off = [&](int e, int h) -> int {
    if (!bulk) return 0;
    return (int)((reinterpret_cast<uintptr_t>(g_rb + (size_t)e * env_stride + k.ids[h][1] * kRow) & 15u) >> 2);
  };
  auto row0_off = [&](int e, int h) -> int {
    if (!bulk) return 0;
    return (int)((reinterpret_cast<uintptr_t>(g_rb + (size_t)e * env_stride + k.ids[h][0] * kRow) & 15u) >> 2);
  };

  // ---- stage ---------------------------------------------------------------------------------------
  if (bulk) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (warp == 0) {
      if (lane == 0) {
        mbar_arrive_expect_tx(bar, L::kTxBytes);
        bulk_g2s(root_s, k.root + (size_t)env0 * L::kRootEnv, 4u * TILE * L::kRootEnv, bar);
        bulk_g2s(dof_s, k.dof + (size_t)env0 * 2 * D, 4u * TILE * 2 * D, bar);
        bulk_g2s(force_s, k.force + (size_t)env0 * D, 4u * TILE * D, bar);
      }
      __syncwarp();
      for (int u = lane; u < TILE * H; u += 32) {       // one (env, humanoid) pair per lane
        const int e = u / H, h = u - e * H;
        const float* row = g_rb + (size_t)e * env_stride;
        const uintptr_t a1 = reinterpret_cast<uintptr_t>(row + k.ids[h][1] * kRow) & ~(uintptr_t)15;
        const uintptr_t a0 = reinterpret_cast<uintptr_t>(row + k.ids[h][0] * kRow) & ~(uintptr_t)15;
        bulk_g2s(span_s + u * L::kSpanFloats, reinterpret_cast<const void*>(a1), 4u * L::kSpanFloats, bar);
        bulk_g2s(row0_s + u * L::kRow0Floats, reinterpret_cast<const void*>(a0), 4u * L::kRow0Floats, bar);
      }
    }
  } else {
    // generic path (tail tile, unaligned tensors, non-consecutive ids): plain loads, same layout
    for (int f = threadIdx.x; f < TILE * H * J * kRow; f += kFamilyThreads) {
      const int u = f / (J * kRow), r = f - u * (J * kRow);
      const int e = u / H, h = u - e * H;
      const int j = r / kRow, c = r - j * kRow;
      float v = (e < nvalid) ? g_rb[(size_t)e * env_stride + k.ids[h][j] * kRow + c] : 0.0f;
      if (j == 0) row0_s[u * L::kRow0Floats + c] = v;
      else span_s[u * L::kSpanFloats + (j - 1) * kRow + c] = v;
    }
    for (int f = threadIdx.x; f < TILE * L::kRootEnv; f += kFamilyThreads)
      root_s[f] = (f < nvalid * L::kRootEnv) ? k.root[(size_t)env0 * L::kRootEnv + f] : 0.0f;
    for (int f = threadIdx.x; f < TILE * 2 * D; f += kFamilyThreads)
      dof_s[f] = (f < nvalid * 2 * D) ? k.dof[(size_t)env0 * 2 * D + f] : 0.0f;
    for (int f = threadIdx.x; f < TILE * D; f += kFamilyThreads)
      force_s[f] = (f < nvalid * D) ? k.force[(size_t)env0 * D + f] : 0.0f;
    __syncthreads();
  }

  if (warp != 0) {
    // ================= warps 1, 2: body observations =============================================
    if (!(phases & PPK_PHASE_OBS)) return;
    if (bulk) mbar_wait(bar, 0);
    float* hdr_s = smem + L::kOffHdr + (warp - 1) * (TILE * H * L::kHdr);
    // heading frames, lane = (env, humanoid)
    for (int u = lane; u < TILE * H; u += 32) {
      const int e = u / H, h = u - e * H;
      const float* r0 = row0_s + u * L::kRow0Floats + row0_off(e, h);
      Heading hq = heading_quat_inv(r0[3], r0[4], r0[5], r0[6]);
      float* hd = hdr_s + u * L::kHdr;
      hd[0] = r0[0]; hd[1] = r0[1]; hd[2] = r0[2]; hd[3] = hq.sz; hd[4] = hq.cw;
    }
    __syncwarp();
    constexpr int P = 32 / J;                       // envs per pass
    constexpr int kPasses = (TILE + P - 1) / P;
    const int a = lane / J, t = lane - a * J;
    const bool lane_on = a < P;
    float* g_obs = k.obs + (size_t)env0 * H * L::kObs;
#pragma unroll 1
    for (int unit = warp - 1; unit < kPasses * H; unit += 2) {
      const int pass = unit / H, h = unit - pass * H;
      const int e = pass * P + a;
      const bool ok = lane_on && (e < nvalid);
      const int u = (ok ? e : 0) * H + h;
      const float* hd = hdr_s + u * L::kHdr;
      Heading hq; hq.sz = hd[3]; hq.cw = hd[4];
      const int tt = lane_on ? t : 0;
      const float* row = (tt == 0) ? (row0_s + u * L::kRow0Floats + row0_off(ok ? e : 0, h))
                                   : (span_s + u * L::kSpanFloats + span_off(ok ? e : 0, h) + (tt - 1) * kRow);
      float lp[3], lv[3];
      rotate_heading(hq, row[0] - hd[0], row[1] - hd[1], row[2] - hd[2], lp[0], lp[1], lp[2]);
      rotate_heading(hq, row[7], row[8], row[9], lv[0], lv[1], lv[2]);
      float* orow = g_obs + ((size_t)e * H + h) * L::kObs;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        // output position o = t + i*J of this env's 3J-float segment comes from body o/3, component o%3
        const int o = tt + i * J;
        const int src = (lane_on ? a * J : 0) + o / 3, comp = o - (o / 3) * 3;
        float x = __shfl_sync(full, lp[0], src), y = __shfl_sync(full, lp[1], src), z = __shfl_sync(full, lp[2], src);
        float pv = comp == 0 ? x : (comp == 1 ? y : z);
        x = __shfl_sync(full, lv[0], src); y = __shfl_sync(full, lv[1], src); z = __shfl_sync(full, lv[2], src);
        float vv = comp == 0 ? x : (comp == 1 ? y : z);
        if (ok) { st_stream(orow + o, pv); st_stream(orow + 3 * J + o, vv); }
      }
    }
    return;
  }

  // ================= warp 0: reward / reset / tail, lane = env ==========================================
  const bool lane_env = lane < nvalid;
  const long long env = env0 + lane;
  long long prog = 0, reset_prev = 0;
  float pre_vx = 0.0f, pre_vz = 0.0f;
  constexpr int NF = (V == PPK_TILT) ? 3 : (V == PPK_A4) ? 6 : (V == PPK_NES) ? 2 : (V == PPK_ALIGN) ? 1 : 0;
  bool flag[NF > 0 ? NF : 1];
  if (lane_env) {       // per-env scalars come straight from global and overlap the bulk copies
    prog = k.progress[env];
    if (!(phases & PPK_PHASE_REWARD)) reset_prev = k.reset[env];
    if (phases & PPK_PHASE_REWARD) {
      const float* p = k.pre + (size_t)env * k.pre_stride;
      pre_vx = ld_stream(p + k.pre_vx);
      if (V == PPK_ALIGN) pre_vz = ld_stream(p + k.pre_vz);
#pragma unroll
      for (int i = 0; i < NF; ++i) flag[i] = k.flags[i][env] != 0;
    }
  }
  if (bulk) mbar_wait(bar, 0);

  const int le = (lane < TILE) ? lane : 0;     // smem row this lane reads (idle lanes read row 0)
  const float* my_root = root_s + le * L::kRootEnv;
  const float* ball = my_root + k.ball * kRow;
  float bx = ball[0], by = ball[1], bz = ball[2];
  float bvx = ball[7], bvy = ball[8], bvz = ball[9];
  float hx[H];
#pragma unroll
  for (int h = 0; h < H; ++h) hx[h] = my_root[k.hum[h] * kRow];
  float dofv[2 * D];
#pragma unroll
  for (int i = 0; i < 2 * D; ++i) dofv[i] = dof_s[le * 2 * D + i];
  float power = 0.0f;
  if (phases & PPK_PHASE_REWARD) {
#pragma unroll
    for (int d = 0; d < D; ++d) power += fabsf(force_s[le * D + d] * dofv[2 * d + 1]);
  }

  long long p_new = prog + ((phases & PPK_PHASE_PROGRESS) ? 1 : 0);
  bool is_reset = reset_prev != 0;
  float rew[H];
#pragma unroll
  for (int h = 0; h < H; ++h) rew[h] = 0.0f;

  if (phases & PPK_PHASE_REWARD) {
    bool die = false;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      Scene s;
      s.bx = bx; s.by = by; s.bz = bz; s.vx = bvx; s.vz = bvz;
      s.pre_vx = pre_vx; s.pre_vz = pre_vz;
      const int pj = k.paddle_j[h];
      const float* pd;
      if (pj == 0) pd = row0_s + (le * H + h) * L::kRow0Floats + row0_off(le, h);
      else if (pj > 0) pd = span_s + (le * H + h) * L::kSpanFloats + span_off(le, h) + (pj - 1) * kRow;
      else pd = k.rb + ((size_t)(lane_env ? env : env0) * k.B + k.paddle_body[h]) * kRow;
      s.px = pd[0]; s.py = pd[1]; s.pz = pd[2];
      s.hx = hx[h];
      s.power_reward = (-k.power_coef) * power;
      s.progress = p_new;
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

  // ---- predicated reset (TILT:847-906): lanes whose env resets rewrite its root / DOF rows -------
  const bool do_reset = (phases & PPK_PHASE_RESET) && is_reset && lane_env;
  if (do_reset) {
    const float* ir = k.init_root + (size_t)env * L::kRootEnv;
    float* gr = k.root + (size_t)env * L::kRootEnv;
    const float* rv = k.reset_vel + (size_t)env * 3;
    float nvx = rv[0], nvy = rv[1], nvz = rv[2];
#pragma unroll
    for (int a = 0; a < A; ++a) {
#pragma unroll
      for (int c = 0; c < 7; ++c) gr[a * kRow + c] = ir[a * kRow + c];
#pragma unroll
      for (int c = 7; c < kRow; ++c) gr[a * kRow + c] = 0.0f;
    }
    gr[k.ball * kRow + 7] = nvx; gr[k.ball * kRow + 8] = nvy; gr[k.ball * kRow + 9] = nvz;
    bx = ir[k.ball * kRow + 0]; by = ir[k.ball * kRow + 1]; bz = ir[k.ball * kRow + 2];
    bvx = nvx; bvy = nvy; bvz = nvz;
    if (k.reset_dof) {
      const float* id = k.init_dof + (size_t)env * 2 * D;
      float* gd = k.dof + (size_t)env * 2 * D;
#pragma unroll
      for (int i = 0; i < 2 * D; ++i) { dofv[i] = id[i]; gd[i] = dofv[i]; }
    }
    p_new = 0;
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
  if (!(phases & PPK_PHASE_OBS)) return;

  // ---- ball in the heading frame; tail of the obs row goes through smem -----------------------------
  float tail_ball[H][6];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const float* r0 = row0_s + (le * H + h) * L::kRow0Floats + row0_off(le, h);
    Heading hq = heading_quat_inv(r0[3], r0[4], r0[5], r0[6]);
    rotate_heading(hq, bx - r0[0], by - r0[1], bz - r0[2], tail_ball[h][0], tail_ball[h][1], tail_ball[h][2]);
    rotate_heading(hq, bvx, bvy, bvz, tail_ball[h][3], tail_ball[h][4], tail_ball[h][5]);
  }
  __syncwarp();   // every lane of this warp is done with root_s / dof_s / force_s (only warp 0 reads them)
  if (lane < TILE) {
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float* t = tail_s + (le * H + h) * L::kSTail;
#pragma unroll
      for (int d = 0; d < D; ++d) { t[d] = dofv[2 * d]; t[D + d] = dofv[2 * d + 1] * 0.1f; }
#pragma unroll
      for (int c = 0; c < 6; ++c) t[2 * D + c] = tail_ball[h][c];
    }
  }
  __syncwarp();
  // tail: dof_pos, 0.1*dof_vel, ball local pos/vel -- kTail contiguous floats per (env, humanoid)
  float* g_obs = k.obs + (size_t)env0 * H * L::kObs;
  constexpr int kTailIt = (TILE * H * L::kTail + 31) / 32;
#pragma unroll 4
  for (int i = 0; i < kTailIt; ++i) {
    int f = i * 32 + lane;
    int eh = f / L::kTail, kk = f - eh * L::kTail;
    int e = eh / H;
    if (e < nvalid) st_stream(g_obs + (size_t)eh * L::kObs + 6 * J + kk, tail_s[eh * L::kSTail + kk]);
  }
}

}  // namespace ppk
