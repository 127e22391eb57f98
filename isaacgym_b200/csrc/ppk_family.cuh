// Fused task step for the 3-/4-actor variants (A3, TILT, NES, ALIGN, A4).
//
// One warp owns one tile of TILE consecutive envs and runs the whole step for it:
//   stage   coalesced loads of the tile's rigid-body rows (ids[]), root rows, DOF state and
//           DOF forces into the warp's shared-memory slice (the state tensors are AoS with
//           52-byte rows, so per-env vector loads are impossible; lanes walk the flat float
//           index instead and every 32-byte sector is fetched exactly once);
//   phase R lane = env: progress+1, reward, die/time-out mask, flag updates, statistics,
//           predicated reset (root/DOF rows rewritten from the initial tensors), heading frame,
//           ball in the heading frame; the obs "tail" (dof_pos, 0.1*dof_vel, ball) goes to smem;
//   phase O lane = (env, body): rotate pos/vel of the J bodies into the heading frame, transpose
//           inside the warp with shuffles and store each obs row segment contiguously.
// No block-level synchronisation: warps are independent (only __syncwarp).
#pragma once
#include "ppk_device.cuh"

namespace ppk {

template <int H, int J, int D, int A, int TILE>
struct FamilyLayout {
  static constexpr int kRbEnv = H * J * kRow;          // staged rigid-body floats per env
  static constexpr int kRootEnv = A * kRow;
  static constexpr int kSRb = kRbEnv | 1;              // odd strides: lane = env reads are conflict-free
  static constexpr int kSRoot = kRootEnv | 1;
  static constexpr int kSDof = (2 * D) | 1;
  static constexpr int kSForce = D | 1;
  static constexpr int kHdr = 5;                       // root pos (3) + heading quat (sz, cw)
  static constexpr int kTail = 2 * D + 6;              // dof_pos, 0.1*dof_vel, ball local pos, vel
  static constexpr int kSTail = kTail | 1;
  static constexpr int kObs = 6 * J + kTail;           // 80 (D=7) / 94 (D=14)
  static constexpr int kOffRoot = TILE * kSRb;
  static constexpr int kOffDof = kOffRoot + TILE * kSRoot;
  static constexpr int kOffForce = kOffDof + TILE * kSDof;
  static constexpr int kOffHdr = kOffForce + TILE * kSForce;
  static constexpr int kWarpFloats = kOffHdr + TILE * H * kHdr;
  // the tail aliases the root/dof/force staging, all of which phase R has consumed by then
  static_assert(TILE * H * kSTail <= kOffHdr - kOffRoot, "tail does not fit its alias region");
  static_assert((TILE * kRbEnv) % 32 == 0 && (TILE * kRootEnv) % 32 == 0 && (TILE * 2 * D) % 32 == 0 &&
                    (TILE * D) % 32 == 0, "flat staging loops assume whole warps");
};

constexpr int kFamilyWarps = 4;   // warps (= tiles) per CTA

template <int V, int H, int J, int D, int A, int TILE>
__global__ void __launch_bounds__(kFamilyWarps * 32)
family_step_kernel(const __grid_constant__ KArgs k) {
  using L = FamilyLayout<H, J, D, A, TILE>;
  extern __shared__ float smem[];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long tile = (long long)blockIdx.x * kFamilyWarps + warp;
  const long long env0 = tile * TILE;
  if (env0 >= k.n) return;
  const int nvalid = (int)min((long long)TILE, k.n - env0);

  float* rb_s = smem + (size_t)warp * L::kWarpFloats;
  float* root_s = rb_s + L::kOffRoot;
  float* dof_s = rb_s + L::kOffDof;
  float* force_s = rb_s + L::kOffForce;
  float* hdr_s = rb_s + L::kOffHdr;
  float* tail_s = root_s;

  const int phases = k.phases;
  const bool lane_env = lane < nvalid;          // this lane owns env0 + lane in the lane = env phases
  const long long env = env0 + lane;

  // ---- per-env scalars straight from global (already one value per env) ----------------------
  long long prog = 0;
  long long reset_prev = 0;
  float pre_vx = 0.0f, pre_vz = 0.0f;
  constexpr int NF = (V == PPK_TILT) ? 3 : (V == PPK_A4) ? 6 : (V == PPK_NES) ? 2 : (V == PPK_ALIGN) ? 1 : 0;
  bool flag[NF > 0 ? NF : 1];
  if (lane_env) {
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

  // ---- stage the tile --------------------------------------------------------------------------
  {
    // root / dof / force: the tile's slice of each tensor is one contiguous run
    constexpr int kRootIt = TILE * L::kRootEnv / 32, kDofIt = TILE * 2 * D / 32, kForceIt = TILE * D / 32;
    const float* g_root = k.root + (size_t)env0 * L::kRootEnv;
    const float* g_dof = k.dof + (size_t)env0 * 2 * D;
    const float* g_force = k.force + (size_t)env0 * D;
    float v[kRootIt + kDofIt + kForceIt];
#pragma unroll
    for (int i = 0; i < kRootIt; ++i) {
      int f = i * 32 + lane;
      v[i] = (f < nvalid * L::kRootEnv) ? ld_stream(g_root + f) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < kDofIt; ++i) {
      int f = i * 32 + lane;
      v[kRootIt + i] = (f < nvalid * 2 * D) ? ld_stream(g_dof + f) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < kForceIt; ++i) {
      int f = i * 32 + lane;
      v[kRootIt + kDofIt + i] = (f < nvalid * D) ? ld_stream(g_force + f) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < kRootIt; ++i) {
      int f = i * 32 + lane;
      int e = f / L::kRootEnv, r = f - e * L::kRootEnv;
      root_s[e * L::kSRoot + r] = v[i];
    }
#pragma unroll
    for (int i = 0; i < kDofIt; ++i) {
      int f = i * 32 + lane;
      int e = f / (2 * D), r = f - e * (2 * D);
      dof_s[e * L::kSDof + r] = v[kRootIt + i];
    }
#pragma unroll
    for (int i = 0; i < kForceIt; ++i) {
      int f = i * 32 + lane;
      int e = f / D, r = f - e * D;
      force_s[e * L::kSForce + r] = v[kRootIt + kDofIt + i];
    }
  }
  {
    // rigid-body rows ids[h][0..J): lanes walk the flat (env, humanoid, body, column) index, so
    // consecutive ids (31..39) give consecutive addresses and fully coalesced requests
    constexpr int kIt = TILE * L::kRbEnv / 32;    // 130
    constexpr int kBatch = (kIt % 65 == 0) ? 65 : (kIt % 26 == 0) ? 26 : (kIt % 13 == 0 ? 13 : 1);
    const int my_id0 = (lane < J) ? k.ids[0][lane] : 0;
    const int my_id1 = (H > 1 && lane < J) ? k.ids[1][lane] : 0;
    const float* g_rb = k.rb + (size_t)env0 * k.B * kRow;
    const int env_stride = k.B * kRow;
#pragma unroll 1
    for (int it0 = 0; it0 < kIt; it0 += kBatch) {
      float v[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        int f = (it0 + u) * 32 + lane;
        int e = f / L::kRbEnv, r = f - e * L::kRbEnv;
        int h = r / (J * kRow), rr = r - h * (J * kRow);
        int j = rr / kRow, c = rr - j * kRow;
        // every lane must offer both lists: the source lane's own h says nothing about ours
        int id = __shfl_sync(full, my_id0, j);
        if (H > 1) {
          int id1 = __shfl_sync(full, my_id1, j);
          if (h) id = id1;
        }
        v[u] = (e < nvalid) ? ld_stream(g_rb + (size_t)e * env_stride + id * kRow + c) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        int f = (it0 + u) * 32 + lane;
        int e = f / L::kRbEnv, r = f - e * L::kRbEnv;
        rb_s[e * L::kSRb + r] = v[u];
      }
    }
  }
  __syncwarp();

  // ---- phase R: lane = env ---------------------------------------------------------------------
  const int le = (lane < TILE) ? lane : 0;     // smem row this lane reads (idle lanes read row 0)
  const float* my_root = root_s + le * L::kSRoot;
  const float* ball = my_root + k.ball * kRow;
  float bx = ball[0], by = ball[1], bz = ball[2];
  float bvx = ball[7], bvy = ball[8], bvz = ball[9];
  float dofv[2 * D];
#pragma unroll
  for (int i = 0; i < 2 * D; ++i) dofv[i] = dof_s[le * L::kSDof + i];

  long long p_new = prog + ((phases & PPK_PHASE_PROGRESS) ? 1 : 0);
  bool is_reset = reset_prev != 0;
  float rew[H];
#pragma unroll
  for (int h = 0; h < H; ++h) rew[h] = 0.0f;

  if (phases & PPK_PHASE_REWARD) {
    float power = 0.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) power += fabsf(force_s[le * L::kSForce + d] * dofv[2 * d + 1]);
    bool die = false;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      Scene s;
      s.bx = bx; s.by = by; s.bz = bz; s.vx = bvx; s.vz = bvz;
      s.pre_vx = pre_vx; s.pre_vz = pre_vz;
      if (k.paddle_j[h] >= 0) {
        const float* pd = rb_s + le * L::kSRb + (h * J + k.paddle_j[h]) * kRow;
        s.px = pd[0]; s.py = pd[1]; s.pz = pd[2];
      } else {
        const float* pd = k.rb + ((size_t)(lane_env ? env : env0) * k.B + k.paddle_body[h]) * kRow;
        s.px = pd[0]; s.py = pd[1]; s.pz = pd[2];
      }
      s.hx = my_root[k.hum[h] * kRow];
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
    double s_rew = lane_env ? (double)rew[0] : 0.0;
    double s_prog = lane_env ? (double)p_new : 0.0;
    double s_rst = (lane_env && is_reset) ? 1.0 : 0.0;
    s_rew = warp_sum(s_rew); s_prog = warp_sum(s_prog); s_rst = warp_sum(s_rst);
    if (lane == 0) {
      double* slot = k.stats + (size_t)(tile % PPK_STATS_SLOTS) * PPK_NUM_STATS;
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

  // ---- heading frame + ball in the frame; tail of the obs row goes through smem ------------------
  float tail_ball[H][6];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const float* r0 = rb_s + le * L::kSRb + h * J * kRow;   // body ids[h][0]: the heading / root body
    float rx = r0[0], ry = r0[1], rz = r0[2];
    Heading hq = heading_quat_inv(r0[3], r0[4], r0[5], r0[6]);
    if (lane < TILE) {
      float* hd = hdr_s + (le * H + h) * L::kHdr;
      hd[0] = rx; hd[1] = ry; hd[2] = rz; hd[3] = hq.sz; hd[4] = hq.cw;
    }
    rotate_heading(hq, bx - rx, by - ry, bz - rz, tail_ball[h][0], tail_ball[h][1], tail_ball[h][2]);
    rotate_heading(hq, bvx, bvy, bvz, tail_ball[h][3], tail_ball[h][4], tail_ball[h][5]);
  }
  __syncwarp();   // every lane is done with root_s / dof_s / force_s: the tail may overwrite them
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

  // ---- phase O: lane = (env a of the pass, body t) ---------------------------------------------
  constexpr int P = 32 / J;                       // envs per pass
  const int a = lane / J, t = lane - a * J;
  float* g_obs = k.obs + (size_t)env0 * H * L::kObs;
#pragma unroll 1
  for (int pass = 0; pass * P < TILE; ++pass) {
    const int e = pass * P + a;
    const bool ok = (a < P) && (e < nvalid);
    const int ec = ok ? e : 0;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float* hd = hdr_s + (ec * H + h) * L::kHdr;
      Heading hq; hq.sz = hd[3]; hq.cw = hd[4];
      const float* row = rb_s + ec * L::kSRb + (h * J + (a < P ? t : 0)) * kRow;
      float lp[3], lv[3];
      rotate_heading(hq, row[0] - hd[0], row[1] - hd[1], row[2] - hd[2], lp[0], lp[1], lp[2]);
      rotate_heading(hq, row[7], row[8], row[9], lv[0], lv[1], lv[2]);
      float* orow = g_obs + ((size_t)e * H + h) * L::kObs;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        // output position o = t + i*J of this env's 3J-float segment comes from body o/3, component o%3
        const int o = t + i * J;
        const int src = (a < P ? a * J : 0) + o / 3, comp = o - (o / 3) * 3;
        float x = __shfl_sync(full, lp[0], src), y = __shfl_sync(full, lp[1], src), z = __shfl_sync(full, lp[2], src);
        float pv = comp == 0 ? x : (comp == 1 ? y : z);
        x = __shfl_sync(full, lv[0], src); y = __shfl_sync(full, lv[1], src); z = __shfl_sync(full, lv[2], src);
        float vv = comp == 0 ? x : (comp == 1 ? y : z);
        if (ok) { st_stream(orow + o, pv); st_stream(orow + 3 * J + o, vv); }
      }
    }
  }
  // tail: dof_pos, 0.1*dof_vel, ball local pos/vel -- kTail contiguous floats per (env, humanoid)
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
