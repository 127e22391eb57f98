// Fused task step of the 27-DOF variant (humanoid_pingpong_3_actor_all_dof.py, ADOF).
//
// Per env the step consumes rows 0..39 of the rigid-body tensor (2080 contiguous bytes), rows
// 0..27 of the initial (reference-pose) rigid-body tensor, the root block, 27 DOF states, their
// reference and the DOF forces, and writes a 313-float observation row.  One CTA of four warps owns
// a tile of 8 envs:
//   stage      1-D bulk async copies (cp.async.bulk / TMA engine) on one mbarrier: per env the
//              16-byte aligned windows around its live rows (2096 B) and reference rows (1472 B),
//              per tile the contiguous root / DOF / reference-DOF / force slices.  Tail tiles and
//              misaligned tensors take an LDG path into the same layout.  COMPACT: the reference pose
//              is a constant of the task (ADOF:196-200), so the caller may hand it over repacked once
//              at init as [N,23,6] (pos, linvel of the balance bodies): 552 B per env in one bulk
//              copy per tile instead of a 1472 B window per env.
//   warps 1-4  one env per pass, lane = balance body (23) and lane = DOF (27): imitation diffs, the
//              seven per-env reductions of compute_imitation_reward (ADOF:1313-1418) in one fused
//              butterfly, heading frame, imitation observation segments (ADOF:1891-1927) and the ten
//              ping-pong bodies in the heading frame (ADOF:1849-1888, lane = output float);
//   warp 0     lane = env: compute_pingpong_reward_nv
//              (ADOF:1440-1690) + compute_gradient_penalty (ADOF:1245-1301), flags, counters,
//              time-out mask, predicated reset (ADOF:965-1028); then the dof / ball / reference-dof
//              segments of the obs row, lane = element.
#pragma once
#include <stdlib.h>

#include "ppk_async.cuh"
#include "ppk_device.cuh"

namespace ppk {

constexpr int kAdofTile = 8;
constexpr int kAdofObsWarps = 4;            // warps 1..4 reduce / rotate the balance bodies, two envs each
constexpr int kAdofThreads = 32 * (1 + kAdofObsWarps);
constexpr int kAdofD = 27;
constexpr int kAdofJ = 10;       // ping-pong bodies
constexpr int kAdofNB = 23;      // balance bodies
constexpr int kAdofRbRows = 40;  // rows 0..39 staged from the live tensor
constexpr int kAdofInitRows = 28;  // rows 0..27 staged from the reference pose
constexpr int kAdofObs = 6 * kAdofJ + 2 * kAdofD + 7 + 6 * kAdofNB + 2 * kAdofD;   // 313

template <bool COMPACT>
struct AdofLayout {
  static constexpr int kRbEnv = kAdofRbRows * kRow;      // 520 floats used
  static constexpr int kInitEnv = COMPACT ? kAdofNB * 6 : kAdofInitRows * kRow;  // 138 | 364
  static constexpr int kSRb = ((kRbEnv + 3 + 3) / 4) * 4;      // 524: window incl. alignment slack
  static constexpr int kSInit = COMPACT ? kInitEnv : ((kInitEnv + 3 + 3) / 4) * 4;  // 138 (dense) | 368
  static constexpr int kRoot = 3 * kRow;                 // 39, dense
  static constexpr int kDof = 2 * kAdofD;                // 54, dense
  static constexpr int kSHdr = 24;
  static constexpr int kOffInit = kAdofTile * kSRb;
  static constexpr int kOffRoot = kOffInit + kAdofTile * kSInit;
  static constexpr int kOffDof = kOffRoot + kAdofTile * kRoot;
  static constexpr int kOffIDof = kOffDof + kAdofTile * kDof;
  static constexpr int kOffForce = kOffIDof + kAdofTile * kDof;
  static constexpr int kOffHdr = kOffForce + kAdofTile * kAdofD;
  static constexpr int kOffBar = kOffHdr + kAdofTile * kSHdr;
  static constexpr int kFloats = kOffBar + 4;
  static constexpr uint32_t kTxBytes =
      4u * (kAdofTile * (kSRb + kSInit) + kAdofTile * kRoot + 2 * kAdofTile * kDof + kAdofTile * kAdofD);
  static_assert((kAdofTile * kRoot) % 4 == 0 && (kAdofTile * kDof) % 4 == 0 && (kAdofTile * kAdofD) % 4 == 0, "bulk sizes");
  static_assert(kOffInit % 4 == 0 && kOffRoot % 4 == 0 && kOffDof % 4 == 0 && kOffIDof % 4 == 0 && kOffForce % 4 == 0,
                "bulk destinations are 16-byte aligned");
  static_assert(kOffBar % 2 == 0, "mbarrier alignment");
};
// hdr slots per env
enum { H_A0 = 0, H_SZ, H_CW, H_RX, H_RY, H_RZ,        // ping-pong heading frame (body ids[0])
       H_SUM_DP2, H_SUM_DV2, H_SUM_NORM,              // balance-body reductions
       H_SUM_DQ22, H_SUM_DQ5, H_SUM_DQD22, H_POWER,   // DOF reductions
       H_BALL0, H_BALL1, H_BALL2, H_BALL3, H_BALL4, H_BALL5, H_BALL6 };

// named barrier 2: warps 1-3 arrive when the per-env reductions / frames are in hdr_s, warp 0 waits
__device__ __forceinline__ void adof_arrive() { asm volatile("bar.arrive 2, %0;" ::"n"(kAdofThreads) : "memory"); }
__device__ __forceinline__ void adof_wait() { asm volatile("bar.sync 2, %0;" ::"n"(kAdofThreads) : "memory"); }

constexpr uint32_t kPhaseDeferCounterClear = 1u << 8;   // internal (host session): clear counters once per shard

#ifndef PPK_ADOF_MINB
#define PPK_ADOF_MINB 8        // CTAs per SM with the compact reference pose (27 KB smem, 48 registers)
#endif
template <bool COMPACT, bool CLIP = false>
__global__ void __launch_bounds__(kAdofThreads, COMPACT ? PPK_ADOF_MINB : 6)
adof_step_kernel(const __grid_constant__ KArgs k) {
  using L = AdofLayout<COMPACT>;
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float clip = CLIP ? k.clip_obs : 0.0f;      // VecTask.step's observation clamp: its own instantiation, free when off
  const long long env0 = (long long)blockIdx.x * kAdofTile;
  const int nvalid = (int)min((long long)kAdofTile, k.n - env0);
  constexpr int D = kAdofD, J = kAdofJ, NB = kAdofNB, T = kAdofTile;

  float* rb_s = smem;
  float* init_s = smem + L::kOffInit;
  float* root_s = smem + L::kOffRoot;
  float* dof_s = smem + L::kOffDof;
  float* idof_s = smem + L::kOffIDof;
  float* force_s = smem + L::kOffForce;
  float* hdr_s = smem + L::kOffHdr;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  const int phases = k.phases;
  const bool bulk = k.bulk_ok && (nvalid == T);
  const int env_stride = k.B * kRow;
  const float* g_rb = k.rb + (size_t)env0 * env_stride;
  const float* g_init = COMPACT ? k.init_bal + (size_t)env0 * L::kInitEnv : k.init_rb + (size_t)env0 * env_stride;

  // position of env e's row 0 inside its 16-byte aligned staging window (floats)
  const int off_rb0 = (int)((reinterpret_cast<uintptr_t>(g_rb) & 15u) >> 2);
  const int off_init0 = (int)((reinterpret_cast<uintptr_t>(g_init) & 15u) >> 2);
  auto win_off = [&](const float* base, int e) -> int {
    if (COMPACT && base != g_rb) return 0;
    return bulk ? (((base == g_rb ? off_rb0 : off_init0) + e * env_stride) & 3) : 0;
  };

  // ---- stage -------------------------------------------------------------------------------
  if (bulk) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
    }
    __syncthreads();
    // soft start of the first wave (see ppk_family.cuh): CTA b delays its copies by (b / #SMs) * stagger cycles
    if (k.stagger > 0 && (int)blockIdx.x < k.first_wave && (int)blockIdx.x >= k.num_sms) {
      const long long wait = (long long)((int)blockIdx.x / k.num_sms) * k.stagger;
      const long long t0 = clock64();
      while (clock64() - t0 < wait) __nanosleep(64);
    }
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(bar, L::kTxBytes);
      bulk_g2s(root_s, k.root + (size_t)env0 * L::kRoot, 4u * T * L::kRoot, bar);
      bulk_g2s(dof_s, k.dof + (size_t)env0 * L::kDof, 4u * T * L::kDof, bar);
      bulk_g2s(idof_s, k.init_dof + (size_t)env0 * L::kDof, 4u * T * L::kDof, bar);
      bulk_g2s(force_s, k.force + (size_t)env0 * D, 4u * T * D, bar);
      if (COMPACT) bulk_g2s(init_s, g_init, 4u * T * L::kInitEnv, bar);
    }
    if (warp >= 1) {       // 16 windows, four per reducing warp: warp-uniform addresses, one elected lane issues
      const int wu = __shfl_sync(0xffffffffu, warp, 0);
#pragma unroll
      for (int q = 0; q < T / kAdofObsWarps; ++q) {
        const int e = (wu - 1) * (T / kAdofObsWarps) + q;
        const uintptr_t a_rb = reinterpret_cast<uintptr_t>(g_rb + (size_t)e * env_stride) & ~(uintptr_t)15;
        const uintptr_t a_in = reinterpret_cast<uintptr_t>(g_init + (size_t)e * env_stride) & ~(uintptr_t)15;
        if (elect_one()) {
          bulk_g2s(rb_s + e * L::kSRb, reinterpret_cast<const void*>(a_rb), 4u * L::kSRb, bar);
          if (!COMPACT) bulk_g2s(init_s + e * L::kSInit, reinterpret_cast<const void*>(a_in), 4u * L::kSInit, bar);
        }
      }
    }
  } else {
    for (int f = threadIdx.x; f < T * L::kRbEnv; f += kAdofThreads) {
      const int e = f / L::kRbEnv, r = f - e * L::kRbEnv;
      rb_s[e * L::kSRb + r] = (e < nvalid) ? g_rb[(size_t)e * env_stride + r] : 0.0f;
    }
    for (int f = threadIdx.x; f < T * L::kInitEnv; f += kAdofThreads) {
      const int e = f / L::kInitEnv, r = f - e * L::kInitEnv;
      init_s[e * L::kSInit + r] = (e < nvalid) ? g_init[(size_t)e * (COMPACT ? L::kInitEnv : env_stride) + r] : 0.0f;
    }
    for (int f = threadIdx.x; f < T * L::kRoot; f += kAdofThreads)
      root_s[f] = (f < nvalid * L::kRoot) ? k.root[(size_t)env0 * L::kRoot + f] : 0.0f;
    for (int f = threadIdx.x; f < T * L::kDof; f += kAdofThreads) {
      dof_s[f] = (f < nvalid * L::kDof) ? k.dof[(size_t)env0 * L::kDof + f] : 0.0f;
      idof_s[f] = (f < nvalid * L::kDof) ? k.init_dof[(size_t)env0 * L::kDof + f] : 0.0f;
    }
    for (int f = threadIdx.x; f < T * D; f += kAdofThreads)
      force_s[f] = (f < nvalid * D) ? k.force[(size_t)env0 * D + f] : 0.0f;
    __syncthreads();
  }

  float* g_obs = k.obs + (size_t)env0 * kAdofObs;
  const int pp_root = k.ids[0][0];

  if (warp != 0) {
    // ================= warps 1-3: one env per pass; lane = balance body and lane = DOF ==============
    if (bulk) mbar_wait(bar, 0);
    const int bal_id = (lane < NB) ? k.bal_ids[lane] : k.bal_ids[0];
    const int bal_root = k.bal_ids[0];
    // heading frames of the tile's 8 envs: warp 1 computes them once (lane = env; atan2f / sinf / cosf are ~140
    // instructions whatever the number of active lanes) and publishes them in hdr_s for the other obs warps
    if (warp == 1 && lane < T) {
      const float* r0 = rb_s + lane * L::kSRb + win_off(g_rb, lane) + pp_root * kRow;
      const Heading hq = heading_quat_inv(r0[3], r0[4], r0[5], r0[6]);
      float* hd = hdr_s + lane * L::kSHdr;
      hd[H_A0] = 2.0f * (hq.cw * hq.cw) - 1.0f; hd[H_SZ] = hq.sz; hd[H_CW] = hq.cw;
    }
    asm volatile("bar.sync 3, %0;" ::"n"(32 * kAdofObsWarps) : "memory");
    // Pass 1, both envs of this warp: imitation diffs and the per-env reductions warp 0's reward waits for; then the
    // warp ARRIVES, and only then (pass 2) spends its time on the observation segments -- the reward phase of warp 0
    // overlaps the obs stores instead of queueing behind them.
    constexpr int EPW = T / kAdofObsWarps;        // envs per obs warp
    static_assert(T % kAdofObsWarps == 0, "every obs warp owns the same number of envs");
    const bool body_on = lane < NB;
    float dpx[EPW], dpy[EPW], dpz[EPW], dvx[EPW], dvy[EPW], dvz[EPW];
#pragma unroll
    for (int qi = 0; qi < EPW; ++qi) {
      const int e = (warp - 1) + qi * kAdofObsWarps;
      const float* rb_e = rb_s + e * L::kSRb + win_off(g_rb, e);
      const float* in_e = init_s + e * L::kSInit + win_off(g_init, e);
      const float* cur = rb_e + bal_id * kRow;
      const float* ref = COMPACT ? in_e + (body_on ? lane : 0) * 6 : in_e + bal_id * kRow;
      const float* refv = ref + (COMPACT ? 3 : 7);
      // imitation diffs: ref - cur (ADOF:1345,1349 / ADOF:1908-1909)
      dpx[qi] = ref[0] - cur[0]; dpy[qi] = ref[1] - cur[1]; dpz[qi] = ref[2] - cur[2];
      dvx[qi] = refv[0] - cur[7]; dvy[qi] = refv[1] - cur[8]; dvz[qi] = refv[2] - cur[9];
      // per-body sums of squares; the /3 of the inner mean is applied once per env in phase R
      float s_dp2 = body_on ? (dpx[qi] * dpx[qi] + dpy[qi] * dpy[qi] + dpz[qi] * dpz[qi]) : 0.0f;
      float s_dv2 = body_on ? (dvx[qi] * dvx[qi] + dvy[qi] * dvy[qi] + dvz[qi] * dvz[qi]) : 0.0f;
      // has_fallen uses the norm of cur - ref (ADOF:1412): cur - ref == -(ref - cur) exactly and squares drop
      // the sign, so it is the square root of the same sum, bit for bit
      float s_nrm = sqrtf(s_dp2);
      // DOF terms, lane = DOF index
      const bool dof_on = lane < D;
      const int dl = dof_on ? lane : 0;
      float q = dof_s[e * L::kDof + 2 * dl], qd = dof_s[e * L::kDof + 2 * dl + 1];
      float rq = idof_s[e * L::kDof + 2 * dl], rqd = idof_s[e * L::kDof + 2 * dl + 1];
      float dq = rq - q, dqd = rqd - qd;
      float s_dq22 = (dof_on && lane < 22) ? dq * dq : 0.0f;
      float s_dq5 = (dof_on && lane >= 22) ? dq * dq : 0.0f;
      float s_dqd22 = (dof_on && lane < 22) ? dqd * dqd : 0.0f;
      float s_pow = dof_on ? fabsf(force_s[e * D + dl] * qd) : 0.0f;
      // all seven sums in one butterfly: lane L ends up with the total of value (L >> 2) & 7
      const float red = warp_sum8(s_dp2, s_dv2, s_nrm, s_dq22, s_dq5, s_dqd22, s_pow, 0.0f, lane);
      const float* r0 = rb_e + pp_root * kRow;
      float* hd = hdr_s + e * L::kSHdr;
      if ((lane & 3) == 0 && lane < 28) hd[H_SUM_DP2 + (lane >> 2)] = red;     // slots are consecutive
      if (lane == 31) { hd[H_RX] = r0[0]; hd[H_RY] = r0[1]; hd[H_RZ] = r0[2]; }
    }
    adof_arrive();
    if (!(phases & PPK_PHASE_OBS)) return;
    // Pass 2: observation segments of the same envs
#pragma unroll
    for (int qi = 0; qi < EPW; ++qi) {
      const int e = (warp - 1) + qi * kAdofObsWarps;
      if (e >= nvalid) continue;
      const float* rb_e = rb_s + e * L::kSRb + win_off(g_rb, e);
      const float* r0 = rb_e + pp_root * kRow;
      const float* hd = hdr_s + e * L::kSHdr;
      // heading frames: ping-pong root body and balance root body (both row 0 in the shipped config)
      Heading hq_pp;
      hq_pp.sz = hd[H_SZ];
      hq_pp.cw = hd[H_CW];
      const float* b0 = rb_e + bal_root * kRow;
      Heading hq_bal = (bal_root == pp_root) ? hq_pp : heading_quat_inv(b0[3], b0[4], b0[5], b0[6]);
      {
        // ping-pong bodies [0,60), lane = output float o: body o/3, component o%3 (ADOF:1849-1888);
        //   out_c = v_c*a0 + ((s1*v_o)*m)*2 == my_quat_rotate((0,0,sz,cw), v) component by component
        if (lane < 3 * J) {
          const int j = lane / 3, c = lane - j * 3;
          const int oth = (c == 0) ? 1 : (c == 1 ? 0 : 2);
          const float a0 = 2.0f * (hq_pp.cw * hq_pp.cw) - 1.0f;
          const float s1 = (c == 0) ? -hq_pp.sz : hq_pp.sz, m = (c == 2) ? hq_pp.sz : hq_pp.cw;
          const float* row = rb_e + k.ids[0][j] * kRow;
          const float pc = row[c] - r0[c], po = row[oth] - r0[oth];
          float* orow = g_obs + (size_t)e * kAdofObs;
          st_stream(orow + lane, clip_opt(pc * a0 + ((s1 * po) * m) * 2.0f, clip));
          st_stream(orow + 3 * J + lane, clip_opt(row[7 + c] * a0 + ((s1 * row[7 + oth]) * m) * 2.0f, clip));
        }
        // imitation observation segments [121,190) = 10*R(dP), [190,259) = R(dV)
        float lp[3], lv[3];
        rotate_heading(hq_bal, dpx[qi], dpy[qi], dpz[qi], lp[0], lp[1], lp[2]);
        rotate_heading(hq_bal, dvx[qi], dvy[qi], dvz[qi], lv[0], lv[1], lv[2]);
        lp[0] *= 10.0f; lp[1] *= 10.0f; lp[2] *= 10.0f;
        float* orow = g_obs + (size_t)e * kAdofObs + (6 * J + 2 * D + 7);
        // [body][xyz] -> 69 consecutive floats: through shared memory.  This env's reference rows have been
        // consumed (diffs are in registers), their staging area is the scratch; 18 shuffles + selects cost 3x more.
        float* tr = init_s + e * L::kSInit;
        __syncwarp();
        if (body_on) {
          tr[3 * lane] = lp[0]; tr[3 * lane + 1] = lp[1]; tr[3 * lane + 2] = lp[2];
          tr[3 * NB + 3 * lane] = lv[0]; tr[3 * NB + 3 * lane + 1] = lv[1]; tr[3 * NB + 3 * lane + 2] = lv[2];
        }
        __syncwarp();
#pragma unroll
        for (int o = lane; o < 3 * NB; o += 32) {
          st_stream(orow + o, clip_opt(tr[o], clip));
          st_stream(orow + 3 * NB + o, clip_opt(tr[3 * NB + o], clip));
        }
      }
    }
    return;
  }

  // ================= warp 0 =================================================================
  // per-env scalars straight from global, overlapping the bulk copies (lane = env)
  const bool lane_env = lane < nvalid;
  const int le = (lane < T) ? lane : 0;
  const long long env = env0 + (lane_env ? lane : 0);
  long long prog = 0, reset_prev = 0;
  float pre_vx = 0.0f;
  bool f_pcc = false, f_htc = false, f_dpc = false, f_hdc = false;
  bool c_closer = false, c_hitp = false, c_net = false, c_table = false, c_fall = false;
  if (lane_env) {
    prog = k.progress[env];
    if (!(phases & PPK_PHASE_REWARD)) reset_prev = k.reset[env];
    if (phases & PPK_PHASE_REWARD) {
      pre_vx = ld_stream(k.pre + (size_t)env * k.pre_stride + k.pre_vx);
      f_pcc = k.flags[0][env] != 0; f_htc = k.flags[1][env] != 0; f_dpc = k.flags[2][env] != 0; f_hdc = k.flags[3][env] != 0;
      c_closer = k.flags[4][env] != 0; c_hitp = k.flags[5][env] != 0; c_net = k.flags[6][env] != 0;
      c_table = k.flags[7][env] != 0; c_fall = k.flags[8][env] != 0;
    }
  }
  if (bulk) mbar_wait(bar, 0);

  // ---- phase R: lane = env ---------------------------------------------------------------------
  adof_wait();          // reductions and frames of all 8 envs are in hdr_s
  float* hd = hdr_s + le * L::kSHdr;
  const float* my_root = root_s + le * L::kRoot;
  const float* ball = my_root + k.ball * kRow;
  float bx = ball[0], by = ball[1], bz = ball[2], bvx = ball[7], bvy = ball[8], bvz = ball[9];
  long long p_new = prog + ((phases & PPK_PHASE_PROGRESS) ? 1 : 0);
  bool is_reset = reset_prev != 0;
  float reward = 0.0f;

  if (phases & PPK_PHASE_REWARD) {
    // compute_imitation_reward, is_g1 branch (ADOF:1330-1418)
    float r_body_pos = expf(-50.0f * ((hd[H_SUM_DP2] / 3.0f) / (float)NB));
    float r_body_vel = expf(-4.0f * ((hd[H_SUM_DV2] / 3.0f) / (float)NB));
    float first22 = 10.0f * expf(-2500.0f * (hd[H_SUM_DQ22] / 22.0f));
    float last5 = 0.2f * expf(-5.0f * (hd[H_SUM_DQ5] / 5.0f));
    float r_dof_vel = expf(-0.05f * (hd[H_SUM_DQD22] / 22.0f));
    float ref_reward = (((first22 + last5) + 0.2f * r_dof_vel) + 0.4f * r_body_pos) + 0.2f * r_body_vel;
    bool has_fallen = (hd[H_SUM_NORM] / (float)NB) > k.term_dist;
    if (has_fallen) ref_reward = 1.0f * -50.0f;
    c_fall = c_fall || has_fallen;

    const float* rb_e = rb_s + le * L::kSRb + win_off(g_rb, le);
    const float* pd = rb_e + k.paddle_body[0] * kRow;
    float px = pd[0], py = pd[1], pz = pd[2];
    float pelvis_z = rb_e[k.pelvis_body * kRow + 2];
    float hx = my_root[k.hum[0] * kRow];
    bool x_close = fabsf(bx - px) < 0.2f;
    bool first_close = x_close && !f_pcc;
    float dy = by - py, dz = bz - pz;
    float yz = sqrtf(dy * dy + dz * dz);
    bool in_circle = yz < 0.15f;
    float pos_reward = (first_close && !f_hdc) ? (in_circle ? k.hit_paddle : k.miss_coef * yz) : 0.0f;
    c_closer = c_closer || (first_close && in_circle);
    bool hit = (pre_vx < 0.0f) && (bvx > 1.5f);
    c_hitp = c_hitp || hit;
    float vel_reward = (hit && !f_pcc && !f_hdc) ? k.alpha * fabsf(bvx) : 0.0f;
    f_pcc = f_pcc || x_close;
    float time_penalty = ((bx > hx) && (bvx < 0.0f)) ? -0.01f * (float)p_new : 0.0f;
    // compute_gradient_penalty (ADOF:1245-1301)
    bool z_in = (bz >= 0.82f) && (bz <= 0.83f) && (bvx > 0.0f);
    float ddx = bx - 2.5f, ddy = by - 0.0f;
    float dist = sqrtf(ddx * ddx + ddy * ddy);
    bool in_range = (bx >= 1.9f) && (bx <= 3.1f) && (by >= -0.6f) && (by <= 0.6f);
    c_table = c_table || (z_in && in_range);
    float table = (z_in && !f_htc && !f_hdc) ? (in_range ? k.hit_table : k.not_hit * dist) : 0.0f;
    f_htc = f_htc || z_in;
    // net (ADOF:1619-1650)
    bool over_net = (bx > 1.72f) && (bx < 1.78f) && (bvx > 0.0f);
    bool suitable = (bz > 0.96f) && (bz < 1.25f);
    float over_h = !suitable ? ((bz > 1.25f) ? (bz - 1.25f) : (0.96f - bz)) : 0.0f;
    float net = (over_net && !f_hdc) ? (suitable ? k.cross_net : -400.0f * over_h) : 0.0f;
    c_net = c_net || (net > 0.0f);
    float power_reward = (-k.power_coef) * hd[H_POWER];
    bool low = bz < 0.78f;
    float die_pen = (low && !f_dpc && !f_hdc) ? k.die_penalty : 0.0f;
    f_dpc = f_dpc || low;
    f_hdc = f_hdc || (pelvis_z < 0.97f);
    reward = 0.0f + (((((((pos_reward + power_reward) + vel_reward) + table) + net) + die_pen) + time_penalty) + ref_reward);
    is_reset = p_new >= k.max_len - 1;          // die stays 0 (ADOF:1688)
    if (lane_env) {
      k.rew[env] = reward;
      k.reset[env] = is_reset ? 1 : 0;
    }
  }

  if (phases & PPK_PHASE_STATS) {
    double v[PPK_NUM_STATS];
    v[PPK_STAT_REWARD] = lane_env ? (double)reward : 0.0;
    v[PPK_STAT_PROGRESS] = lane_env ? (double)p_new : 0.0;
    v[PPK_STAT_RESETS] = (lane_env && is_reset) ? 1.0 : 0.0;
    v[PPK_STAT_FALL_DOWN] = (lane_env && c_fall) ? 1.0 : 0.0;
    v[PPK_STAT_CLOSER] = (lane_env && c_closer) ? 1.0 : 0.0;
    v[PPK_STAT_HIT_PADDLE] = (lane_env && c_hitp) ? 1.0 : 0.0;
    v[PPK_STAT_CROSS_NET] = (lane_env && c_net) ? 1.0 : 0.0;
    v[PPK_STAT_HIT_TABLE] = (lane_env && c_table) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < PPK_NUM_STATS; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
      double* slot = k.stats + (size_t)(blockIdx.x % PPK_STATS_SLOTS) * PPK_NUM_STATS;
#pragma unroll
      for (int i = 0; i < PPK_NUM_STATS; ++i) atomicAdd(slot + i, v[i]);
    }
  }

  // ---- predicated reset (ADOF:965-1028) ----------------------------------------------------------
  const bool do_reset = (phases & PPK_PHASE_RESET) && is_reset && lane_env;
  if (phases & PPK_PHASE_RESET) append_reset_indices(k, do_reset, env, lane);
  if (k.timeout != nullptr && lane_env && (phases & (PPK_PHASE_REWARD | PPK_PHASE_PROGRESS)))
    k.timeout[env] = (p_new >= k.max_len - 1) ? 1 : 0;
  if (do_reset) {
    const float* ir = k.init_root + (size_t)env * 39;
    float* gr = k.root_out + (size_t)env * 39;
    const float* rv = k.reset_vel + (size_t)env * 3;
    const float* ryz = k.reset_yz + (size_t)env * 2;
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int c = 0; c < 7; ++c) gr[a * kRow + c] = ir[a * kRow + c];
#pragma unroll
      for (int c = 7; c < kRow; ++c) gr[a * kRow + c] = 0.0f;
    }
    bx = ir[k.ball * kRow + 0]; by = ryz[0]; bz = ryz[1];
    bvx = rv[0]; bvy = rv[1]; bvz = rv[2];
    gr[k.ball * kRow + 1] = by; gr[k.ball * kRow + 2] = bz;
    gr[k.ball * kRow + 7] = bvx; gr[k.ball * kRow + 8] = bvy; gr[k.ball * kRow + 9] = bvz;
    if (k.reset_dof) {
      float* gd = k.dof_out + (size_t)env * 2 * D;
      for (int i = 0; i < 2 * D; ++i) {
        float v = idof_s[le * L::kDof + i];
        dof_s[le * L::kDof + i] = v;
        gd[i] = v;
      }
    }
    p_new = 0;
    k.scratch[0] = 1u;      // some env of the shard reset: the counters get cleared after the step
  }
  if (lane_env) {
    if (phases & (PPK_PHASE_PROGRESS | PPK_PHASE_RESET)) k.progress[env] = p_new;
    if (phases & PPK_PHASE_REWARD) {
      k.flags[0][env] = f_pcc; k.flags[1][env] = f_htc; k.flags[2][env] = f_dpc; k.flags[3][env] = f_hdc;
      k.flags[4][env] = c_closer; k.flags[5][env] = c_hitp; k.flags[6][env] = c_net; k.flags[7][env] = c_table;
      k.flags[8][env] = c_fall;
    }
    if (do_reset) { k.flags[0][env] = 0; k.flags[1][env] = 0; k.flags[2][env] = 0; k.flags[3][env] = 0; }
  }
  if (!(phases & PPK_PHASE_OBS)) return;

  // ball in the heading frame (+ y-intersect, ADOF:1833-1839)
  if (lane < T) {
    const float a0 = hd[H_A0], sz = hd[H_SZ], cw = hd[H_CW];
    const float rx = bx - hd[H_RX], ry = by - hd[H_RY], rz = bz - hd[H_RZ];
    const float lp0 = rx * a0 + ((-(sz * ry)) * cw) * 2.0f;
    const float lp1 = ry * a0 + ((sz * rx) * cw) * 2.0f;
    const float lp2 = rz * a0 + (sz * (sz * rz)) * 2.0f;
    const float lv0 = bvx * a0 + ((-(sz * bvy)) * cw) * 2.0f;
    const float lv1 = bvy * a0 + ((sz * bvx) * cw) * 2.0f;
    const float lv2 = bvz * a0 + (sz * (sz * bvz)) * 2.0f;
    const float yi = lp1 + (lv1 / (-lv0 + 1e-6f)) * lp0;
    hd[H_BALL0] = lp0; hd[H_BALL1] = lp1; hd[H_BALL2] = lp2;
    hd[H_BALL3] = lv0; hd[H_BALL4] = lv1; hd[H_BALL5] = lv2; hd[H_BALL6] = yi;
  }
  __syncwarp();
  // ---- tail segments, lane = element ------------------------------------------------------------
  // [60,121): dof_pos, 0.1*dof_vel, ball local pos/vel, y_intersect; [259,313): reference dof pos/vel
  constexpr int kSegA = 2 * D + 7, kSegB = 2 * D;
#pragma unroll
  for (int l = lane; l < kSegA + kSegB; l += 32) {
    // per-lane source: which staged array, which element, which scale
    const float* src;
    int stride, oo;
    float scale = 1.0f;
    if (l < D) { src = dof_s + 2 * l; stride = L::kDof; }
    else if (l < 2 * D) { src = dof_s + 2 * (l - D) + 1; stride = L::kDof; scale = 0.1f; }
    else if (l < kSegA) { src = hdr_s + H_BALL0 + (l - 2 * D); stride = L::kSHdr; }
    else if (l < kSegA + D) { src = idof_s + 2 * (l - kSegA); stride = L::kDof; }
    else { src = idof_s + 2 * (l - kSegA - D) + 1; stride = L::kDof; }
    oo = (l < kSegA) ? (6 * J + l) : (6 * J + kSegA + 6 * NB + (l - kSegA));
#pragma unroll
    for (int e = 0; e < T; ++e)
      if (e < nvalid) st_stream(g_obs + (size_t)e * kAdofObs + oo, clip_opt(src[e * stride] * scale, clip));
  }
}

inline int launch_adof_clear(unsigned int* scratch, unsigned char* const* flags, long long n, cudaStream_t s) {
  long long cb = (n / 4 + 255) / 256;
  if (cb < 1) cb = 1;
  if (cb > sm_count() * 8) cb = sm_count() * 8;
  if (launch_pdl(adof_clear_counters_kernel, (unsigned)cb, 256u, 0, s, scratch, flags[4], flags[5], flags[6], flags[7], flags[8], n) != cudaSuccess) {
    cudaGetLastError();
    return PPK_ERR_LAUNCH;
  }
  return PPK_OK;
}

}  // namespace ppk
