// Fused task step of the 27-DOF variant (ADOF), second design: the full fused step on full, aligned tiles with the
// compact reference pose.  Everything else (other phase subsets, tail envs, misaligned tensors, the uncompacted reference
// pose, ids that are not in rows 0..39) stays with adof_step_kernel (ppk_adof.cuh).
//
// Round 1's kernel spends 535 warp instructions per env: one env per warp pass with lane = balance body (23 of 32 lanes)
// or lane = output float, seven butterfly reductions per env, and an 8-lane reward warp.  Here one CTA of six warps owns
// a tile of 8 envs and every per-item loop runs with lane = (sub, env): env = lane & 7, and the items of an env (23
// balance bodies, 10 ping-pong bodies, 27 DOFs) are dealt round-robin to its four lanes sub = lane >> 3.  All 32 lanes
// work in every pass, a per-env sum is two xor-shuffles (8, 16) once per warp instead of a butterfly per env, and the obs
// row is assembled in shared memory and written back with ONE bulk copy per tile (10 016 B, full sectors).
//   stage   cp.async.bulk (TMA engine): the tile's rigid-body block is contiguous (8 x 2184 B: one copy), so are the
//           compact reference pose (8 x 552 B), root, DOF-state, reference-DOF and DOF-force slices; the eight 64-byte
//           windows around row ids[0] go out FIRST on their own mbarrier so that the heading frames (atan2f / sinf / cosf)
//           are ready when the bulk of the tile lands.
//   warp 0  lane = env (8): heading frames early; after the reductions arrive: compute_pingpong_reward_nv
//           (ADOF:1440-1690) + compute_gradient_penalty (ADOF:1245-1301) + the is_g1 branch of compute_imitation_reward
//           (ADOF:1313-1418), flags, counters, statistics, time-out mask, reset bookkeeping
//   warps 1, 2  balance bodies: imitation diffs, their three per-env sums, imitation observation segments (ADOF:1891-1927)
//   warp 3  ping-pong bodies in the heading frame (ADOF:1849-1888)
//   warp 4  DOFs: the four DOF sums, dof_pos / 0.1*dof_vel / reference-dof observation segments
//   warp 5  the predicated reset (ADOF:965-1028; time-outs only, decidable from progress_buf alone), the ball in the heading
//           frame and y_intersect (ADOF:1811-1846)
#pragma once
#include "ppk_adof.cuh"

namespace ppk {

constexpr int kAdof2Warps = 6;
constexpr int kAdof2Threads = 32 * kAdof2Warps;

struct Adof2Layout {
  static constexpr int T = kAdofTile;                       // 8 envs
  static constexpr int kRbEnv = 42 * kRow;                  // 546: the whole env block (rows 0..41)
  static constexpr int kInitEnv = kAdofNB * 6;              // 138
  static constexpr int kRoot = 3 * kRow, kDof = 2 * kAdofD;
  static constexpr int kOffRb = 0;
  static constexpr int kOffInit = kOffRb + T * kRbEnv;                  // 4368
  static constexpr int kOffRoot = kOffInit + T * kInitEnv;              // 5472
  static constexpr int kOffDof = kOffRoot + T * kRoot;                  // 5784
  static constexpr int kOffIDof = kOffDof + T * kDof;                   // 6216
  static constexpr int kOffForce = kOffIDof + T * kDof;                 // 6648
  static constexpr int kOffHd = kOffForce + T * kAdofD;                 // 6864: [T][8] a0, sz, cw, -, root pos xyz, -
  static constexpr int kOffObs = kOffHd + T * 8;                        // 6928
  static constexpr int kOffBar = kOffObs + T * kAdofObs;                // 9432
  static constexpr int kFloats = kOffBar + 4;                           // 37 744 B: six CTAs per SM
  // Two regions live inside others (224 floats less = the sixth resident CTA):
  //  * the [T][16] windows around row ids[0] sit in env 0's imitation columns of the obs tile: they are consumed by
  //    warp 0 before it arrives on the heading barriers 1..4, and the warps that write those columns (1, 2) wait on them;
  //  * the per-env partial sums of warps 1, 2, 4 overwrite the first floats of what that warp alone has just consumed
  //    (its half of the env's reference bodies / the env's DOF forces), after a __syncwarp.
  static constexpr int kOffRow0 = kOffObs + 128;
  static constexpr int kPartW2 = 12 * 6;                                // warp 2's reference bodies start here (body 12)
  static_assert(6 * 11 + 2 * kAdofD + 7 <= 128 && 128 + T * 16 <= 6 * 11 + 2 * kAdofD + 7 + 6 * kAdofNB,
                "the row windows stay inside the imitation columns");
  static constexpr uint32_t kTx = 4u * T * (kRbEnv + kInitEnv + kRoot + 2 * kDof + kAdofD);
  static_assert(kOffInit % 4 == 0 && kOffRoot % 4 == 0 && kOffDof % 4 == 0 && kOffIDof % 4 == 0 && kOffForce % 4 == 0 &&
                    kOffRow0 % 4 == 0 && kOffHd % 4 == 0 && kOffObs % 4 == 0 && kOffBar % 2 == 0,
                "16-byte bulk copy destinations");
  static_assert((T * kRbEnv) % 4 == 0 && (T * kInitEnv) % 4 == 0 && (T * kRoot) % 4 == 0 && (T * kDof) % 4 == 0 &&
                    (T * kAdofD) % 4 == 0 && (T * kAdofObs) % 4 == 0,
                "16-byte bulk copy sizes");
};

template <bool CLIP>
__global__ void __launch_bounds__(kAdof2Threads, 6)
adof2_step_kernel(const __grid_constant__ KArgs k) {
  using L = Adof2Layout;
  constexpr int T = L::T, D = kAdofD, J = kAdofJ, NB = kAdofNB;
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e8 = lane & 7, sub = lane >> 3;                 // lane = (sub, env)
  const long long env0 = (long long)blockIdx.x * T;
  float* rb_s = smem + L::kOffRb;
  float* init_s = smem + L::kOffInit;
  float* root_s = smem + L::kOffRoot;
  float* dof_s = smem + L::kOffDof;
  float* idof_s = smem + L::kOffIDof;
  float* force_s = smem + L::kOffForce;
  float* row0_s = smem + L::kOffRow0;
  float4* hd_s = reinterpret_cast<float4*>(smem + L::kOffHd);
  float* obs_s = smem + L::kOffObs;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  const float clip = CLIP ? k.clip_obs : 0.0f;
  const int root_id = k.ids[0][0];
  const bool stats = (k.phases & PPK_PHASE_STATS) != 0;
  gdc_wait();             // programmatic dependent launch: everything above ran while the previous kernel drained
  // every warp that needs the env's progress reads it HERE, before the barrier below: warp 0 rewrites it at the end
  const long long prog_in = k.progress[env0 + e8];

  // ---- stage ---------------------------------------------------------------------------------------
  PPK_STAMP(0);
  if (warp == 0) {
    if (elect_one()) {
      mbar_init(bar, 1);
      mbar_init(bar + 1, 1);
      mbar_fence_init();
      if (k.stagger > 0 && (int)blockIdx.x < k.first_wave && (int)blockIdx.x >= k.num_sms) {
        const long long wait = (long long)((int)blockIdx.x / k.num_sms) * k.stagger;
        const long long t0 = clock64();
        while (clock64() - t0 < wait) __nanosleep(64);
      }
      const float* g_rb = k.rb + (size_t)env0 * L::kRbEnv;
      mbar_arrive_expect_tx(bar, 4u * T * 16);
#pragma unroll
      for (int e = 0; e < T; ++e) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(g_rb + (size_t)e * L::kRbEnv + root_id * kRow) & ~(uintptr_t)15;
        bulk_g2s_in(row0_s + e * 16, reinterpret_cast<const void*>(a), 64u, bar);
      }
      mbar_arrive_expect_tx(bar + 1, L::kTx);
      bulk_g2s_in(rb_s, g_rb, 4u * T * L::kRbEnv, bar + 1);
      bulk_g2s_in(init_s, k.init_bal + (size_t)env0 * L::kInitEnv, 4u * T * L::kInitEnv, bar + 1);
      bulk_g2s_in(root_s, k.root + (size_t)env0 * L::kRoot, 4u * T * L::kRoot, bar + 1);
      bulk_g2s_in(dof_s, k.dof + (size_t)env0 * L::kDof, 4u * T * L::kDof, bar + 1);
      bulk_g2s_in(idof_s, k.init_dof + (size_t)env0 * L::kDof, 4u * T * L::kDof, bar + 1);
      bulk_g2s_in(force_s, k.force + (size_t)env0 * D, 4u * T * D, bar + 1);
    }
    __syncwarp();
  }
  __syncthreads();        // the initialised barriers are visible to every waiter
  PPK_STAMP(1);
  gdc_launch_dependents();      // every CTA of the grid is scheduled by the time the last one gets here: the next kernel of the
                                // stream may fill the slots the tail leaves empty and run its prologue (blocks in gdc_wait)

  // named barriers: 1..4 "heading table ready" (warp 0 arrives, ONE waiting warp each: 1, 2, 3, 5);
  //                 5 "per-env sums ready" (warps 1, 2, 4 arrive, warp 0 waits)
  const float* rb_e = rb_s + e8 * L::kRbEnv;

  if (warp == 0) {
    // ================= warp 0: frames, then reward (lane = env) ==========================================
    const bool on = lane < T;
    const long long env = env0 + e8;
    long long prog = 0;
    float pre_vx = 0.0f;
    bool f_pcc = false, f_htc = false, f_dpc = false, f_hdc = false;
    bool c_closer = false, c_hitp = false, c_net = false, c_table = false, c_fall = false;
    if (on) {
      prog = prog_in;
      pre_vx = ld_stream(k.pre + (size_t)env * k.pre_stride + k.pre_vx);
      f_pcc = k.flags[0][env] != 0; f_htc = k.flags[1][env] != 0; f_dpc = k.flags[2][env] != 0; f_hdc = k.flags[3][env] != 0;
      c_closer = k.flags[4][env] != 0; c_hitp = k.flags[5][env] != 0; c_net = k.flags[6][env] != 0;
      c_table = k.flags[7][env] != 0; c_fall = k.flags[8][env] != 0;
    }
    mbar_wait(bar, 0);
    if (on) {
      // where row ids[0] sits inside its 16-byte aligned window
      const int off = (int)((reinterpret_cast<uintptr_t>(k.rb + (size_t)env * L::kRbEnv + root_id * kRow) & 15u) >> 2);
      const float* r0 = row0_s + e8 * 16 + off;
      const Heading hq = heading_quat_inv(r0[3], r0[4], r0[5], r0[6]);
      hd_s[e8 * 2] = make_float4(2.0f * (hq.cw * hq.cw) - 1.0f, hq.sz, hq.cw, 0.0f);
      hd_s[e8 * 2 + 1] = make_float4(r0[0], r0[1], r0[2], 0.0f);
    }
    bar_arrive(1, 64); bar_arrive(2, 64); bar_arrive(3, 64); bar_arrive(4, 64);
    __syncwarp();
    mbar_wait(bar + 1, 0);
    PPK_STAMP(2);
    const float* my_root = root_s + e8 * L::kRoot;
    const float* ball = my_root + k.ball * kRow;
    const float bx = ball[0], by = ball[1], bz = ball[2], bvx = ball[7];
    const float* pd = rb_e + k.paddle_body[0] * kRow;
    const float px = pd[0], py = pd[1], pz = pd[2];
    const float pelvis_z = rb_e[k.pelvis_body * kRow + 2];
    const float hx = my_root[k.hum[0] * kRow];
    const long long p_new = prog + 1;
    // everything of the reward that does not need the per-env sums first: the wait below is then followed by five
    // expf and the final additions only
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
    bool low = bz < 0.78f;
    float die_pen = (low && !f_dpc && !f_hdc) ? k.die_penalty : 0.0f;
    f_dpc = f_dpc || low;
    f_hdc = f_hdc || (pelvis_z < 0.97f);
    bar_wait(5, 128);       // the sums of warps 1, 2 (balance bodies) and 4 (DOFs)
    const float* p1 = init_s + e8 * L::kInitEnv;
    const float* p2 = init_s + e8 * L::kInitEnv + L::kPartW2;
    const float* p4 = force_s + e8 * D;
    const float sum_dp2 = p1[0] + p2[0], sum_dv2 = p1[1] + p2[1], sum_nrm = p1[2] + p2[2];
    const float sum_dq22 = p4[0], sum_dq5 = p4[1], sum_dqd22 = p4[2], sum_pow = p4[3];
    // compute_imitation_reward, is_g1 branch (ADOF:1330-1418)
    float r_body_pos = expf(-50.0f * ((sum_dp2 / 3.0f) / (float)NB));
    float r_body_vel = expf(-4.0f * ((sum_dv2 / 3.0f) / (float)NB));
    float first22 = 10.0f * expf(-2500.0f * (sum_dq22 / 22.0f));
    float last5 = 0.2f * expf(-5.0f * (sum_dq5 / 5.0f));
    float r_dof_vel = expf(-0.05f * (sum_dqd22 / 22.0f));
    float ref_reward = (((first22 + last5) + 0.2f * r_dof_vel) + 0.4f * r_body_pos) + 0.2f * r_body_vel;
    bool has_fallen = (sum_nrm / (float)NB) > k.term_dist;
    if (has_fallen) ref_reward = 1.0f * -50.0f;
    c_fall = c_fall || has_fallen;
    float power_reward = (-k.power_coef) * sum_pow;
    const float reward = 0.0f + (((((((pos_reward + power_reward) + vel_reward) + table) + net) + die_pen) + time_penalty) + ref_reward);
    const bool is_reset = p_new >= k.max_len - 1;          // die stays 0 (ADOF:1688)
    if (on) {
      k.rew[env] = reward;
      k.reset[env] = is_reset ? 1 : 0;
    }
    if (stats) {
      double v[PPK_NUM_STATS];
      v[PPK_STAT_REWARD] = on ? (double)reward : 0.0;
      v[PPK_STAT_PROGRESS] = on ? (double)p_new : 0.0;
      v[PPK_STAT_RESETS] = (on && is_reset) ? 1.0 : 0.0;
      v[PPK_STAT_FALL_DOWN] = (on && c_fall) ? 1.0 : 0.0;
      v[PPK_STAT_CLOSER] = (on && c_closer) ? 1.0 : 0.0;
      v[PPK_STAT_HIT_PADDLE] = (on && c_hitp) ? 1.0 : 0.0;
      v[PPK_STAT_CROSS_NET] = (on && c_net) ? 1.0 : 0.0;
      v[PPK_STAT_HIT_TABLE] = (on && c_table) ? 1.0 : 0.0;
#pragma unroll
      for (int i = 0; i < PPK_NUM_STATS; ++i) v[i] = warp_sum(v[i]);
      if (lane == 0) {
        double* slot = k.stats + (size_t)(blockIdx.x % PPK_STATS_SLOTS) * PPK_NUM_STATS;
#pragma unroll
        for (int i = 0; i < PPK_NUM_STATS; ++i) atomicAdd(slot + i, v[i]);
      }
    }
    const bool do_reset = is_reset && on;
    append_reset_indices(k, do_reset, env, lane);
    if (k.timeout != nullptr && on) k.timeout[env] = is_reset ? 1 : 0;
    if (on) {
      k.progress[env] = do_reset ? 0 : p_new;
      k.flags[0][env] = f_pcc; k.flags[1][env] = f_htc; k.flags[2][env] = f_dpc; k.flags[3][env] = f_hdc;
      k.flags[4][env] = c_closer; k.flags[5][env] = c_hitp; k.flags[6][env] = c_net; k.flags[7][env] = c_table;
      k.flags[8][env] = c_fall;
      if (do_reset) { k.flags[0][env] = 0; k.flags[1][env] = 0; k.flags[2][env] = 0; k.flags[3][env] = 0; }
    }
    PPK_STAMP(3);
  } else if (warp == 1 || warp == 2) {
    // ================= warps 1, 2: balance bodies, lane = (sub, env), body k = 4*pass + sub ==============
    mbar_wait(bar + 1, 0);
    PPK_STAMP(2);
    const float* in_e = init_s + e8 * L::kInitEnv;
    float* o_e = obs_s + e8 * kAdofObs + (6 * J + 2 * D + 7);
    constexpr int kPasses = (NB + 3) / 4;                  // 6 passes of 4 bodies per env, three per warp
    constexpr int kMine = kPasses / 2;
    static_assert(kPasses % 2 == 0, "the two balance warps take the same number of passes");
    const int p0 = (warp == 1) ? 0 : kMine;
    // every load of the warp's passes first (shared-memory stores of one pass would otherwise order the loads of the next),
    // then the sums -- the reward of warp 0 waits for them -- and only then the rotations and the obs stores
    float dp[kMine][3], dv[kMine][3];
#pragma unroll
    for (int i = 0; i < kMine; ++i) {
      const int kb = (p0 + i) * 4 + sub;
      const int kk = (kb < NB) ? kb : 0;
      const float* cur = rb_e + k.bal_ids[kk] * kRow;
      const float* ref = in_e + kk * 6;
      // imitation diffs: ref - cur (ADOF:1345,1349 / ADOF:1908-1909)
#pragma unroll
      for (int c = 0; c < 3; ++c) { dp[i][c] = ref[c] - cur[c]; dv[i][c] = ref[3 + c] - cur[7 + c]; }
    }
    float s_dp2 = 0.0f, s_dv2 = 0.0f, s_nrm = 0.0f;
#pragma unroll
    for (int i = 0; i < kMine; ++i) {
      if ((p0 + i) * 4 + sub < NB) {
        const float sq = dp[i][0] * dp[i][0] + dp[i][1] * dp[i][1] + dp[i][2] * dp[i][2];
        s_dp2 += sq;
        s_dv2 += dv[i][0] * dv[i][0] + dv[i][1] * dv[i][1] + dv[i][2] * dv[i][2];
        // has_fallen uses the norm of cur - ref (ADOF:1412): the square root of the same sum of squares, bit for bit
        s_nrm += sqrtf(sq);
      }
    }
    s_dp2 += __shfl_xor_sync(0xffffffffu, s_dp2, 8); s_dp2 += __shfl_xor_sync(0xffffffffu, s_dp2, 16);
    s_dv2 += __shfl_xor_sync(0xffffffffu, s_dv2, 8); s_dv2 += __shfl_xor_sync(0xffffffffu, s_dv2, 16);
    s_nrm += __shfl_xor_sync(0xffffffffu, s_nrm, 8); s_nrm += __shfl_xor_sync(0xffffffffu, s_nrm, 16);
    __syncwarp();                                         // every lane has its reference bodies in registers
    if (lane < T) {
      float* p = init_s + e8 * L::kInitEnv + ((warp == 1) ? 0 : L::kPartW2);
      p[0] = s_dp2; p[1] = s_dv2; p[2] = s_nrm;
    }
    bar_arrive(5, 128);
    if (warp == 1) bar_wait(1, 64);                       // heading table (long ready: the frames are computed early)
    else bar_wait(2, 64);
    const float4 fa = hd_s[e8 * 2];
    Heading hq; hq.sz = fa.y; hq.cw = fa.z;
    const float a0 = fa.x;
#pragma unroll
    for (int i = 0; i < kMine; ++i) {
      const int kb = (p0 + i) * 4 + sub;
      if (kb < NB) {
        // rotate_heading, with the frame's a0 computed once
        const float lpx = dp[i][0] * a0 + ((-(hq.sz * dp[i][1])) * hq.cw) * 2.0f;
        const float lpy = dp[i][1] * a0 + ((hq.sz * dp[i][0]) * hq.cw) * 2.0f;
        const float lpz = dp[i][2] * a0 + (hq.sz * (hq.sz * dp[i][2])) * 2.0f;
        const float lvx = dv[i][0] * a0 + ((-(hq.sz * dv[i][1])) * hq.cw) * 2.0f;
        const float lvy = dv[i][1] * a0 + ((hq.sz * dv[i][0]) * hq.cw) * 2.0f;
        const float lvz = dv[i][2] * a0 + (hq.sz * (hq.sz * dv[i][2])) * 2.0f;
        o_e[3 * kb] = clip_opt(lpx * 10.0f, clip); o_e[3 * kb + 1] = clip_opt(lpy * 10.0f, clip); o_e[3 * kb + 2] = clip_opt(lpz * 10.0f, clip);
        o_e[3 * NB + 3 * kb] = clip_opt(lvx, clip); o_e[3 * NB + 3 * kb + 1] = clip_opt(lvy, clip); o_e[3 * NB + 3 * kb + 2] = clip_opt(lvz, clip);
      }
    }
    PPK_STAMP(3);
  } else if (warp == 3) {
    // ================= warp 3: ping-pong bodies in the heading frame, body j = 4*pass + sub ==============
    mbar_wait(bar + 1, 0);
    PPK_STAMP(2);
    bar_wait(3, 64);
    const float4 fa = hd_s[e8 * 2], fb = hd_s[e8 * 2 + 1];
    const float a0 = fa.x, sz = fa.y, cw = fa.z;
    float* o_e = obs_s + e8 * kAdofObs;
    constexpr int kPp = (J + 3) / 4;       // 3 passes
    float pv[kPp][6];
#pragma unroll
    for (int pass = 0; pass < kPp; ++pass) {
      const int j = pass * 4 + sub;
      const float* row = rb_e + k.ids[0][(j < J) ? j : 0] * kRow;
#pragma unroll
      for (int c = 0; c < 3; ++c) { pv[pass][c] = row[c]; pv[pass][3 + c] = row[7 + c]; }
    }
#pragma unroll
    for (int pass = 0; pass < kPp; ++pass) {
      const int j = pass * 4 + sub;
      const float px = pv[pass][0] - fb.x, py = pv[pass][1] - fb.y, pz = pv[pass][2] - fb.z;
      const float vx = pv[pass][3], vy = pv[pass][4], vz = pv[pass][5];
      if (j < J) {
        o_e[3 * j] = clip_opt(px * a0 + ((-(sz * py)) * cw) * 2.0f, clip);
        o_e[3 * j + 1] = clip_opt(py * a0 + ((sz * px) * cw) * 2.0f, clip);
        o_e[3 * j + 2] = clip_opt(pz * a0 + (sz * (sz * pz)) * 2.0f, clip);
        o_e[3 * J + 3 * j] = clip_opt(vx * a0 + ((-(sz * vy)) * cw) * 2.0f, clip);
        o_e[3 * J + 3 * j + 1] = clip_opt(vy * a0 + ((sz * vx) * cw) * 2.0f, clip);
        o_e[3 * J + 3 * j + 2] = clip_opt(vz * a0 + (sz * (sz * vz)) * 2.0f, clip);
      }
    }
    PPK_STAMP(3);
  } else if (warp == 4) {
    // ================= warp 4: DOFs, dof d = 4*pass + sub ================================================
    // A time-out reset puts the DOF state back to initial_dof_states (ADOF:990), which is the staged reference: the
    // observation segments of a resetting env read that instead of the live state; the sums use the live state.
    const bool is_reset = prog_in + 1 >= k.max_len - 1;
    const bool use_init = is_reset && k.reset_dof;
    mbar_wait(bar + 1, 0);
    PPK_STAMP(2);
    constexpr int kDp = (D + 3) / 4;       // 7 passes
    float qv[kDp][4], fv[kDp];
#pragma unroll
    for (int pass = 0; pass < kDp; ++pass) {
      const int d = pass * 4 + sub;
      const int dd = (d < D) ? d : 0;
      qv[pass][0] = dof_s[e8 * L::kDof + 2 * dd]; qv[pass][1] = dof_s[e8 * L::kDof + 2 * dd + 1];
      qv[pass][2] = idof_s[e8 * L::kDof + 2 * dd]; qv[pass][3] = idof_s[e8 * L::kDof + 2 * dd + 1];
      fv[pass] = force_s[e8 * D + dd];
    }
    float s_dq22 = 0.0f, s_dq5 = 0.0f, s_dqd22 = 0.0f, s_pow = 0.0f;
#pragma unroll
    for (int pass = 0; pass < kDp; ++pass) {
      const int d = pass * 4 + sub;
      if (d < D) {
        const float dq = qv[pass][2] - qv[pass][0], dqd = qv[pass][3] - qv[pass][1];
        if (d < 22) { s_dq22 += dq * dq; s_dqd22 += dqd * dqd; }
        else s_dq5 += dq * dq;
        s_pow += fabsf(fv[pass] * qv[pass][1]);
      }
    }
    s_dq22 += __shfl_xor_sync(0xffffffffu, s_dq22, 8); s_dq22 += __shfl_xor_sync(0xffffffffu, s_dq22, 16);
    s_dq5 += __shfl_xor_sync(0xffffffffu, s_dq5, 8); s_dq5 += __shfl_xor_sync(0xffffffffu, s_dq5, 16);
    s_dqd22 += __shfl_xor_sync(0xffffffffu, s_dqd22, 8); s_dqd22 += __shfl_xor_sync(0xffffffffu, s_dqd22, 16);
    s_pow += __shfl_xor_sync(0xffffffffu, s_pow, 8); s_pow += __shfl_xor_sync(0xffffffffu, s_pow, 16);
    __syncwarp();                                         // every lane has its forces in registers
    if (lane < T) {
      float* p = force_s + e8 * D;
      p[0] = s_dq22; p[1] = s_dq5; p[2] = s_dqd22; p[3] = s_pow;
    }
    bar_arrive(5, 128);
    float* o_e = obs_s + e8 * kAdofObs;
#pragma unroll
    for (int pass = 0; pass < kDp; ++pass) {
      const int d = pass * 4 + sub;
      if (d < D) {
        o_e[6 * J + d] = clip_opt(use_init ? qv[pass][2] : qv[pass][0], clip);
        o_e[6 * J + D + d] = clip_opt((use_init ? qv[pass][3] : qv[pass][1]) * 0.1f, clip);
        o_e[6 * J + 2 * D + 7 + 6 * NB + d] = clip_opt(qv[pass][2], clip);
        o_e[6 * J + 2 * D + 7 + 6 * NB + D + d] = clip_opt(qv[pass][3], clip);
      }
    }
    PPK_STAMP(3);
  } else {
    // ================= warp 5: predicated reset, ball in the heading frame ================================
    const bool on = lane < T;
    const long long env = env0 + e8;
    const bool is_reset = on && (prog_in + 1 >= k.max_len - 1);
    const unsigned pending = __ballot_sync(0xffffffffu, is_reset);
    // the post-reset ball of a resetting env (ADOF:975-988): initial x, sampled y / z, sampled launch velocity
    float nbx = 0.0f, nby = 0.0f, nbz = 0.0f, nvx = 0.0f, nvy = 0.0f, nvz = 0.0f;
    if (is_reset) {
      nbx = __ldg(k.init_root + (size_t)env * L::kRoot + k.ball * kRow);
      nby = __ldg(k.reset_yz + (size_t)env * 2); nbz = __ldg(k.reset_yz + (size_t)env * 2 + 1);
      nvx = __ldg(k.reset_vel + (size_t)env * 3); nvy = __ldg(k.reset_vel + (size_t)env * 3 + 1); nvz = __ldg(k.reset_vel + (size_t)env * 3 + 2);
    }
    // root rows of the resetting envs, all lanes (f = lane, lane + 32 < 39): requested before the tile lands
    float rr[2][2];
    int re[2] = {-1, -1};
    unsigned todo = pending;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (todo != 0u) {
        const int e = __ffs(todo) - 1;
        todo &= todo - 1u;
        re[i] = e;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int f = lane + 32 * half;
          float v = 0.0f;
          if (f < L::kRoot && (f % kRow) < 7) v = __ldg(k.init_root + (size_t)(env0 + e) * L::kRoot + f);
          rr[i][half] = v;
        }
      }
    }
    mbar_wait(bar + 1, 0);
    PPK_STAMP(2);
    bar_wait(4, 64);
    for (;;) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int e = re[i];
        if (e >= 0) {
          const long long ge = env0 + e;
          // ball row: y, z and the launch velocity come from the lane that owns env e
          const float by_e = __shfl_sync(0xffffffffu, nby, e), bz_e = __shfl_sync(0xffffffffu, nbz, e);
          const float vx_e = __shfl_sync(0xffffffffu, nvx, e), vy_e = __shfl_sync(0xffffffffu, nvy, e), vz_e = __shfl_sync(0xffffffffu, nvz, e);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int f = lane + 32 * half;
            if (f < L::kRoot) {
              float v = rr[i][half];
              const int a = f / kRow, c = f - a * kRow;
              if (a == k.ball) {
                if (c == 1) v = by_e;
                if (c == 2) v = bz_e;
                if (c == 7) v = vx_e;
                if (c == 8) v = vy_e;
                if (c == 9) v = vz_e;
              }
              k.root_out[(size_t)ge * L::kRoot + f] = v;
            }
          }
          if (k.reset_dof) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int f = lane + 32 * half;
              if (f < L::kDof) k.dof_out[(size_t)ge * L::kDof + f] = idof_s[e * L::kDof + f];
            }
          }
        }
      }
      if (todo == 0u) break;
      // more than two resetting envs in the tile (rare): next batch, loads and stores back to back
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        re[i] = -1;
        if (todo != 0u) {
          const int e = __ffs(todo) - 1;
          todo &= todo - 1u;
          re[i] = e;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int f = lane + 32 * half;
            float v = 0.0f;
            if (f < L::kRoot && (f % kRow) < 7) v = __ldg(k.init_root + (size_t)(env0 + e) * L::kRoot + f);
            rr[i][half] = v;
          }
        }
      }
    }
    if (pending != 0u && lane == 0) k.scratch[0] = 1u;      // some env of the shard reset: the counters get cleared after the step
    if (on) {
      const float* ball = root_s + e8 * L::kRoot + k.ball * kRow;
      const float bx = is_reset ? nbx : ball[0], by = is_reset ? nby : ball[1], bz = is_reset ? nbz : ball[2];
      const float bvx = is_reset ? nvx : ball[7], bvy = is_reset ? nvy : ball[8], bvz = is_reset ? nvz : ball[9];
      const float4 fa = hd_s[e8 * 2], fb = hd_s[e8 * 2 + 1];
      const float a0 = fa.x, sz = fa.y, cw = fa.z;
      const float rx = bx - fb.x, ry = by - fb.y, rz = bz - fb.z;
      const float lp0 = rx * a0 + ((-(sz * ry)) * cw) * 2.0f;
      const float lp1 = ry * a0 + ((sz * rx) * cw) * 2.0f;
      const float lp2 = rz * a0 + (sz * (sz * rz)) * 2.0f;
      const float lv0 = bvx * a0 + ((-(sz * bvy)) * cw) * 2.0f;
      const float lv1 = bvy * a0 + ((sz * bvx) * cw) * 2.0f;
      const float lv2 = bvz * a0 + (sz * (sz * bvz)) * 2.0f;
      const float yi = lp1 + (lv1 / (-lv0 + 1e-6f)) * lp0;      // ADOF:1839
      float* o = obs_s + e8 * kAdofObs + 6 * J + 2 * D;
      o[0] = clip_opt(lp0, clip); o[1] = clip_opt(lp1, clip); o[2] = clip_opt(lp2, clip);
      o[3] = clip_opt(lv0, clip); o[4] = clip_opt(lv1, clip); o[5] = clip_opt(lv2, clip);
      o[6] = clip_opt(yi, clip);
    }
    PPK_STAMP(3);
  }

  // ---- store the tile's obs rows: one bulk copy -----------------------------------------------------------
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    bulk_s2g_out(k.obs + (size_t)env0 * kAdofObs, obs_s, 4u * T * kAdofObs);
    bulk_commit();
    bulk_wait_read();
  }
}

// shift every per-env pointer of `k` by e0 envs (a shard view of the same buffers)
inline void offset_envs(KArgs& k, long long e0) {
  const size_t e = (size_t)e0;
  k.rb += e * k.B * kRow; k.root += e * k.A * kRow; k.dof += e * 2 * k.D; k.root_out += e * k.A * kRow; k.dof_out += e * 2 * k.D;
  k.force += e * k.D;
  if (k.pre) k.pre += e * k.pre_stride;
  if (k.init_root) k.init_root += e * k.A * kRow;
  if (k.init_dof) k.init_dof += e * 2 * k.D;
  if (k.init_rb) k.init_rb += e * k.B * kRow;
  if (k.init_bal) k.init_bal += e * kAdofNB * 6;
  if (k.reset_vel) k.reset_vel += e * 3;
  if (k.reset_yz) k.reset_yz += e * 2;
  if (k.obs) k.obs += e * kAdofObs;
  if (k.rew) k.rew += e;
  k.reset += e; k.progress += e;
  for (int i = 0; i < PPK_MAX_FLAGS; ++i)
    if (k.flags[i]) k.flags[i] += e;
  if (k.timeout) k.timeout += e;
  if (k.actor_idx) k.actor_idx += e * k.A;
  if (k.dof_idx) k.dof_idx += e * k.dof_per_env;
  k.n -= e0;
}

template <class Kernel>
inline void adof_soft_start(KArgs& k, Kernel kern, int threads, size_t smem, long long tiles, int& occ, int& sms) {
  if (occ == 0) {
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, threads, smem) == cudaSuccess && o > 0) { occ = o; sms = sm_count(); }
    else cudaGetLastError();
  }
  static int stag = -2;
  if (stag == -2) {
    const char* e = getenv("PPK_STAGGER_ADOF");
    stag = e ? atoi(e) : -1;
  }
  k.num_sms = sms > 0 ? sms : 1;
  k.first_wave = occ * sms;
  k.stagger = (occ > 0 && tiles > (long long)occ * sms) ? (stag >= 0 ? stag : 600) : 0;
}

inline int launch_adof_v1(KArgs k, cudaStream_t s) {
  // staged row windows: live rows 0..39, reference rows 0..27 (+ one following row for the window slack)
  const bool compact = k.init_bal != nullptr;
  const void* al[] = {k.rb, compact ? (const void*)k.init_bal : (const void*)k.init_rb, k.root, k.dof, k.init_dof, k.force};
  bool bulk = k.B > kAdofRbRows;
  for (const void* p : al) bulk = bulk && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0);
  k.bulk_ok = bulk ? 1 : 0;
  const size_t smem = (size_t)(compact ? AdofLayout<true>::kFloats : AdofLayout<false>::kFloats) * sizeof(float);
  static SmemOptIn opt_full, opt_compact, opt_full_c, opt_compact_c;
  if (!opt_full.ensure(adof_step_kernel<false, false>, AdofLayout<false>::kFloats * sizeof(float)) ||
      !opt_compact.ensure(adof_step_kernel<true, false>, AdofLayout<true>::kFloats * sizeof(float)) ||
      !opt_full_c.ensure(adof_step_kernel<false, true>, AdofLayout<false>::kFloats * sizeof(float)) ||
      !opt_compact_c.ensure(adof_step_kernel<true, true>, AdofLayout<true>::kFloats * sizeof(float)))
    return PPK_ERR_LAUNCH;
  const long long tiles = (k.n + kAdofTile - 1) / kAdofTile;
  static int occ_c = 0, occ_f = 0, sms = 0;
  if (compact) adof_soft_start(k, adof_step_kernel<true, false>, kAdofThreads, smem, tiles, occ_c, sms);
  else adof_soft_start(k, adof_step_kernel<false, false>, kAdofThreads, smem, tiles, occ_f, sms);
  const bool clip = k.clip_obs > 0.0f;
  if (compact && !clip) adof_step_kernel<true, false><<<(unsigned)tiles, kAdofThreads, smem, s>>>(k);
  else if (compact) adof_step_kernel<true, true><<<(unsigned)tiles, kAdofThreads, smem, s>>>(k);
  else if (!clip) adof_step_kernel<false, false><<<(unsigned)tiles, kAdofThreads, smem, s>>>(k);
  else adof_step_kernel<false, true><<<(unsigned)tiles, kAdofThreads, smem, s>>>(k);
  return cudaGetLastError() == cudaSuccess ? PPK_OK : PPK_ERR_LAUNCH;
}

// PPK_ADOF_V1=1 keeps every env on the first design (A/B runs, tests of the fallback path)
inline bool adof_force_v1() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PPK_ADOF_V1");
    v = (e && atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}

inline int launch_adof(const KArgs& k0, cudaStream_t s) {
  KArgs k = k0;
  if (k.B < kAdofRbRows) return PPK_ERR_SHAPE;
  for (int j = 0; j < kAdofJ; ++j)
    if (k.ids[0][j] >= kAdofRbRows) return PPK_ERR_SHAPE;
  const bool compact = k.init_bal != nullptr;
  for (int j = 0; j < kAdofNB; ++j)
    if (k.bal_ids[j] >= (compact ? kAdofRbRows : kAdofInitRows)) return PPK_ERR_SHAPE;
  if (k.paddle_body[0] >= kAdofRbRows || k.pelvis_body >= kAdofRbRows) return PPK_ERR_SHAPE;
  const bool fused_reset = (k.phases & PPK_PHASE_RESET) != 0 && !(k.phases & kPhaseDeferCounterClear);
  // the second design takes the full fused step on full, 16-byte aligned tiles of the shipped layout (42 bodies, compact
  // reference pose, one root body for both id lists); the first design takes everything else, tail envs included
  const int need = PPK_PHASE_PROGRESS | PPK_PHASE_REWARD | PPK_PHASE_RESET | PPK_PHASE_OBS;
  bool v2 = !adof_force_v1() && compact && (k.phases & need) == need && k.B == 42 && k.A == 3 && k.ids[0][0] == k.bal_ids[0] &&
            k.n >= kAdofTile && k.reset_yz != nullptr && k.pre != nullptr;
  const void* al[] = {k.rb, k.init_bal, k.root, k.dof, k.init_dof, k.force, k.obs};
  for (const void* p : al) v2 = v2 && p != nullptr && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0);
  long long done = 0;
  if (v2) {
    constexpr size_t smem2 = (size_t)Adof2Layout::kFloats * sizeof(float);
    static SmemOptIn opt2, opt2c;
    if (!opt2.ensure(adof2_step_kernel<false>, smem2) || !opt2c.ensure(adof2_step_kernel<true>, smem2)) return PPK_ERR_LAUNCH;
    const long long tiles = k.n / kAdofTile;
    KArgs k2 = k;
    k2.n = tiles * kAdofTile;
    static int occ2 = 0, sms2 = 0;
    adof_soft_start(k2, adof2_step_kernel<false>, kAdof2Threads, smem2, tiles, occ2, sms2);
    const cudaError_t le = (k.clip_obs > 0.0f) ? launch_pdl(adof2_step_kernel<true>, (unsigned)tiles, kAdof2Threads, smem2, s, k2)
                                               : launch_pdl(adof2_step_kernel<false>, (unsigned)tiles, kAdof2Threads, smem2, s, k2);
    if (le != cudaSuccess) { cudaGetLastError(); return PPK_ERR_LAUNCH; }
    done = k2.n;
  }
  if (done < k.n) {
    KArgs kt = k;
    if (done > 0) offset_envs(kt, done);
    const int rc = launch_adof_v1(kt, s);
    if (rc != PPK_OK) return rc;
  }
  if (fused_reset) return launch_adof_clear(k.scratch, k.flags, k.n, s);
  return PPK_OK;
}

}  // namespace ppk
