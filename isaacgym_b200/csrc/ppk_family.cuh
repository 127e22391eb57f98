// Fused task step for the 3-/4-actor variants (A3, TILT, NES, ALIGN, A4).
//
// One CTA of three warps owns one tile of TILE consecutive envs.
//   stage    The PhysX tensors are AoS with 52-byte rows, so per-env vector loads are impossible.
//            Instead every env's rigid-body rows ids[1..J) (one contiguous 468-byte run) and row
//            ids[0] are fetched with 1-D bulk async copies (cp.async.bulk, the TMA engine) of the
//            enclosing 16-byte-aligned windows (480 B / 64 B), and the tile's slices of the DOF
//            and DOF-force tensors (contiguous across envs) with one bulk copy each.  All copies
//            complete on one mbarrier; no registers or LSU issue slots are spent on staging.  The
//            few per-env scalars (progress, flags, saved ball velocity, ball / humanoid root
//            fields) are plain loads issued before the wait.  Tail tiles / misaligned tensors
//            use an LDG path into the same layout.
//   warp 0   lane = env: progress+1, reward, die/time-out mask, flag updates, statistics, the
//            predicated reset (root/DOF rows rewritten from the initial tensors), ball in the
//            heading frame, and the obs "tail" (dof_pos, 0.1*dof_vel, ball).
//   warps 1,2  heading frames of their half of the (env, humanoid) units, then lane = output
//            float: every lane produces one component of a rotated body position and velocity,
//            so each obs row segment is stored as one contiguous 120-byte run.
#pragma once
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

// warps 1..OW rotate bodies, warp 0 does the reward; OW is chosen per variant (register budget)

template <int H, int J, int D, int A, int TILE, int OW>
struct FamilyLayout {
  static constexpr int kThreads = 32 * (1 + OW);
  static constexpr int kUnits = TILE * H;                               // (env, humanoid) pairs
  static constexpr int kSpanRows = J - 1;                               // rows ids[1..J)
  static constexpr int kSpanFloats = ((kSpanRows * kRow + 3 + 3) / 4) * 4;   // 120: run + alignment slack
  static constexpr int kRootEnv = A * kRow;
  static constexpr int kTail = 2 * D + 6;              // dof_pos, 0.1*dof_vel, ball local pos, vel
  static constexpr int kObs = 6 * J + kTail;           // 80 (D=7) / 94 (D=14)
  static constexpr int kHdr = 16;                      // a0, sz, cw, span offset, then row ids[0]: pos3 quat4 vel3
  static constexpr int kHdrRow = 4;                    // where the copy of row ids[0] starts inside a header
  // float offsets inside the CTA's shared memory
  static constexpr int kOffDof = kUnits * kSpanFloats;
  static constexpr int kOffForce = kOffDof + TILE * 2 * D;
  static constexpr int kOffHdr = kOffForce + TILE * D;
  static constexpr int kOffBar = kOffHdr + kUnits * kHdr;            // 8-byte mbarrier
  static constexpr int kFloats = kOffBar + 4;
  static constexpr uint32_t kTxBytes = 4u * (kUnits * kSpanFloats + TILE * 2 * D + TILE * D);
  static_assert((TILE * 2 * D) % 4 == 0 && (TILE * D) % 4 == 0, "16-byte bulk sizes");
  static_assert(kOffBar % 2 == 0, "mbarrier alignment");
  static_assert(3 * J <= 32, "one lane per output float of a body segment");
  static_assert(kUnits % OW == 0 && kUnits / OW <= 32, "one heading per lane of an obs warp");
  static_assert(H * 6 <= D, "ball-in-frame scratch aliases the force staging");
};

// named barrier 1: obs warps arrive once their heading tables are written, warp 0 waits on it
__device__ __forceinline__ void hdr_arrive(int threads) { asm volatile("bar.arrive 1, %0;" ::"r"(threads) : "memory"); }
__device__ __forceinline__ void hdr_wait(int threads) { asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory"); }

template <int V, int H, int J, int D, int A, int TILE, int OW>
__global__ void __launch_bounds__(32 * (1 + OW), OW <= 2 ? 10 : 6)
family_step_kernel(const __grid_constant__ KArgs k) {
  using L = FamilyLayout<H, J, D, A, TILE, OW>;
  constexpr int kFamilyThreads = L::kThreads;
  constexpr int kFamilyWarps = 1 + OW;
  constexpr int kObsWarps = OW;
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long env0 = (long long)blockIdx.x * TILE;
  const int nvalid = (int)min((long long)TILE, k.n - env0);

  float* span_s = smem;
  float* dof_s = smem + L::kOffDof;
  float* force_s = smem + L::kOffForce;
  float* hdr_s = smem + L::kOffHdr;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  float* ball_s = force_s;          // [unit][6], written by warp 0 after it consumed the forces

  const int phases = k.phases;
  const bool bulk = k.bulk_ok && (nvalid == TILE);
  const int env_stride = k.B * kRow;
  const float* g_rb = k.rb + (size_t)env0 * env_stride;

  // ---- stage ---------------------------------------------------------------------------------------
  PPK_STAMP(0);
  if (bulk) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(bar, L::kTxBytes);
      bulk_g2s(dof_s, k.dof + (size_t)env0 * 2 * D, 4u * TILE * 2 * D, bar);
      bulk_g2s(force_s, k.force + (size_t)env0 * D, 4u * TILE * D, bar);
    }
    // The 2 x kUnits row windows: every warp walks its share of the units with warp-uniform
    // addresses and one elected lane issues the copies (the copy instruction takes uniform operands;
    // letting each lane issue its own makes the compiler serialise over the lanes at ~16
    // instructions per copy).
    const int wu = __shfl_sync(0xffffffffu, warp, 0);       // provably warp-uniform
#pragma unroll 2
    for (int u = wu; u < L::kUnits; u += kFamilyWarps) {
      const int e = u / H, h = u - e * H;
      const float* row = g_rb + (size_t)e * env_stride;
      const uintptr_t a1 = reinterpret_cast<uintptr_t>(row + k.ids[h][1] * kRow) & ~(uintptr_t)15;
      if (elect_one()) bulk_g2s(span_s + u * L::kSpanFloats, reinterpret_cast<const void*>(a1), 4u * L::kSpanFloats, bar);
    }
  } else {
    // generic path (tail tile, unaligned tensors, non-consecutive ids): plain loads, same layout
    for (int f = threadIdx.x; f < L::kUnits * J * kRow; f += kFamilyThreads) {
      const int u = f / (J * kRow), r = f - u * (J * kRow);
      const int e = u / H, h = u - e * H;
      const int j = r / kRow, c = r - j * kRow;
      float v = (e < nvalid) ? g_rb[(size_t)e * env_stride + k.ids[h][j] * kRow + c] : 0.0f;
      if (j == 0) { if (c < 10) hdr_s[u * L::kHdr + L::kHdrRow + c] = v; }
      else span_s[u * L::kSpanFloats + (j - 1) * kRow + c] = v;
    }
    for (int f = threadIdx.x; f < TILE * 2 * D; f += kFamilyThreads)
      dof_s[f] = (f < nvalid * 2 * D) ? k.dof[(size_t)env0 * 2 * D + f] : 0.0f;
    for (int f = threadIdx.x; f < TILE * D; f += kFamilyThreads)
      force_s[f] = (f < nvalid * D) ? k.force[(size_t)env0 * D + f] : 0.0f;
    __syncthreads();
  }

  PPK_STAMP(1);
  if (warp != 0) {
    // ================= warps 1, 2: body observations =============================================
    if (!(phases & PPK_PHASE_OBS)) return;
    // Row ids[0] (the heading / root body: 10 floats) is one plain load per float by the lane that
    // owns the unit's frame, issued before the wait so it overlaps the bulk copies; a second TMA
    // window per env would double the number of copy descriptors for 40 useful bytes.
    const bool frame_lane = lane < L::kUnits / kObsWarps;
    const int fu = kObsWarps * (frame_lane ? lane : 0) + (warp - 1);
    float r0[10];
    if (bulk && frame_lane) {
      const int e = fu / H, h = fu - e * H;
      const float* g0 = g_rb + (size_t)e * env_stride + k.ids[h][0] * kRow;
#pragma unroll
      for (int c = 0; c < 10; ++c) r0[c] = ld_stream(g0 + c);
    }
    // On the bulk path the frames need only the plain-loaded root rows: compute them while the row windows are
    // still in flight and wait for those afterwards.
    if (!bulk) PPK_STAMP(2);
    // heading frames of this warp's units u = OW*lane + (warp-1)
    if (frame_lane) {
      const int e = fu / H, h = fu - e * H;
      float* hd = hdr_s + fu * L::kHdr;
      int off1 = 0;
      if (bulk) {   // where the run sits inside its 16-byte aligned staging window
        off1 = (int)((reinterpret_cast<uintptr_t>(g_rb + (size_t)e * env_stride + k.ids[h][1] * kRow) & 15u) >> 2);
#pragma unroll
        for (int c = 0; c < 10; ++c) hd[L::kHdrRow + c] = r0[c];
      } else {
#pragma unroll
        for (int c = 0; c < 10; ++c) r0[c] = hd[L::kHdrRow + c];     // staged by the generic path
      }
      Heading hq = heading_quat_inv(r0[3], r0[4], r0[5], r0[6]);
      hd[0] = 2.0f * (hq.cw * hq.cw) - 1.0f;
      hd[1] = hq.sz; hd[2] = hq.cw;
      hd[3] = __int_as_float(off1);
    }
    hdr_arrive(kFamilyThreads);       // warp 0 needs the frames for the ball
    if (bulk) {
      mbar_wait(bar, 0);
      PPK_STAMP(2);
    }
    __syncwarp();
    // lane = output float o of a 3J-float segment: body j = o/3, component c = o%3.
    //   out_c = v_c*a0 + ((s1*v_o)*m)*2 with (s1, o, m) = (-sz, y, cw) / (sz, x, cw) / (sz, z, sz)
    // which is my_quat_rotate((0,0,sz,cw), v) component by component, same operation order.
    const bool lane_on = lane < 3 * J;
    const int o = lane_on ? lane : 0;
    const int j = o / 3, c = o - j * 3;
    const int oth = (c == 0) ? 1 : (c == 1 ? 0 : 2);
    const float sgn = (c == 0) ? -1.0f : 1.0f;
    float* g_obs = k.obs + (size_t)env0 * H * L::kObs;
#pragma unroll 2
    for (int u = warp - 1; u < L::kUnits; u += kObsWarps) {
      const int e = u / H;
      const float* hd = hdr_s + u * L::kHdr;
      const float a0 = hd[0], sz = hd[1], cw = hd[2];
      const float s1 = sgn * sz, m = (c == 2) ? sz : cw;
      const float* row = (j == 0) ? (hd + L::kHdrRow)
                                  : (span_s + u * L::kSpanFloats + __float_as_int(hd[3]) + (j - 1) * kRow);
      const float pc = row[c] - hd[L::kHdrRow + c], po = row[oth] - hd[L::kHdrRow + oth];
      const float out_p = pc * a0 + ((s1 * po) * m) * 2.0f;
      const float out_v = row[7 + c] * a0 + ((s1 * row[7 + oth]) * m) * 2.0f;
      if (lane_on && e < nvalid) {
        float* orow = g_obs + (size_t)u * L::kObs;
        st_stream(orow + o, out_p);
        st_stream(orow + 3 * J + o, out_v);
      }
    }
    PPK_STAMP(3);
    return;
  }

  // ================= warp 0: reward / reset / tail, lane = env ==========================================
  const bool lane_env = lane < nvalid;
  const long long env = env0 + (lane_env ? lane : 0);
  long long prog = 0, reset_prev = 0;
  float pre_vx = 0.0f, pre_vz = 0.0f;
  constexpr int NF = (V == PPK_TILT) ? 3 : (V == PPK_A4) ? 6 : (V == PPK_NES) ? 2 : (V == PPK_ALIGN || V == PPK_ALIGN2) ? 1 : 0;
  long long last_hitter = 2;
  bool flag[NF > 0 ? NF : 1];
  // per-env scalars come straight from global and overlap the bulk copies
  const float* g_root = k.root + (size_t)env * L::kRootEnv;
  const float* g_ball = g_root + k.ball * kRow;
  float bx = g_ball[0], by = g_ball[1], bz = g_ball[2];
  float bvx = g_ball[7], bvy = g_ball[8], bvz = g_ball[9];
  float hx[H];
#pragma unroll
  for (int h = 0; h < H; ++h) hx[h] = g_root[k.hum[h] * kRow];
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
  // Time-out and ball-height termination are already decidable: pull the reset sources of those
  // envs towards L2 now, so the predicated reset below does not pay a DRAM round trip.
  if ((phases & PPK_PHASE_RESET) && lane_env) {
    const bool timeout = prog + ((phases & PPK_PHASE_PROGRESS) ? 1 : 0) >= k.max_len - 1;
    const bool low_ball = (V != PPK_NES) && (bz < 0.1f);
    if (timeout || low_ball || reset_prev != 0) {
      const char* ir = reinterpret_cast<const char*>(k.init_root + (size_t)env * L::kRootEnv);
      prefetch_l2(ir);
      prefetch_l2(ir + 4 * L::kRootEnv - 4);
      prefetch_l2(k.reset_vel + (size_t)env * 3);
      if (k.reset_dof) {
        const char* id = reinterpret_cast<const char*>(k.init_dof + (size_t)env * 2 * D);
        prefetch_l2(id);
        prefetch_l2(id + 8 * D - 4);
      }
    }
  }
  if (bulk) mbar_wait(bar, 0);
  PPK_STAMP(2);

  const int le = (lane < TILE) ? lane : 0;     // smem row this lane reads (idle lanes read row 0)
  float dofv[2 * D];
#pragma unroll
  for (int i = 0; i < 2 * D; ++i) dofv[i] = dof_s[le * 2 * D + i];

  long long p_new = prog + ((phases & PPK_PHASE_PROGRESS) ? 1 : 0);
  bool is_reset = reset_prev != 0;
  float rew[H];
#pragma unroll
  for (int h = 0; h < H; ++h) rew[h] = 0.0f;

  if (phases & PPK_PHASE_REWARD) {
    float power = 0.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) power += fabsf(force_s[le * D + d] * dofv[2 * d + 1]);
    bool die = false;
    Scene sc[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      Scene& s = sc[h];
      s.bx = bx; s.by = by; s.bz = bz; s.vx = bvx; s.vz = bvz;
      s.pre_vx = pre_vx; s.pre_vz = pre_vz;
      const int pj = k.paddle_j[h];
      const int u = le * H + h;
      const float* pd;
      if (pj > 0) {
        const int off = bulk ? (int)((reinterpret_cast<uintptr_t>(g_rb + (size_t)le * env_stride + k.ids[h][1] * kRow) & 15u) >> 2) : 0;
        pd = span_s + u * L::kSpanFloats + off + (pj - 1) * kRow;
      } else {    // paddle is the root body or not among the observation bodies: read it in place
        pd = k.rb + ((size_t)env * k.B + k.paddle_body[h]) * kRow;
      }
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

  // ---- predicated reset (TILT:847-906): lanes whose env resets rewrite its root / DOF rows -------
  const bool do_reset = (phases & PPK_PHASE_RESET) && is_reset && lane_env;
  if (phases & PPK_PHASE_RESET) append_reset_indices(k, do_reset, env, lane);
  if (k.timeout != nullptr && lane_env && (phases & (PPK_PHASE_REWARD | PPK_PHASE_PROGRESS)))
    k.timeout[env] = (p_new >= k.max_len - 1) ? 1 : 0;
  if (do_reset) {
    const float* ir = k.init_root + (size_t)env * L::kRootEnv;
    float* gr = k.root_out + (size_t)env * L::kRootEnv;
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
      float* gd = k.dof_out + (size_t)env * 2 * D;
#pragma unroll
      for (int i = 0; i < 2 * D; ++i) {
        const float v = id[i];
        gd[i] = v;
        dof_s[le * 2 * D + i] = v;      // the observation tail below reads the staged copy
      }
    }
    p_new = 0;
  }
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
  if (!(phases & PPK_PHASE_OBS)) return;

  // ---- ball in the heading frame (frames come from the obs warps) -------------------------------------
  hdr_wait(kFamilyThreads);
  if (lane < TILE) {
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float* hd = hdr_s + (le * H + h) * L::kHdr;
      const float a0 = hd[0], sz = hd[1], cw = hd[2];
      const float rx = bx - hd[L::kHdrRow], ry = by - hd[L::kHdrRow + 1], rz = bz - hd[L::kHdrRow + 2];
      float* bs = ball_s + (le * H + h) * 6;
      bs[0] = rx * a0 + ((-(sz * ry)) * cw) * 2.0f;
      bs[1] = ry * a0 + ((sz * rx) * cw) * 2.0f;
      bs[2] = rz * a0 + (sz * (sz * rz)) * 2.0f;
      bs[3] = bvx * a0 + ((-(sz * bvy)) * cw) * 2.0f;
      bs[4] = bvy * a0 + ((sz * bvx) * cw) * 2.0f;
      bs[5] = bvz * a0 + (sz * (sz * bvz)) * 2.0f;
    }
  }
  __syncwarp();
  // tail of the obs row, lane = element: dof_pos (D), 0.1*dof_vel (D), ball local pos/vel (6)
  float* g_obs = k.obs + (size_t)env0 * H * L::kObs + 6 * J;
#pragma unroll
  for (int l = lane; l < L::kTail; l += 32) {
    const float scale = (l >= D && l < 2 * D) ? 0.1f : 1.0f;
    const int dsrc = (l < D) ? 2 * l : 2 * (l - D) + 1;
#pragma unroll 4
    for (int u = 0; u < L::kUnits; ++u) {
      const int e = u / H;
      const float v = (l < 2 * D) ? dof_s[e * 2 * D + dsrc] * scale : ball_s[u * 6 + (l - 2 * D)];
      if (e < nvalid) st_stream(g_obs + (size_t)u * L::kObs + l, v);
    }
  }
  PPK_STAMP(3);
}

}  // namespace ppk
