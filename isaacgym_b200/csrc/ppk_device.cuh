// Device-side building blocks of the ping-pong task step (sm_100a).
//
// Compiled with -fmad=false: every flag of the reference is a comparison of fp32
// values produced by single correctly-rounded ATen ops, so contraction into FMA
// would change which side of a threshold a value lands on.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/ppk.h"

namespace ppk {

// The opt-in dynamic shared-memory size is an attribute of (kernel, device): one of these per launcher
// remembers the devices that already have it.  (A benign race sets the attribute twice.)
struct SmemOptIn {
  unsigned long long done = 0;
  template <class Kernel>
  bool ensure(Kernel kern, size_t bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (dev >= 0 && dev < 64 && ((done >> dev) & 1ull)) return true;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    if (dev >= 0 && dev < 64) done |= 1ull << dev;
    return true;
  }
};

// SM count of the current device (grids are sized in multiples of it), queried once per device
inline int sm_count() {
  static int cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return 148; }
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
    cached[dev] = n;
  }
  return cached[dev];
}

// PPK_PDL=0 turns programmatic dependent launch of the step kernels off (A/B runs); default on
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PPK_PDL");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

// Launch `kern` so that it may start while the previous kernel of the stream drains (the kernel itself calls gdc_wait()
// before its first global access); falls back to a plain launch when PDL is off.
template <class... KP, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KP...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KP(args)...);
}

constexpr int kRow = 13;  // floats per rigid-body / root-state row: pos3 quat4 linvel3 angvel3

// Kernel arguments: the two C structs flattened, by value in param space.
struct KArgs {
  const float* rb;
  float* root;
  float* dof;
  float* root_out;   // where the predicated reset writes root / DOF rows (= root / dof unless the caller splits them)
  float* dof_out;
  const float* force;
  const float* pre;
  const float* init_root;
  const float* init_dof;
  const float* init_rb;
  const float* init_bal;     // optional compact reference pose [N, n_balance, 6]
  const float* reset_vel;
  const float* reset_yz;
  float* obs;
  float* rew;
  long long* reset;
  long long* progress;
  unsigned char* flags[PPK_MAX_FLAGS];
  double* stats;
  unsigned int* scratch;
  long long n;
  long long max_len;
  int pre_stride, pre_vx, pre_vz;
  int A, B, D;
  int hum[2];
  int ball;
  int paddle_body[2];
  int paddle_j[2];   // index of the paddle row inside ids (or -1: load it from global)
  int pelvis_body;
  int ids[2][PPK_MAX_BODY_IDS];
  int bal_ids[PPK_MAX_BODY_IDS];
  float alpha, power_coef, penalty, hit_table, not_hit, cross_net, die_penalty, hit_paddle, miss_coef;
  float term_dist;
  int phases;
  int write_flags;
  int reset_dof;
  int bulk_ok;   // tensors 16-byte aligned and ids[1..J) consecutive: bulk async staging is legal
  long long* last_hitter;   // ALIGN2
  // VecTask.step envelope (all optional)
  long long* timeout;
  const long long* actor_idx;
  const long long* dof_idx;
  int dof_per_env;
  int* reset_count;
  int* reset_actor_out;
  int* reset_dof_out;
  double* moments;           // PPK_PHASE_MOMENTS: [PPK_MOMENT_SLOTS][2 * obs width] fp64 column sums / sums of squares
  float clip_obs;            // > 0: observations are clamped to +-clip_obs where they are produced (VecTask.step)
  // tensor-map staging of the rigid-body rows (family kernel): inner coordinates (floats, 16-byte aligned) of the
  // boxes [humanoid][env parity] and where the wanted rows start inside them
  int span_c[2][2], span_off[2][2];
  int row0_c[2][2], row0_off[2][2];
  int row0_box;              // floats per staged ids[0] row (12 or 16)
  // soft start of the first wave (profiles/r2_staging_probe.md): CTA b < first_wave delays its loads by
  // (b / num_sms) * stagger cycles
  int stagger, first_wave, num_sms;
};

// Append env `env`'s actor / dof indices to the compacted reset lists (TILT:876-877).  Called by a
// fully converged warp; `resets` marks the lanes whose env is being reset this step.
__device__ __forceinline__ void append_reset_indices(const KArgs& k, bool resets, long long env, int lane) {
  if (k.reset_count == nullptr) return;
  const unsigned m = __ballot_sync(0xffffffffu, resets);
  if (m == 0u) return;
  const int leader = __ffs(m) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(k.reset_count, __popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (resets) {
    const int slot = base + __popc(m & ((1u << lane) - 1u));
    for (int a = 0; a < k.A; ++a) k.reset_actor_out[(size_t)slot * k.A + a] = (int)k.actor_idx[env * k.A + a];
    for (int d = 0; d < k.dof_per_env; ++d)
      k.reset_dof_out[(size_t)slot * k.dof_per_env + d] = (int)k.dof_idx[env * k.dof_per_env + d];
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Eight warp-wide sums for the price of ~one: each butterfly level halves the number of values a
// lane carries.  Returns, in lane L, the total over all lanes of value number (L >> 2) & 7.
__device__ __forceinline__ float warp_sum8(float v0, float v1, float v2, float v3, float v4, float v5, float v6,
                                           float v7, int lane) {
  const unsigned full = 0xffffffffu;
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float w0 = (b4 ? v4 : v0) + __shfl_xor_sync(full, b4 ? v0 : v4, 16);
  float w1 = (b4 ? v5 : v1) + __shfl_xor_sync(full, b4 ? v1 : v5, 16);
  float w2 = (b4 ? v6 : v2) + __shfl_xor_sync(full, b4 ? v2 : v6, 16);
  float w3 = (b4 ? v7 : v3) + __shfl_xor_sync(full, b4 ? v3 : v7, 16);
  float x0 = (b3 ? w2 : w0) + __shfl_xor_sync(full, b3 ? w0 : w2, 8);
  float x1 = (b3 ? w3 : w1) + __shfl_xor_sync(full, b3 ? w1 : w3, 8);
  float y = (b2 ? x1 : x0) + __shfl_xor_sync(full, b2 ? x0 : x1, 4);
  y += __shfl_xor_sync(full, y, 2);
  y += __shfl_xor_sync(full, y, 1);
  return y;
}

// VecTask.step's observation clamp (torch.clamp(obs_buf, -c, c): NaN passes through), applied where an
// observation is produced; c <= 0 means no clamp (the upstream default is inf)
__device__ __forceinline__ float clip_to(float v, float c) { return (v < -c) ? -c : ((v > c) ? c : v); }
__device__ __forceinline__ float clip_opt(float v, float c) { return (c > 0.0f) ? clip_to(v, c) : v; }

// streaming loads/stores: every byte of state is touched once per step
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }

// ---- heading frame -------------------------------------------------------------------------
// calc_heading_quat_inv(q) of torch_jit_utils, restated op by op (oracle/jit_utils_restated.py):
//   rot = my_quat_rotate(q, (1,0,0)); heading = atan2(rot.y, rot.x);
//   hq  = quat_unit([0, 0, sin(-heading/2), cos(-heading/2)])
// The heading quaternion only has z and w components (x, y are exact zeros).
struct Heading {
  float sz, cw;
};

__device__ __forceinline__ Heading heading_quat_inv(float qx, float qy, float qz, float qw) {
  // a = v*(2 w^2 - 1) with v = (1,0,0);  b = cross(q_vec, v)*w*2 = (0, qz*w*2, -qy*w*2);
  // c = q_vec*(q_vec . v)*2 = (qx*qx*2, qy*qx*2, qz*qx*2);  rot = (a + b) + c
  float a0 = 2.0f * (qw * qw) - 1.0f;
  float rx = (a0 + 0.0f) + (qx * qx) * 2.0f;
  float ry = (0.0f + (qz * qw) * 2.0f) + (qy * qx) * 2.0f;
#ifndef PPK_LIBM_HEADING
  // sin / cos of -atan2(ry, rx) / 2 from the half-angle identities: two IEEE square roots and three divisions instead of
  // atan2f + sinf + cosf (~60 instead of ~250 dependent instructions: the frames sit at the head of every CTA's chain).
  // Well conditioned on both half planes, atan2's signed-zero cases kept.  Against the CPU oracle (torch: SLEEF atan2 /
  // sin / cos) at 65 536 envs the rotated fields deviate by at most 5.2e-7 of the row scale, mean 8.1e-9; the libm
  // chain below (-DPPK_LIBM_HEADING) by 5.3e-7 / 8.5e-9: neither is the reference's own rounding, both sit 20x inside the
  // stated tolerance (tools/heading_error.py).
  float s, c;
  {
    const float r = sqrtf(rx * rx + ry * ry);
    if (r == 0.0f) {
      // atan2(+-0, +0) = +-0, atan2(+-0, -0) = +-pi
      const bool neg_x = signbit(rx);
      c = neg_x ? 0.0f : 1.0f;
      s = neg_x ? (signbit(ry) ? 1.0f : -1.0f) : 0.0f;
    } else {
      const float ch = rx / r, sh = ry / r;
      if (rx >= 0.0f) {
        c = sqrtf((1.0f + ch) * 0.5f);
        s = -(sh / (2.0f * c));
      } else {
        const float s2 = copysignf(sqrtf((1.0f - ch) * 0.5f), ry);
        c = sh / (2.0f * s2);
        s = -s2;
      }
    }
  }
#else
  float heading = atan2f(ry, rx);
  float half = (-heading) / 2.0f;
  float s = sinf(half);
  float c = cosf(half);
#endif
  float nrm = sqrtf(s * s + c * c);
  nrm = fmaxf(nrm, 1e-9f);
  Heading h;
  h.sz = s / nrm;
  h.cw = c / nrm;
  return h;
}

// my_quat_rotate((0,0,sz,cw), v):  a + b + c with the zero components dropped
__device__ __forceinline__ void rotate_heading(const Heading& h, float vx, float vy, float vz, float& ox, float& oy,
                                               float& oz) {
  float a0 = 2.0f * (h.cw * h.cw) - 1.0f;
  float bx = ((-(h.sz * vy)) * h.cw) * 2.0f;
  float by = ((h.sz * vx) * h.cw) * 2.0f;
  float cz = (h.sz * (h.sz * vz)) * 2.0f;
  ox = vx * a0 + bx;
  oy = vy * a0 + by;
  oz = vz * a0 + cz;
}

// ---- rewards --------------------------------------------------------------------------------
struct Scene {  // what the reward functions read, one env
  float bx, by, bz, vx, vz;   // ball position / velocity (root row)
  float pre_vx, pre_vz;       // ball velocity saved by pre_physics_step
  float px, py, pz;           // paddle position (rigid-body row)
  float hx;                   // humanoid root x
  float power_reward;         // -power_coefficient * sum|force*dof_vel|
  long long progress;         // after the +1
};

__device__ __forceinline__ float dist3(const Scene& s) {
  float dx = s.px - s.bx, dy = s.py - s.by, dz = s.pz - s.bz;
  return sqrtf(dx * dx + dy * dy + dz * dz);
}

// A3:1080-1173
__device__ __forceinline__ float reward_a3(const Scene& s, const KArgs& k, bool& die) {
  float d = dist3(s);
  float pos = 1.0f / (1.0f + 1.5f * d * d);
  bool hit = (s.pre_vx < 0.0f) && (s.vx > 0.0f);
  float vel = hit ? k.alpha * fabsf(s.vx) : 0.0f;
  float r = (pos + s.power_reward) + vel;
  bool missed = s.bx < s.px - 1e-3f;
  if (missed) r = r + k.penalty;
  die = missed || (s.bz < 0.1f);
  return r;
}

// TILT:1105-1270 / A4:1113-1278 (MIRROR=false), A4:1280-1439 (MIRROR=true)
template <bool MIRROR>
__device__ __forceinline__ float reward_tilt(const Scene& s, const KArgs& k, bool& cc, bool& rc, bool& nbbh, bool& die) {
  float d = dist3(s);
  float pos = 1.0f / (1.0f + 1.5f * d * d);
  bool outgoing = MIRROR ? (s.vx < 0.0f) : (s.vx > 0.0f);
  bool cond = (MIRROR ? (s.pre_vx > 0.0f) : (s.pre_vx < 0.0f)) && outgoing;
  float vel = (cond && !cc) ? k.alpha * fabsf(s.vx) : 0.0f;
  cc = cc || cond;
  bool missed = MIRROR ? (s.bx > s.hx + 0.05f) : (s.bx < s.hx - 0.05f);
  float r = missed ? (0.0f + k.penalty) : 0.0f;
  bool bounce_up = (s.bz < 0.83f) && outgoing && (s.by < 0.6f) && (s.by > -0.6f);
  bool near_half = MIRROR ? (s.bx > 1.06f) : (s.bx < 2.44f);
  bool far_table = MIRROR ? ((s.bx < 1.06f) && (s.bx > 0.4f)) : ((s.bx > 2.44f) && (s.bx < 3.1f));
  bool beyond = MIRROR ? (s.bx <= 0.4f) : (s.bx >= 3.1f);
  bool stage_a = near_half && bounce_up;
  float hit = (stage_a && !rc) ? k.not_hit : 0.0f;
  rc = rc || stage_a;
  nbbh = nbbh && !stage_a;
  bool stage_b = far_table && bounce_up && nbbh;
  if (stage_b && !rc) hit = k.hit_table;
  rc = rc || stage_b;
  if (beyond && outgoing && !rc) hit = k.not_hit;
  rc = rc || beyond;
  bool over_net = (s.bx > 1.7f) && (s.bx < 1.8f) && outgoing && (s.by < 0.4f) && (s.by > -0.4f) && (s.bz > 0.98f) &&
                  (s.bz < 1.14f);
  float net = over_net ? 400.0f : 0.0f;
  r = r + ((((pos + s.power_reward) + vel) + hit) + net);
  die = s.bz < 0.1f;
  return r;
}

// NES:1115-1322
__device__ __forceinline__ float reward_nes(const Scene& s, const KArgs& k, bool& pcc, bool& mbc, bool& die) {
  bool hit = (s.pre_vx < 0.0f) && (s.vx > 1.0f);
  bool behind = s.bx < s.hx - 0.05f;
  bool missed = behind || (s.bx < s.px - 0.1f);
  float r = (!mbc && missed) ? (0.0f + k.penalty) : 0.0f;
  mbc = mbc || missed;
  float dy = s.py - s.by, dz = s.pz - s.bz;
  float d = sqrtf(dy * dy + dz * dz);
  float pos = (!pcc || behind) ? 1.0f * expf((-20.0f * d) * d) : 0.0f;
  float vel = (hit && !pcc) ? k.alpha * fabsf(s.vx) : 0.0f;
  pcc = pcc || hit;
  r = r + ((pos + s.power_reward) + vel);
  if (s.bz < 0.1f) r = -800.0f + r;
  die = false;
  return r;
}

// ALIGN:1097-1230
__device__ __forceinline__ float reward_align(const Scene& s, const KArgs& k, bool& rc, bool& die) {
  float d = dist3(s);
  float pos = 1.0f / (1.0f + 1.5f * d * d);
  bool cond = (s.pre_vx < 0.0f) && (s.vx > 0.0f);
  float vel = cond ? k.alpha * fabsf(s.vx) : 0.0f;
  bool in_table = (s.bx > 2.2f) && (s.bx < 3.1f);
  bool bounce_up = (s.pre_vz < 0.0f) && (s.vz > 0.0f);
  bool no_bounce_before_half = (s.bx < 2.2f) && !bounce_up;
  bool award = in_table && bounce_up && no_bounce_before_half;
  float hit = (award && !rc) ? k.hit_table : 0.0f;
  rc = rc || award;
  if ((s.bx >= 3.1f) && (s.vx > 0.0f) && !rc) hit = k.not_hit;
  rc = rc || (s.bx >= 3.1f);
  float r = ((pos + s.power_reward) + vel) + hit;
  if (s.bx < s.hx - 0.05f) r = r + k.penalty;
  die = s.bz < 0.1f;
  return r;
}

// ALIGN:1233-1351 (two humanoids; returns both rewards, updates last_hitter).  `rc` is the caller's
// reward_calculated flag; the function is TorchScript, so its `|=` updates stay local (D16).
__device__ __forceinline__ void reward_align2(const Scene& s1, const Scene& s2, const KArgs& k, bool rc, long long& last_hitter,
                                              float& r1, float& r2, bool& die) {
  const float d1 = dist3(s1), d2 = dist3(s2);
  const float pos1 = 1.0f / (1.0f + 1.5f * d1 * d1), pos2 = 1.0f / (1.0f + 1.5f * d2 * d2);
  const bool c1 = (s1.pre_vx < 0.0f) && (s1.vx > 0.0f);
  const bool c2 = (s1.pre_vx > 0.0f) && (s1.vx < 0.0f);
  const float vel1 = c1 ? k.alpha * fabsf(s1.vx) : 0.0f;
  const float vel2 = c2 ? k.alpha * fabsf(s1.vx) : 0.0f;
  const bool range1 = (s1.bx > 2.2f) && (s1.bx < 3.1f);
  const bool range2 = (s1.bx < 1.3f) && (s1.bx > 0.4f);
  const bool bounce_up = (s1.pre_vz < 0.0f) && (s1.vz > 0.0f);
  float hit1 = (range1 && bounce_up && last_hitter == 1 && !rc) ? k.hit_table : 0.0f;
  float hit2 = (range2 && bounce_up && last_hitter == 2 && !rc) ? k.hit_table : 0.0f;
  rc = rc || (range1 && bounce_up) || (range2 && bounce_up);
  if ((s1.bx >= 3.1f) && last_hitter == 1 && !rc) hit1 = k.not_hit;
  if ((s1.bx <= -3.1f) && last_hitter == 2 && !rc) hit2 = k.not_hit;
  r1 = ((pos1 + s1.power_reward) + vel1) + hit1;
  r2 = ((pos2 + s1.power_reward) + vel2) + hit2;
  if (s1.bx < s1.hx - 0.05f) r1 = r1 + k.penalty;
  if (s1.bx > s2.hx + 0.05f) r2 = r2 + k.penalty;
  die = s1.bz < 0.1f;
  if (c1) last_hitter = 1;
  if (c2) last_hitter = 2;
}

}  // namespace ppk
