// The learner side of the path: what consumes obs_buf right after the task step.
//
//   rl_games `normalize_input: True` (cfg/train/HumanoidPingpongTiltG1PPO.yaml:51) wraps the network input in
//   RunningMeanStd (rl_games/algos_torch/running_mean_std.py, un-vendored): fp64 running mean / variance /
//   count, batch moments merged with the parallel-variance formula, y = clamp((x - mean) / sqrt(var + eps), +-5).
//   The first layer of the MLP (`units: [2048, ...]`, `activation: elu`, yaml:29-30) runs under
//   `mixed_precision: True` (yaml:50): fp16 operands, fp32 accumulation, fp16 result, ELU on the fp16 tensor.
//
// Kernels:
//   rms_moments_kernel   per-column sum / sum of squares of a batch, fp64, HBM-bound streaming reduction; the
//                        last CTA merges the batch into the running statistics (one launch per update)
//   rms_merge_kernel     stand-alone merge (after data-parallel ranks all-reduced the moments)
//   rms_apply_kernel     stand-alone normalisation, fp32 out
//   linear_pack_kernel   nn.Linear weight [units, width] (+ bias) fp32 -> fp16 tensor-core operand tiles (at init)
//   first_layer_kernel   obs -> clamp -> normalise -> fp16 -> tcgen05.mma (fp32 accumulators in TMEM, bias inside
//                        the K padding) -> fp16 -> ELU -> fp16 tiles -> TMA tensor stores.  The only GEMM on
//                        either side of the path; bound by the [rows, units] fp16 write, not by the math.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "ppk_tc.cuh"

namespace ppk {

// ---------------------------------------------------------------------------------------------------
// running mean / std
// ---------------------------------------------------------------------------------------------------
struct RmsArgs {
  int width;
  float eps, clip;        // clip <= 0: no observation clamp (VecTask clipObservations, upstream default inf)
  double* mean;           // [width] running_mean
  double* var;            // [width] running_var
  double* count;          // [1]
};

__device__ __forceinline__ float clamp_obs(float x, float clip) { return clip > 0.0f ? fminf(fmaxf(x, -clip), clip) : x; }

// RunningMeanStd._update_mean_var_count_from_moments for one column, batch mean / unbiased batch variance
// from the fp64 sums; clears the sums for the next batch.
__device__ __forceinline__ void rms_merge_column(const RmsArgs& a, double* sums, double batch_rows, double count, int c) {
  const double S = sums[c], SS = sums[a.width + c];
  const double bmean = S / batch_rows;
  const double bvar = (SS - S * bmean) / (batch_rows - 1.0);   // torch.var: correction = 1 (nan for one row, as torch)
  const double delta = bmean - a.mean[c];
  const double tot = count + batch_rows;
  const double m2 = a.var[c] * count + bvar * batch_rows + delta * delta * count * batch_rows / tot;
  a.mean[c] = a.mean[c] + delta * batch_rows / tot;
  a.var[c] = m2 / tot;
  sums[c] = 0.0;
  sums[a.width + c] = 0.0;
}

constexpr int kRmsMaxCtas = 512;      // upper bound of the grid (two CTAs per SM, 2 * sm_count() at launch)
constexpr int kRmsSlots = 16;         // copies of the 2W accumulators the CTAs' atomics are spread over

// Column sums and sums of squares of obs [rows, width] in fp64.  VEC = 4: threadIdx.x owns four adjacent
// columns (one 16-byte load per row), threadIdx.y walks the rows; eight rows in flight per thread.
// Every CTA adds its 2W partial sums into one of kRmsSlots copies of the accumulators (same-address fp64
// atomics serialise at ~50 ns each: 148 onto one address cost 7 us, 19 cost 1 us); the last CTA to finish
// (ticket) folds the copies into sums[0..2W) (+=) and clears them; with merge != 0 it then folds the batch
// into the running statistics: one launch per update.
// scratch layout: [0,2W) sums | [2W] ticket | [2W+1 + slot*2W ...) accumulator copies
template <int VEC>
__global__ void rms_moments_kernel(RmsArgs a, const float* __restrict__ obs, long long rows, double* sums, int merge,
                                   double batch_rows) {
  extern __shared__ double red[];            // [2 * VEC][blockDim.y][blockDim.x]
  const int cx = threadIdx.x, ry = threadIdx.y, R = blockDim.y, WX = blockDim.x;
  const int width = a.width;
  const bool on = cx * VEC < width;
  double s[VEC], ss[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) s[v] = ss[v] = 0.0;
  if (on) {
    const long long step = (long long)gridDim.x * R;
    constexpr int U = 8;
    for (long long r0 = (long long)blockIdx.x * R + ry; r0 < rows; r0 += U * step) {
      float x[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long r = r0 + u * step;
#pragma unroll
        for (int v = 0; v < VEC; ++v) x[u][v] = 0.0f;
        if (r < rows) {
          const float* p = obs + r * width + cx * VEC;
          if (VEC == 4) { const float4 t = __ldcs(reinterpret_cast<const float4*>(p)); x[u][0] = t.x; x[u][1] = t.y; x[u][2] = t.z; x[u][VEC - 1] = t.w; }
          else x[u][0] = __ldcs(p);
          if (a.clip > 0.0f) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) x[u][v] = clamp_obs(x[u][v], a.clip);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)         // rows past the end contribute exact zeros
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const double d = (double)x[u][v];
          s[v] += d;
          ss[v] += d * d;
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      red[((2 * v) * R + ry) * WX + cx] = s[v];
      red[((2 * v + 1) * R + ry) * WX + cx] = ss[v];
    }
  }
  __syncthreads();
  // thread (cx, ry < 2*VEC) folds one of the 2*VEC partial arrays of column group cx into the CTA's scratch row
  double* slot = sums + 2 * width + 1 + (size_t)(blockIdx.x % kRmsSlots) * 2 * width;
  if (on && ry < 2 * VEC) {
    double t = 0.0;
    for (int y = 0; y < R; ++y) t += red[(ry * R + y) * WX + cx];
    const int v = ry >> 1;
    atomicAdd(slot + ((ry & 1) ? width : 0) + cx * VEC + v, t);
  }
  __shared__ unsigned int last;
  __threadfence();
  __syncthreads();
  if (cx == 0 && ry == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(sums + 2 * width);
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    if (last) *ticket = 0u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  const double count = *a.count;
  __syncthreads();
  double* slots = sums + 2 * width + 1;
  for (int c = ry * WX + cx; c < 2 * width; c += R * WX) {
    double t[kRmsSlots];
#pragma unroll
    for (int b = 0; b < kRmsSlots; ++b) t[b] = __ldcg(slots + (size_t)b * 2 * width + c);     // written by L2 atomics
    double tot = 0.0;
#pragma unroll
    for (int b = 0; b < kRmsSlots; ++b) { tot += t[b]; slots[(size_t)b * 2 * width + c] = 0.0; }
    sums[c] += tot;
  }
  if (!merge) return;
  __threadfence_block();
  __syncthreads();
  for (int c = ry * WX + cx; c < width; c += R * WX) rms_merge_column(a, sums, batch_rows, count, c);
  if (cx == 0 && ry == 0) *a.count = count + batch_rows;
}

// stand-alone merge (after the ranks all-reduced the moments).  One CTA.
__global__ void rms_merge_kernel(RmsArgs a, double* sums, double batch_rows) {
  const double count = *a.count;
  __syncthreads();
  for (int c = threadIdx.x; c < a.width; c += blockDim.x) rms_merge_column(a, sums, batch_rows, count, c);
  if (threadIdx.x == 0) *a.count = count + batch_rows;
}

// The column moments a step kernel left in PPK_MOMENT_SLOTS slot copies (PPK_PHASE_MOMENTS): fold them into sums[0..2W),
// clear the slots, and (merge != 0) fold the batch into the running statistics.  One CTA: 2W <= 1024 sums of 64 doubles.
// Launched as a programmatic dependent of the step kernel: it is resident, with the running statistics already in
// registers, when the step's last CTA retires, and every slot value of a thread is requested at once -- the kernel is a
// chain of L2 round trips (was: launch, count, three batches of slot loads, mean / var).
__global__ void __launch_bounds__(1024)
rms_fold_step_kernel(RmsArgs a, double* sums, double* slots, double batch_rows, int merge) {
  // thread (g, c): column c of the 2W sums, slot copies g, g + G, ...
  __shared__ double part[1024];
  const int width = a.width, cols = 2 * width;
  const int G = max(1, min((int)blockDim.x / cols, PPK_MOMENT_SLOTS));
  const int per = blockDim.x / G;                 // columns handled per pass
  // the running statistics were written by an earlier rms kernel of the stream, which the step kernel ahead of this one
  // has already waited for: they may be read before the dependency wait
  const double count = *a.count;
  double old_mean = 0.0, old_var = 0.0;
  if (merge && (int)threadIdx.x < width) { old_mean = a.mean[threadIdx.x]; old_var = a.var[threadIdx.x]; }
  gdc_wait();
  for (int c0 = 0; c0 < cols; c0 += per) {
    const int g = threadIdx.x / per, c = c0 + threadIdx.x % per;
    double tot = 0.0;
    if (g < G && c < cols) {
      constexpr int kBatch = 16;
      for (int b0 = g; b0 < PPK_MOMENT_SLOTS; b0 += kBatch * G) {
        double t[kBatch];
#pragma unroll
        for (int i = 0; i < kBatch; ++i) {
          const int b = b0 + i * G;
          t[i] = b < PPK_MOMENT_SLOTS ? __ldcg(slots + (size_t)b * cols + c) : 0.0;       // written by L2 atomics of the step kernel
        }
#pragma unroll
        for (int i = 0; i < kBatch; ++i) {
          const int b = b0 + i * G;
          if (b < PPK_MOMENT_SLOTS) { tot += t[i]; slots[(size_t)b * cols + c] = 0.0; }
        }
      }
    }
    part[threadIdx.x] = tot;
    __syncthreads();
    if (g == 0 && c < cols) {
      for (int k = 1; k < G; ++k) tot += part[threadIdx.x + k * per];
      sums[c] += tot;
    }
    __syncthreads();
  }
  if (!merge) return;
  __threadfence_block();
  __syncthreads();
  if (width <= (int)blockDim.x) {
    const int c = threadIdx.x;
    if (c < width) {            // rms_merge_column with the old statistics already here
      const double S = sums[c], SS = sums[width + c];
      const double bmean = S / batch_rows;
      const double bvar = (SS - S * bmean) / (batch_rows - 1.0);
      const double delta = bmean - old_mean;
      const double tot = count + batch_rows;
      const double m2 = old_var * count + bvar * batch_rows + delta * delta * count * batch_rows / tot;
      a.mean[c] = old_mean + delta * batch_rows / tot;
      a.var[c] = m2 / tot;
      sums[c] = 0.0;
      sums[width + c] = 0.0;
    }
  } else {
    for (int c = threadIdx.x; c < width; c += blockDim.x) rms_merge_column(a, sums, batch_rows, count, c);
  }
  if (threadIdx.x == 0) *a.count = count + batch_rows;
}

// y = clamp((clamp(x) - float(mean)) / sqrt(float(var) + eps), -5, 5), fp32; same thread layout as the moments
template <int VEC>
__global__ void rms_apply_kernel(RmsArgs a, const float* __restrict__ obs, long long rows, float* __restrict__ out) {
  const int cx = threadIdx.x, ry = threadIdx.y, R = blockDim.y;
  if (cx * VEC >= a.width) return;
  float m[VEC], d[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    m[v] = (float)a.mean[cx * VEC + v];
    d[v] = sqrtf((float)a.var[cx * VEC + v] + a.eps);
  }
  const long long step = (long long)gridDim.x * R;
  constexpr int U = 4;                       // rows in flight per thread
  for (long long r0 = (long long)blockIdx.x * R + ry; r0 < rows; r0 += U * step) {
    float x[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * step;
      if (r < rows) {
        const long long off = r * a.width + cx * VEC;
        if (VEC == 4) { const float4 t = __ldcs(reinterpret_cast<const float4*>(obs + off)); x[u][0] = t.x; x[u][1] = t.y; x[u][2] = t.z; x[u][VEC - 1] = t.w; }
        else x[u][0] = __ldcs(obs + off);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * step;
      if (r >= rows) break;
      const long long off = r * a.width + cx * VEC;
#pragma unroll
      for (int v = 0; v < VEC; ++v) x[u][v] = fminf(fmaxf((clamp_obs(x[u][v], a.clip) - m[v]) / d[v], -5.0f), 5.0f);
      if (VEC == 4) __stcs(reinterpret_cast<float4*>(out + off), make_float4(x[u][0], x[u][1], x[u][2], x[u][VEC - 1]));
      else __stcs(out + off, x[u][0]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// first layer
// ---------------------------------------------------------------------------------------------------
#ifndef PPK_FL_DBG
#define PPK_FL_DBG 0        // A/B builds only: 1 no epilogue math, 2 no global stores, 4 no MMA, 8 no prep loads, 16 no weight loads
#endif
constexpr int kFlM = 128;            // rows per tile = TMEM lanes
constexpr int kFlN = 256;            // units per chunk = fp32 accumulator columns per buffer
constexpr int kFlPrepWarps = 8;          // latency-bound (global loads) and competing with 16 epilogue warps for issue slots
constexpr int kFlEpiWarps = 16;          // four per TMEM lane quarter: the epilogue is latency-bound per warp
constexpr int kFlEpiCols = kFlN / (kFlEpiWarps / 4);   // accumulator columns per epilogue warp and unit
constexpr int kFlThreads = 32 * (2 + kFlPrepWarps + kFlEpiWarps);   // producer, mma, prep, epilogue
constexpr int kFlALbo = kFlM * 16 + 16;   // +16: the prep warps write 16-byte pieces of different K chunks
constexpr int kFlBLbo = kFlN * 16;
constexpr int kFlStageRow = kFlEpiCols * 2;        // 128 B: one row of a warp's [32 x 64] fp16 output tile (128B-swizzled)
static_assert(kFlStageRow == 128, "the output staging tile is laid out for the 128-byte TMA swizzle");

// K is padded to a multiple of 16 with at least ONE spare column: column `width` of the row tile is the
// constant 1 and column `width` of the weights is the bias, so the bias add happens inside the MMA (fp32
// accumulation either way) instead of costing the epilogue an add and a load per element.
__host__ __device__ constexpr int fl_kpad(int width) { return (width + 1 + 15) / 16 * 16; }
// Inputs up to 95 columns keep the whole K of a weight chunk in one ring stage and two row-tile buffers.  Wider
// inputs (ADOF: 313 -> KP = 320) stream the weights in K blocks of 64 and keep ONE row-tile buffer (82 KB).
__host__ __device__ constexpr int fl_kblk(int kp) { return kp <= 96 ? kp : 64; }

template <int KP, int KBLK>
struct FlLayout {
  static_assert(KP % KBLK == 0 && KBLK % 16 == 0, "K blocking");
  static constexpr int kKc = KP / 8;
  static constexpr int kNkb = KP / KBLK;                               // K blocks per unit
  static constexpr int kABufs = (KP == KBLK) ? 2 : 1;
  static constexpr int kABytes = kKc * kFlALbo;
  static constexpr int kBBytes = (KBLK / 8) * kFlBLbo;                 // one ring stage = one K block of a chunk
  static constexpr int kOffStage = 0;                                  // 1024-byte aligned tiles for the swizzle
  static constexpr int kOffA = 32 * kFlEpiWarps * kFlStageRow;
  static constexpr int kOffB = kOffA + kABufs * kABytes;
  static constexpr int kOffCst = kOffB + 2 * kBBytes;
  static constexpr int kOffBar = kOffCst + 3 * KP * 4;      // mean, den, 1/den per column
  static constexpr int kBytes = kOffBar + 12 * 8 + 16;
  static_assert(kOffA % 16 == 0 && kOffB % 16 == 0 && kOffCst % 16 == 0 && kOffBar % 8 == 0, "alignment");
  static_assert(kBytes <= 227 * 1024, "shared memory budget");
};

// packed weight blob: [units/256 chunks][K blocks][KBLK/8][256][8] fp16 (one contiguous ring stage per chunk and K
// block); column `width` holds the bias, the rest of the pad is 0
__global__ void linear_pack_kernel(const float* __restrict__ w, const float* __restrict__ bias, int units, int width, int kp,
                                   __half* __restrict__ packed) {
  const long long total = (long long)units * kp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i enumerates the packed order
    const int j = (int)(i & 7);
    const long long t = i >> 3;
    const int nl = (int)(t % kFlN);
    const long long t2 = t / kFlN;                 // = (chunk * nkb + kb) * (kblk / 8) + kc_local, and kb * (kblk / 8) + kc_local
    const int kc = (int)(t2 % (kp / 8));           //   is the global 16-byte K piece: the blocked order IS the plain order
    const int chunk = (int)(t2 / (kp / 8));
    const int n = chunk * kFlN + nl, k = kc * 8 + j;
    float v = 0.0f;
    if (k < width) v = w[(size_t)n * width + k];
    else if (k == width && bias != nullptr) v = bias[n];
    packed[i] = __float2half_rn(v);
  }
}

#ifdef PPK_TRACE
__device__ __forceinline__ void fl_stamp(int slot) {
  if (g_trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace[(size_t)blockIdx.x * 16 + slot] = t;
  }
}
#define FL_STAMP(cond, slot) do { if (cond) fl_stamp(slot); } while (0)
#else
#define FL_STAMP(cond, slot)
#endif

struct FlArgs {
  const float* obs;          // [rows, width]
  long long rows;
  int width, units, activation;   // activation: 0 none, 1 ELU
  RmsArgs rms;               // rms.mean == nullptr: no normalisation
  const unsigned char* packed;
  __half* out;               // [rows, units]
};

// 2^x, one MUFU (the non-ftz form costs two more multiplies and a compare per element for denormal
// results, which the following "- 1" discards anyway)
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// n / d with r = RN(1/d) computed once per column: quotient estimate plus two exact-residual corrections
// (the refinement div.rn itself performs, minus its range checks: |n| and d are ordinary normal numbers
// here).  Five instructions instead of the dozen of a generic IEEE division, same result (0 mismatches against
// n / d on 2^33 random operand pairs: tools/div_by_const_check.cu).
__device__ __forceinline__ float div_by_const(float n, float d, float r) {
  float q = n * r;
  q = fmaf(fmaf(-d, q, n), r, q);
  q = fmaf(fmaf(-d, q, n), r, q);
  return q;
}

// One operand tile: rows [mt*128, +128) of obs -> clamp -> normalise -> fp16 -> K-major core-matrix layout at
// `dst`.  An item is one row's 8 consecutive columns (one 16-byte operand piece); consecutive threads take
// consecutive pieces of a row, so the global loads coalesce.  Columns >= width load 0 and have mean 0 / den 1,
// rows >= rows produce values nobody stores: no per-element selects.
template <int KP, int BATCH>
__device__ __forceinline__ void fl_prep_tile(const FlArgs& k, const float* cst, long long mt, unsigned char* dst, int t,
                                             int nthreads) {
  constexpr int kKc = KP / 8, kItems = kFlM * kKc;
  const bool vec = (k.width % 4 == 0) && ((reinterpret_cast<uintptr_t>(k.obs) & 15u) == 0);
  const float inf = __int_as_float(0x7f800000);
  const float clip = k.rms.clip > 0.0f ? k.rms.clip : inf;            // VecTask clipObservations
  const float lim = k.rms.mean != nullptr ? 5.0f : inf;               // RunningMeanStd output clamp
  for (int base = t; base < kItems; base += nthreads * BATCH) {
    float x[BATCH][8];
    // all loads of the batch first (independent, in flight together), then the arithmetic
#pragma unroll
    for (int bi = 0; bi < BATCH; ++bi) {
      const int item = base + bi * nthreads;
      const int r = item / kKc, kc = item - r * kKc;
      const long long row = mt * kFlM + r;
#pragma unroll
      for (int i = 0; i < 8; ++i) x[bi][i] = 0.0f;
      if (item < kItems && row < k.rows && !(PPK_FL_DBG & 8)) {
        const float* src = k.obs + row * k.width + kc * 8;
        if (vec) {
          if (kc * 8 < k.width) { float4 v = __ldcs(reinterpret_cast<const float4*>(src)); x[bi][0] = v.x; x[bi][1] = v.y; x[bi][2] = v.z; x[bi][3] = v.w; }
          if (kc * 8 + 4 < k.width) { float4 v = __ldcs(reinterpret_cast<const float4*>(src) + 1); x[bi][4] = v.x; x[bi][5] = v.y; x[bi][6] = v.z; x[bi][7] = v.w; }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (kc * 8 + i < k.width) x[bi][i] = __ldcs(src + i);
        }
      }
    }
#pragma unroll
    for (int bi = 0; bi < BATCH; ++bi) {
      const int item = base + bi * nthreads;
      if (item >= kItems) break;
      const int r = item / kKc, kc = item - r * kKc;
      const float4* m4 = reinterpret_cast<const float4*>(cst + kc * 8);
      const float4* d4 = reinterpret_cast<const float4*>(cst + KP + kc * 8);
      const float4* r4 = reinterpret_cast<const float4*>(cst + 2 * KP + kc * 8);
      const float4 ma = m4[0], mb = m4[1], da = d4[0], db = d4[1], ra = r4[0], rb = r4[1];
      const float m[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
      const float d[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
      const float rc[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float y = x[bi][i];
        if (k.rms.clip > 0.0f) y = fminf(fmaxf(y, -clip), clip);
        x[bi][i] = fminf(fmaxf(div_by_const(y - m[i], d[i], rc[i]), -lim), lim);
      }
      uint4 pk;
      pk.x = pack_half2(x[bi][0], x[bi][1]); pk.y = pack_half2(x[bi][2], x[bi][3]);
      pk.z = pack_half2(x[bi][4], x[bi][5]); pk.w = pack_half2(x[bi][6], x[bi][7]);
      *reinterpret_cast<uint4*>(dst + kc * kFlALbo + r * 16) = pk;
    }
  }
}

template <int KP, int KBLK, int ACT>
__global__ void __launch_bounds__(kFlThreads, 1)
first_layer_kernel(const __grid_constant__ FlArgs k, const __grid_constant__ CUtensorMap out_map) {
  using L = FlLayout<KP, KBLK>;
  constexpr int NKB = L::kNkb, ABUFS = L::kABufs;
  extern __shared__ __align__(1024) unsigned char fl_smem[];
  unsigned char* a_s = fl_smem + L::kOffA;
  unsigned char* b_s = fl_smem + L::kOffB;
  unsigned char* stage_s = fl_smem + L::kOffStage;
  float* cst = reinterpret_cast<float*>(fl_smem + L::kOffCst);
  uint64_t* bars = reinterpret_cast<uint64_t*>(fl_smem + L::kOffBar);
  uint64_t *b_full = bars, *b_empty = bars + 2, *a_full = bars + 4, *a_empty = bars + 6, *acc_full = bars + 8,
           *acc_empty = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NC = k.units / kFlN;
  const long long MT = (k.rows + kFlM - 1) / kFlM;
  const long long U = MT * NC;
  const long long u_begin = U * blockIdx.x / gridDim.x, u_end = U * (blockIdx.x + 1) / gridDim.x;
  const long long mt_begin = u_begin / NC;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1);
      mbar_init(a_full + i, 32 * kFlPrepWarps); mbar_init(a_empty + i, 1);
      mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 32 * kFlEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, 512);
  // normalisation constants: float(mean), sqrt(float(var) + eps)
  for (int c = threadIdx.x; c < KP; c += kFlThreads) {
    const bool on = k.rms.mean != nullptr && c < k.width;
    cst[c] = on ? (float)k.rms.mean[c] : (c == k.width ? -1.0f : 0.0f);      // column `width`: (0 - -1) / 1 = the bias column's 1
    const float den = on ? sqrtf((float)k.rms.var[c] + k.rms.eps) : 1.0f;
    cst[KP + c] = den;
    cst[2 * KP + c] = 1.0f / den;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  FL_STAMP(threadIdx.x == 0, 0);
  // weight stage `bi` of this CTA's sequence (unit-major, K blocks inside a unit)
  auto load_weights = [&](long long bi) {
    const long long u = u_begin + bi / NKB;
    const int kb = (int)(bi % NKB), nc = (int)(u % NC), s = (int)(bi & 1);
    mbar_arrive_expect_tx(b_full + s, (uint32_t)L::kBBytes);
    bulk_g2s(b_s + s * L::kBBytes, k.packed + ((size_t)nc * NKB + kb) * L::kBBytes, (uint32_t)L::kBBytes, b_full + s);
  };
  const long long n_stages = (u_end - u_begin) * NKB;
  // both ring stages are requested before the first row tile is prepared, so the two latencies overlap
  if (threadIdx.x == 0 && !(PPK_FL_DBG & 16))
    for (long long bi = 0; bi < 2 && bi < n_stages; ++bi) load_weights(bi);
  // the CTA's first row tile is on the critical path of everything: all warps prepare it together
  fl_prep_tile<KP, 2>(k, cst, mt_begin, a_s, threadIdx.x, kFlThreads);
  tc::fence_proxy_async_smem();
  __syncthreads();

  if (warp == 0) {
    // ===== weight producer: one bulk copy per unit and K block into the 2-deep ring =====
    if (lane == 0) {
      for (long long bi = (PPK_FL_DBG & 16) ? 0 : 2; bi < n_stages; ++bi) {
        const int s = (int)(bi & 1), ph = (int)((bi >> 1) & 1);
        mbar_wait_relaxed(b_empty + s, ph ^ 1);
        if (PPK_FL_DBG & 16) { tc::mbar_arrive(b_full + s); continue; }
        load_weights(bi);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: a single thread =====
    if (lane == 0) {
      constexpr uint32_t idesc = tc::instr_desc_f16_f32(kFlM, kFlN);
      for (long long u = u_begin; u < u_end; ++u) {
        const long long it = u - u_begin;
        const int s = (int)(it & 1), ph = (int)((it >> 1) & 1);
        const long long mt = u / NC;
        const int nc = (int)(u - mt * NC);
        const long long j = mt - mt_begin;
        const int ab = (int)(j % ABUFS), aph = (int)((j / ABUFS) & 1);
        if (it == 0 || nc == 0) mbar_wait(a_full + ab, aph);
        mbar_wait(acc_empty + s, ph ^ 1);
        const uint32_t a0 = smem_u32(a_s + ab * L::kABytes);
        for (int kb = 0; kb < NKB; ++kb) {
          const long long bi = it * NKB + kb;
          const int sb = (int)(bi & 1), phb = (int)((bi >> 1) & 1);
          mbar_wait(b_full + sb, phb);
          tc::fence_after_sync();
          const uint32_t b0 = smem_u32(b_s + sb * L::kBBytes);
#pragma unroll
          for (int kk = 0; kk < ((PPK_FL_DBG & 4) ? 0 : KBLK / 16); ++kk)
            tc::mma_f16(tmem_base + s * kFlN, tc::smem_desc(a0 + (kb * (KBLK / 16) + kk) * 2 * kFlALbo, kFlALbo, 128),
                        tc::smem_desc(b0 + kk * 2 * kFlBLbo, kFlBLbo, 128), idesc, (kb | kk) > 0);
          tc::mma_commit(b_empty + sb);
        }
        if (u == u_end - 1 || nc == NC - 1) tc::mma_commit(a_empty + ab);
        tc::mma_commit(acc_full + s);
        FL_STAMP(it == 0, 3);
        FL_STAMP(u == u_end - 1, 6);
      }
    }
  } else if (warp < 2 + kFlPrepWarps) {
    // ===== row-tile preparation (tiles after the CTA's first): one tile ahead of the MMA =====
    const int t = threadIdx.x - 64;     // warps 2..9
    const long long mt_last = (u_end - 1) / NC;
    for (long long mt = mt_begin; mt <= mt_last; ++mt) {
      const long long j = mt - mt_begin;
      const int ab = (int)(j % ABUFS), aph = (int)((j / ABUFS) & 1);
      if (mt < mt_last) {               // pull the tile after this one into L2 while this one is worked on
        const long long r0 = (mt + 1) * kFlM;
        const long long bytes = (min((long long)kFlM, k.rows - r0)) * k.width * 4;
        const char* p = reinterpret_cast<const char*>(k.obs + r0 * k.width);
        for (long long off = (long long)t * 128; off < bytes; off += 128LL * 32 * kFlPrepWarps) prefetch_l2(p + off);
      }
      if (j > 0) {                      // tile 0 was prepared by the whole CTA above
        mbar_wait_relaxed(a_empty + ab, aph ^ 1);
        fl_prep_tile<KP, 3>(k, cst, mt, a_s + ab * L::kABytes, t, 32 * kFlPrepWarps);
        tc::fence_proxy_async_smem();
      }
      tc::mbar_arrive(a_full + ab);
      FL_STAMP(t == 0 && j == 0, 1);
      FL_STAMP(t == 0 && j == 1, 2);
    }
  } else {
    // ===== epilogue: TMEM -> registers -> fp16 -> activation -> fp16 -> swizzled warp tile -> tensor store =====
    const int ew = warp - (2 + kFlPrepWarps);
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int cq = ew >> 2;                 // which kFlEpiCols-wide slice of the chunk
    // the warp's [32 rows x 64 units] fp16 tile, 128B-swizzled: 16-byte piece j of row r sits at piece j ^ (r & 7)
    unsigned char* wstage = stage_s + (size_t)ew * 32 * kFlStageRow;
    unsigned char* my_row = wstage + lane * kFlStageRow;
    const int sw = lane & 7;
    for (long long u = u_begin; u < u_end; ++u) {
      const long long it = u - u_begin;
      const int s = (int)(it & 1), ph = (int)((it >> 1) & 1);
      const long long mt = u / NC;
      const int nc = (int)(u - mt * NC);
      const int col0 = nc * kFlN + cq * kFlEpiCols;
      mbar_wait(acc_full + s, ph);
      tc::fence_after_sync();
      FL_STAMP(ew == 0 && lane == 0 && it == 0, 4);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * kFlN + cq * kFlEpiCols);
      // the tensor store of the previous unit has finished reading the tile
      if (lane == 0) tc::bulk_wait_read0();
      __syncwarp();
#pragma unroll 1
      for (int g = 0; g < kFlEpiCols / 32; ++g) {
        uint32_t v[32];
        tc::tmem_ld32(taddr + g * 32, v);
        tc::tmem_ld_wait();
        if (g == kFlEpiCols / 32 - 1) {     // accumulator drained: the next MMA into this buffer may start
          tc::fence_before_sync();
          tc::mbar_arrive(acc_empty + s);
        }
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float f[4] = {__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                              __uint_as_float(v[4 * i + 3])};       // bias already inside (K padding column)
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            __half2 lin = __floats2half2_rn(f[2 * p], f[2 * p + 1]);     // the fp16 output of the linear layer
            if (ACT == 1 && !(PPK_FL_DBG & 1)) {
              const float2 xf = __half22float2(lin);
              const float e0 = xf.x > 0.0f ? xf.x : ex2_ftz(xf.x * 1.4426950408889634f) - 1.0f;
              const float e1 = xf.y > 0.0f ? xf.y : ex2_ftz(xf.y * 1.4426950408889634f) - 1.0f;
              lin = __floats2half2_rn(e0, e1);
            }
            o[2 * i + p] = *reinterpret_cast<uint32_t*>(&lin);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(my_row + (((g * 4 + i) ^ sw) << 4)) = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
      }
      // one tensor store per warp and unit (rows past the end of the batch are clipped by the copy engine).
      // (One 1-D bulk store per ROW was tried first: the engine serves ~1 operation per 46 cycles per SM, and
      // 256 of them per unit cost 4x the HBM budget; staged, coalesced st.global cost 146 instructions here.)
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && !(PPK_FL_DBG & 2)) {
        tc::tensor_store_2d(&out_map, col0, (int)(mt * kFlM + q * 32), wstage);
        tc::bulk_commit();
      }
      FL_STAMP(ew == 0 && lane == 0 && it == 0, 5);
      FL_STAMP(ew == 0 && lane == 0 && it == 8, 9);
      FL_STAMP(ew == 0 && lane == 0 && it == 16, 10);
      FL_STAMP(ew == 0 && lane == 0 && u == u_end - 1, 7);
    }
    if (lane == 0) tc::bulk_wait_all();
  }

  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  FL_STAMP(threadIdx.x == 0, 8);
  if (warp == 2) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace ppk
