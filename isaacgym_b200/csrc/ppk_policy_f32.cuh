// The fp32 variant of the first policy layer: what the ROLLOUT forward computes (rl_games `get_action_values`, under
// no_grad and outside autocast: act(linear(running_mean_std(obs))) in fp32), where ppk_policy.cuh implements the learner's
// autocast(fp16) semantics.
//
// fp32 products on the 5th-generation tensor cores: each operand is split into two TF32 numbers, x = hi + lo with
// hi = RN_tf32(x), lo = RN_tf32(x - hi) (x - hi is exact in fp32), and
//     x * w  ~=  hi_x * hi_w + lo_x * hi_w + hi_x * lo_w            (the dropped lo_x * lo_w is <= 2^-22 |x w|)
// is three `tcgen05.mma.kind::tf32` into the same fp32 accumulator in TMEM: the result differs from an fp32 FMA chain by
// ~1e-6 of sum_k |x_k w_k| (a different, not a worse, rounding pattern than cuBLAS SGEMM's summation order).
//
// Structure (as first_layer_kernel): persistent CTAs, unit = [128 rows x 256 units] per CTA; warp 0 streams the packed
// weight stages (one K step of 8: hi and lo pieces) through a ring; warp 1 (one thread) issues the MMAs; warps 2-5 prepare
// the row tile K step by K step (cp.async raw ring -> clamp -> normalise -> split -> K-major core matrices; the single
// tile buffer is released and refilled per K step while the previous tile's last unit is still being multiplied); warps
// 6-21 are the epilogue: TMEM -> ELU in fp32 -> 128B-swizzled [32 x 32] fp32 tile -> TMA tensor store.
// CL = 1 (default): single CTAs.
// CL = 2 (PPK_FL32_CLUSTER=2): the two CTAs of a cluster form a CTA pair (`cta_group::2`): ONE MMA of M = 256 covers both
// CTAs' row tiles, and each CTA holds only HALF of the B operand (128 of the chunk's 256 units), so the weight bytes every
// SM pulls from L2 -- 8 per element, re-streamed for every row tile, ~5 TB/s chip-wide with single CTAs -- halve.  The
// leader (rank 0) issues the MMAs; the peer's MMA warp forwards its CTA's "stage landed + K step of the row tile ready /
// accumulator drained" events to barriers in the leader's shared memory, and the leader's commits are multicast to both
// CTAs' barriers.  Parity-green, but measured slower (a pair MMA takes twice a single one with this operand layout), so
// it is the A/B variant, not the default.  (Raw issue rate without operand waits, `-DPPK_F32_DBG=336`: 178 cycles per
// M128 N256 K8 MMA, 367 per M256 pair MMA: the same per SM, ~70 % of the tensor pipe's rate, 85-88 us for the whole layer;
// polling the remotely signalled barriers with test_wait instead of try_wait made the pair slower, not faster.)
// Measured, 65 536 x 80 -> 2048, ELU (`gpurun_out/r2_f32_ab*.log`): 202 us (no activation: 178) = 2.7 TB/s of output; the
// parts alone: epilogue + stores 129 us, weights + MMAs + row tiles 138 us; they share L2 bandwidth (180 KB of weights in,
// 128 KB of output out per unit) and overlap to 6.4 us per unit.
#pragma once
#include "ppk_policy.cuh"

#ifndef PPK_F32_DBG
#define PPK_F32_DBG 0       // A/B builds only: 2 no global stores, 4 no MMA, 16 no weight loads, 32 no staging writes, 64 no epilogue, 256 no operand waits, 512 epilogue stores straight from registers (2.5x slower than the staged tensor stores)
#endif

namespace ppk {

constexpr int kF32EpiWarps = 16;                    // four per TMEM lane quarter, 64 accumulator columns each (latency-bound per warp)
constexpr int kF32EpiCols = kFlN / (kF32EpiWarps / 4);
constexpr int kF32PrepWarps = 4;                    // 93 registers per thread for the epilogue's two TMEM loads in flight
constexpr int kF32PrepFirst = 2;                    // warp 0: weight producer, warp 1: MMA issuer
constexpr int kF32EpiFirst = kF32PrepFirst + kF32PrepWarps;
constexpr int kF32Threads = 32 * (kF32EpiFirst + kF32EpiWarps);
constexpr int kF32StepBytes = 2 * 2 * kFlN * 16;    // one K step (8 columns) of a chunk: [hi | lo][2 pieces][256 units][16 B]
constexpr int kF32RingSteps = 3;                    // K steps of weights in flight (a fourth changes nothing: the stream is bandwidth-, not latency-bound)
constexpr int kF32RingBytes = kF32RingSteps * kF32StepBytes;
constexpr int kF32EpiTile = 32 * 128;               // one epilogue warp's [32 rows x 32 units] fp32 tile
constexpr int kF32MaxRing = 2 * kF32RingSteps;      // stages of a CTA pair are half K steps
constexpr int kF32MaxNkb = 12;                      // K steps of the widest row tile (KP = 96)

__host__ __device__ constexpr int f32_kpad(int width) { return (width + 1 + 7) / 8 * 8; }

template <int KP>
struct F32Layout {
  static_assert(KP % 8 == 0, "K step of kind::tf32");
  static constexpr int kPieces = KP / 4;                                // 16-byte K pieces per row
  static constexpr int kNkb = KP / 8;                                   // weight stages per unit
  static constexpr int kAPart = kPieces * kFlALbo;                      // hi (or lo) part of the row tile
  static constexpr int kOffStage = 0;                                   // 1024-byte aligned tiles for the swizzle
  static constexpr int kOffA = kF32EpiWarps * kF32EpiTile;
  static constexpr int kOffB = kOffA + 2 * kAPart;
  static constexpr int kRawDepth = KP <= 88 ? (kNkb < 5 ? kNkb : 5) : 3;  // raw K steps of the next rows in flight (4 KB each)
  static constexpr int kOffRaw = kOffB + kF32RingBytes;
  static constexpr int kOffCst = kOffRaw + kRawDepth * kFlM * 32;
  static constexpr int kOffBar = kOffCst + 3 * KP * 4;
  static constexpr int kBytes = kOffBar + (3 * kF32MaxRing + 2 * kF32MaxNkb + 6) * 8 + 16;
  static_assert(kNkb <= kF32MaxNkb, "barrier slots");
  static_assert(kOffA % 16 == 0 && kOffB % 16 == 0 && kOffCst % 16 == 0 && kOffBar % 8 == 0, "alignment");
  static_assert(kOffRaw % 16 == 0, "alignment");
  static_assert(kBytes <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// packed weight blob: [units/256 chunks][KP/8 K steps][cl halves of the chunk's units][hi | lo][2 pieces][256/cl][4] fp32: one
// contiguous stage per chunk, K step and CTA of the pair (cl = 2: each CTA holds 128 of the 256 units of the MMA's B
// operand); column `width` holds the bias, the rest of the pad is 0
__global__ void linear_pack_f32_kernel(const float* __restrict__ w, const float* __restrict__ bias, int units, int width, int kp,
                                       int cl, float* __restrict__ packed) {
  const long long total = (long long)units * kp * 2;
  const int nh = kFlN / cl;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i & 3);
    long long t = i >> 2;
    int nl = (int)(t % nh); t /= nh;
    const int piece = (int)(t & 1); t >>= 1;
    const int part = (int)(t & 1); t >>= 1;
    const int half = (int)(t % cl); t /= cl;
    nl += half * nh;
    const int kb = (int)(t % (kp / 8));
    const int chunk = (int)(t / (kp / 8));
    const int n = chunk * kFlN + nl, k = kb * 8 + piece * 4 + j;
    float v = 0.0f;
    if (k < width) v = w[(size_t)n * width + k];
    else if (k == width && bias != nullptr) v = bias[n];
    const float hi = to_tf32(v);
    packed[i] = part == 0 ? hi : to_tf32(v - hi);
  }
}

#ifdef PPK_TRACE
// debug builds: g_trace[((block * 4 + role) * 32 + unit) * 4 + k] = globaltimer ns; roles: 0 MMA thread (k: 0 waits done,
// 1 MMAs issued), 1 epilogue warp 0 (0 accumulator full, 1 TMEM loaded, 2 first store issued, 3 second store issued),
// 2 weight producer (0 first stage of the unit issued), 3 row-tile preparation (0 start, 1 buffer released, 2 tile ready)
__device__ __forceinline__ void f32_stamp(int role, long long unit, int kk) {
  if (g_trace != nullptr && unit < 32) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace[(((size_t)blockIdx.x * 4 + role) * 32 + unit) * 4 + kk] = t;
  }
}
#define F32_STAMP(cond, role, unit, kk) do { if (cond) f32_stamp(role, unit, kk); } while (0)
#else
#define F32_STAMP(cond, role, unit, kk)
#endif

struct F32Args {
  const float* obs;          // [rows, width]
  long long rows;
  int width, units;
  RmsArgs rms;               // rms.mean == nullptr: no normalisation
  const unsigned char* packed;
  float* out;                // [rows, units]
};

namespace tc {
// Instruction descriptor, kind::tf32: D fp32 (bits [4,6) = 1), A and B TF32 (format 2 at [7,10) and [10,13)), both K-major,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t instr_desc_tf32_f32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// ---- CTA pair (cta_group::2): one MMA of M = 256 over the two CTAs of a cluster; the leader (rank 0) issues it ----
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// A rows [0,128) come from the leader's shared memory, [128,256) from the peer's, same offsets; each CTA holds N/2 of B;
// each CTA's tensor memory receives its 128 rows of D
__device__ __forceinline__ void mma_tf32_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// one arrival at the barrier at this shared-memory offset in BOTH CTAs when the pair's MMAs issued so far have completed
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
// wait on a local barrier whose arrivals come from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t done, spins = 0;
  do {
    if (++spins > (1u << 26)) __trap();
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
}  // namespace tc

// ELU in fp32: x > 0 ? x : expm1(x), relative error < 1e-6: exp2 - 1 where the subtraction cancels at most two bits,
// the Taylor polynomial (degree 6, next term x^7 / 5040 <= 5e-8 |x|) near zero.
__device__ __forceinline__ float elu_f32(float x) {
  const float t = ex2_ftz(x * 1.4426950408889634f) - 1.0f;
  float p = fmaf(x, 1.0f / 720.0f, 1.0f / 120.0f);
  p = fmaf(p, x, 1.0f / 24.0f);
  p = fmaf(p, x, 1.0f / 6.0f);
  p = fmaf(p, x, 0.5f);
  p = fmaf(p, x, 1.0f);
  p *= x;
  const float em1 = x < -0.25f ? t : p;
  return x > 0.0f ? x : em1;
}

// Ampere-style asynchronous copies (LDGSTS): per-thread groups, `wait_group N` = all but the N most recent groups landed
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  const uint32_t n = valid ? BYTES : 0;       // bytes read from the source; the rest of the destination is zero-filled
  if (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(n) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The row tiles of this CTA, K step by K step: rows [mt*128, +128) of obs -> clamp -> normalise -> (hi, lo) TF32 -> K-major
// core matrices.  A K step (8 columns = two 16-byte operand pieces per row) is 256 items, two per thread (rows t/2 and
// 64 + t/2, piece t & 1: a warp reads 32 contiguous bytes of 16 rows per copy and its stores are bank-conflict free).
// The tile has ONE buffer: K step kb of the previous tile is released (`slice_free[kb]`) as soon as the MMAs of that
// tile's last unit have read it, so the refill overlaps the rest of that unit's MMAs and the next tile's first unit
// starts on `slice_full[0]`.  A global load takes 2-3 us while the chip streams its output; register-staged loads end up
// waited for one round trip per K step (the compiler's scoreboard waits cover everything outstanding), so the raw values
// travel by cp.async into the thread's own slots of a small ring, kRawDepth K steps ahead across tile boundaries, and
// `cp.async.wait_group` is the only wait.
template <int KP, bool VEC>
__device__ __forceinline__ void f32_prep_tiles(const F32Args& k, const float* cst, unsigned char* raw, unsigned char* a_hi, int t,
                                               long long mt_first, long long mt_step, long long tiles, uint64_t* slice_free,
                                               uint64_t* slice_full) {
  using L = F32Layout<KP>;
  constexpr int NKB = L::kNkb, D = L::kRawDepth;
  const float inf = __int_as_float(0x7f800000);
  const float clip = k.rms.clip > 0.0f ? k.rms.clip : inf;
  const float lim = k.rms.mean != nullptr ? 5.0f : inf;
  const int piece = t & 1, r0 = t >> 1;
  float* my_raw = reinterpret_cast<float*>(raw) + t * 4;        // + (slot * 2 + h) * 128 * 4
  // request K step kb of tile j into ring slot `slot` (one group, possibly empty: the group count per iteration is fixed)
  auto request = [&](long long j, int kb, int slot) {
    if (j < tiles) {
      const int c = (2 * kb + piece) * 4;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long row = (mt_first + j * mt_step) * kFlM + r0 + 64 * h;
        const bool live = row < k.rows;
        const float* rowp = k.obs + (live ? row : 0) * k.width;
        float* dst = my_raw + (slot * 2 + h) * (kFlM * 4);
        if (VEC) {          // width % 4 == 0: a piece is inside the row or outside it
          const bool in = live && c < k.width;
          cp_async_zfill<16>(dst, rowp + (in ? c : 0), in);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool in = live && c + i < k.width;
            cp_async_zfill<4>(dst + i, rowp + (in ? c + i : 0), in);
          }
        }
      }
    }
    cp_async_commit();
  };
  // sequence position q = j * NKB + kb; slot = q % D
  {
    long long j = 0; int kb = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) { request(j, kb, d); if (++kb == NKB) { kb = 0; ++j; } }
  }
  long long jn = D / NKB; int kbn = D % NKB;     // the next position to request
  int slot = 0;
  for (long long j = 0; j < tiles; ++j) {
    F32_STAMP(t == 0, 3, j, 0);
    if (j + 1 < tiles) {              // pull the next tile into L2: most of its K steps are requested only as this one is released
      const long long row0 = (mt_first + (j + 1) * mt_step) * kFlM;
      if (row0 < k.rows) {
        const long long bytes = (min((long long)kFlM, k.rows - row0)) * k.width * 4;
        const char* p = reinterpret_cast<const char*>(k.obs + row0 * k.width);
        for (long long off = (long long)t * 128; off < bytes; off += 128LL * 32 * kF32PrepWarps) prefetch_l2(p + off);
      }
    }
#pragma unroll
    for (int kb = 0; kb < NKB; ++kb) {
      cp_async_wait<D - 1>();
      const float4 x0 = *reinterpret_cast<const float4*>(my_raw + (slot * 2 + 0) * (kFlM * 4));
      const float4 x1 = *reinterpret_cast<const float4*>(my_raw + (slot * 2 + 1) * (kFlM * 4));
      F32_STAMP(t == 0 && j == 1, 3, 16 + kb, 0);
      if (j > 0) {
        if (kb == 0) mbar_wait_relaxed(slice_free, (uint32_t)((j - 1) & 1));      // a long wait by design: most of a tile's units
        else mbar_wait(slice_free + kb, (uint32_t)((j - 1) & 1));
      }
      F32_STAMP(t == 0 && j == 1, 3, 16 + kb, 1);
      const int pc = 2 * kb + piece;
      const float4 m = *reinterpret_cast<const float4*>(cst + pc * 4);
      const float4 d = *reinterpret_cast<const float4*>(cst + KP + pc * 4);
      const float4 rc = *reinterpret_cast<const float4*>(cst + 2 * KP + pc * 4);
      const float mm[4] = {m.x, m.y, m.z, m.w}, dd[4] = {d.x, d.y, d.z, d.w}, rr[4] = {rc.x, rc.y, rc.z, rc.w};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 xv = h == 0 ? x0 : x1;
        const float v[4] = {xv.x, xv.y, xv.z, xv.w};
        float hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float y = v[i];
          if (k.rms.clip > 0.0f) y = fminf(fmaxf(y, -clip), clip);
          y = fminf(fmaxf(div_by_const(y - mm[i], dd[i], rr[i]), -lim), lim);
          hi[i] = to_tf32(y);
          lo[i] = to_tf32(y - hi[i]);
        }
        unsigned char* dst = a_hi + pc * kFlALbo + (r0 + 64 * h) * 16;
        *reinterpret_cast<float4*>(dst) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(dst + L::kAPart) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
      F32_STAMP(t == 0 && j == 1, 3, 16 + kb, 2);
      tc::fence_proxy_async_smem();
      __syncwarp();
      if ((t & 31) == 0) tc::mbar_arrive(slice_full + kb);
      request(jn, kbn, slot);               // the slot this thread has just read
      if (++kbn == NKB) { kbn = 0; ++jn; }
      if (++slot == D) slot = 0;
      F32_STAMP(t == 0 && j == 1, 3, 16 + kb, 3);
    }
    F32_STAMP(t == 0, 3, j, 2);
  }
  cp_async_wait<0>();
}

template <int KP, int ACT, int CL>
__global__ void __launch_bounds__(kF32Threads, 1)
first_layer_f32_kernel(const __grid_constant__ F32Args k, const __grid_constant__ CUtensorMap out_map) {
  using L = F32Layout<KP>;
  constexpr int NKB = L::kNkb;
  constexpr int kStage = kF32StepBytes / CL;                // this CTA's part of a K step: hi | lo, 2 pieces, 256/CL units
  constexpr int R = kF32RingBytes / kStage;
  constexpr int kBLbo = kFlN / CL * 16;
  static_assert(R <= kF32MaxRing, "barrier slots");
  extern __shared__ __align__(1024) unsigned char fl_smem[];
  unsigned char* a_s = fl_smem + L::kOffA;
  unsigned char* b_s = fl_smem + L::kOffB;
  unsigned char* stage_s = fl_smem + L::kOffStage;
  float* cst = reinterpret_cast<float*>(fl_smem + L::kOffCst);
  uint64_t* bars = reinterpret_cast<uint64_t*>(fl_smem + L::kOffBar);
  constexpr int X = kF32MaxRing, Y = kF32MaxNkb;
  uint64_t *b_full = bars, *b_empty = bars + X, *peer_b_full = bars + 2 * X, *a_full = bars + 3 * X, *a_free = bars + 3 * X + Y,
           *acc_full = bars + 3 * X + 2 * Y, *acc_empty = acc_full + 2, *peer_acc_empty = acc_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NC = k.units / kFlN;
  const long long MT = (k.rows + kFlM - 1) / kFlM;
  // a cluster works on row-tile GROUPS (CL tiles, one per CTA): its CTAs walk the same units
  const uint32_t rank = CL > 1 ? tc::cluster_ctarank() : 0u;
  const long long cluster = blockIdx.x / CL, clusters = gridDim.x / CL;
  const long long G = (MT + CL - 1) / CL;
  const long long U = G * NC;
  const long long u_begin = U * cluster / clusters, u_end = U * (cluster + 1) / clusters;
  const long long g_begin = u_begin / NC;

  if (threadIdx.x == 0) {
    for (int i = 0; i < R; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); mbar_init(peer_b_full + i, 1); }
    for (int i = 0; i < NKB; ++i) { mbar_init(a_full + i, kF32PrepWarps); mbar_init(a_free + i, 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 32 * kF32EpiWarps); mbar_init(peer_acc_empty + i, 1);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    if (CL > 1) tc::tmem_alloc_pair(tmem_slot, 512);
    else tc::tmem_alloc(tmem_slot, 512);
  }
  for (int c = threadIdx.x; c < KP; c += kF32Threads) {
    const bool on = k.rms.mean != nullptr && c < k.width;
    cst[c] = on ? (float)k.rms.mean[c] : (c == k.width ? -1.0f : 0.0f);      // column `width`: (0 - -1) / 1 = the bias column's 1
    const float den = on ? sqrtf((float)k.rms.var[c] + k.rms.eps) : 1.0f;
    cst[KP + c] = den;
    cst[2 * KP + c] = 1.0f / den;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (CL > 1) tc::cluster_sync();       // the peer's barriers exist before anything signals them
  const uint32_t tmem_base = *tmem_slot;

  // weight stage `bi` of the cluster's sequence (unit-major, K steps inside a unit): this CTA's part
  auto load_weights = [&](long long bi) {
    const long long u = u_begin + bi / NKB;
    const int kb = (int)(bi % NKB), nc = (int)(u % NC), s = (int)(bi % R);
    if (PPK_F32_DBG & 16) { tc::mbar_arrive(b_full + s); return; }
    mbar_arrive_expect_tx(b_full + s, (uint32_t)kStage);
    bulk_g2s(b_s + s * kStage, k.packed + (((size_t)nc * NKB + kb) * CL + rank) * kStage, (uint32_t)kStage, b_full + s);
  };
  const long long n_stages = (u_end - u_begin) * NKB;
  if (threadIdx.x == 0)
    for (long long bi = 0; bi < R && bi < n_stages; ++bi) load_weights(bi);

  if (warp == 0) {
    // ===== weight producer =====
    if (lane == 0) {
      for (long long bi = R; bi < n_stages; ++bi) {
        const int s = (int)(bi % R), ph = (int)((bi / R) & 1);
        mbar_wait_relaxed(b_empty + s, ph ^ 1);         // the MMAs that read the stage have completed
        load_weights(bi);
        F32_STAMP(bi % NKB == 0, 2, bi / NKB, 0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread of the leader CTA) / event forwarder (one thread of the peer CTA) =====
    // Both walk the same sequence of waits.  The leader waits for its own barrier and for the peer's copy of it, then
    // issues; the peer waits for its own barrier and arrives on the leader's peer_* barrier.
    if (lane == 0) {
      constexpr uint32_t idesc = tc::instr_desc_tf32_f32(kFlM * CL, kFlN);
      const uint32_t a_hi = smem_u32(a_s), a_lo = a_hi + L::kAPart;
      const bool leader = rank == 0;
      for (long long u = u_begin; u < u_end; ++u) {
        const long long it = u - u_begin;
        const int s = (int)(it & 1), ph = (int)((it >> 1) & 1);
        const long long g = u / NC;
        const int nc = (int)(u - g * NC);
        const long long j = g - g_begin;
        const bool first_of_tile = it == 0 || nc == 0;            // the row tile arrives K step by K step
        const bool last_of_tile = u == u_end - 1 || nc == NC - 1;   // ... and is released K step by K step
        if (it >= 2) {                                  // the accumulator buffer has been drained by the epilogue warps
          mbar_wait(acc_empty + s, ph ^ 1);
          if (CL > 1) { if (leader) tc::mbar_wait_cluster(peer_acc_empty + s, ph ^ 1); else tc::mbar_arrive_remote(peer_acc_empty + s, 0); }
        }
        F32_STAMP(true, 0, it, 0);
        for (int kb = 0; kb < NKB; ++kb) {
          const long long bi = it * NKB + kb;
          const int sb = (int)(bi % R), phb = (int)((bi / R) & 1);
          if (PPK_F32_DBG & 256) { if (!leader) continue; }     // raw MMA issue rate: no operand waits at all
          else {
            if (first_of_tile) mbar_wait(a_full + kb, (uint32_t)(j & 1));
            mbar_wait(b_full + sb, phb);
          }
          if (CL > 1 && !(PPK_F32_DBG & 256)) {     // the peer's arrival says: its weight stage AND its K step of the row tile are in place
            if (!leader) { tc::mbar_arrive_remote(peer_b_full + sb, 0); continue; }
            tc::mbar_wait_cluster(peer_b_full + sb, phb);
          }
          tc::fence_after_sync();
          const uint32_t b_hi = smem_u32(b_s + sb * kStage), b_lo = b_hi + kStage / 2;
          const uint64_t da_hi = tc::smem_desc(a_hi + kb * 2 * kFlALbo, kFlALbo, 128);
          const uint64_t da_lo = tc::smem_desc(a_lo + kb * 2 * kFlALbo, kFlALbo, 128);
          const uint64_t db_hi = tc::smem_desc(b_hi, kBLbo, 128), db_lo = tc::smem_desc(b_lo, kBLbo, 128);
          const uint32_t d = tmem_base + s * kFlN;
          if (!(PPK_F32_DBG & 4)) {
            if (CL > 1) {
              tc::mma_tf32_pair(d, da_hi, db_hi, idesc, kb > 0);
              tc::mma_tf32_pair(d, da_lo, db_hi, idesc, true);
              tc::mma_tf32_pair(d, da_hi, db_lo, idesc, true);
            } else {
              tc::mma_tf32(d, da_hi, db_hi, idesc, kb > 0);
              tc::mma_tf32(d, da_lo, db_hi, idesc, true);
              tc::mma_tf32(d, da_hi, db_lo, idesc, true);
            }
          }
          if (CL > 1) tc::mma_commit_pair(b_empty + sb);
          else tc::mma_commit(b_empty + sb);
          if (last_of_tile) { if (CL > 1) tc::mma_commit_pair(a_free + kb); else tc::mma_commit(a_free + kb); }
        }
        F32_STAMP(true, 0, it, 1);
        if (!leader) continue;
        if (CL > 1) tc::mma_commit_pair(acc_full + s);
        else tc::mma_commit(acc_full + s);
      }
    }
  } else if (warp < kF32EpiFirst) {
    // ===== row-tile preparation =====
    const int t = threadIdx.x - 32 * kF32PrepFirst;
    const long long g_last = (u_end - 1) / NC;
    const bool vec = (k.width % 4 == 0) && ((reinterpret_cast<uintptr_t>(k.obs) & 15u) == 0);
    unsigned char* raw = fl_smem + L::kOffRaw;
    if (vec) f32_prep_tiles<KP, true>(k, cst, raw, a_s, t, g_begin * CL + rank, CL, g_last - g_begin + 1, a_free, a_full);
    else f32_prep_tiles<KP, false>(k, cst, raw, a_s, t, g_begin * CL + rank, CL, g_last - g_begin + 1, a_free, a_full);
  } else {
    // ===== epilogue: TMEM -> registers -> activation -> swizzled [32 x 32] fp32 tile -> tensor store =====
    const int ew = warp - kF32EpiFirst;
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int cq = ew >> 2;                 // which kF32EpiCols-wide slice of the chunk
    unsigned char* wstage = stage_s + (size_t)ew * kF32EpiTile;
    unsigned char* my_row = wstage + lane * 128;
    const int sw = lane & 7;
    constexpr int kGroups = kF32EpiCols / 32;
    static_assert(kGroups == 2, "the TMEM loads alternate between two register buffers");
    for (long long u = u_begin; u < u_end; ++u) {
      const long long it = u - u_begin;
      const int s = (int)(it & 1), ph = (int)((it >> 1) & 1);
      const long long g = u / NC;
      const int nc = (int)(u - g * NC);
      const long long mt = g * CL + rank;
      const int col0 = nc * kFlN + cq * kF32EpiCols;
      mbar_wait(acc_full + s, ph);
      tc::fence_after_sync();
      F32_STAMP(ew == 0 && lane == 0, 1, it, 0);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * kFlN + cq * kF32EpiCols);
      if (PPK_F32_DBG & 64) { tc::fence_before_sync(); tc::mbar_arrive(acc_empty + s); continue; }
      // one group of 32 columns: activation in place, then the tile and its tensor store
      auto emit = [&](uint32_t (&v)[32], int gi) {
        float (&o)[32] = reinterpret_cast<float (&)[32]>(v);     // bias already inside (K padding column)
        if (ACT == 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = elu_f32(o[i]);
        }
        if (PPK_F32_DBG & 512) {          // A/B: straight from the registers, 16-byte pieces of the lane's own output row
          const long long row = mt * kFlM + q * 32 + lane;
          if (row < k.rows) {
            float4* dst = reinterpret_cast<float4*>(k.out + row * k.units + col0 + gi * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) __stcs(dst + i, make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]));
          }
          return;
        }
        // the tensor store of the previous group has finished reading the tile
        if (lane == 0) tc::bulk_wait_read0();
        __syncwarp();
        if (!(PPK_F32_DBG & 32)) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(my_row + ((i ^ sw) << 4)) = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        } else {
          float acc = 0.0f;
#pragma unroll
          for (int i = 0; i < 32; ++i) acc += o[i];
          if (acc == 1.2345e-30f) *reinterpret_cast<float*>(my_row) = acc;
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && mt < MT && !(PPK_F32_DBG & 2)) {
          tc::tensor_store_2d(&out_map, col0 + gi * 32, (int)(mt * kFlM + q * 32), wstage);
          tc::bulk_commit();
        }
      };
      // both TMEM loads first: the accumulator buffer is released before any of the arithmetic
      uint32_t va[32], vb[32];
      tc::tmem_ld32(taddr, va);
      tc::tmem_ld32(taddr + 32, vb);
      tc::tmem_ld_wait();
      tc::fence_before_sync();
      tc::mbar_arrive(acc_empty + s);
      F32_STAMP(ew == 0 && lane == 0, 1, it, 1);
      emit(va, 0);
      F32_STAMP(ew == 0 && lane == 0, 1, it, 2);
      emit(vb, 1);
      F32_STAMP(ew == 0 && lane == 0, 1, it, 3);
    }
    if (lane == 0) tc::bulk_wait_all();
  }

  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (CL > 1) tc::cluster_sync();       // no CTA leaves (or frees tensor memory) while the pair's MMAs or signals may still touch it
  if (warp == 2) {
    if (CL > 1) tc::tmem_dealloc_pair(tmem_base, 512);
    else tc::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ppk
