// Probe for the staging design of the family step kernel (not part of the product path).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_probe tools/tma_probe.cu
//   ./tma_probe [envs=65536] [sets=4] [iters=200] [ncu=0]
//
// Questions it answers on a B200:
//   1. Does a 2-D tensor-map load (cp.async.bulk.tensor) deliver the rigid-body rows of 16 env PAIRS
//      (pair = 2 x 2184 B = 4368 B, a multiple of 16 as TMA strides must be) when the inner coordinate
//      starts at an arbitrary float (row 31 = float 403, 1612 B: not 16-byte aligned)?
//   2. How fast can the TILT step's inputs be brought on chip and its outputs written back, with no
//      compute at all: per-env 1-D bulk windows (round-1 staging) vs tensor loads, for different L2
//      promotion settings of the tensor map and cudaLimitMaxL2FetchGranularity values.
// With ncu=1 every configuration is launched twice only (for a metrics pass: dram bytes per launch).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define CK(x)                                                                                 \
  do {                                                                                        \
    cudaError_t e_ = (x);                                                                     \
    if (e_ != cudaSuccess) {                                                                  \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                                \
    }                                                                                         \
  } while (0)

constexpr int kRow = 13, kB = 42, kEnvF = kB * kRow;   // 546 floats = 2184 B per env
constexpr int kTile = 32, kPairs = kTile / 2;   // verification kernel; the floors are templated on the tile
constexpr int kSpanF = 120;   // rows 31..39 = 117 floats, box padded to 16 B
constexpr int kRow0F = 12;    // row 0: 10 floats used, box 48 B
constexpr int kObs = 80;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done, spins = 0;
  do {
    if (++spins > (1u << 22)) __trap();
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

struct Coords {   // inner coordinates (floats) of the four boxes and where the wanted data starts inside each box
  int span[2], row0[2];
  int stagger_cycles;   // soft start: CTA b of the first wave delays its loads by (b / num_sms) * stagger_cycles
  int first_wave, num_sms;
};

struct Set {
  float *rb, *root, *dof, *force, *pre, *obs, *rew;
  long long *progress, *reset;
  unsigned char* flags[3];
};

template <int TILE>
struct SmemLayout {   // floats
  static constexpr int pairs = TILE / 2;
  static constexpr int span = 0;                                  // [2][pairs][120]
  static constexpr int row0 = span + 2 * pairs * kSpanF;          // [2][pairs][12]
  static constexpr int root = row0 + 2 * pairs * kRow0F;          // [TILE][39]
  static constexpr int dof = root + TILE * 39;                    // [TILE][14]
  static constexpr int force = dof + TILE * 14;                   // [TILE][7]
  static constexpr int pre = force + TILE * 7;                    // [TILE][2]
  static constexpr int prog = pre + TILE * 2;                     // [TILE] i64
  static constexpr int flags = prog + TILE * 2;                   // 3 x TILE B
  static constexpr int obs = flags + 3 * TILE / 4;                // [TILE][80]
  static constexpr int bar = obs + TILE * kObs;
  static constexpr int total = bar + 4;
  static constexpr uint32_t tx = 4u * (2 * pairs * kSpanF + 2 * pairs * kRow0F + TILE * (39 + 14 + 7 + 2 + 2) + 3 * TILE / 4);
  static_assert(row0 % 32 == 0 && (row0 + pairs * kRow0F) % 32 == 0 && (pairs * kSpanF) % 32 == 0, "128-byte tensor destinations");
  static_assert(root % 4 == 0 && dof % 4 == 0 && force % 4 == 0 && pre % 4 == 0 && prog % 4 == 0 && flags % 4 == 0 && obs % 4 == 0, "align");
};

// The TILT step's traffic with no compute: tensor loads + bulk loads of one tile, then a bulk store of the obs rows and
// plain stores of the per-env scalars.  PERSIST: grid = resident CTAs, each walks tiles blockIdx.x, +gridDim.x, ...
__device__ unsigned long long* g_trace = nullptr;     // [tile][4]: start, data arrived, end, smid
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int TILE, bool PERSIST, bool DO_LOAD, bool DO_STORE>
__global__ void __launch_bounds__(64) probe_kernel(Set s, const __grid_constant__ CUtensorMap m_span, const __grid_constant__ CUtensorMap m_row0,
                                                   long long n, float* sink, Coords co) {
  extern __shared__ __align__(128) float smem[];
  using L = SmemLayout<TILE>;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::bar);
  const int tid = threadIdx.x;
  const long long ntiles = n / TILE;
  if (DO_LOAD) {
    if (tid == 0) {
      mbar_init(bar, 1);
      fence_init();
    }
    __syncthreads();
  }
  if (co.stagger_cycles > 0 && (int)blockIdx.x < co.first_wave && blockIdx.x >= (unsigned)co.num_sms) {
    const long long wait = (long long)(blockIdx.x / co.num_sms) * co.stagger_cycles;
    const long long t = clock64();
    while (clock64() - t < wait) {}
  }
  uint32_t phase = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += PERSIST ? gridDim.x : ntiles) {
    const long long env0 = tile * TILE;
    float acc = 0.0f;
    unsigned long long* tr = (g_trace != nullptr && tid == 0) ? g_trace + tile * 4 : nullptr;
    if (tr) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      tr[0] = gtime();
      tr[3] = smid;
    }
    if (DO_LOAD) {
      if (tid == 0) {
        expect_tx(bar, L::tx);
        const int p0 = (int)(env0 >> 1);
        tma_2d(smem + L::span, &m_span, co.span[0], p0, bar);
        tma_2d(smem + L::span + L::pairs * kSpanF, &m_span, co.span[1], p0, bar);
        tma_2d(smem + L::row0, &m_row0, co.row0[0], p0, bar);
        tma_2d(smem + L::row0 + L::pairs * kRow0F, &m_row0, co.row0[1], p0, bar);
      }
      if (tid == 32) {
        bulk_g2s(smem + L::root, s.root + (size_t)env0 * 39, 4u * TILE * 39, bar);
        bulk_g2s(smem + L::dof, s.dof + (size_t)env0 * 14, 4u * TILE * 14, bar);
        bulk_g2s(smem + L::force, s.force + (size_t)env0 * 7, 4u * TILE * 7, bar);
        bulk_g2s(smem + L::pre, s.pre + (size_t)env0 * 2, 4u * TILE * 2, bar);
        bulk_g2s(smem + L::prog, s.progress + env0, 8u * TILE, bar);
        for (int i = 0; i < 3; ++i) bulk_g2s(smem + L::flags + (TILE / 4) * i, s.flags[i] + env0, TILE, bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      if (tr) tr[1] = gtime();
      for (int i = tid; i < L::obs; i += 64) acc += smem[i];      // touch the staged data
    }
    if (DO_STORE) {
      for (int i = tid; i < TILE * kObs; i += 64) smem[L::obs + i] = acc + (float)i;
      fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        bulk_s2g(s.obs + (size_t)env0 * kObs, smem + L::obs, 4u * TILE * kObs);
        bulk_commit();
      }
      if (tid < TILE) {
        const long long e = env0 + tid;
        s.rew[e] = acc;
        s.reset[e] = (long long)(acc > 1e30f);
        s.progress[e] = (long long)tid;
        for (int i = 0; i < 3; ++i) s.flags[i][e] = (unsigned char)(tid & 1);
      }
      if (tid == 0) bulk_wait_read();
      if (tr) tr[2] = gtime();
      __syncthreads();        // the staging area is free again (persistent loop)
    } else {
      if (acc == 123.456f) sink[blockIdx.x] = acc;
    }
  }
}

__global__ void __launch_bounds__(64) empty_kernel(float* sink) {
  if (sink == nullptr && threadIdx.x == 1234) sink[0] = 0.0f;
}

// one tile's staged span / row-0 boxes copied out for verification
__global__ void verify_kernel(const __grid_constant__ CUtensorMap m_span, const __grid_constant__ CUtensorMap m_row0, int p0, float* out, Coords co) {
  __shared__ __align__(128) float sm[2 * kPairs * kSpanF + 2 * kPairs * kRow0F];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    expect_tx(&bar, 4u * (2 * kPairs * kSpanF + 2 * kPairs * kRow0F));
    tma_2d(sm, &m_span, co.span[0], p0, &bar);
    tma_2d(sm + kPairs * kSpanF, &m_span, co.span[1], p0, &bar);
    tma_2d(sm + 2 * kPairs * kSpanF, &m_row0, co.row0[0], p0, &bar);
    tma_2d(sm + 2 * kPairs * kSpanF + kPairs * kRow0F, &m_row0, co.row0[1], p0, &bar);
  }
  // bounded wait without a trap: report a copy that never completes instead of killing the context
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 16) && !done; ++spin)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  const int nout = 2 * kPairs * kSpanF + 2 * kPairs * kRow0F;
  if (threadIdx.x == 0) out[nout] = done ? 1.0f : 0.0f;
  if (!done) return;
  for (int i = threadIdx.x; i < nout; i += blockDim.x) out[i] = sm[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess) { fprintf(stderr, "no cuTensorMapEncodeTiled\n"); exit(1); }
  return (EncodeTiledFn)p;
}

static int make_map(CUtensorMap* m, const float* rb, long long n, int box0, CUtensorMapL2promotion promo) {
  static EncodeTiledFn enc = get_encode();
  const cuuint64_t dims[2] = {(cuuint64_t)2 * kEnvF, (cuuint64_t)(n / 2)};
  const cuuint64_t strides[1] = {(cuuint64_t)2 * kEnvF * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)kPairs};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)rb, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
  return (int)r;
}

int main(int argc, char** argv) {
  const long long n = argc > 1 ? atoll(argv[1]) : 65536;
  const int sets = argc > 2 ? atoi(argv[2]) : 4;
  const int iters = argc > 3 ? atoi(argv[3]) : 200;
  const bool ncu = argc > 4 && atoi(argv[4]) != 0;
  if (n % kTile != 0) { fprintf(stderr, "n must be a multiple of %d\n", kTile); return 1; }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, L2 %d MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20);
  size_t gran = 0;
  CK(cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity));
  printf("default cudaLimitMaxL2FetchGranularity = %zu\n", gran);

  std::vector<Set> S(sets);
  std::vector<float> h_rb((size_t)n * kEnvF);
  for (size_t i = 0; i < h_rb.size(); ++i) h_rb[i] = (float)(i % 1000003) * 0.5f;
  for (int s = 0; s < sets; ++s) {
    Set& t = S[s];
    CK(cudaMalloc(&t.rb, (size_t)n * kEnvF * 4));
    CK(cudaMemcpy(t.rb, h_rb.data(), (size_t)n * kEnvF * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&t.root, (size_t)n * 39 * 4)); CK(cudaMemset(t.root, 0, (size_t)n * 39 * 4));
    CK(cudaMalloc(&t.dof, (size_t)n * 14 * 4)); CK(cudaMemset(t.dof, 0, (size_t)n * 14 * 4));
    CK(cudaMalloc(&t.force, (size_t)n * 7 * 4)); CK(cudaMemset(t.force, 0, (size_t)n * 7 * 4));
    CK(cudaMalloc(&t.pre, (size_t)n * 2 * 4)); CK(cudaMemset(t.pre, 0, (size_t)n * 2 * 4));
    CK(cudaMalloc(&t.obs, (size_t)n * kObs * 4));
    CK(cudaMalloc(&t.rew, (size_t)n * 4));
    CK(cudaMalloc(&t.progress, (size_t)n * 8)); CK(cudaMemset(t.progress, 0, (size_t)n * 8));
    CK(cudaMalloc(&t.reset, (size_t)n * 8));
    for (int i = 0; i < 3; ++i) { CK(cudaMalloc(&t.flags[i], (size_t)n)); CK(cudaMemset(t.flags[i], 0, (size_t)n)); }
  }
  float* sink;
  CK(cudaMalloc(&sink, (size_t)(n / kTile) * 4));

  // ---- 1. correctness of the tensor loads ----
  // aligned: each box starts at the 16-byte boundary at or below the first wanted float (per parity of the env
  // inside its pair); unaligned (argv[5]=1, run as its own process): the box starts at the wanted float itself.
  const bool unaligned = argc > 5 && atoi(argv[5]) != 0;
  Coords co, in_box;
  co.stagger_cycles = 0; co.first_wave = 0; co.num_sms = prop.multiProcessorCount;
  for (int q = 0; q < 2; ++q) {
    const int s0 = q * kEnvF + 31 * kRow, r0 = q * kEnvF;
    co.span[q] = unaligned ? s0 : (s0 & ~3);
    co.row0[q] = unaligned ? r0 : (r0 & ~3);
    in_box.span[q] = s0 - co.span[q];
    in_box.row0[q] = r0 - co.row0[q];
  }
  printf("box coordinates (floats): span %d/%d (+%d/+%d), row0 %d/%d (+%d/+%d)\n", co.span[0], co.span[1], in_box.span[0],
         in_box.span[1], co.row0[0], co.row0[1], in_box.row0[0], in_box.row0[1]);
  {
    CUtensorMap ms, m0;
    if (make_map(&ms, S[0].rb, n, kSpanF, CU_TENSOR_MAP_L2_PROMOTION_NONE) || make_map(&m0, S[0].rb, n, kRow0F, CU_TENSOR_MAP_L2_PROMOTION_NONE)) return 1;
    const int nout = 2 * kPairs * kSpanF + 2 * kPairs * kRow0F;
    float* d_out;
    CK(cudaMalloc(&d_out, (nout + 1) * 4));
    std::vector<float> h_out(nout + 1);
    long long bad = 0;
    const long long tiles[3] = {0, 7, n / kTile - 1};
    for (long long tile : tiles) {
      verify_kernel<<<1, 128>>>(ms, m0, (int)(tile * kPairs), d_out, co);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h_out.data(), d_out, (nout + 1) * 4, cudaMemcpyDeviceToHost));
      if (h_out[nout] != 1.0f) { printf("tile %lld: the copies never completed\n", tile); bad += 1000000; continue; }
      for (int par = 0; par < 2; ++par)
        for (int p = 0; p < kPairs; ++p) {
          const long long env = tile * kTile + 2 * p + par;
          for (int f = 0; f < 117; ++f)
            if (h_out[(par * kPairs + p) * kSpanF + in_box.span[par] + f] != h_rb[(size_t)env * kEnvF + 31 * kRow + f]) ++bad;
          for (int f = 0; f < 10; ++f)
            if (h_out[2 * kPairs * kSpanF + (par * kPairs + p) * kRow0F + in_box.row0[par] + f] != h_rb[(size_t)env * kEnvF + f]) ++bad;
        }
    }
    printf("tensor-load verification (%s coordinates): %lld mismatches over 3 tiles\n", unaligned ? "UNALIGNED" : "aligned", bad);
    CK(cudaFree(d_out));
    if (unaligned) return 0;
    if (bad) return 2;
  }

  // ---- 2. floors ----
  typedef void (*KernFn)(Set, const CUtensorMap, const CUtensorMap, long long, float*, Coords);
  struct Variant { const char* name; KernFn fn; int tile; bool persist; bool ld, st; size_t smem; };
  std::vector<Variant> vars = {
      {"load  tile32      ", probe_kernel<32, false, true, false>, 32, false, true, false, SmemLayout<32>::total * 4},
      {"store tile32      ", probe_kernel<32, false, false, true>, 32, false, false, true, SmemLayout<32>::total * 4},
      {"copy  tile32      ", probe_kernel<32, false, true, true>, 32, false, true, true, SmemLayout<32>::total * 4},
      {"copy  tile16      ", probe_kernel<16, false, true, true>, 16, false, true, true, SmemLayout<16>::total * 4},
      {"copy  tile32 pers ", probe_kernel<32, true, true, true>, 32, true, true, true, SmemLayout<32>::total * 4},
      {"copy  tile16 pers ", probe_kernel<16, true, true, true>, 16, true, true, true, SmemLayout<16>::total * 4},
  };
  for (auto& v : vars) CK(cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double algo_in = 383.0, algo_out = 343.0;
  const CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
  std::vector<CUtensorMap> ms32(sets), m032(sets), ms16(sets), m016(sets);
  auto make_map_t = [&](CUtensorMap* m, const float* rb, int box0, int pairs) {
    static EncodeTiledFn enc = get_encode();
    const cuuint64_t dims[2] = {(cuuint64_t)2 * kEnvF, (cuuint64_t)(n / 2)};
    const cuuint64_t strides[1] = {(cuuint64_t)2 * kEnvF * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)pairs};
    const cuuint32_t estr[2] = {1, 1};
    return (int)enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)rb, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  for (int s = 0; s < sets; ++s) {
    if (make_map_t(&ms32[s], S[s].rb, kSpanF, 16) || make_map_t(&m032[s], S[s].rb, kRow0F, 16) ||
        make_map_t(&ms16[s], S[s].rb, kSpanF, 8) || make_map_t(&m016[s], S[s].rb, kRow0F, 8)) return 1;
  }
  auto time_graph = [&](auto launch, int reps) -> double {
    for (int i = 0; i < 3; ++i) launch(i);
    CK(cudaStreamSynchronize(st));
    cudaGraph_t g;
    cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < reps; ++i) launch(i);
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, st));      // warm replay
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms_total = 0;
    CK(cudaEventElapsedTime(&ms_total, e0, e1));
    CK(cudaGraphExecDestroy(ge));
    CK(cudaGraphDestroy(g));
    return 1e3 * ms_total / reps;
  };
  const int reps = ncu ? 2 : iters;
  {
    const double us0 = time_graph([&](int) { empty_kernel<<<1, 64, 0, st>>>(sink); }, reps);
    const double us1 = time_graph([&](int) { empty_kernel<<<(unsigned)(n / 32), 64, 34 * 1024, st>>>(sink); }, reps);
    printf("empty kernel node: grid 1: %.2f us, grid %lld x 34 KB smem: %.2f us\n", us0, n / 32, us1);
  }
  const int pads_kb[3] = {0, 14, 40};    // extra dynamic smem: fewer resident CTAs per SM
  for (auto& v : vars) {
    for (int pad = 0; pad < 3; ++pad) {
      if (ncu && pad != 0) continue;
      const size_t sm = v.smem + (size_t)pads_kb[pad] * 1024;
      int occ = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v.fn, 64, sm));
      const long long ntiles = n / v.tile;
      const unsigned grid = v.persist ? (unsigned)std::min<long long>(ntiles, (long long)occ * prop.multiProcessorCount) : (unsigned)ntiles;
      auto& ms = v.tile == 32 ? ms32 : ms16;
      auto& m0 = v.tile == 32 ? m032 : m016;
      const int staggers[5] = {0, 200, 400, 600, 900};
      for (int si = 0; si < ((v.ld && v.st && !ncu) ? 5 : 1); ++si) {
        Coords c2 = co;
        c2.stagger_cycles = staggers[si];
        c2.first_wave = occ * prop.multiProcessorCount;
        const double us = time_graph([&](int i) { v.fn<<<grid, 64, sm, st>>>(S[i % sets], ms[i % sets], m0[i % sets], n, sink, c2); }, reps);
        const double bytes = (v.ld ? algo_in : 0.0) + (v.st ? algo_out : 0.0);
        printf("%s smem=%3zuKB occ=%2d grid=%6u stagger=%3d cyc  %8.2f us/launch  algo %.0f B/env -> %.0f GB/s algorithmic\n", v.name, sm >> 10,
               occ, grid, staggers[si], us, bytes, bytes * n / us * 1e-3);
        fflush(stdout);
      }
    }
  }
  // ---- 3. per-tile timeline of one launch of variant argv[6] (globaltimer stamps) ----
  if (argc > 6) {
    const int vi = atoi(argv[6]);
    auto& v = vars[vi];
    const long long ntiles = n / v.tile;
    unsigned long long* d_tr;
    CK(cudaMalloc(&d_tr, ntiles * 4 * 8));
    CK(cudaMemset(d_tr, 0, ntiles * 4 * 8));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v.fn, 64, v.smem));
    const unsigned grid = v.persist ? (unsigned)std::min<long long>(ntiles, (long long)occ * prop.multiProcessorCount) : (unsigned)ntiles;
    auto& ms = v.tile == 32 ? ms32 : ms16;
    auto& m0 = v.tile == 32 ? m032 : m016;
    for (int i = 0; i < 4; ++i) v.fn<<<grid, 64, v.smem, st>>>(S[i % sets], ms[i % sets], m0[i % sets], n, sink, co);
    CK(cudaStreamSynchronize(st));
    co.stagger_cycles = argc > 7 ? atoi(argv[7]) : 0;
    co.first_wave = occ * prop.multiProcessorCount;
    CK(cudaMemcpyToSymbol(g_trace, &d_tr, sizeof(d_tr)));
    unsigned long long* h_t0;   // a host-visible "launch issued" stamp is not available: use the earliest CTA start as t = 0
    (void)h_t0;
    v.fn<<<grid, 64, v.smem, st>>>(S[0], ms[0], m0[0], n, sink, co);
    CK(cudaStreamSynchronize(st));
    std::vector<unsigned long long> h(ntiles * 4);
    CK(cudaMemcpy(h.data(), d_tr, ntiles * 4 * 8, cudaMemcpyDeviceToHost));
    char path[256];
    snprintf(path, sizeof(path), "gpurun_out/probe_trace_%d_%lld_s%d.bin", vi, n, co.stagger_cycles);
    FILE* f = fopen(path, "wb");
    if (f) { fwrite(h.data(), 8, h.size(), f); fclose(f); printf("trace of '%s' written to %s\n", v.name, path); }
  }
  printf("done\n");
  return 0;
}
