// Does div_by_const (ppk_policy.cuh) equal IEEE n / d?  Random and structured operands in the ranges the
// normalisation sees: |n| <= 1e4, d in [1e-3, 1e3].
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float div_by_const(float n, float d, float r) {
  float q = n * r;
  q = fmaf(fmaf(-d, q, n), r, q);
  q = fmaf(fmaf(-d, q, n), r, q);
  return q;
}
__device__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__global__ void k(unsigned long long* bad, unsigned long long* bad1, unsigned long long total) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned long long nb = 0, nb1 = 0;
  for (; i < total; i += stride) {
    uint32_t a = hash((uint32_t)i), b = hash((uint32_t)(i >> 32) ^ a ^ 0x9e3779b9U), c = hash(a ^ 0x85ebca6bU);
    // d: random mantissa, exponent in [-10, 10]; n: random mantissa, exponent in [-20, 13], random sign
    float d = __uint_as_float(((117u + (b % 21u)) << 23) | (a & 0x7fffffu));
    float n = __uint_as_float(((c & 1u) << 31) | ((107u + ((c >> 1) % 34u)) << 23) | (b & 0x7fffffu));
    float r = 1.0f / d;
    float q = div_by_const(n, d, r), w = n / d;
    if (q != w) { ++nb; float q1 = n * r; q1 = fmaf(fmaf(-d, q1, n), r, q1); if (q1 != w) ++nb1; }
  }
  atomicAdd(bad, nb); atomicAdd(bad1, nb1);
}
int main() {
  unsigned long long *bad, *bad1, h[2] = {0, 0};
  cudaMalloc(&bad, 8); cudaMalloc(&bad1, 8); cudaMemset(bad, 0, 8); cudaMemset(bad1, 0, 8);
  unsigned long long total = 1ull << 33;
  k<<<148 * 16, 256>>>(bad, bad1, total);
  cudaDeviceSynchronize();
  cudaMemcpy(&h[0], bad, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&h[1], bad1, 8, cudaMemcpyDeviceToHost);
  printf("pairs %llu  mismatches(two corrections) %llu  (of those, one correction also wrong: %llu)  err %s\n", total, h[0], h[1], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
