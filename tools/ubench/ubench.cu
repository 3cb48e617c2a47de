// Micro-benchmarks of primitives the codec kernels lean on (development tool).  nvcc -arch=sm_100a -O3 ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_sync(long long* out, int iters) {
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (t1 - t0) / iters;
}
__global__ void k_match(long long* out, int iters, unsigned seed) {
  unsigned v = threadIdx.x * 2654435761u + seed, acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { acc += __match_any_sync(0xffffffffu, (v >> (i & 7)) & 0x7fff); v = v * 1664525u + 1013904223u; }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[1] = (t1 - t0) / iters; out[7] = acc; }
}
__global__ void k_ballot15(long long* out, int iters, unsigned seed) {
  unsigned v = threadIdx.x * 2654435761u + seed, acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    unsigned h = (v >> (i & 7)) & 0x7fff, m = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 15; b++) { unsigned bit = (h >> b) & 1; unsigned bl = __ballot_sync(0xffffffffu, bit); m &= bit ? bl : ~bl; }
    acc += m; v = v * 1664525u + 1013904223u;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[2] = (t1 - t0) / iters; out[7] = acc; }
}
__global__ void k_shfl(long long* out, int iters) {
  unsigned v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) v += __shfl_xor_sync(0xffffffffu, v, 1 + (i & 15));
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[3] = (t1 - t0) / iters; out[7] = v; }
}
__global__ void k_lds_chain(long long* out, int iters) {
  __shared__ unsigned short tab[16384];
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) tab[i] = (unsigned short)((i * 7919 + 13) & 16383);
  __syncthreads();
  unsigned p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) p = tab[p];
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[4] = (t1 - t0) / iters; out[7] = p; }
}
__global__ void k_sts_lds(long long* out, int iters) {
  __shared__ unsigned short tab[16384];
  unsigned p = threadIdx.x * 37;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { tab[p & 16383] = (unsigned short)i; __syncwarp(); p += tab[(p + 64) & 16383] + 1; __syncwarp(); }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[5] = (t1 - t0) / iters; out[7] = p; }
}
__global__ void k_atoms(long long* out, int iters) {
  __shared__ unsigned h[320];
  if (threadIdx.x < 320) h[threadIdx.x] = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) atomicAdd(&h[(threadIdx.x & 3) + 288], 1u);   // 4 hot addresses, like dist symbols
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[6] = (t1 - t0) / iters; out[7] = h[288]; }
}
int main() {
  long long* d; cudaMalloc(&d, 64); long long h[8];
  const int it = 2000;
  k_sync<<<1, 1024>>>(d, it); k_match<<<1, 32>>>(d, it, 1); k_ballot15<<<1, 32>>>(d, it, 1); k_shfl<<<1, 32>>>(d, it);
  k_lds_chain<<<1, 32>>>(d, it); k_sts_lds<<<1, 32>>>(d, it);
  cudaDeviceSynchronize(); cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("cycles per op, single warp unless noted:\n syncthreads(1024 thr) %lld\n match_any(15-bit keys) %lld\n 15-ballot emulation %lld\n shfl dependent %lld\n LDS dependent chain %lld\n STS+syncwarp+LDS+syncwarp %lld\n", h[0], h[1], h[2], h[3], h[4], h[5]);
  k_atoms<<<1, 1024>>>(d, it); cudaDeviceSynchronize(); cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf(" smem atomicAdd, 1024 threads on 4 hot addresses: %lld cycles per warp-instruction round (all 32 warps)\n", h[6]);
  k_match<<<1, 1024>>>(d, it, 1); cudaDeviceSynchronize(); cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf(" match_any with 32 warps resident: %lld\n", h[1]);
  return 0;
}
