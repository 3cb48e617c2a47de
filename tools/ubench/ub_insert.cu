// Micro-benchmark of lz_insert_step in isolation (development tool).
#include <cstdio>
#include "../../mtscomp_b200/csrc/deflate.cuh"
using namespace mts;
__global__ void k_ins(long long* out, int iters, int mode) {
  extern __shared__ __align__(16) unsigned char sm[];
  unsigned short* head = (unsigned short*)sm;                 // 64 KB
  unsigned short* prev = (unsigned short*)(sm + 65536);       // 32 KB
  unsigned short* hbuf = (unsigned short*)(sm + 98304);       // 960 x u16
  for (int i = threadIdx.x; i < 32768; i += blockDim.x) head[i] = 0;
  unsigned v = 12345;
  for (int i = threadIdx.x; i < 960; i += blockDim.x) { unsigned x = (i * 2654435761u) ^ (i >> 3); hbuf[i] = mode == 1 ? (unsigned short)(x % 1021) : (unsigned short)(x & 0x7fff); }
  __syncthreads();
  if (threadIdx.x < 32) {
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) lz_insert_step<2, true>(hbuf, head, prev, it * 960, 960, threadIdx.x);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[mode] = (t1 - t0) / iters;
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64); long long h[8];
  cudaFuncSetAttribute(k_ins, cudaFuncAttributeMaxDynamicSharedMemorySize, 110000);
  k_ins<<<1, 64, 110000>>>(d, 200, 0);
  k_ins<<<1, 64, 110000>>>(d, 200, 1);
  cudaDeviceSynchronize(); cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("lz_insert_step<2,true>, 960 units (30 batches): %lld cycles (distinct hashes), %lld cycles (1021 distinct values: many in-batch collisions)\n", h[0], h[1]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
