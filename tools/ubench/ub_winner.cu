// Which lane wins when several lanes of one STS store to the same shared-memory address?  (development tool)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned* out, int iters) {
  __shared__ unsigned short tab[1024];
  unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned v = (threadIdx.x + 1) * 2654435761u;
  unsigned bad = 0, dups = 0, bad2 = 0;
  for (int it = 0; it < iters; it++) {
    v = v * 1664525u + 1013904223u;
    unsigned nbits = 1 + (it % 6);
    unsigned a = (v >> 7) & ((1u << nbits) - 1);          // few distinct addresses -> many collisions
    unsigned short* p = &tab[w * 64 + a];
    __syncwarp();
    *p = (unsigned short)lane;
    __syncwarp();
    unsigned r = *p;
    // highest lane with the same address
    unsigned want = lane;
    for (int l = 31; l > (int)lane; l--) if (__shfl_sync(0xffffffffu, a, l) == a) { want = l; break; }
    unsigned hi = 0;
    for (int l = 0; l < 32; l++) { unsigned al = __shfl_sync(0xffffffffu, a, l); if (al == a && (unsigned)l > hi) hi = l; }
    unsigned lo = 31;
    for (int l = 31; l >= 0; l--) { unsigned al = __shfl_sync(0xffffffffu, a, l); if (al == a) lo = l; }
    if (r != hi) bad++;
    if (r != lo) bad2++;
    if (hi != lane) dups++;
  }
  atomicAdd(&out[0], bad);
  atomicAdd(&out[1], dups);
  atomicAdd(&out[2], bad2);
}
int main() {
  unsigned* d; cudaMalloc(&d, 16); cudaMemset(d, 0, 16); unsigned h[4];
  k<<<4, 256>>>(d, 20000);
  cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("same-address STS.U16: %u lanes saw a winner other than the highest colliding lane (of %u colliding lane-stores)\n", h[0], h[1]);
  printf("winner != lowest colliding lane: %u\n", h[2]);
  return 0;
}
