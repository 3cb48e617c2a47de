// common.cuh — shared definitions for the sm_100a kernels of the mtscomp per-chunk codec.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef MTSCOMP_EMU
#include "emu/cuda_emu.h"
#else
#include <cuda_runtime.h>
#define MTS_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define MTS_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace mts {

// One row of the chunk table that every batched kernel receives (device array, one entry per chunk of the batch).
struct ChunkDesc {
  long long elem_off;   // offset (in elements) of the chunk's first sample in the raw / transformed buffers
  int ns;               // samples (rows) in this chunk
  int first_seg;        // index of the chunk's first deflate segment in the segment table
  int n_seg;            // number of deflate segments
  int pad_;
};

enum { FLAG_TIME_DIFF = 1, FLAG_SPATIAL_DIFF = 2, FLAG_ORDER_C = 4, FLAG_FLOAT = 8 };   // FLAG_FLOAT: IEEE float32 / float64 elements

static const uint32_t ADLER_BASE = 65521u;

// Barrier among `count` threads (a multiple of 32) of the CTA on hardware barrier `id` (1..15; 0 is __syncthreads).
__device__ __forceinline__ void named_barrier(int id, int count) {
#ifdef MTSCOMP_EMU
  __emu_named_barrier(id, count);
#else
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}

// Non-blocking arrival at the same kind of barrier (the other `count - 32k` threads use named_barrier on this id).
__device__ __forceinline__ void named_barrier_arrive(int id, int count) {
#ifdef MTSCOMP_EMU
  __emu_named_barrier_arrive(id, count);
#else
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
#endif
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned warp_id() { return threadIdx.x >> 5; }

template <class T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <class T> __device__ __forceinline__ T warp_incl_scan(T v) {
  unsigned l = lane_id();
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T n = __shfl_up_sync(0xffffffffu, v, o);
    if (l >= (unsigned)o) v += n;
  }
  return v;
}

}  // namespace mts
