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

// ---- TMA bulk copies (cp.async.bulk) completing on an mbarrier: one elected thread arms the barrier with the byte
//      count and issues the copy, every consumer waits for the phase.
#ifdef MTSCOMP_EMU
// host emulation: the "asynchronous" copy completes at once
typedef unsigned long long* mbar_t;
__device__ __forceinline__ mbar_t mbar_addr(unsigned long long* bar) { return bar; }
__device__ __forceinline__ void mbar_init(mbar_t bar) { *bar = 0; }
__device__ __forceinline__ void mbar_expect_tx(mbar_t, unsigned) {}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, mbar_t) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void mbar_wait(mbar_t, unsigned) {}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void bulk_s2g_wait() {}
__device__ __forceinline__ void fence_async_smem() {}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
__device__ __forceinline__ void cp_async_commit() {}
template <int N> __device__ __forceinline__ void cp_async_wait() {}
#else
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
typedef unsigned mbar_t;                         // shared-window address of an mbarrier
__device__ __forceinline__ mbar_t mbar_addr(unsigned long long* bar) { return smem_u32(bar); }
__device__ __forceinline__ void mbar_init(mbar_t bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(mbar_t bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (16-byte aligned addresses, size a multiple of 16); completes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, mbar_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(mbar_t bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// TMA bulk copy shared -> global (same alignment rules); bulk_s2g_wait() returns when the sources may be reused.
// Shared memory written by ordinary stores must be fenced (fence_async_smem) before the copy engine reads it.
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_s2g_wait() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Per-thread asynchronous copies of 16 bytes global -> shared (no destination register, so nothing in the issuing warp
// waits for the load until cp_async_wait<N>: all but the thread's N most recent groups are complete).
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

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
