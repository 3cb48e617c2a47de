// deflate.cuh — K2: parallel DEFLATE encoder (RFC 1951) producing one valid zlib stream (RFC 1950) per chunk.
//
// Replaces `zlib.compress(chunkd.tobytes(order))` at mtscomp.py:394.  The reference Reader only requires that
// zlib.decompress() accepts the stream and returns the transformed bytes (mtscomp.py:619), so the encoder
// is free to pick its own parse; the contract is decodability + size <= 1.031x zlib level 6 (north star).
//
// Each chunk's transformed bytes are cut into independent "segments" (fresh 32 KB window, own dynamic-Huffman block,
// closed by an empty stored block so the next segment starts byte-aligned).  Four kernels, split by parallelism shape:
//   lz77_kernel     one CTA / segment : shared-memory hash-chain match finder over a 64 KB data ring, every position
//                                       matched in parallel, lazy parse by pointer doubling, token + histogram output
//   huff_kernel     one warp / segment: length-limited Huffman codes (<=15 / <=7 bits), RLE'd dynamic header, sizes
//   scan_kernel     one CTA           : exclusive prefix of segment byte sizes -> final offsets (packed .cbin layout)
//   encode_kernel   one CTA / segment : code lookup, prefix-sum bit offsets, bit-pack straight into the final stream,
//                                       zlib header / stored fallback / final block / adler32 trailer
#pragma once
#include "common.cuh"

namespace mts {

// ------------------------------------------------------------------------------------------------ tables
struct DeflateSeg {
  long long in_off;     // byte offset of the segment's input in the transformed buffer
  long long tok_off;    // element offset of the segment's token area (capacity = in_len u16 elements)
  int in_len;           // input bytes (> 0)
  int chunk;            // owning chunk (index within the batch)
  int flags;            // SEG_FIRST | SEG_LAST
  int pad_;
};
enum { SEG_FIRST = 1, SEG_LAST = 2 };
enum { MODE_STORED = 0, MODE_DYNAMIC = 1, MODE_FIXED = 2 };

struct DeflateSegOut {   // written by huff_kernel / scan_kernel, read by encode_kernel
  unsigned n_tok;        // u16 token elements produced by lz77_kernel
  unsigned hdr_bits;     // dynamic header length in bits
  unsigned mode;
  unsigned body_bytes;   // bytes of this segment's deflate data (without zlib header/trailer)
  long long out_off;     // byte offset in the packed output
};

static const int LL_SYMS = 286, D_SYMS = 30, BL_SYMS = 19;
static const int HIST_STRIDE = 320;          // u32 per segment: [0,286) lit/len, [288,318) dist
static const int HDR_WORDS = 192;            // dynamic header capacity (u32) per segment
static const int CODE_STRIDE = 320;          // u32 per segment: code | len<<16, same layout as the histogram

struct LzParams {
  int max_chain;    // candidates examined per position
  int nice_len;     // stop searching at this length
  int far4;         // a 4-byte match is accepted only if dist <= far4
  int far5;         // a 5-byte match is accepted only if dist <= far5
  int far6;         // a 6-byte match is accepted only if dist <= far6
  int lazy;         // 1: defer to a longer match at the next position
};

__device__ __forceinline__ void len_symbol(unsigned len, unsigned& sym, unsigned& nb, unsigned& ev) {
  unsigned l = len - 3;
  if (l < 8) { sym = 257 + l; nb = 0; ev = 0; }
  else if (len == 258) { sym = 285; nb = 0; ev = 0; }
  else { nb = 29 - __clz((int)l); sym = 261 + 4 * nb + ((l >> nb) & 3); ev = l & ((1u << nb) - 1); }
}
__device__ __forceinline__ void dist_symbol(unsigned dist, unsigned& sym, unsigned& nb, unsigned& ev) {
  unsigned d = dist - 1;
  if (d < 4) { sym = d; nb = 0; ev = 0; }
  else { nb = 30 - __clz((int)d); sym = 2 * nb + 2 + ((d >> nb) & 1); ev = d & ((1u << nb) - 1); }
}
__device__ __forceinline__ unsigned ll_extra_bits(unsigned sym) {   // sym in [0,286)
  if (sym < 265 || sym == 285) return 0;
  return (sym - 261) >> 2;
}
__device__ __forceinline__ unsigned d_extra_bits(unsigned sym) { return sym < 4 ? 0 : (sym >> 1) - 1; }

// ------------------------------------------------------------------------------------------------ lz77_kernel
// One persistent CTA (1024 threads, 1 per SM) per segment.  STRIDE = bytes between indexed positions: 2 for int16
// streams (matches are searched at sample boundaries, the byte in between inherits the next sample's match extended
// backwards), 1 generic.  Two match finders share the 64 KB data ring:
//   L ("long")   hash of the 6 bytes at a unit -> headL (2^15 x u16) + prevL chain over the whole window; the searcher
//                walks up to max_chain candidates and keeps the longest match (>= 6 bytes)
//   S ("short")  hash of the 4 bytes at a unit -> headS (2^14 x u16); only the NEAREST previous unit with the same 4
//                bytes is kept (prevS, two steps deep) and is used when L finds nothing
// The segment is processed in steps of 960 units.  Warps 0 and 1 are the INSERTERS (S resp. L table; warp 0 also
// prefetches input): they thread the units of step s+1 into the tables, in position order, while warps 2..31 SEARCH
// step s, one unit per thread.  Then all warps parse the step (greedy, pointer doubling) and emit tokens + histogram.
static const int LZ_THREADS = 1024;
static const int LZ_UNITS = 960;              // units searched per step = 30 searcher warps x 32 lanes
static const int LZ_RING = 65536;
static const int LZ_MIRROR = 288;             // ring[65536 + i] mirrors ring[i]: reads of up to 258 + 16 + 4 bytes never wrap
static const int LZ_HASHS_BITS = 14;
static const unsigned LZ_BIAS = 32768;

template <int STRIDE> struct LzSmem {
  static const int SEG = LZ_UNITS * STRIDE;   // bytes per step
  static const int PREV_N = 32768 / STRIDE;
  static const int HL_BITS = STRIDE == 1 ? 14 : 15;   // headL size: what still fits next to the 64 KB prev ring of STRIDE 1
  // the inserters run one step ahead, so chain entries older than PREV_N - 2 steps may already be recycled
  static const int MAXD_UNITS = PREV_N - 2 * LZ_UNITS - 8;
  static const size_t ring_off = 0;
  static const size_t headl_off = LZ_RING + LZ_MIRROR;
  static const size_t heads_off = headl_off + (size_t)(1 << HL_BITS) * 2;
  static const size_t prevl_off = heads_off + (size_t)(1 << LZ_HASHS_BITS) * 2;
  static const size_t prevs_off = prevl_off + (size_t)PREV_N * 2;
  static const size_t hbuf_off = prevs_off + (size_t)(2 * LZ_UNITS) * 2;   // [2 steps][S,L][LZ_UNITS] u16 hashes
  static const size_t mlen_off = hbuf_off + (size_t)(4 * LZ_UNITS) * 2;
  static const size_t mdist_off = mlen_off + (size_t)(LZ_UNITS + 8) * 2;       // per unit: match length | back-extension flag
  static const size_t jump_off = (mdist_off + (size_t)(LZ_UNITS + 8) * 2 + 3) & ~(size_t)3;
  static const size_t hist_off = (jump_off + (size_t)(LZ_UNITS + 8) * 4 + 15) & ~(size_t)15;
  static const size_t misc_off = hist_off + (size_t)HIST_STRIDE * 4;
  static const size_t total = misc_off + 512;
};

__device__ __forceinline__ unsigned ring_load4(const unsigned char* ring, unsigned r) {   // r: any ring coordinate
  const unsigned m = r & 0xffffu;
  const unsigned* w = (const unsigned*)(ring + (m & 0xfffcu));
  return __funnelshift_r(w[0], w[1], m << 3);              // w[1] may lie in the mirror
}
// store 16 bytes at ring offset o (multiple of 16, < LZ_RING), keeping the mirror in step
__device__ __forceinline__ void ring_store16(unsigned char* ring, unsigned o, uint4 v) {
  *(uint4*)(ring + o) = v;
  if (o < (unsigned)LZ_MIRROR) *(uint4*)(ring + LZ_RING + o) = v;
}
__device__ __forceinline__ unsigned lz_hash4(unsigned w) { return (w * 0x9E3779B1u) >> (32 - LZ_HASHS_BITS); }
template <int BITS> __device__ __forceinline__ unsigned lz_hash6(unsigned w0, unsigned w1) {
  return ((w0 * 0x9E3779B1u) ^ ((w1 & 0xffffu) * 0x85EBCA6Bu)) >> (32 - BITS);
}

// Hashes of the unit at position p for both tables, 0xffff where the key would run past the end of the segment.
// (Computed by the searcher threads two steps ahead of their use, so the serial inserters only touch the tables.)
template <int STRIDE>
__device__ __forceinline__ void lz_unit_hashes(const unsigned char* ring, unsigned p, unsigned n, unsigned off0,
                                               unsigned short& hs, unsigned short& hl) {
  hs = 0xffff; hl = 0xffff;
  if (p + 4 <= n) {
    const unsigned w0 = ring_load4(ring, p + off0);
    hs = (unsigned short)lz_hash4(w0);
    if (p + 6 <= n) hl = (unsigned short)lz_hash6<LzSmem<STRIDE>::HL_BITS>(w0, ring_load4(ring, p + off0 + 4));
  }
}

// Work a searcher thread does for LATER steps: hashes of its unit two steps ahead (position p2) into hbuf, and the
// store of its prefetched 16 input bytes into the ring.
template <int STRIDE>
__device__ __forceinline__ void lz_ahead(unsigned char* ring, unsigned short* hbuf, unsigned step, unsigned tid, unsigned p2,
                                         unsigned n, unsigned off0, bool pf, unsigned pf_rc, uint4 pf_v) {
  unsigned short hs, hl;
  lz_unit_hashes<STRIDE>(ring, p2, n, off0, hs, hl);
  unsigned short* hb = hbuf + (step & 1) * 2 * LZ_UNITS;
  hb[tid] = hs;
  hb[LZ_UNITS + tid] = hl;
  if (pf) ring_store16(ring, pf_rc & 0xffffu, pf_v);
}

// Thread `units` consecutive units starting at unit u0 into one hash table (one warp), batches of 64 units (two
// consecutive units per lane) in position order.  Every unit links to the table head as it was BEFORE its batch — or to
// its lane's first unit when both have the same hash; units of different lanes of one batch never link to each other
// (such matches are < 64 units away and later batches find them anyway).  Then the batch's LAST unit of each hash
// becomes the new head: colliding lanes re-store until the highest one has won, so the result does not depend on how
// the hardware orders same-address stores.  No __match_any_sync: it costs > 1000 cycles per call when many warps use it.
// LONG: links go to the prev ring (index u & pmask); otherwise to a two-step array.
template <int STRIDE, bool LONG>
__device__ __forceinline__ void lz_insert_step(const unsigned short* hbuf, unsigned short* head, unsigned short* prev,
                                               unsigned u0, unsigned units, unsigned lane) {
  const unsigned PM = LzSmem<STRIDE>::PREV_N - 1;
  const unsigned pbase = LONG ? 0 : (u0 % (2 * LZ_UNITS));     // S links live in a two-step array
  // software pipeline: the next batch's hashes are fetched before, and its head lookups right after, this batch's stores
  // have settled
  unsigned hA = (2 * lane < units) ? hbuf[2 * lane] : 0xffffu;
  unsigned hB = (2 * lane + 1 < units) ? hbuf[2 * lane + 1] : 0xffffu;
  unsigned short oldA = hA != 0xffffu ? head[hA] : (unsigned short)0;
  unsigned short oldB = hB != 0xffffu ? head[hB] : (unsigned short)0;
  for (unsigned b = 0; b < units; b += 64) {
    const unsigned iA = b + 2 * lane, iB = iA + 1;
    const unsigned short ubA = (unsigned short)(u0 + iA + LZ_BIAS), ubB = (unsigned short)(ubA + 1);
    const bool vA = hA != 0xffffu, vB = hB != 0xffffu;
    const bool same = vA && hA == hB;                          // the lane's second unit follows its first
    const unsigned nA = (iA + 64 < units) ? hbuf[iA + 64] : 0xffffu;
    const unsigned nB = (iB + 64 < units) ? hbuf[iB + 64] : 0xffffu;
    __syncwarp();
    if (vA && !same) head[hA] = ubA;
    if (vB) head[hB] = ubB;
    __syncwarp();
    // a later unit of this batch must end up as the head: re-store while an earlier one is visible
    bool wA = vA && !same && (unsigned short)(ubA - head[hA] - 1) < 63;
    bool wB = vB && (unsigned short)(ubB - head[hB] - 1) < 63;
    while (__any_sync(0xffffffffu, wA || wB)) {
      if (wA) head[hA] = ubA;
      if (wB) head[hB] = ubB;
      __syncwarp();
      wA = vA && !same && (unsigned short)(ubA - head[hA] - 1) < 63;
      wB = vB && (unsigned short)(ubB - head[hB] - 1) < 63;
    }
    // only now (the heads are settled, whichever store the hardware let win first) look up the next batch's links;
    // they are not needed before the end of the next iteration, so this load is off the critical path
    const unsigned short noA = nA != 0xffffu ? head[nA] : (unsigned short)0;
    const unsigned short noB = nB != 0xffffu ? head[nB] : (unsigned short)0;
    if (iA < units) prev[LONG ? ((u0 + iA) & PM) : (pbase + iA)] = vA ? oldA : ubA;   // invalid: distance 0 = none
    if (iB < units) prev[LONG ? ((u0 + iB) & PM) : (pbase + iB)] = vB ? (same ? ubA : oldB) : ubB;
    hA = nA; hB = nB; oldA = noA; oldB = noB;
  }
  __syncwarp();
}

// Length of the match between the position whose first 16 bytes are w0..w3 (ring offset pm < LZ_RING) and the candidate
// at ring offset qm < LZ_RING, up to lim.  All reads run forward without wrapping (mirror).
__device__ __forceinline__ unsigned lz_match_len(const unsigned char* ring, unsigned pm, unsigned qm, unsigned lim,
                                                 unsigned w0, unsigned w1, unsigned w2, unsigned w3) {
  const unsigned* qw = (const unsigned*)(ring + (qm & 0xfffcu));
  const unsigned sh = qm << 3;
  unsigned len = 0;
  unsigned t0 = qw[0], t1 = qw[1];
  unsigned x = __funnelshift_r(t0, t1, sh) ^ w0;
  if (!x) { len = 4; t0 = qw[2]; x = __funnelshift_r(t1, t0, sh) ^ w1;
    if (!x) { len = 8; t1 = qw[3]; x = __funnelshift_r(t0, t1, sh) ^ w2;
      if (!x) { len = 12; t0 = qw[4]; x = __funnelshift_r(t1, t0, sh) ^ w3;
        if (!x) { len = 16;
          while (len < lim) {
            x = ring_load4(ring, pm + len) ^ ring_load4(ring, qm + len);
            if (x) break;
            len += 4;
          } } } } }
  if (x) len += (unsigned)(__ffs((int)x) - 1) >> 3;
  return min(len, lim);
}

#if defined(MTS_LZ_PROFILE) && !defined(MTSCOMP_EMU)
// development instrumentation: cycles per phase, summed over steps and CTAs
// [0] steps [1] A/B phase (thread 0) [2] thread 0's own search [3] S inserter [4] L inserter [5] parse+emit phase
__device__ unsigned long long g_lz_prof[16];
#define LZ_PROF_T(var) long long var = clock64()
#define LZ_PROF_ADD(i, v) do { if (blockIdx.x == 0) atomicAdd(&g_lz_prof[i], (unsigned long long)(v)); } while (0)
#else
#define LZ_PROF_T(var)
#define LZ_PROF_ADD(i, v)
#endif

template <int STRIDE>
__global__ void __launch_bounds__(LZ_THREADS, 1) lz77_kernel(const unsigned char* __restrict__ tbuf,
                                                             const DeflateSeg* __restrict__ segs, int n_segs,
                                                             unsigned short* __restrict__ tokens,
                                                             unsigned* __restrict__ hist, DeflateSegOut* __restrict__ so,
                                                             LzParams prm) {
  typedef LzSmem<STRIDE> L;
  const unsigned SEG = L::SEG;
  const unsigned NSW = 30;                          // searcher / parser warps; warps 30 and 31 are the inserters
  MTS_DYN_SMEM(sm);
  unsigned char* ring = sm + L::ring_off;
  unsigned short* headL = (unsigned short*)(sm + L::headl_off);
  unsigned short* headS = (unsigned short*)(sm + L::heads_off);
  unsigned short* prevL = (unsigned short*)(sm + L::prevl_off);
  unsigned short* prevS = (unsigned short*)(sm + L::prevs_off);
  unsigned short* hbuf = (unsigned short*)(sm + L::hbuf_off);   // hbuf[(step & 1) * 2 * LZ_UNITS + (LONG ? LZ_UNITS : 0) + k]
  unsigned short* mlen = (unsigned short*)(sm + L::mlen_off);
  unsigned short* mdist = (unsigned short*)(sm + L::mdist_off);
  unsigned* xe = (unsigned*)(sm + L::jump_off);                   // per unit: exit of its stretch | token elements << 16
  unsigned* shist = (unsigned*)(sm + L::hist_off);
  unsigned* misc = (unsigned*)(sm + L::misc_off);   // [0..29] element base of each stretch, [33],[34] carry, [35] step total
  unsigned* ent = misc + 36;                         // [0..29] parse entry position of each stretch
  const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned PM = L::PREV_N - 1;

  for (int sidx = blockIdx.x; sidx < n_segs; sidx += gridDim.x) {
    const DeflateSeg sg = segs[sidx];
    const unsigned n = (unsigned)sg.in_len;
    const unsigned char* in = tbuf + sg.in_off;
    const unsigned off0 = (unsigned)((uintptr_t)in & 15);
    const uint4* in16 = (const uint4*)(in - off0);
    unsigned short* tok = tokens + sg.tok_off;
    const unsigned n_ring = n + off0;                    // ring coordinates [off0, n_ring) are real input
    unsigned run_tok = 0;                                // token elements emitted so far (uniform across threads)

    // reset tables (headL and headS are contiguous); initial load: ring coordinates [0, 3*SEG + 32) (the hashes of
    // step 2 read up to 3*SEG + 8 + 15)
    for (unsigned i = tid; i < ((1u << L::HL_BITS) + (1u << LZ_HASHS_BITS)) / 2; i += LZ_THREADS) ((unsigned*)headL)[i] = 0;
    for (unsigned i = tid; i < HIST_STRIDE; i += LZ_THREADS) shist[i] = 0;
    if (tid == 0) { misc[33] = 0; misc[34] = 0; }
    for (unsigned v = tid; v * 16 < 3 * SEG + 32; v += LZ_THREADS)
      if (v * 16 < n_ring) ring_store16(ring, v * 16, in16[v]);
    __syncthreads();
    // hashes of the units of steps 0 and 1 (later steps: computed two steps ahead by the searchers)
    for (unsigned k = tid; k < 2 * LZ_UNITS; k += LZ_THREADS) {
      unsigned short hs, hl;
      lz_unit_hashes<STRIDE>(ring, k * STRIDE, n, off0, hs, hl);
      const unsigned st = k / LZ_UNITS, kk = k % LZ_UNITS;
      hbuf[st * 2 * LZ_UNITS + kk] = hs;
      hbuf[st * 2 * LZ_UNITS + LZ_UNITS + kk] = hl;
    }
    __syncthreads();
    {
      const unsigned units0 = min((unsigned)LZ_UNITS, (n + STRIDE - 1) / STRIDE);
      if (wid == 30) lz_insert_step<STRIDE, false>(hbuf, headS, prevS, 0, units0, lane);
      if (wid == 31) lz_insert_step<STRIDE, true>(hbuf + LZ_UNITS, headL, prevL, 0, units0, lane);
    }
    __syncthreads();

    const unsigned n_steps = (n + SEG - 1) / SEG;
    for (unsigned step = 0; step < n_steps; step++) {
      const unsigned s0 = step * SEG;                    // first position of this step
      const unsigned slen = min(SEG, n - s0);

      LZ_PROF_T(t_step);
      // prefetch of ring coordinates [(step+3)*SEG + 32, (step+4)*SEG + 32) by the first SEG/16 searcher threads: loaded
      // before the search, stored (with the hashes of the units two steps ahead) where the thread has nothing else to do
      const unsigned pf_rc = (step + 3) * SEG + 32 + tid * 16;
      const bool pf = wid < NSW && tid * 16 < SEG && pf_rc < n_ring;
      uint4 pf_v = make_uint4(0, 0, 0, 0);
      if (wid >= NSW) {
        // ---- (A) inserters (the two highest warp ids: the issue arbiter favours them over the searchers): prefetch
        //          ring coordinates [(step+3)*SEG, (step+4)*SEG) and thread step+1's units into the tables
        const unsigned s1 = s0 + SEG;
        if (s1 < n) {
          const unsigned units1 = (min(SEG, n - s1) + STRIDE - 1) / STRIDE;
          const unsigned short* hb = hbuf + ((step + 1) & 1) * 2 * LZ_UNITS;
          if (wid == 30) lz_insert_step<STRIDE, false>(hb, headS, prevS, s1 / STRIDE, units1, lane);
          else lz_insert_step<STRIDE, true>(hb + LZ_UNITS, headL, prevL, s1 / STRIDE, units1, lane);
        }
        if (lane == 0) { LZ_PROF_T(t_i); LZ_PROF_ADD(wid == 30 ? 3 : 4, t_i - t_step); }
      } else {
        // ---- (B) searchers: one unit per thread.  The first SEG/16 of them also prefetch ring coordinates
        //      [(step+3)*SEG + 32, (step+4)*SEG + 32): load now, store after the search has hidden the latency.
        const unsigned li = tid * STRIDE;                 // local position in the step
        const unsigned p = s0 + li;
        unsigned best = 0, bdist = 0;
        if (pf) pf_v = in16[pf_rc >> 4];
        if (li < slen && p + 4 <= n && prm.lazy != 2) {
          const unsigned u = p / STRIDE;
          const unsigned lim = min(258u, n - p);
          const unsigned pr = (p + off0) & 0xffffu;       // ring offset of this unit
          unsigned w0, w1, w2, w3;                        // the first 16 bytes at p stay in registers
          {
            const unsigned* pw = (const unsigned*)(ring + (pr & 0xfffcu));
            const unsigned sh = pr << 3;
            const unsigned t0 = pw[0], t1 = pw[1], t2 = pw[2], t3 = pw[3], t4 = pw[4];
            w0 = __funnelshift_r(t0, t1, sh); w1 = __funnelshift_r(t1, t2, sh);
            w2 = __funnelshift_r(t2, t3, sh); w3 = __funnelshift_r(t3, t4, sh);
          }
          const unsigned ub = (u + LZ_BIAS) & 0xffffu;
          if (lim >= 6) {
            // L chain: candidates share (the hash of) 6 bytes; keep the longest
            best = 5;
            unsigned wq = __funnelshift_r(w0, w1, 16);     // bytes [best-3, best] of p
            unsigned short cand = prevL[u & PM];
            unsigned lastd = 0;
            for (int depth = prm.max_chain; depth > 0; depth--) {
              const unsigned du = (ub - cand) & 0xffffu;
              if (du - 1 >= (unsigned)L::MAXD_UNITS || du <= lastd || du > u) break;
              lastd = du;
              const unsigned dist = du * STRIDE;
              const unsigned qr = (pr - dist) & 0xffffu;
              cand = prevL[(cand - LZ_BIAS) & PM];
              if (bdist && ring_load4(ring, qr + best - 3) != wq) continue;   // cannot beat the match in hand
              const unsigned len = lz_match_len(ring, pr, qr, lim, w0, w1, w2, w3);
              if (len > best) {
                best = len; bdist = dist;
                if (len >= (unsigned)prm.nice_len || len >= lim) break;
                wq = ring_load4(ring, pr + best - 3);
              }
            }
          }
          if (!bdist) {
            // S: the nearest previous unit with the same 4 bytes
            best = 0;
            const unsigned du = (ub - prevS[u % (2 * LZ_UNITS)]) & 0xffffu;
            if (du - 1 < (unsigned)L::MAXD_UNITS && du <= u) {
              const unsigned dist = du * STRIDE;
              const unsigned len = lz_match_len(ring, pr, (pr - dist) & 0xffffu, lim, w0, w1, w2, w3);
              if (len >= 4) { best = len; bdist = dist; }
            }
          }
        }
        // hashes of this thread's unit two steps ahead (consumed by the inserters during the next step) and the prefetch
        // store: warp 0 does them here, the other warps while warp 0 chains the stretches (they would only wait there)
        if (wid == 0) lz_ahead<STRIDE>(ring, hbuf, step, tid, s0 + li + 2 * SEG, n, off0, pf, pf_rc, pf_v);
        if (tid == 0) { LZ_PROF_T(t_s); LZ_PROF_ADD(2, t_s - t_step); }
        unsigned bext = 0;
        if (STRIDE == 2 && bdist) {
          // The parse works on units, so matches cover whole units (length truncated to even: a sample whose low byte
          // matched nearly always matches in its high byte too).  Bit 15 = the match also covers the byte BEFORE this
          // unit (the previous sample's high byte); used when the previous unit of the same stretch turns out to be a literal.
          best &= ~1u;
          const unsigned q = p - 1;
          bext = (lane > 0 && best < 258 && q >= bdist &&
                  ring[(q + off0) & 0xffffu] == ring[(q + off0 - bdist) & 0xffffu]) ? 0x8000u : 0u;
        }
        mlen[tid] = (unsigned short)(bdist ? (best | bext) : 0);
        mdist[tid] = (unsigned short)bdist;
      }
      // The 30 searcher warps now parse and emit the step among themselves (hardware barrier 1, 960 threads); the two
      // inserter warps keep threading step+1 into the tables and only rejoin at the end of the step.
      if (wid < NSW) {
      named_barrier(1, NSW * 32);
      LZ_PROF_T(t_p0);
      if (tid == 0) { LZ_PROF_ADD(0, 1); LZ_PROF_ADD(1, t_p0 - t_step); }
      // ---- (C) greedy parse over units, hierarchical; thread = unit, warp = stretch of 32 units.
      //      Stage 1 (registers, shuffles): pointer doubling gives every unit the exit of the token chain that starts
      //      there (where it leaves the stretch), the token elements it emits on the way and the mask of units it visits.
      //      Stage 2: one thread chains the 30 stretches from the carried start (entry + element base per stretch).
      //      Stage 3: the warp picks the mask of its entry unit and emits those tokens (ballot prefix -> offsets).
      const unsigned nu = (slen + STRIDE - 1) / STRIDE;   // units in this step
      const unsigned start = misc[33 + (step & 1)];       // local start unit carried from the previous step
      if (start < nu) {
        const unsigned sb = wid * 32, se = min(sb + 32, nu);
        const unsigned sel = se > sb ? se - sb : 0;        // valid lanes of this stretch
        const unsigned m = tid < nu ? mlen[tid] : 0;
        const unsigned mnext = tid + 1 < nu ? mlen[tid + 1] : 0;
        const unsigned l = m & 0x7fffu;
        unsigned el = 0;                                   // token elements of the token starting at this unit
        if (tid < nu) {
          if (l) el = 2;
          else {
            el = (STRIDE == 2 && s0 + tid * STRIDE + 1 < n) ? 2 : 1;       // a literal per byte ...
            if (mnext & 0x8000u) el--;                                     // ... unless the next match takes the last one
          }
        }
        unsigned v = (lane + (l ? l / STRIDE : 1u)) | (el << 16);   // next unit (stretch-local) | elements
        unsigned M = 1u << lane;
        for (unsigned r = 0; r < 5; r++) {
          const unsigned j = v & 0xffffu;
          const unsigned tv = __shfl_sync(0xffffffffu, v, j & 31);
          const unsigned tM = __shfl_sync(0xffffffffu, M, j & 31);
          if (j < sel) { v = (tv & 0xffffu) | ((v & 0xffff0000u) + (tv & 0xffff0000u)); M |= tM; }
        }
        xe[tid] = (sb + (v & 0xffffu)) | (v & 0xffff0000u);
        named_barrier(1, NSW * 32);
        if (tid == 0) { LZ_PROF_T(t_b); LZ_PROF_ADD(6, t_b - t_p0); }
        if (wid != 0) lz_ahead<STRIDE>(ring, hbuf, step, tid, s0 + tid * STRIDE + 2 * SEG, n, off0, pf, pf_rc, pf_v);
        if (wid == 0) {
          // Stage 2, by relaxation in one warp (lane = stretch): every lane guesses that its stretch is entered at
          // its first unit, looks up where that chain leaves, and hands the exit to the next lane as ITS entry; repeat
          // until no entry changes.  Lane w is certainly right after w rounds, but greedy parses that start a few
          // units apart merge almost at once, so the exits barely depend on the entries: 2-4 rounds instead of a
          // 30-step serial walk.
          const unsigned my_se = min((lane + 1) * 32, nu);
          unsigned e = lane == 0 ? start : lane * 32, t = 0;
          for (;;) {
            t = (lane < NSW && e < my_se) ? xe[e] : e;              // exit | elements << 16 (entry beyond the stretch: pass)
            unsigned ne = __shfl_up_sync(0xffffffffu, t & 0xffffu, 1);
            if (lane == 0) ne = start;
            const bool ch = lane < NSW && ne != e;
            e = ne;
            if (!__any_sync(0xffffffffu, ch)) break;
          }
          unsigned acc = lane < NSW ? t >> 16 : 0u;                  // elements emitted by my stretch; inclusive scan
          for (int d = 1; d < 32; d <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, acc, d); if ((int)lane >= d) acc += v; }
          if (lane < NSW) { ent[lane] = e; misc[lane] = acc - (t >> 16); }
          if (lane == NSW - 1) {
            misc[35] = acc;                                         // elements emitted by this step
            misc[33 + ((step + 1) & 1)] = (t & 0xffffu) - nu;       // >= 0: where the last token of this step ends
          }
        }
        named_barrier(1, NSW * 32);
        if (tid == 0) { LZ_PROF_T(t_b); LZ_PROF_ADD(7, t_b - t_p0); }
        const unsigned entry = ent[wid];
        if (entry < se) {                                   // warp-uniform
          const unsigned reach = __shfl_sync(0xffffffffu, M, entry - sb);
          const bool isr = (reach >> lane) & 1u;
          const unsigned mine = isr ? el : 0u;
          const unsigned lt = (1u << lane) - 1;
          const unsigned excl = __popc(__ballot_sync(0xffffffffu, mine & 1) & lt) +
                                2 * __popc(__ballot_sync(0xffffffffu, mine & 2) & lt);
          const unsigned pos = run_tok + misc[wid] + excl;
          if (isr) {
            if (l) {
              const unsigned d = mdist[tid];
              // back-extended over the last byte of the previous unit when that one is a (literal) token start
              const unsigned le = l + ((m >> 15) & (reach >> ((lane + 31) & 31)) & 1u);
              tok[pos] = (unsigned short)(0x8000u | le);
              tok[pos + 1] = (unsigned short)(d - 1);
              unsigned sym, nb, ev;
              len_symbol(le, sym, nb, ev);
              atomicAdd(&shist[sym], 1u);
              dist_symbol(d, sym, nb, ev);
              atomicAdd(&shist[288 + sym], 1u);
            } else {
              const unsigned r0 = s0 + tid * STRIDE + off0;
              const unsigned b0 = ring[r0 & 0xffffu];
              tok[pos] = (unsigned short)b0;
              atomicAdd(&shist[b0], 1u);
              if (mine == 2) {
                const unsigned b1 = ring[(r0 + 1) & 0xffffu];
                tok[pos + 1] = (unsigned short)b1;
                atomicAdd(&shist[b1], 1u);
              }
            }
          }
        }
        if (tid == 0) { LZ_PROF_T(t_b); LZ_PROF_ADD(8, t_b - t_p0); }
        run_tok += misc[35];
      } else {
        if (tid == 0) misc[33 + ((step + 1) & 1)] = start - nu;
        if (wid != 0) lz_ahead<STRIDE>(ring, hbuf, step, tid, s0 + tid * STRIDE + 2 * SEG, n, off0, pf, pf_rc, pf_v);
      }
      }   // wid < NSW
      __syncthreads();
      if (tid == 0) { LZ_PROF_T(t_end); LZ_PROF_ADD(5, t_end - t_step); }
    }

    // ---- segment done: publish histogram + token count
    for (unsigned i = tid; i < HIST_STRIDE; i += LZ_THREADS) hist[(size_t)sidx * HIST_STRIDE + i] = shist[i];
    if (tid == 0) so[sidx].n_tok = run_tok;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ huff_kernel
// One warp per segment.  Package: sort used symbols by frequency (rank sort across lanes), two-queue Huffman merge,
// depth count, zlib-style overflow repair to the length limit, canonical codes (bit-reversed for LSB-first packing).
struct HuffScratch {
  unsigned weight[2 * LL_SYMS];
  unsigned short parent[2 * LL_SYMS];
  unsigned short sorted[LL_SYMS];
  unsigned char depth[2 * LL_SYMS];
  unsigned char lens[LL_SYMS + D_SYMS + 8];
  unsigned char rle_sym[LL_SYMS + D_SYMS + 8];
  unsigned char rle_ext[LL_SYMS + D_SYMS + 8];
  unsigned bl_freq[BL_SYMS];
  unsigned bl_code[BL_SYMS];
  unsigned bl_count[16];
  unsigned next_code[16];
  unsigned hdr[HDR_WORDS];
};

// Builds code lengths for `nsym` symbols with frequencies f[] (u32, >= 2 non-zero entries) limited to maxbits, then
// canonical bit-reversed codes: out[s] = code | len << 16.  All lanes call; lane 0 does the serial parts.
__device__ void build_codes(const unsigned* f, int nsym, int maxbits, unsigned* out, unsigned char* lens_out,
                            HuffScratch* S) {
  const unsigned lane = lane_id();
  // rank sort of used symbols by (freq, symbol)
  int nused = 0;
  for (int s0 = 0; s0 < nsym; s0 += 32) {
    int s = s0 + lane;
    bool used = s < nsym && f[s] > 0;
    nused += __popc(__ballot_sync(0xffffffffu, used));
  }
  for (int s = lane; s < nsym; s += 32) {
    unsigned fs = f[s];
    lens_out[s] = 0;
    if (!fs) continue;
    int rank = 0;
    for (int t = 0; t < nsym; t++) {
      unsigned ft = f[t];
      rank += (ft > 0) && (ft < fs || (ft == fs && t < s));
    }
    S->sorted[rank] = (unsigned short)s;
  }
  __syncwarp();
  if (lane == 0) {
    const int n = nused;
    for (int i = 0; i < n; i++) S->weight[i] = f[S->sorted[i]];
    // two-queue merge: leaves [0,n), internal nodes [n, 2n-1)
    int a = 0, b = n, e = n;
    for (; e < 2 * n - 1; e++) {
      unsigned w = 0;
      for (int k = 0; k < 2; k++) {
        int pick;
        if (a < n && (b >= e || S->weight[a] <= S->weight[b])) pick = a++; else pick = b++;
        w += S->weight[pick];
        S->parent[pick] = (unsigned short)e;
      }
      S->weight[e] = w;
    }
    for (int i = 0; i <= maxbits; i++) S->bl_count[i] = 0;
    const int root = 2 * n - 2;
    S->depth[root] = 0;
    for (int i = root - 1; i >= 0; i--) {
      int d = S->depth[S->parent[i]] + 1;
      if (i < n) S->bl_count[min(d, maxbits)]++;
      S->depth[i] = (unsigned char)min(d, 255);
    }
    // length limiting: clamping deep leaves to maxbits over-subscribes the code by `excess` units of 2^-maxbits;
    // each zlib-style move (one leaf one level down, one clamped leaf becomes its sibling) removes exactly one unit
    long long excess = -(1ll << maxbits);
    for (int b = 1; b <= maxbits; b++) excess += (long long)S->bl_count[b] << (maxbits - b);
    while (excess > 0) {
      int bits = maxbits - 1;
      while (S->bl_count[bits] == 0) bits--;
      S->bl_count[bits]--;
      S->bl_count[bits + 1] += 2;
      S->bl_count[maxbits]--;
      excess--;
    }
    // least frequent symbols take the longest codes
    int i = 0;
    for (int bits = maxbits; bits >= 1; bits--)
      for (unsigned k = 0; k < S->bl_count[bits]; k++) lens_out[S->sorted[i++]] = (unsigned char)bits;
    unsigned code = 0;
    S->bl_count[0] = 0;
    for (int bits = 1; bits <= maxbits; bits++) {
      code = (code + S->bl_count[bits - 1]) << 1;
      S->next_code[bits] = code;
    }
    for (int s = 0; s < nsym; s++) {
      unsigned l = lens_out[s];
      unsigned c = 0;
      if (l) c = __brev(S->next_code[l]++) >> (32 - l);
      out[s] = c | (l << 16);
    }
  }
  __syncwarp();
}

struct BitW {   // serial LSB-first bit writer (lane 0 of huff_kernel)
  unsigned* w;
  unsigned nbits;
  __device__ void put(unsigned v, unsigned n) {
    if (!n) return;
    unsigned i = nbits >> 5, sh = nbits & 31;
    w[i] |= v << sh;
    if (sh + n > 32) w[i + 1] |= v >> (32 - sh);
    nbits += n;
  }
};

__global__ void __launch_bounds__(32) huff_kernel(const DeflateSeg* __restrict__ segs, int n_segs,
                                                  unsigned* __restrict__ hist, unsigned* __restrict__ codes,
                                                  unsigned* __restrict__ hdrs, DeflateSegOut* __restrict__ so) {
  __shared__ HuffScratch S;
  __shared__ unsigned f[HIST_STRIDE];
  __shared__ unsigned c[CODE_STRIDE];
  __shared__ unsigned char bl_len_sh[BL_SYMS + 5];
  __shared__ int sh_nr, sh_hlit, sh_hdist;
  const int sidx = blockIdx.x;
  if (sidx >= n_segs) return;
  const unsigned lane = lane_id();
  for (int i = lane; i < HIST_STRIDE; i += 32) f[i] = hist[(size_t)sidx * HIST_STRIDE + i];
  for (int i = lane; i < HDR_WORDS; i += 32) S.hdr[i] = 0;
  __syncwarp();
  if (lane == 0) {
    f[256] += 1;                                          // end-of-block
    // at least two used symbols per tree (zlib does the same so that every tree is complete)
    int used = 0;
    for (int i = 0; i < LL_SYMS; i++) used += f[i] > 0;
    for (int i = 0; used < 2; i++) if (!f[i]) { f[i] = 1; used++; }
    used = 0;
    for (int i = 0; i < D_SYMS; i++) used += f[288 + i] > 0;
    for (int i = 0; used < 2; i++) if (!f[288 + i]) { f[288 + i] = 1; used++; }
  }
  __syncwarp();
  build_codes(f, LL_SYMS, 15, c, S.lens, &S);
  build_codes(f + 288, D_SYMS, 15, c + 288, S.lens + LL_SYMS, &S);

  // body size in bits under the dynamic codes
  unsigned long long bits = 0;
  for (int s = lane; s < LL_SYMS; s += 32) bits += (unsigned long long)hist[(size_t)sidx * HIST_STRIDE + s] * ((c[s] >> 16) + ll_extra_bits(s));
  for (int s = lane; s < D_SYMS; s += 32) bits += (unsigned long long)hist[(size_t)sidx * HIST_STRIDE + 288 + s] * ((c[288 + s] >> 16) + d_extra_bits(s));
  bits = warp_sum(bits);
  bits += c[256] >> 16;

  if (lane == 0) {
    int hlit = LL_SYMS, hdist = D_SYMS;
    while (hlit > 257 && S.lens[hlit - 1] == 0) hlit--;
    while (hdist > 1 && S.lens[LL_SYMS + hdist - 1] == 0) hdist--;
    // concatenated length sequence, run-length coded with symbols 16/17/18 (RFC 1951 3.2.7)
    unsigned char seq[LL_SYMS + D_SYMS];
    int nseq = 0;
    for (int i = 0; i < hlit; i++) seq[nseq++] = S.lens[i];
    for (int i = 0; i < hdist; i++) seq[nseq++] = S.lens[LL_SYMS + i];
    for (int i = 0; i < BL_SYMS; i++) S.bl_freq[i] = 0;
    int nr = 0;
    for (int i = 0; i < nseq;) {
      int v = seq[i], run = 1;
      while (i + run < nseq && seq[i + run] == v) run++;
      i += run;
      if (v == 0) {
        while (run >= 11) { int r = min(run, 138); S.rle_sym[nr] = 18; S.rle_ext[nr++] = (unsigned char)(r - 11); run -= r; }
        if (run >= 3) { S.rle_sym[nr] = 17; S.rle_ext[nr++] = (unsigned char)(run - 3); run = 0; }
        while (run-- > 0) { S.rle_sym[nr] = 0; S.rle_ext[nr++] = 0; }
      } else {
        S.rle_sym[nr] = (unsigned char)v; S.rle_ext[nr++] = 0; run--;
        while (run >= 3) { int r = min(run, 6); S.rle_sym[nr] = 16; S.rle_ext[nr++] = (unsigned char)(r - 3); run -= r; }
        while (run-- > 0) { S.rle_sym[nr] = (unsigned char)v; S.rle_ext[nr++] = 0; }
      }
    }
    for (int i = 0; i < nr; i++) S.bl_freq[S.rle_sym[i]]++;
    int used = 0;
    for (int i = 0; i < BL_SYMS; i++) used += S.bl_freq[i] > 0;
    for (int i = 0; used < 2; i++) if (!S.bl_freq[i]) { S.bl_freq[i] = 1; used++; }
    sh_nr = nr; sh_hlit = hlit; sh_hdist = hdist;
  }
  __syncwarp();
  const int nr = sh_nr, hlit = sh_hlit, hdist = sh_hdist;
  build_codes(S.bl_freq, BL_SYMS, 7, S.bl_code, bl_len_sh, &S);

  if (lane == 0) {
    static const unsigned char order[BL_SYMS] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = BL_SYMS;
    while (hclen > 4 && bl_len_sh[order[hclen - 1]] == 0) hclen--;
    BitW bw{S.hdr, 0};
    bw.put(0, 1);             // BFINAL = 0 (the stream is closed by a separate empty final block)
    bw.put(2, 2);             // BTYPE = 10 dynamic
    bw.put((unsigned)(hlit - 257), 5);
    bw.put((unsigned)(hdist - 1), 5);
    bw.put((unsigned)(hclen - 4), 4);
    for (int i = 0; i < hclen; i++) bw.put(bl_len_sh[order[i]], 3);
    for (int i = 0; i < nr; i++) {
      unsigned s = S.rle_sym[i];
      bw.put(S.bl_code[s] & 0xffff, S.bl_code[s] >> 16);
      if (s == 16) bw.put(S.rle_ext[i], 2);
      else if (s == 17) bw.put(S.rle_ext[i], 3);
      else if (s == 18) bw.put(S.rle_ext[i], 7);
    }
    const unsigned n = (unsigned)segs[sidx].in_len;
    unsigned long long dyn_bits = bw.nbits + bits;
    // + empty stored block that byte-aligns the next segment: 3 header bits, pad, 00 00 FF FF
    unsigned dyn_bytes = (unsigned)((dyn_bits + 3 + 7) >> 3) + 4;
    unsigned stored_bytes = n + 5 * ((n + 65534) / 65535);
    DeflateSegOut o = so[sidx];
    o.hdr_bits = bw.nbits;
    if (dyn_bytes < stored_bytes) { o.mode = MODE_DYNAMIC; o.body_bytes = dyn_bytes; }
    else { o.mode = MODE_STORED; o.body_bytes = stored_bytes; }
    so[sidx] = o;
  }
  __syncwarp();
  for (int i = lane; i < CODE_STRIDE; i += 32) codes[(size_t)sidx * CODE_STRIDE + i] = c[i];
  for (int i = lane; i < HDR_WORDS; i += 32) hdrs[(size_t)sidx * HDR_WORDS + i] = S.hdr[i];
}

// ------------------------------------------------------------------------------------------------ scan_kernel
// Single CTA: exclusive prefix of (zlib header + body + trailer) sizes over all segments of the batch, in order.
// Segments of a chunk are contiguous, so chunk_off[] (n_chunks + 1 entries) falls out of the same scan.
__global__ void __launch_bounds__(1024) scan_kernel(const DeflateSeg* __restrict__ segs, int n_segs,
                                                    DeflateSegOut* __restrict__ so, long long* __restrict__ chunk_off,
                                                    int n_chunks, const ChunkDesc* __restrict__ chunks,
                                                    int write_index) {
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long carry, tile_total;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_segs; base += 1024) {
    int i = base + threadIdx.x;
    unsigned long long sz = 0;
    int fl = 0;
    if (i < n_segs) {
      fl = segs[i].flags;
      sz = so[i].body_bytes + ((fl & SEG_FIRST) ? 2 : 0) + ((fl & SEG_LAST) ? 6 : 0);
      if ((fl & SEG_LAST) && write_index) sz += 4ull * (unsigned)chunks[segs[i].chunk].n_seg + 16;
    }
    unsigned long long incl = warp_incl_scan(sz);
    if (lane_id() == 31) wsum[warp_id()] = incl;
    __syncthreads();
    if (warp_id() == 0) {
      unsigned long long t = wsum[lane_id()];
      unsigned long long ti = warp_incl_scan(t);
      wsum[lane_id()] = ti - t;
      if (lane_id() == 31) tile_total = ti;
    }
    __syncthreads();
    unsigned long long excl = carry + wsum[warp_id()] + incl - sz;
    if (i < n_segs) {
      so[i].out_off = (long long)excl;
      if (fl & SEG_FIRST) chunk_off[segs[i].chunk] = (long long)excl;
      if (i == n_segs - 1) chunk_off[n_chunks] = (long long)(excl + sz);
    }
    __syncthreads();
    if (threadIdx.x == 0) carry += tile_total;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ encode_kernel
static const int ENC_THREADS = 256;
static const int ENC_PER = 8;                              // token elements per thread per tile
static const int ENC_TILE = ENC_THREADS * ENC_PER;
static const int ENC_STAGE_WORDS = ENC_TILE + 64;          // >= tile * 28 bits / 32 + header slack

__device__ __forceinline__ void stage_or(unsigned* stage, unsigned bitpos, unsigned v, unsigned n) {
  if (!n) return;
  unsigned i = bitpos >> 5, sh = bitpos & 31;
  atomicOr(&stage[i], v << sh);
  if (sh + n > 32) atomicOr(&stage[i + 1], v >> (32 - sh));
}

// Flush stage words [0, nwords) to global memory starting at the 4-byte aligned address `gw`; bytes before
// `lo_byte` (absolute address) or at/after `hi_byte` are not touched (they belong to neighbouring segments).
__device__ __forceinline__ void stage_flush(const unsigned* stage, unsigned nwords, unsigned char* gw,
                                            const unsigned char* lo_byte, const unsigned char* hi_byte) {
  for (unsigned i = threadIdx.x; i < nwords; i += blockDim.x) {
    unsigned char* a = gw + 4 * (size_t)i;
    unsigned v = stage[i];
    if (a >= lo_byte && a + 4 <= hi_byte) *(unsigned*)a = v;
    else
      for (int b = 0; b < 4; b++)
        if (a + b >= lo_byte && a + b < hi_byte) a[b] = (unsigned char)(v >> (8 * b));
  }
}

__global__ void __launch_bounds__(ENC_THREADS) encode_kernel(const unsigned char* __restrict__ tbuf,
                                                             const DeflateSeg* __restrict__ segs, int n_segs,
                                                             const unsigned short* __restrict__ tokens,
                                                             const unsigned* __restrict__ codes,
                                                             const unsigned* __restrict__ hdrs,
                                                             const DeflateSegOut* __restrict__ so,
                                                             const unsigned* __restrict__ chunk_adler,
                                                             unsigned char* __restrict__ dst,
                                                             const ChunkDesc* __restrict__ chunks, int write_index) {
  __shared__ unsigned stage[ENC_STAGE_WORDS];
  __shared__ unsigned code[CODE_STRIDE];
  __shared__ unsigned wtot[ENC_THREADS / 32];
  __shared__ unsigned tile_bits;
  const int sidx = blockIdx.x;
  if (sidx >= n_segs) return;
  const DeflateSeg sg = segs[sidx];
  const DeflateSegOut o = so[sidx];
  const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  unsigned char* out = dst + o.out_off;
  const unsigned total = o.body_bytes + ((sg.flags & SEG_FIRST) ? 2 : 0) + ((sg.flags & SEG_LAST) ? 6 : 0);
  unsigned char* const out_end = out + total;   // end of the zlib stream part written by this CTA
  if (sg.flags & SEG_FIRST) {
    if (tid == 0) { out[0] = 0x78; out[1] = 0x9c; }
    out += 2;
  }
  unsigned char* body_end = out + o.body_bytes;
  if ((sg.flags & SEG_LAST) && tid < 6) {
    // empty final fixed-Huffman block (03 00) + adler32 of the whole transformed chunk, big-endian
    unsigned a = chunk_adler[sg.chunk];
    unsigned char t[6] = {0x03, 0x00, (unsigned char)(a >> 24), (unsigned char)(a >> 16), (unsigned char)(a >> 8), (unsigned char)a};
    body_end[tid] = t[tid];
  }
  if ((sg.flags & SEG_LAST) && write_index) {
    // segment index after the zlib stream (zlib.decompress ignores trailing bytes): k compressed segment
    // lengths, then {segment bytes, k, magic "MTSB", sum of the lengths}, all little-endian u32
    const ChunkDesc cd = chunks[sg.chunk];
    unsigned char* ix = body_end + 6;
    const unsigned k = (unsigned)cd.n_seg;
    unsigned sum = 0;
    for (unsigned j = tid; j < k; j += blockDim.x) {
      unsigned v = so[cd.first_seg + j].body_bytes;
      for (int b = 0; b < 4; b++) ix[4 * j + b] = (unsigned char)(v >> (8 * b));
    }
    if (tid == 0) {
      for (unsigned j = 0; j < k; j++) sum += so[cd.first_seg + j].body_bytes;
      unsigned tail[4] = {(unsigned)cd.pad_, k, 0x4253544Du, sum};
      for (int q = 0; q < 4; q++)
        for (int b = 0; b < 4; b++) ix[4 * k + 4 * q + b] = (unsigned char)(tail[q] >> (8 * b));
    }
  }
  const unsigned char* in = tbuf + sg.in_off;
  if (o.mode == MODE_STORED) {
    unsigned n = (unsigned)sg.in_len;
    for (unsigned b0 = 0, k = 0; b0 < n; b0 += 65535, k++) {
      unsigned len = min(65535u, n - b0);
      unsigned char* q = out + b0 + 5 * (size_t)k;
      if (tid == 0) { q[0] = 0; q[1] = (unsigned char)len; q[2] = (unsigned char)(len >> 8); q[3] = (unsigned char)~len; q[4] = (unsigned char)(~len >> 8); }
      for (unsigned i = tid; i < len; i += blockDim.x) q[5 + i] = in[b0 + i];
    }
    return;
  }
  // ---- dynamic block
  for (unsigned i = tid; i < CODE_STRIDE; i += blockDim.x) code[i] = codes[(size_t)sidx * CODE_STRIDE + i];
  for (unsigned i = tid; i < ENC_STAGE_WORDS; i += blockDim.x) stage[i] = 0;
  __syncthreads();
  unsigned char* gw = (unsigned char*)((uintptr_t)out & ~(uintptr_t)3);   // global address of stage word 0
  unsigned cur = 8 * (unsigned)((uintptr_t)out & 3);                       // bit cursor inside the stage
  // header
  {
    const unsigned* h = hdrs + (size_t)sidx * HDR_WORDS;
    const unsigned hb = o.hdr_bits;
    for (unsigned j = tid; j * 32 < hb; j += blockDim.x) stage_or(stage, cur + 32 * j, h[j], min(32u, hb - 32 * j));
    __syncthreads();
    cur += hb;
    unsigned nw = cur >> 5;
    stage_flush(stage, nw, gw, out, out_end);
    __syncthreads();
    unsigned keep = stage[nw];
    __syncthreads();
    for (unsigned i = tid; i <= nw; i += blockDim.x) stage[i] = 0;
    __syncthreads();
    if (tid == 0) stage[0] = keep;
    gw += 4 * (size_t)nw;
    cur &= 31;
    __syncthreads();
  }
  const unsigned short* tok = tokens + sg.tok_off;
  const unsigned ntok = o.n_tok;
  for (unsigned base = 0; base < ntok; base += ENC_TILE) {
    unsigned v[ENC_PER], nb[ENC_PER], mine = 0;
    unsigned i0 = base + tid * ENC_PER;
    unsigned prev_el = (i0 > 0 && i0 <= ntok) ? tok[i0 - 1] : 0;
    for (int j = 0; j < ENC_PER; j++) {
      unsigned i = i0 + j;
      v[j] = 0; nb[j] = 0;
      if (i < ntok) {
        unsigned e = tok[i];
        if (prev_el & 0x8000u) {            // distance element (follows a length element)
          unsigned sym, xb, ev;
          dist_symbol(e + 1, sym, xb, ev);
          unsigned c = code[288 + sym];
          v[j] = (c & 0xffff) | (ev << (c >> 16));
          nb[j] = (c >> 16) + xb;
          prev_el = 0;
        } else if (e & 0x8000u) {           // length element
          unsigned sym, xb, ev;
          len_symbol(e & 0x1ff, sym, xb, ev);
          unsigned c = code[sym];
          v[j] = (c & 0xffff) | (ev << (c >> 16));
          nb[j] = (c >> 16) + xb;
          prev_el = e;
        } else {                            // literal
          unsigned c = code[e];
          v[j] = c & 0xffff;
          nb[j] = c >> 16;
          prev_el = e;
        }
        mine += nb[j];
      }
    }
    unsigned incl = warp_incl_scan(mine);
    if (lane == 31) wtot[wid] = incl;
    __syncthreads();
    if (tid == 0) {
      unsigned run = 0;
      for (int w = 0; w < ENC_THREADS / 32; w++) { unsigned t = wtot[w]; wtot[w] = run; run += t; }
      tile_bits = run;
    }
    __syncthreads();
    unsigned bp = cur + wtot[wid] + incl - mine;
    for (int j = 0; j < ENC_PER; j++) { stage_or(stage, bp, v[j], nb[j]); bp += nb[j]; }
    __syncthreads();
    cur += tile_bits;
    unsigned nw = cur >> 5;
    stage_flush(stage, nw, gw, out, out_end);
    __syncthreads();
    unsigned keep = stage[nw];
    __syncthreads();
    for (unsigned i = tid; i <= nw + 1; i += blockDim.x) stage[i] = 0;
    __syncthreads();
    if (tid == 0) stage[0] = keep;
    gw += 4 * (size_t)nw;
    cur &= 31;
    __syncthreads();
  }
  // end-of-block, then the empty stored block: 3 zero bits, pad to a byte boundary, 00 00 FF FF
  if (tid == 0) {
    unsigned c = code[256];
    stage_or(stage, cur, c & 0xffff, c >> 16);
    unsigned e = cur + (c >> 16) + 3;
    e = (e + 7) & ~7u;
    stage_or(stage, e + 16, 0xffffu, 16);
    tile_bits = e + 32;
  }
  __syncthreads();
  unsigned endbits = tile_bits;
  stage_flush(stage, (endbits + 31) >> 5, gw, out, body_end);
}

}  // namespace mts
