// inflate_par.cuh — block-parallel decoding of plain zlib streams (e.g. chunks written by the reference Writer).
//
// A zlib stream has no index: its DEFLATE blocks (~300 dynamic-Huffman blocks per 23 MB chunk at zlib level 6) can only
// be located by decoding.  One warp per stream (inflate.cuh) therefore leaves the GPU idle.  This path finds the blocks
// speculatively and decodes them all at once:
//   1 par_find_kernel      every bit offset is tested for a plausible dynamic-block header (BTYPE, HLIT/HDIST ranges,
//                          complete code-length code); survivors (~8e-4 of the offsets) are appended to a list
//   2 par_validate_kernel  thread per survivor: full header parse; both Huffman codes must be complete and have an
//                          end-of-block code (what zlib's deflate always emits) -> candidate blocks
//   (host)                 sort the candidates (stream, bit); a block's bit range ends at the next candidate
//   3 par_block_kernel     WARP per candidate: tables in shared memory, the 32 lanes decode 32 sub-chunks of the block
//                          in parallel (self-synchronising Huffman decoding) into 32-bit tokens (literal / length+distance)
//   4 par_lz_kernel        CTA per stream: follows the chain of blocks (each must start where the previous one ended) and
//                          turns their tokens into bytes, a tile of 256 tokens at a time, staged in shared memory
//   5 (inflate.cuh)        one warp per stream decodes the tail serially (usually nothing or the final small block)
// Anything unexpected (no chain, overflow of a list, bad data) falls back to the serial decoder, which also produces
// the error status; this path never decides that a stream is corrupt by itself.
#pragma once
#include "common.cuh"
#include "inflate.cuh"

namespace mts {

struct ParStream {       // one DEFLATE stream: a whole zlib stream, or one indexed segment of a GPU-written chunk
  long long in_off;      // byte offset in the compressed buffer
  long long out_off;     // byte offset of its output in the transformed buffer
  int in_len, out_len;
  unsigned first_bit;    // bit offset of its first block header (16 after a zlib header, 0 for a segment)
  unsigned pad_;
};

struct ParBlk {          // a candidate block in stream order (host-sorted), filled in by par_block_kernel
  unsigned stream, bit;  // owning stream, bit offset of the block header
  unsigned limit;        // bit offset of the next candidate of the stream (or the end of the stream)
  unsigned end_bit;      // OUT: bit offset just after the end-of-block code
  unsigned n_tok;        // OUT: tokens decoded
  unsigned out_len;      // OUT: bytes they produce
  unsigned flags;        // OUT: 1 = decoded cleanly, 2 = BFINAL
  unsigned pad_;
  long long tok_off;     // OUT: first slot in the token buffer
};

struct ParRes {          // per stream, written by par_lz_kernel
  unsigned tail_bit;     // bit offset where the serial decoder has to resume
  unsigned tail_out;     // output bytes produced so far
  unsigned n_done;       // blocks resolved
  unsigned flags;        // 1 = the last resolved block was final, 2 = inconsistent tokens (decode the stream serially),
                         // 4 = `adler` is the adler32 of out[0, tail_out) (par_lz_kernel; not the cells path)
  unsigned adler;
};

// Decoding tables of one block, in SHARED memory, owned by the warp that decodes the block.  Entries are 16 bits:
//   literal/length (PAR_LBITS index bits): code length | kind << 4 | (literal byte or length symbol - 257) << 6
//   distance       (PAR_DBITS index bits): code length | symbol << 4
// A zero entry means "longer code".  Those are resolved with one word per code length l >= table bits:
//   limit | oend << 17,  limit = one past the last l-bit code, left-justified to 16 bits,
//                        oend  = number of symbols with codes of length <= l;
// the symbol is sorted[oend - 1 - ((limit - 1 - code16) >> (16 - l))] for the first l whose limit exceeds code16.
static const int PAR_LBITS = 10, PAR_DBITS = 8;
struct BlkTabs {
  unsigned short ltab[1 << PAR_LBITS];
  unsigned short dtab[1 << PAR_DBITS];
  unsigned llong[16 - PAR_LBITS], dlong[16 - PAR_DBITS];
  unsigned short lsorted[288], dsorted[32];
  unsigned short code[288];        // scratch of the table build: first table index of every short symbol
  unsigned char lens[320];
};

// A token is 32 bits: a literal byte, or 0x80000000 | length << 16 | (distance - 1).

// ---------------------------------------------------------------------------------------------- per-thread bit reader
struct TBits {
  const unsigned* w;     // 4-byte aligned base of the stream
  unsigned sh;           // 8 * (stream address & 3)
  unsigned kmax;         // last readable word
  unsigned k;            // index of the raw word held in `ahead`
  unsigned raw, ahead;   // raw = w[k-1]; ahead = w[k], loaded one refill early so that its latency is hidden
  unsigned lo, hi;       // 64 stream bits
  unsigned pos;          // cursor inside (lo, hi), < 32 after refill
  unsigned base_bit;     // stream bit offset of bit 0 of lo
  __device__ __forceinline__ unsigned next_word() {
    const unsigned v = __funnelshift_r(raw, ahead, sh);
    raw = ahead; k++;
    ahead = w[min(k, kmax)];
    return v;
  }
  __device__ __forceinline__ void init(const unsigned char* in, unsigned in_len, unsigned bit) {
    const unsigned mis = (unsigned)((uintptr_t)in & 3);
    w = (const unsigned*)(in - mis);
    sh = mis * 8;
    kmax = (mis + max(in_len, 1u) - 1) >> 2;
    const unsigned word = bit >> 5;           // stream word that holds `bit`
    k = word;
    raw = w[min(k, kmax)]; k++;
    ahead = w[min(k, kmax)];
    lo = next_word();
    hi = next_word();
    pos = bit & 31;
    base_bit = word << 5;
  }
  // no branch: the lanes of a warp read different streams and must not drift apart (ncu: with `if (pos >= 32)` as a
  // branch the refill was 57 % of the warp instructions of the symbol loops)
  __device__ __forceinline__ void refill() {
    const bool need = pos >= 32;
    const unsigned nw = __funnelshift_r(raw, ahead, sh);
    lo = need ? hi : lo;
    hi = need ? nw : hi;
    raw = need ? ahead : raw;
    pos -= need ? 32u : 0u;
    base_bit += need ? 32u : 0u;
    k += need ? 1u : 0u;
    if (need) ahead = w[min(k, kmax)];
  }
  __device__ __forceinline__ unsigned window() const { return __funnelshift_r(lo, hi, pos); }
  __device__ __forceinline__ void drop(unsigned n) { pos += n; }
  __device__ __forceinline__ unsigned get(unsigned n) { refill(); unsigned v = window() & ((1u << n) - 1); pos += n; return v; }
  __device__ __forceinline__ unsigned bit_pos() const { return base_bit + pos; }
};

// The same reader for the symbol loops that all 32 lanes of a warp run on 32 different streams.  There the loads of
// TBits hurt: a refill selects from `ahead` in every iteration, and the scoreboard of that register is the WARP's, so
// every iteration waited for the global load some lane had issued one iteration earlier (ncu: half of the stall samples
// of the token kernel on that select).  Here the stream words come from a per-lane ring of four 16-byte slots in shared
// memory that cp.async fills two slots ahead: no load with a register destination touches global memory.
// `ring` is the lane's slot 0; slot s lies at ring + 32 * s (slot-major over the lanes of the warp).
struct SBits {
  const uint4* g;        // 16-byte aligned base of the stream
  const unsigned* rw;    // the lane's ring as words: slot s, word i at rw[128 * s + i]
  uint4* ring;
  unsigned bmax;         // last readable 16-byte block
  unsigned k;            // index of the stream word (from g) held in `ahead`
  unsigned ahead, lo, hi;
  unsigned pos;          // cursor inside (lo, hi), < 32 after refill
  unsigned base_bit;     // stream bit offset of bit 0 of lo
  __device__ __forceinline__ void fetch(unsigned b) { cp_async16(ring + 32 * (b & 3), g + min(b, bmax)); cp_async_commit(); }
  __device__ __forceinline__ unsigned word(unsigned i) const { return rw[128 * ((i >> 2) & 3) + (i & 3)]; }
  __device__ __forceinline__ void init(const unsigned char* in, unsigned in_len, unsigned bit, uint4* lane_ring) {
    const unsigned mis = (unsigned)((uintptr_t)in & 15);
    g = (const uint4*)(in - mis);
    ring = lane_ring; rw = (const unsigned*)lane_ring;
    bmax = (mis + max(in_len, 1u) - 1) >> 4;
    const unsigned ab = bit + 8 * mis, w0 = ab >> 5, b0 = w0 >> 2;
    cp_async_wait<0>();          // copies a previous pass left in flight must not land on top of the new ones
    fetch(b0); fetch(b0 + 1); fetch(b0 + 2);
    cp_async_wait<0>();
    lo = word(w0); hi = word(w0 + 1); ahead = word(w0 + 2);
    k = w0 + 2;
    // blocks up to (k >> 2) + 2 must be under way
    if ((k >> 2) != b0) fetch(b0 + 3);
    pos = ab & 31;
    base_bit = (w0 << 5) - 8 * mis;
  }
  __device__ __forceinline__ void refill() {
    const bool need = pos >= 32;
    lo = need ? hi : lo;
    hi = need ? ahead : hi;
    pos -= need ? 32u : 0u;
    base_bit += need ? 32u : 0u;
    if (need) {
      k++;
      if ((k & 3) == 0) { fetch((k >> 2) + 2); cp_async_wait<1>(); }     // entering a block: the next one is complete
      ahead = word(k);
    }
  }
  __device__ __forceinline__ unsigned window() const { return __funnelshift_r(lo, hi, pos); }
  __device__ __forceinline__ void drop(unsigned n) { pos += n; }
  __device__ __forceinline__ unsigned bit_pos() const { return base_bit + pos; }
  __device__ __forceinline__ void finish() { cp_async_wait<0>(); }      // nothing in flight when the ring is given up
};

// ---------------------------------------------------------------------------------------------- header parse
// Canonical decode of one symbol of the 19-symbol code-length code from a 32-bit window (bit-serial, <= 7 bits).
__device__ __forceinline__ int par_cl_decode(unsigned win, const unsigned char* count, const unsigned char* sorted,
                                             unsigned& nbits) {
  unsigned code = 0, first = 0, index = 0;
  for (unsigned l = 1; l <= 7; l++) {
    code |= (win >> (l - 1)) & 1;
    const unsigned c = count[l];
    if (code - first < c) { nbits = l; return sorted[index + (code - first)]; }
    index += c;
    first = (first + c) << 1;
    code <<= 1;
  }
  return -1;
}

// Does a dynamic block header that zlib's deflate could have written start at the reader's position?  The code
// lengths are not stored: a complete literal/length code and a complete (or single-symbol) distance code are checked
// through running Kraft sums (units of 2^-15) — a random bit string over-subscribes one of the codes within a few
// dozen lengths, so false survivors are dropped long before the end of the header — and an end-of-block code must exist.
__device__ bool par_check_header(TBits& br) {
  br.get(1);
  if (br.get(2) != 2) return false;
  const int nl = (int)br.get(5) + 257;
  const int nd = (int)br.get(5) + 1;
  const int ncl = (int)br.get(4) + 4;
  if (nl > 286 || nd > 30) return false;
  const unsigned char order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  unsigned char cl[19], count[8], sorted[19];
  for (int i = 0; i < 19; i++) cl[i] = 0;
  for (int i = 0; i < 8; i++) count[i] = 0;
  for (int i = 0; i < ncl; i++) cl[order[i]] = (unsigned char)br.get(3);
  int left = 1;
  for (int i = 0; i < 19; i++) count[cl[i]]++;
  count[0] = 0;
  unsigned char offs[8];
  offs[1] = 0;
  for (int l = 1; l <= 7; l++) {
    left = (left << 1) - count[l];
    if (left < 0) return false;
    if (l < 7) offs[l + 1] = (unsigned char)(offs[l] + count[l]);
  }
  if (left != 0) return false;
  for (int i = 0; i < 19; i++) if (cl[i]) sorted[offs[cl[i]]++] = (unsigned char)i;
  int idx = 0;
  unsigned prev = 0, kl = 0, kd = 0, ndist = 0;
  bool eob = false;
  while (idx < nl + nd) {
    br.refill();
    unsigned nb;
    const int sym = par_cl_decode(br.window(), count, sorted, nb);
    if (sym < 0) return false;
    br.drop(nb);
    unsigned rep = 1, val = (unsigned)sym;
    if (sym == 16) { if (idx == 0) return false; val = prev; rep = 3 + br.get(2); }
    else if (sym == 17) { val = 0; rep = 3 + br.get(3); }
    else if (sym == 18) { val = 0; rep = 11 + br.get(7); }
    if (idx + (int)rep > nl + nd) return false;
    if (val) {
      const unsigned n_ll = (unsigned)min(max(nl - idx, 0), (int)rep);      // how many of the run are literal/length codes
      kl += n_ll * (32768u >> val);
      kd += (rep - n_ll) * (32768u >> val);
      ndist += rep - n_ll;
      if (kl > 32768u || kd > 32768u) return false;
      if (idx <= 256 && 256 < idx + (int)rep) eob = true;
    }
    idx += (int)rep;
    prev = val;
  }
  if (!eob || kl != 32768u) return false;
  if (kd != 32768u && !(ndist <= 1 && kd <= 16384u)) return false;   // zlib emits >= 2 distance codes; tolerate 0/1
  return true;
}

// Tables of one code, built by the whole warp from lens[0..n) (shared memory).  Lane 0 assigns the canonical codes
// (serial, n <= 288); the replication of the short codes over the fast table is spread over the lanes.
// KIND 1 literal/length, 2 distance.  Returns false for an over-subscribed code (warp-uniform).
template <int KIND>
__device__ bool blk_build(const unsigned char* lens, int n, unsigned short* tab, int tb, unsigned* longw,
                          unsigned short* sorted, unsigned short* code_s, unsigned lane) {
  int ok = 1;
  for (int k = (int)lane; k < (1 << tb); k += 32) tab[k] = 0;
  if (lane == 0) {
    unsigned count[16], first[16], offs[16];
    for (int i = 0; i < 16; i++) count[i] = 0;
    for (int i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    unsigned code = 0, o = 0;
    int left = 1;
    for (int l = 1; l <= 15; l++) {
      code = (code + (l > 1 ? count[l - 1] : 0)) << 1;
      first[l] = code;
      offs[l] = o;
      o += count[l];
      left = (left << 1) - (int)count[l];
      if (left < 0) ok = 0;
      if (l >= tb) longw[l - tb] = ((code + count[l]) << (16 - l)) | (o << 17);
    }
    if (ok)
      for (int s = 0; s < n; s++) {
        const unsigned l = lens[s];
        if (!l) continue;
        sorted[offs[l]++] = (unsigned short)s;
        code_s[s] = (unsigned short)(__brev(first[l]++) >> (32 - l));      // bit-reversed: the table is indexed LSB first
      }
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  __syncwarp();
  if (!ok) return false;
  for (int s = (int)lane; s < n; s += 32) {
    const unsigned l = lens[s];
    if (!l || (int)l > tb) continue;
    unsigned e;
    if (KIND == 1) e = l | ((s < 256 ? (unsigned)K_LIT : s == 256 ? (unsigned)K_EOB : s < 286 ? (unsigned)K_LEN : (unsigned)K_BAD) << 4) |
                       ((unsigned)(s < 256 ? s : s > 256 ? s - 257 : 0) << 6);
    else e = l | ((unsigned)s << 4);
    for (unsigned k = code_s[s]; k < (1u << tb); k += 1u << l) tab[k] = (unsigned short)e;
  }
  __syncwarp();
  return true;
}

// Decode of a code longer than the fast table (TB bits): first length whose limit exceeds the left-justified code.
template <int TB>
__device__ __forceinline__ bool blk_long(unsigned win, const unsigned* longw, const unsigned short* sorted,
                                         unsigned& sym, unsigned& nbits) {
  const unsigned c16 = __brev(win) >> 16;
  unsigned word = 0, prev = 0, l = 0;
  unsigned below = longw[0];                                // length TB: only its symbol count matters
#pragma unroll
  for (int j = 1; j <= 15 - TB; j++) {
    const unsigned wj = longw[j];
    if (!l && c16 < (wj & 0x1ffffu)) { word = wj; prev = below; l = (unsigned)(TB + j); }
    below = wj;
  }
  if (!l) return false;
  const unsigned back = ((word & 0x1ffffu) - c16 - 1) >> (16 - l);      // codes between this one and the last of length l
  const unsigned idx = (word >> 17) - 1 - back;
  if ((int)idx < (int)(prev >> 17)) return false;           // a hole of an incomplete code
  sym = sorted[idx];
  nbits = l;
  return true;
}

// base | extra bits << 12 of length symbol 257 + i; base | extra bits << 16 of distance symbol i (0 = invalid)
__device__ __forceinline__ unsigned par_len_info(unsigned i) {
  if (i < 8) return 3 + i;
  if (i == 28) return 258;
  if (i > 28) return 0;
  const unsigned nb = (i - 4) >> 2;
  return (3 + ((4 + (i & 3)) << nb)) | (nb << 12);
}
__device__ __forceinline__ unsigned par_dist_info(unsigned i) {
  if (i < 4) return i + 1;
  if (i >= 30) return 0;
  const unsigned nb = (i >> 1) - 1;
  return (1 + ((2 + (i & 1)) << nb)) | (nb << 16);
}

// ---------------------------------------------------------------------------------------------- kernels
// 1. Survivors of the cheap header test: one thread per aligned stream word (32 bit offsets).  Phase 1 tests the fixed
//    fields of all 32 offsets at once with word-parallel logic (BTYPE = dynamic: bit 1 clear, bit 2 set; HLIT <= 29 and
//    HDIST <= 29: their upper four bits not all ones); phase 2 visits the ~22 % passing offsets and checks that the
//    code-length code is complete: its 3-bit lengths are summed 3 at a time through a 512-entry table of 2^(7-len),
//    leaving as soon as the sum exceeds 1.
static const unsigned PAR_FIND_WORDS = 8;     // stream words per thread (amortises the table set-up of the CTA)
__global__ void __launch_bounds__(256) par_find_kernel(const unsigned char* __restrict__ comp,
                                                       const ParStream* __restrict__ streams,
                                                       unsigned long long* __restrict__ surv, unsigned cap,
                                                       unsigned* __restrict__ counters) {
  __shared__ unsigned char k9[512];
  for (unsigned i = threadIdx.x; i < 512; i += blockDim.x) {
    unsigned a = i & 7, b2 = (i >> 3) & 7, c2 = i >> 6;
    k9[i] = (unsigned char)((a ? 128u >> a : 0) + (b2 ? 128u >> b2 : 0) + (c2 ? 128u >> c2 : 0));
  }
  __syncthreads();
  const ParStream st = streams[blockIdx.y];
  const unsigned char* in = comp + st.in_off;
  const unsigned mis = (unsigned)((uintptr_t)in & 3);
  const unsigned* w = (const unsigned*)(in - mis);
  if (st.in_len < 16) return;
  const unsigned kmax = (mis + (unsigned)st.in_len - 1) >> 2;
  // valid header bits: [16, 8 * (in_len - 12)] (a dynamic header is at least 17 + 12 bits + two codes: the zlib trailer
  // and the shortest block are ignored)
  const long long first_ok = 16, last_ok = 8ll * ((long long)st.in_len - 12) + 7;
  for (unsigned rep = 0; rep < PAR_FIND_WORDS; rep++) {
    const unsigned j = (blockIdx.x * PAR_FIND_WORDS + rep) * blockDim.x + threadIdx.x;      // aligned word index
    if (j > kmax) break;
    const long long bit0 = 32ll * j - 8ll * mis;                      // stream bit of this word's bit 0 (may be negative)
    if (bit0 + 31 < first_ok || bit0 > last_ok) continue;
    unsigned x[5];
#pragma unroll
    for (unsigned i = 0; i < 5; i++) x[i] = w[min(j + i, kmax)];
    const unsigned long long X = ((unsigned long long)x[1] << 32) | x[0];
    unsigned mask = (unsigned)(~(X >> 1) & (X >> 2) & ~((X >> 4) & (X >> 5) & (X >> 6) & (X >> 7)) &
                               ~((X >> 9) & (X >> 10) & (X >> 11) & (X >> 12)));
    if (bit0 < first_ok) mask &= 0xffffffffu << (unsigned)(first_ok - bit0);
    if (bit0 + 31 > last_ok) mask &= 0xffffffffu >> (unsigned)(bit0 + 31 - last_ok);
    while (mask) {
      const unsigned s = (unsigned)__ffs((int)mask) - 1;
      mask &= mask - 1;
      const unsigned t0 = __funnelshift_r(x[0], x[1], s), t1 = __funnelshift_r(x[1], x[2], s),
                     t2 = __funnelshift_r(x[2], x[3], s);
      const unsigned ncl = ((t0 >> 13) & 15) + 4;
      // stream bits [17, 17 + 3 * ncl) = the code-length code lengths (<= 57 bits)
      unsigned long long f = ((((unsigned long long)t1 << 32) | t0) >> 17) | ((unsigned long long)t2 << 47);
      f &= (1ull << (3 * ncl)) - 1;
      const unsigned lo = (unsigned)f, hi = (unsigned)(f >> 32);
      unsigned kraft = k9[lo & 511] + k9[(lo >> 9) & 511] + k9[(lo >> 18) & 511] + k9[((lo >> 27) | (hi << 5)) & 511];
      if (kraft > 128) continue;
      kraft += k9[(hi >> 4) & 511] + k9[(hi >> 13) & 511] + k9[(hi >> 22) & 511];
      if (kraft != 128) continue;
      const unsigned at = atomicAdd(&counters[0], 1u);
      if (at < cap) surv[at] = ((unsigned long long)blockIdx.y << 32) | (unsigned)(bit0 + s);
    }
  }
}

// 2. Full header validation: the survivors (their number is read from the device counter, so the host does not have
//    to wait for the search) are spread over a fixed grid; the bit offset of a valid one is appended to the bucket of
//    its stream (bcap slots per stream; a stream with more candidates is left to the serial decoder).
__global__ void __launch_bounds__(128) par_validate_kernel(const unsigned char* __restrict__ comp,
                                                           const ParStream* __restrict__ streams,
                                                           const unsigned long long* __restrict__ surv, unsigned surv_cap,
                                                           unsigned* __restrict__ keys, unsigned bcap,
                                                           unsigned* __restrict__ bcount,
                                                           const unsigned* __restrict__ counters) {
  const unsigned n_surv = min(counters[0], surv_cap);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_surv; i += gridDim.x * blockDim.x) {
    const unsigned long long sv = surv[i];
    const unsigned sidx = (unsigned)(sv >> 32), bit = (unsigned)sv;
    const ParStream st = streams[sidx];
    TBits br;
    br.init(comp + st.in_off, (unsigned)st.in_len, bit);
    if (!par_check_header(br)) continue;
    if (br.bit_pos() > (unsigned)st.in_len * 8) continue;
    const unsigned at = atomicAdd(&bcount[sidx], 1u);
    if (at < bcap) keys[(size_t)sidx * bcap + at] = bit;
  }
}

// 2b. One warp per stream: rank-sort the bucket (a few hundred bit offsets) and write the block descriptors in stream
//     order, each block's bit range ending where the next candidate begins; unused slots get the sentinel bit ~0.
__global__ void __launch_bounds__(32) par_sort_kernel(const ParStream* __restrict__ streams,
                                                      const unsigned* __restrict__ keys, unsigned bcap,
                                                      const unsigned* __restrict__ bcount, const unsigned* __restrict__ counters,
                                                      unsigned surv_cap, ParBlk* __restrict__ blks) {
  const unsigned sidx = blockIdx.x, lane = threadIdx.x;
  const ParStream st = streams[sidx];
  const unsigned* k = keys + (size_t)sidx * bcap;
  ParBlk* out = blks + (size_t)sidx * bcap;
  unsigned n = bcount[sidx];
  if (n > bcap || counters[0] > surv_cap) n = 0;              // a list overflowed: candidates may be missing, do not chain
  for (unsigned i = lane; i < bcap; i += 32) {
    ParBlk b;
    b.stream = sidx; b.bit = 0xffffffffu; b.limit = 0; b.end_bit = 0; b.n_tok = 0; b.out_len = 0; b.flags = 0; b.pad_ = 0; b.tok_off = 0;
    if (i >= n) out[i] = b;
  }
  for (unsigned i = lane; i < n; i += 32) {
    const unsigned key = k[i];
    unsigned rank = 0, next = (unsigned)st.in_len * 8u;        // next larger key = where this block's bit range ends
    for (unsigned j = 0; j < n; j++) {
      const unsigned kj = k[j];
      rank += kj < key;
      if (kj > key) next = min(next, kj);
    }
    ParBlk b;
    b.stream = sidx; b.bit = key; b.limit = next; b.end_bit = 0; b.n_tok = 0; b.out_len = 0; b.flags = 0; b.pad_ = 0; b.tok_off = 0;
    out[rank] = b;
  }
}

// 3. One WARP per candidate block.  The warp parses the header and builds the tables in shared memory, then its 32
//    lanes decode 32 consecutive sub-chunks of the block's bit range in parallel.  Only lane 0 knows where its first
//    symbol starts; the others start at a guess (the sub-chunk boundary, usually in the middle of a symbol).  Huffman
//    streams re-synchronise within a few symbols, so after the first pass most lanes END on a true symbol boundary;
//    every lane then restarts from where its predecessor ended, until nothing changes (typically two passes; lane k is
//    certainly right after k passes).  A last pass writes the tokens at the lane's offset (prefix sum of the counts).
struct SpanRes { unsigned end, ntok, nout, flags; };   // flags: 1 end-of-block seen, 2 error, 4 nothing to do (after the EOB)

// The symbol loop is written WITHOUT data-dependent branches on the common paths (a literal goes through the distance
// lookup too and discards it): the 32 lanes decode 32 different bit streams, and any branch on the symbol kind would
// let them drift apart until every lane runs alone (measured: 2.5 active lanes per instruction with a branchy loop).
// Called by ALL lanes of the warp (run = false: nothing to decode, r is left alone).  Every iteration starts with a
// warp vote, which makes the lanes reconverge once per symbol; finished lanes idle until the last one is done.
template <bool EMIT>
__device__ __forceinline__ void blk_span(bool run, const unsigned char* in, unsigned in_len, unsigned start,
                                         unsigned bound, const BlkTabs& T, const unsigned short* lenx,
                                         const unsigned* distx, unsigned* tok, SpanRes& r, uint4* lane_ring,
                                         unsigned cap = 0xffffffffu) {
  const unsigned in_bits = in_len * 8;
  SBits br;
  br.init(in, in_len, run ? start : 0u, lane_ring);
  unsigned ntok = 0, nout = 0, flags = 0;
  bool active = run;
  while (__any_sync(0xffffffffu, active)) {
    if (!active) continue;
    br.refill();
    if (br.bit_pos() >= bound) { active = false; continue; }
    const unsigned win = br.window();
    const unsigned e = T.ltab[win & ((1u << PAR_LBITS) - 1)];
    unsigned cl = e & 15, kind = (e >> 4) & 3, val = e >> 6;
    if (cl == 0) {                                                  // rare: a code longer than the table
      unsigned sym;
      if (!blk_long<PAR_LBITS>(win, T.llong, T.lsorted, sym, cl)) { flags = 2; active = false; continue; }
      kind = sym < 256 ? (unsigned)K_LIT : sym == 256 ? (unsigned)K_EOB : sym < 286 ? (unsigned)K_LEN : (unsigned)K_BAD;
      val = sym < 256 ? sym : sym > 256 ? sym - 257 : 0;
    }
    if (kind >= (unsigned)K_EOB) {                                  // once per block, or an error
      br.drop(cl);
      flags = kind == K_EOB ? 1u : 2u;
      active = false;
      continue;
    }
    const bool isl = kind == K_LEN;
    const unsigned lx = lenx[isl ? val : 0u];
    const unsigned xb = isl ? lx >> 12 : 0u;
    const unsigned len = (lx & 0xfffu) + ((win >> cl) & ((1u << xb) - 1));
    br.drop(cl + xb);
    br.refill();
    const unsigned win2 = br.window();
    const unsigned e2 = T.dtab[win2 & ((1u << PAR_DBITS) - 1)];
    unsigned cl2 = e2 & 15, dsym = e2 >> 4;
    if (isl && cl2 == 0 && !blk_long<PAR_DBITS>(win2, T.dlong, T.dsorted, dsym, cl2)) { flags = 2; active = false; continue; }
    const unsigned dx = distx[dsym & 31];
    if (isl && dx == 0) { flags = 2; active = false; continue; }
    const unsigned xb2 = dx >> 16;
    const unsigned dist = (dx & 0xffffu) + ((win2 >> cl2) & ((1u << xb2) - 1));
    br.drop(isl ? cl2 + xb2 : 0u);
    if (EMIT && ntok < cap) tok[ntok] = isl ? (0x80000000u | (len << 16) | (dist - 1)) : val;
    ntok++;
    nout += isl ? len : 1u;
  }
  br.finish();
  if (!run) return;
  if (br.bit_pos() > in_bits) flags |= 2;
  r.end = br.bit_pos(); r.ntok = ntok; r.nout = nout; r.flags = flags;
}

// Dynamic block header at bit `bit`, parsed by ONE lane with a 7-bit table for the code-length code (tab7: 128 bytes of
// scratch).  On success lens[0..nl) and lens[nl..nl+nd) hold the code lengths and hdr_end the bit of the first symbol.
__device__ bool blk_parse_header(const unsigned char* in, unsigned in_len, unsigned bit, unsigned char* lens,
                                 unsigned char* tab7, int& nl, int& nd, unsigned& final_, unsigned& hdr_end) {
  TBits br;
  br.init(in, in_len, bit);
  final_ = br.get(1);
  if (br.get(2) != 2) return false;
  nl = (int)br.get(5) + 257;
  nd = (int)br.get(5) + 1;
  const int ncl = (int)br.get(4) + 4;
  if (nl > 286 || nd > 30) return false;
  // code-length code: 19 lengths of 3 bits in the order 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15, packed 4 bits each
  const unsigned long long order = 0xf1e2d3c4b5a69780ull;           // nibbles 3..18 of the order (low nibble first)
  unsigned long long cl = 0;                                         // cl nibble s = length of symbol s (s < 16)
  unsigned cl16 = 0, cl17 = 0, cl18 = 0;
  unsigned count[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < ncl; i++) {
    const unsigned v = br.get(3);
    if (i == 0) cl16 = v; else if (i == 1) cl17 = v; else if (i == 2) cl18 = v;
    else cl |= (unsigned long long)v << (4 * (unsigned)((order >> (4 * (i - 3))) & 15));
    count[v]++;
  }
  count[0] = 0;
  unsigned first[8];
  {
    unsigned code = 0;
    int left = 1;
    for (int l = 1; l <= 7; l++) {
      code = (code + (l > 1 ? count[l - 1] : 0)) << 1;
      first[l] = code;
      left = (left << 1) - (int)count[l];
      if (left < 0) return false;
    }
  }
  for (int k = 0; k < 128; k += 4) *(unsigned*)(tab7 + k) = 0;
  for (unsigned sy = 0; sy < 19; sy++) {
    const unsigned l = sy < 16 ? (unsigned)((cl >> (4 * sy)) & 15) : sy == 16 ? cl16 : sy == 17 ? cl17 : cl18;
    if (!l) continue;
    const unsigned rv = __brev(first[l]++) >> (32 - l);
    for (unsigned k = rv; k < 128; k += 1u << l) tab7[k] = (unsigned char)(l | (sy << 3));
  }
  int idx = 0;
  unsigned prev = 0;
  while (idx < nl + nd) {
    br.refill();
    const unsigned win = br.window();
    const unsigned e = tab7[win & 127];
    if (!e) return false;
    const unsigned sy = e >> 3;
    br.drop(e & 7);
    unsigned rep = 1, val = sy;
    if (sy == 16) { if (idx == 0) return false; val = prev; rep = 3 + br.get(2); }
    else if (sy == 17) { val = 0; rep = 3 + br.get(3); }
    else if (sy == 18) { val = 0; rep = 11 + br.get(7); }
    if (idx + (int)rep > nl + nd) return false;
    for (unsigned k = 0; k < rep; k++) lens[idx + k] = (unsigned char)val;
    idx += (int)rep;
    prev = val;
  }
  if (lens[256] == 0) return false;
  hdr_end = br.bit_pos();
  return true;
}

#ifndef MTS_PAR_BLK_WARPS
#define MTS_PAR_BLK_WARPS 4
#endif
static const int PAR_BLK_WARPS = MTS_PAR_BLK_WARPS;   // (tables 4.1 KB + bit-reader ring 2 KB per warp, static shared memory)
static const unsigned PAR_PREROLL_BITS = 768;       // see par_block_kernel
#ifndef MTS_PAR_PIECE_BITS
#define MTS_PAR_PIECE_BITS 8
#endif
static const unsigned PAR_PIECE_BITS = MTS_PAR_PIECE_BITS;   // input bits per token slot of the speculative token area
static const unsigned PAR_RUN_ON_BITS = 1u << 20;   // how far past the next candidate the last lane may look for the end-of-block code
__global__ void __launch_bounds__(PAR_BLK_WARPS * 32) par_block_kernel(const unsigned char* __restrict__ comp,
                                                                       const ParStream* __restrict__ streams,
                                                                       ParBlk* __restrict__ blks, unsigned n,
                                                                       unsigned* __restrict__ tokens,
                                                                       unsigned long long* __restrict__ tok_cursor,
                                                                       unsigned long long tok_capacity,
                                                                       unsigned* __restrict__ pieces) {
  __shared__ BlkTabs tabs[PAR_BLK_WARPS];
  __shared__ uint4 rings[PAR_BLK_WARPS][4 * 32];              // SBits: four 16-byte slots per lane
  __shared__ unsigned short lenx[32];
  __shared__ unsigned distx[32];
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint4* lane_ring = &rings[wid][lane];
  if (threadIdx.x < 32) { lenx[lane] = (unsigned short)par_len_info(lane); distx[lane] = par_dist_info(lane); }
  __syncthreads();
  const unsigned bi = blockIdx.x * PAR_BLK_WARPS + wid;
  if (bi >= n) return;
  BlkTabs& T = tabs[wid];
  ParBlk blk = blks[bi];
  if (blk.bit == 0xffffffffu) return;                         // unused slot of a stream's bucket
  const ParStream st = streams[blk.stream];
  const unsigned char* in = comp + st.in_off;
  const unsigned in_len = (unsigned)st.in_len, in_bits = in_len * 8;
  // ---- header (lane 0) and tables (warp)
  int nl = 0, nd = 0, hok = 0;
  unsigned fin = 0, hdr_end = 0;
  if (lane == 0) hok = blk_parse_header(in, in_len, blk.bit, T.lens, (unsigned char*)T.dtab, nl, nd, fin, hdr_end) ? 1 : 0;
  hok = __shfl_sync(0xffffffffu, hok, 0);
  nl = __shfl_sync(0xffffffffu, nl, 0);
  nd = __shfl_sync(0xffffffffu, nd, 0);
  fin = __shfl_sync(0xffffffffu, fin, 0);
  hdr_end = __shfl_sync(0xffffffffu, hdr_end, 0);
  __syncwarp();
  bool ok = hok && hdr_end < in_bits;
  ok = ok && blk_build<1>(T.lens, nl, T.ltab, PAR_LBITS, T.llong, T.lsorted, T.code, lane);
  ok = ok && blk_build<2>(T.lens + nl, nd, T.dtab, PAR_DBITS, T.dlong, T.dsorted, T.code, lane);
  SpanRes r;
  r.end = 0; r.ntok = 0; r.nout = 0; r.flags = 2;
  unsigned start = 0, total_tok = 0, total_out = 0;
  if (ok) {
    // ---- sub-chunks of the bit range [hdr_end, limit); the last lane may run on to the end-of-block code
    const unsigned limit = min(max(blk.limit, hdr_end), in_bits);
    const unsigned sc = max(64u, (limit - hdr_end + 31) / 32);
    start = hdr_end + lane * sc;
    // (the "next candidate" can be a false positive inside this block: the last lane then decodes the rest alone, and
    // the chain walk skips the false candidate because it starts before this block's end)
    const unsigned bound = lane == 31 ? (unsigned)min((unsigned long long)in_bits, (unsigned long long)limit + PAR_RUN_ON_BITS)
                                      : hdr_end + (lane + 1) * sc;
    // pre-roll: lanes 1..31 start decoding PAR_PREROLL_BITS before their sub-chunk and adopt the first symbol boundary
    // inside it as their start; by then the decoder has usually re-synchronised, so the start already equals the
    // point where the predecessor will end and the whole second pass below is skipped (it stays as the safety net)
    {
      const unsigned guess = start;
      const unsigned back = min(guess - hdr_end, PAR_PREROLL_BITS);
      SpanRes pr;
      pr.end = guess; pr.ntok = 0; pr.nout = 0; pr.flags = 0;
      blk_span<false>(lane > 0 && back > 0, in, in_len, guess - back, guess, T, lenx, distx, nullptr, pr, lane_ring);
      if (lane > 0 && pr.flags == 0 && pr.end >= guess && pr.end < bound) start = pr.end;
    }
    // The counting pass already writes its tokens, into the lane's PIECE of a scratch area (piece = the slots of the bits of the
    // lane's sub-chunk: typical streams spend 9..15 bits per token); if every lane's tokens fit its piece, the tokens are then COPIED to their final place instead of being
    // decoded a second time.  (One slot per PAR_PIECE_BITS input bits, addressed by the bit's offset in the buffer.)
    unsigned* piece = nullptr;
    unsigned pcap = 0;
    if (pieces) {
      const unsigned long long b0 = ((unsigned long long)st.in_off * 8 + hdr_end + lane * sc) / PAR_PIECE_BITS;
      const unsigned long long b1 = ((unsigned long long)st.in_off * 8 + min(hdr_end + (lane + 1) * sc, limit)) / PAR_PIECE_BITS;
      piece = pieces + b0;
      pcap = b1 > b0 ? (unsigned)(b1 - b0) : 0u;
    }
    if (pieces) blk_span<true>(true, in, in_len, start, bound, T, lenx, distx, piece, r, lane_ring, pcap);
    else blk_span<false>(true, in, in_len, start, bound, T, lenx, distx, nullptr, r, lane_ring);
    for (int it = 0; it < 34; it++) {
      const unsigned pe = __shfl_up_sync(0xffffffffu, r.end, 1), pf = __shfl_up_sync(0xffffffffu, r.flags, 1);
      bool ch = false, rerun = false;
      if (lane > 0) {
        if (pf & 1) {                                   // the block ended before my sub-chunk
          if (!(r.flags & 4) || r.end != pe) { r.end = pe; r.ntok = 0; r.nout = 0; r.flags = 1 | 4; start = pe; ch = true; }
        } else if (pe != start || (r.flags & 4)) {
          start = pe;
          rerun = true;
          ch = true;
        }
      }
      if (__any_sync(0xffffffffu, rerun)) {
        if (pieces) blk_span<true>(rerun, in, in_len, start, bound, T, lenx, distx, piece, r, lane_ring, pcap);
        else blk_span<false>(rerun, in, in_len, start, bound, T, lenx, distx, nullptr, r, lane_ring);
      }
      if (!__any_sync(0xffffffffu, ch)) break;
    }
    // ---- totals and token offsets
    const unsigned bad = __ballot_sync(0xffffffffu, (r.flags & 2) != 0);
    const unsigned eob = __shfl_sync(0xffffffffu, r.flags, 31) & 1;
    unsigned pre = r.ntok, sum_out = r.nout;
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, pre, d);
      if ((int)lane >= d) pre += v;
    }
    for (int d = 16; d >= 1; d >>= 1) sum_out += __shfl_xor_sync(0xffffffffu, sum_out, d);
    total_tok = __shfl_sync(0xffffffffu, pre, 31);
    total_out = sum_out;
    ok = !bad && eob;
    // token slots are handed out from one cursor once the count is known (false candidates take some too)
    unsigned long long toff = 0;
    if (ok && lane == 0) toff = atomicAdd(tok_cursor, (unsigned long long)total_tok);
    toff = __shfl_sync(0xffffffffu, toff, 0);
    if (toff + total_tok > tok_capacity) ok = false;
    blk.tok_off = (long long)toff;
    const bool in_pieces = pieces && __all_sync(0xffffffffu, (r.flags & 4) || r.ntok <= pcap);
    if (ok && in_pieces) {
      for (int l = 0; l < 32; l++) {
        const unsigned n = __shfl_sync(0xffffffffu, (r.flags & 4) ? 0u : r.ntok, l);
        const unsigned off = __shfl_sync(0xffffffffu, pre - r.ntok, l);
        const unsigned long long pb = __shfl_sync(0xffffffffu, (unsigned long long)(piece - pieces), l);
        const unsigned* src = pieces + pb;
        unsigned* dst = tokens + blk.tok_off + off;
        for (unsigned i = lane; i < n; i += 32) dst[i] = src[i];
      }
    } else {
      const bool emit = ok && !(r.flags & 4) && r.ntok;
      if (__any_sync(0xffffffffu, emit)) {
        SpanRes r2 = r;
        blk_span<true>(emit, in, in_len, start, bound, T, lenx, distx, tokens + blk.tok_off + (pre - r.ntok), r2, lane_ring);
        if (emit && (r2.ntok != r.ntok || r2.end != r.end)) ok = false;           // cannot happen
      }
    }
    ok = __all_sync(0xffffffffu, ok);
  }
  if (lane == 31) {
    ParBlk* o = blks + bi;
    o->end_bit = r.end;
    o->n_tok = total_tok;
    o->out_len = total_out;
    o->flags = (ok ? 1u : 0u) | (fin ? 2u : 0u);
    o->tok_off = blk.tok_off;
  }
}

// 5. Tokens -> bytes: one CTA per stream walks the stream's tokens in order, one tile of up to PAR_LZ_THREADS tokens /
//    PAR_LZ_CAP output bytes at a time (thread = token).  A block-wide scan of the token lengths gives every token its
//    position; the tile's output is assembled in shared memory and then stored coalesced.  Literals and the bytes that
//    matches copy from before the tile (global memory, written by earlier tiles) are placed at once; bytes copied from
//    inside the tile wait, without block barriers, until the PENDING bitmap (one bit per staged byte, set by the
//    matches that still have to produce it) is clear over their source range.
// Two shapes: 256 threads / 4 KB (up to 8 CTAs per SM: many streams) and 1024 threads / 16 KB (few streams: a stream
// is a serial chain of tiles, so its latency is what counts).
__device__ __forceinline__ unsigned par_bits(unsigned a, unsigned b, unsigned w) {   // bits of [a, b) that fall in word w
  const unsigned lo = max(a, w * 32), hi = min(b, w * 32 + 32);
  if (hi <= lo) return 0;
  return (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1) << (lo & 31));
}
// Apply OP (0 = or, 1 = and-not, 2 = test) to the bits [a, b) of the bitmap; ranges of up to 33 bits (the usual case)
// touch at most two words and get their masks from one 64-bit shift.
template <int OP>
__device__ __forceinline__ bool par_range(unsigned* bm, unsigned a, unsigned b) {
  bool hit = false;
  if (b - a <= 33) {
    const unsigned long long m = ((1ull << (b - a)) - 1) << (a & 31);
    const unsigned m0 = (unsigned)m, m1 = (unsigned)(m >> 32), w = a >> 5;
    if (OP == 0) { atomicOr(&bm[w], m0); if (m1) atomicOr(&bm[w + 1], m1); }
    else if (OP == 1) { atomicAnd(&bm[w], ~m0); if (m1) atomicAnd(&bm[w + 1], ~m1); }
    else { hit = (((volatile unsigned*)bm)[w] & m0) != 0; if (m1 && !hit) hit = (((volatile unsigned*)bm)[w + 1] & m1) != 0; }
  } else {
    for (unsigned w = a >> 5; w <= (b - 1) >> 5; w++) {
      const unsigned mk = par_bits(a, b, w);
      if (OP == 0) atomicOr(&bm[w], mk);
      else if (OP == 1) atomicAnd(&bm[w], ~mk);
      else if (((volatile unsigned*)bm)[w] & mk) { hit = true; break; }
    }
  }
  return hit;
}
static const unsigned PAR_LZ_MIRROR = 32768;          // dynamic shared memory of par_lz_kernel
template <int PAR_LZ_THREADS, int PAR_LZ_CAP>
__global__ void __launch_bounds__(PAR_LZ_THREADS) par_lz_kernel(const ParStream* __restrict__ streams,
                                                                const ParBlk* __restrict__ blks,
                                                                unsigned bstride,
                                                                const unsigned* __restrict__ tokens,
                                                                unsigned char* out_base, ParRes* __restrict__ res) {
  const int NT = PAR_LZ_THREADS;
  // The last 32 KB of the stream's finished output are mirrored in (dynamic) shared memory: what a match copies from
  // before the tile comes from there instead of from global memory, whose round trip (stores of the previous tile,
  // then dependent loads) was on the critical path of every tile (600 reference chunks: 178 -> 167 ms of inflate).
  const bool MIRROR = true;
  const unsigned MR = PAR_LZ_MIRROR;
  MTS_DYN_SMEM(mirror);
  __shared__ unsigned ob_w[PAR_LZ_CAP / 4 + 1];
  __shared__ unsigned pend_w[PAR_LZ_CAP / 32 + 1];
  __shared__ unsigned wsum[NT / 32];
  __shared__ unsigned s_total;
  __shared__ unsigned long long s_ad[2 * (NT / 32)];
  unsigned char* ob = (unsigned char*)ob_w;
  const ParStream st = streams[blockIdx.x];
  unsigned char* out = out_base + st.out_off;
  const unsigned out_cap = (unsigned)st.out_len;
  const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (unsigned i = tid; i < PAR_LZ_CAP / 32 + 1; i += NT) pend_w[i] = 0;
  unsigned obase = 0;
  bool fail = false;
  // adler32 of the stream, from the bytes that pass through the storing threads' registers (as seg_resolve_kernel):
  // ad_a = sum of bytes, ad_b = sum of (out_cap - position) * byte
  unsigned long long ad_a = 0, ad_b = 0;
  // ---- walk the chain of blocks: the first block starts right after the zlib header, every next one where its
  //      predecessor ended; the walk stops at the first block that is missing or was not decoded cleanly
  unsigned cur_bit = st.first_bit, n_done = 0, fin = 0;
  unsigned bj = blockIdx.x * bstride;                        // this stream's blocks: [bj, bend), unused slots have bit ~0
  const unsigned bend = bj + bstride;
  for (;;) {
    while (bj < bend && blks[bj].bit < cur_bit) bj++;
    if (bj >= bend) break;
    const ParBlk blk = blks[bj];
    if (blk.bit != cur_bit || !(blk.flags & 1) || blk.end_bit <= cur_bit || (unsigned long long)obase + blk.out_len > out_cap) break;
    const unsigned* tk = tokens + blk.tok_off;
    const unsigned T = blk.n_tok;
    const unsigned block_end = obase + blk.out_len;
  unsigned t0 = 0;
  unsigned nxt = tid < T ? tk[tid] : 0;
  while (t0 < T) {
    const unsigned t = nxt;
    const bool has = t0 + tid < T;
    if (t0 + NT + tid < T) nxt = tk[t0 + NT + tid];                 // assumes that the whole tile fits (the usual case)
    const bool isM = has && (t >> 31);
    const unsigned L = has ? (isM ? (t >> 16) & 0x1ffu : 1u) : 0u;
    unsigned incl = L;
    for (int d = 1; d < 32; d <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += v; }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();                                                // also orders the previous tile's stores before the loads below
    unsigned ws = lane < (unsigned)(NT / 32) ? wsum[lane] : 0u;     // every warp scans the warp totals itself
    for (int d = 1; d < NT / 32; d <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, ws, d); if ((int)lane >= d) ws += v; }
    const unsigned woff = __shfl_sync(0xffffffffu, ws, (wid + 31) & 31) * (wid > 0);
    const unsigned rel = woff + incl - L;                           // position inside the tile
    const bool fits = has && rel + L <= (unsigned)PAR_LZ_CAP;       // monotone: the tile is the longest fitting prefix
    const unsigned ncut = (unsigned)__syncthreads_count(fits);
    if (fits && tid + 1 == ncut) s_total = rel + L;
    const unsigned dist = (t & 0x7fffu) + 1;
    bool m = fits && isM;
    if (m && dist > obase + rel) { fail = true; m = false; }
    if (fits && obase + rel + L > block_end) { fail = true; m = false; }
    else if (fits && !isM) ob[rel] = (unsigned char)t;
    const int srel = (int)rel - (int)dist;                          // source position inside the tile (negative: before it)
    unsigned n_old = 0;
    if (m) {
      n_old = srel < 0 ? min(L, (unsigned)(-srel)) : 0u;
      if (n_old < L) {                                              // has an in-tile part: its bytes are pending
        par_range<0>(pend_w, rel, rel + L);
      }
      if (MIRROR) {
        const unsigned m0 = obase + (unsigned)srel;                 // (srel < 0 here whenever n_old > 0; wraps as intended)
        for (unsigned j = 0; j < n_old; j++) ob[rel + j] = mirror[(m0 + j) & (MR - 1)];
      }
      // bytes from before the tile: aligned 32-bit loads (the output buffer is 4-byte aligned and padded), 8 bytes a turn
      const unsigned char* sp = out + obase + srel;
      const unsigned mis = (unsigned)((uintptr_t)sp & 3);
      const unsigned* sw = (const unsigned*)(sp - mis);
      for (unsigned j0 = 0; !MIRROR && j0 < n_old; j0 += 8) {
        const unsigned w0 = sw[0], w1 = (mis + n_old - j0 > 4) ? sw[1] : 0u, w2 = (mis + n_old - j0 > 8) ? sw[2] : 0u;
        unsigned v0 = __funnelshift_r(w0, w1, mis * 8), v1 = __funnelshift_r(w1, w2, mis * 8);
        unsigned char* dp = ob + rel + j0;
#pragma unroll
        for (unsigned j = 0; j < 4; j++) { if (j0 + j < n_old) dp[j] = (unsigned char)v0; v0 >>= 8; }
#pragma unroll
        for (unsigned j = 4; j < 8; j++) { if (j0 + j < n_old) dp[j] = (unsigned char)v1; v1 >>= 8; }
        sw += 2;
      }
      if (n_old == L) m = false;
    }
    __syncthreads();                                                // literals, old bytes, pending bits and s_total are visible
    const unsigned a = (unsigned)max(srel, 0), b = min((unsigned)(srel + (int)L), rel);   // in-tile source bytes outside my own output
    while (__any_sync(0xffffffffu, m)) {
      if (m) {
        const bool clear = !(b > a && par_range<2>(pend_w, a, b));
        if (clear) {
          __threadfence_block();
          for (unsigned j = n_old; j < L; j++) ob[rel + j] = ob[(unsigned)(srel + (int)j)];   // in order: may read my own bytes
          __threadfence_block();
          par_range<1>(pend_w, rel, rel + L);
          m = false;
        }
      }
    }
    __syncthreads();
    const unsigned total = s_total;
    {
      // coalesced store of the tile: bytes up to the first 16-byte boundary, 16-byte vectors (assembled from the
      // staging words, whose alignment differs), bytes of the rest
      unsigned char* gp = out + obase;
      const unsigned head = min(total, (unsigned)((16 - ((uintptr_t)gp & 15)) & 15));
      if (tid < head) { const unsigned v = ob[tid]; gp[tid] = (unsigned char)v; ad_a += v; ad_b += (unsigned long long)(out_cap - (obase + tid)) * v; }
      const unsigned nvec = (total - head) >> 4;
      for (unsigned v = tid; v < nvec; v += NT) {
        const unsigned o = head + v * 16;
        const unsigned* ww = ob_w + (o >> 2);
        const unsigned sh = (o & 3) * 8;
        const unsigned a0 = ww[0], a1 = ww[1], a2 = ww[2], a3 = ww[3], a4 = sh ? ww[4] : 0u;
        uint4 val;
        val.x = __funnelshift_r(a0, a1, sh); val.y = __funnelshift_r(a1, a2, sh);
        val.z = __funnelshift_r(a2, a3, sh); val.w = __funnelshift_r(a3, a4, sh);
        *(uint4*)(gp + o) = val;
        const unsigned wv[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const unsigned w = wv[q];
          const unsigned sum = (w & 0xffu) + ((w >> 8) & 0xffu) + ((w >> 16) & 0xffu) + (w >> 24);
          ad_a += sum;
          ad_b += (unsigned long long)(out_cap - (obase + o + 4 * q)) * sum - (((w >> 8) & 0xffu) + 2 * ((w >> 16) & 0xffu) + 3 * (w >> 24));
        }
      }
      const unsigned rest = head + nvec * 16;
      if (rest + tid < total) {
        const unsigned v = ob[rest + tid];
        gp[rest + tid] = (unsigned char)v;
        ad_a += v; ad_b += (unsigned long long)(out_cap - (obase + rest + tid)) * v;
      }
      if (MIRROR) for (unsigned i = tid; i < total; i += NT) mirror[(obase + i) & (MR - 1)] = ob[i];
    }
    obase += total;
    t0 += ncut;
    if (ncut != (unsigned)NT && t0 < T) nxt = t0 + tid < T ? tk[t0 + tid] : 0;   // the tile was cut short: reload
  }
    if (__syncthreads_or(fail || obase != block_end)) { fail = true; break; }
    cur_bit = blk.end_bit;
    n_done++;
    fin = (blk.flags >> 1) & 1;
    bj++;
    if (fin) break;
  }
  {
    unsigned long long a = warp_sum(ad_a), b2 = warp_sum(ad_b % ADLER_BASE);
    if (lane == 0) { s_ad[2 * wid] = a; s_ad[2 * wid + 1] = b2; }
  }
  __syncthreads();
  if (tid == 0) {
    unsigned long long ta = 0, tb = 0;
    for (int w = 0; w < NT / 32; w++) { ta += s_ad[2 * w]; tb += s_ad[2 * w + 1]; }
    ParRes r;
    r.tail_bit = cur_bit; r.tail_out = obase; r.n_done = n_done; r.flags = fin | (fail ? 2u : 0u) | 4u;
    // (ad_b weighs a byte at position p with out_cap - p; for a prefix of obase bytes the weight is obase - p)
    const unsigned long long short_by = (unsigned long long)(out_cap - obase) % ADLER_BASE * (ta % ADLER_BASE) % ADLER_BASE;
    const unsigned long long s2 = (tb + ADLER_BASE - short_by + obase % ADLER_BASE) % ADLER_BASE;
    r.adler = ((unsigned)s2 << 16) | (unsigned)((1 + ta) % ADLER_BASE);
    res[blockIdx.x] = r;
  }
}


// ---------------------------------------------------------------------------------------------- few streams: low latency
// par_lz_kernel makes every stream a serial chain of tiles (~20 ms for a 23 MB chunk, whatever the batch size).  With
// few streams the blocks of a stream are resolved IN PARALLEL instead, one CTA per block, into 16-bit cells: a byte, or a
// MARKER 0x8000 | index into the 32 KB before the block for what is copied from there (markers are copied like data, so
// every cell ends up as a byte or as a direct reference to the window).  A last, fully parallel pass turns the cells
// into bytes.
//   par_chain_kernel  thread per stream: follow the chain (as par_lz_kernel does), give every chained block its output
//                     offset (kept in ParBlk::limit, which nobody needs any more) and flag 4, write the stream's result
//   par_lzc_kernel    CTA per block slot: tokens -> cells, same tile scheme as par_lz_kernel
//   par_cells_kernel  CTA per block: cells -> bytes, markers chased back through the cells of the earlier blocks
__global__ void __launch_bounds__(64) par_chain_kernel(const ParStream* __restrict__ streams, ParBlk* __restrict__ blks,
                                                       unsigned bstride, unsigned ns, ParRes* __restrict__ res) {
  const unsigned sidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (sidx >= ns) return;
  const ParStream st = streams[sidx];
  unsigned cur_bit = st.first_bit, n_done = 0, fin = 0, obase = 0;
  unsigned bj = sidx * bstride;
  const unsigned bend = bj + bstride;
  for (;;) {
    while (bj < bend && blks[bj].bit < cur_bit) bj++;
    if (bj >= bend) break;
    const ParBlk blk = blks[bj];
    if (blk.bit != cur_bit || !(blk.flags & 1) || blk.end_bit <= cur_bit ||
        (unsigned long long)obase + blk.out_len > (unsigned)st.out_len) break;
    blks[bj].limit = obase;
    blks[bj].flags = blk.flags | 4;
    obase += blk.out_len;
    cur_bit = blk.end_bit;
    n_done++;
    fin = (blk.flags >> 1) & 1;
    bj++;
    if (fin) break;
  }
  ParRes r;
  r.tail_bit = cur_bit; r.tail_out = obase; r.n_done = n_done; r.flags = fin;
  res[sidx] = r;
}

template <int NT, int CAP>
__global__ void __launch_bounds__(NT) par_lzc_kernel(const ParStream* __restrict__ streams, ParBlk* __restrict__ blks,
                                                     const unsigned* __restrict__ tokens,
                                                     unsigned short* cells_base, long long cells_origin) {
  __shared__ unsigned short ob[CAP];
  __shared__ unsigned pend_w[CAP / 32 + 1];
  __shared__ unsigned wsum[NT / 32];
  __shared__ unsigned s_total;
  const ParBlk blk = blks[blockIdx.x];
  if (blk.bit == 0xffffffffu || !(blk.flags & 4)) return;
  const ParStream st = streams[blk.stream];
  const unsigned boff = blk.limit;                               // output offset of the block inside its stream
  unsigned short* cells = cells_base + (st.out_off - cells_origin) + boff;
  const unsigned* tk = tokens + blk.tok_off;
  const unsigned T = blk.n_tok, block_end = blk.out_len;
  const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (unsigned i = tid; i < CAP / 32 + 1; i += NT) pend_w[i] = 0;
  unsigned obase = 0, t0 = 0;
  bool fail = false;
  unsigned nxt = tid < T ? tk[tid] : 0;
  while (t0 < T) {
    const unsigned t = nxt;
    const bool has = t0 + tid < T;
    if (t0 + NT + tid < T) nxt = tk[t0 + NT + tid];
    const bool isM = has && (t >> 31);
    const unsigned L = has ? (isM ? (t >> 16) & 0x1ffu : 1u) : 0u;
    unsigned incl = L;
    for (int d = 1; d < 32; d <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += v; }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    unsigned ws = lane < (unsigned)(NT / 32) ? wsum[lane] : 0u;
    for (int d = 1; d < NT / 32; d <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, ws, d); if ((int)lane >= d) ws += v; }
    const unsigned woff = __shfl_sync(0xffffffffu, ws, (wid + 31) & 31) * (wid > 0);
    const unsigned rel = woff + incl - L;
    const bool fits = has && rel + L <= (unsigned)CAP;
    const unsigned ncut = (unsigned)__syncthreads_count(fits);
    if (fits && tid + 1 == ncut) s_total = rel + L;
    const unsigned dist = (t & 0x7fffu) + 1;
    bool m = fits && isM;
    if (m && dist > boff + obase + rel) { fail = true; m = false; }              // reaches before the stream
    if (fits && obase + rel + L > block_end) { fail = true; m = false; }
    else if (fits && !isM) ob[rel] = (unsigned short)(t & 0xffu);
    const int srel = (int)rel - (int)dist;
    unsigned n_old = 0;
    if (m) {
      n_old = srel < 0 ? min(L, (unsigned)(-srel)) : 0u;
      if (n_old < L) par_range<0>(pend_w, rel, rel + L);
      // cells from before the tile: earlier tiles of this block (global), or markers for what precedes the block
      const int g0 = (int)obase + srel;                          // block-local position of the first source cell
      for (unsigned j0 = 0; j0 < n_old; j0 += 8) {
        unsigned short v[8];
#pragma unroll
        for (unsigned j = 0; j < 8; j++) {
          const int g = g0 + (int)(j0 + j);
          v[j] = (j0 + j < n_old) ? (g >= 0 ? cells[g] : (unsigned short)(0x8000u | (unsigned)(32768 + g))) : (unsigned short)0;
        }
#pragma unroll
        for (unsigned j = 0; j < 8; j++) if (j0 + j < n_old) ob[rel + j0 + j] = v[j];
      }
      if (n_old == L) m = false;
    }
    __syncthreads();
    const unsigned a = (unsigned)max(srel, 0), b = min((unsigned)(srel + (int)L), rel);
    while (__any_sync(0xffffffffu, m)) {
      if (m) {
        const bool clear = !(b > a && par_range<2>(pend_w, a, b));
        if (clear) {
          __threadfence_block();
          for (unsigned j = n_old; j < L; j++) ob[rel + j] = ob[(unsigned)(srel + (int)j)];
          __threadfence_block();
          par_range<1>(pend_w, rel, rel + L);
          m = false;
        }
      }
    }
    __syncthreads();
    const unsigned total = s_total;
    for (unsigned i = tid; i < total; i += NT) cells[obase + i] = ob[i];
    obase += total;
    t0 += ncut;
    if (ncut != (unsigned)NT && t0 < T) nxt = t0 + tid < T ? tk[t0 + tid] : 0;
  }
  if (__syncthreads_or(fail || obase != block_end)) { if (tid == 0) blks[blockIdx.x].flags = blk.flags | 8; }
}

// cells -> bytes, fully parallel: one CTA per block slot, four cells per thread and turn.  A marker is chased through
// the cells of the earlier blocks; the output offsets of the PAR_CELLS_BACK chained blocks before this one are kept in
// shared memory, so a hop costs one global load (every hop goes at least one block back; one or two reach a byte).
static const int PAR_CELLS_BACK = 32;
__global__ void __launch_bounds__(256) par_cells_kernel(const ParStream* __restrict__ streams, const ParBlk* __restrict__ blks,
                                                        unsigned bstride, const unsigned short* __restrict__ cells_base,
                                                        long long cells_origin, unsigned char* __restrict__ out_base,
                                                        ParRes* __restrict__ res) {
  __shared__ unsigned pb[PAR_CELLS_BACK];                       // output offsets of the previous chained blocks, nearest first
  __shared__ int pbs[PAR_CELLS_BACK];                           // ... and their slots
  __shared__ int npb;
  const ParBlk blk = blks[blockIdx.x];
  if (blk.bit == 0xffffffffu || !(blk.flags & 4)) return;
  const ParStream st = streams[blk.stream];
  unsigned char* out = out_base + st.out_off;
  const unsigned short* cl = cells_base + (st.out_off - cells_origin);
  const int slot0 = (int)(blk.stream * bstride);
  const unsigned o = blk.limit, n = blk.out_len;
  if (threadIdx.x == 0) {
    int cnt = 0;
    for (int k = (int)blockIdx.x - 1; k >= slot0 && cnt < PAR_CELLS_BACK; k--)
      if (blks[k].flags & 4) { pb[cnt] = blks[k].limit; pbs[cnt] = k; cnt++; }
    npb = cnt;
  }
  __syncthreads();
  const int np = npb;
  bool bad = (blk.flags & 8) != 0;
  for (unsigned i0 = threadIdx.x * 4; i0 < n; i0 += blockDim.x * 4) {
    unsigned v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) v[q] = i0 + q < n ? cl[o + i0 + q] : 0u;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      unsigned cur_o = o;
      int j = -1, cur_slot = 0;                                 // index in pb of the block being looked at (-1: this block)
      while (v[q] & 0x8000u) {
        const int p = (int)cur_o - 32768 + (int)(v[q] & 0x7fffu);   // position in the stream, before the block looked at
        if (p < 0) { bad = true; v[q] = 0; break; }
        if (j < np) { do { j++; } while (j < np && pb[j] > (unsigned)p); }
        if (j < np) cur_o = pb[j];
        else {
          // further back than the table (long chains through many blocks, e.g. periodic data): walk the slots
          if (j == np) cur_slot = np ? pbs[np - 1] : (int)blockIdx.x;
          int k = cur_slot;
          do { k--; } while (k >= slot0 && (!(blks[k].flags & 4) || blks[k].limit > (unsigned)p));
          if (k < slot0) { bad = true; v[q] = 0; break; }
          cur_slot = k; cur_o = blks[k].limit; j = np + 1;
        }
        v[q] = cl[p];
      }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) if (i0 + q < n) out[o + i0 + q] = (unsigned char)v[q];
  }
  if (bad) atomicOr(&res[blk.stream].flags, 2u);
}

}  // namespace mts
