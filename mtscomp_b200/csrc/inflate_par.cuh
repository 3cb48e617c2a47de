// inflate_par.cuh — block-parallel decoding of plain zlib streams (e.g. chunks written by the reference Writer).
//
// A zlib stream has no index: its DEFLATE blocks (~300 dynamic-Huffman blocks per 23 MB chunk at zlib level 6) can only
// be located by decoding.  One warp per stream (inflate.cuh) therefore leaves the GPU idle.  This path finds the blocks
// speculatively and decodes them with ONE THREAD PER BLOCK (thousands of blocks in flight):
//   1 par_find_kernel      every bit offset is tested for a plausible dynamic-block header (BTYPE, HLIT/HDIST ranges,
//                          complete code-length code); survivors (~8e-4 of the offsets) are appended to a list
//   2 par_validate_kernel  thread per survivor: full header parse; both Huffman codes must be complete and have an
//                          end-of-block code (what zlib's deflate always emits) -> candidate blocks
//   3 par_decode_kernel<0> thread per candidate: dry decode -> end bit offset and output length
//   (host)                 per stream: follow end -> next start from the first block; blocks on that chain get their
//                          output offsets; whatever follows the last chained block is the "tail"
//   4 par_decode_kernel<1> thread per chained block: decode into 16-bit cells; a back-reference that reaches before the
//                          block's own output becomes a MARKER (0x8000 | index into the previous 32 KB), and markers are
//                          copied like data, so every cell ends up as a byte or as a direct reference to the window
//   5 par_resolve_kernel   CTA per stream, blocks in order: cells -> bytes (markers read the already resolved window)
//   6 (inflate.cuh)        one warp per stream decodes the tail serially (usually nothing or the final small block)
// Anything unexpected (no chain, overflow of a list, bad data) falls back to the serial decoder, which also produces
// the error status; this path never decides that a stream is corrupt by itself.
#pragma once
#include "common.cuh"
#include "inflate.cuh"

namespace mts {

struct ParStream {       // one whole zlib stream
  long long in_off;      // byte offset in the compressed buffer
  long long out_off;     // byte offset of its output in the transformed buffer
  int in_len, out_len;
};

struct ParCand {         // a candidate dynamic block
  unsigned stream, bit;  // owning stream, bit offset of the block header inside it
  unsigned end_bit;      // bit offset just after the end-of-block code (dry decode)
  unsigned out_len;      // bytes the block produces
  unsigned out_off;      // offset of those bytes inside the stream's output (set on the host for chained blocks)
  int state;             // 1 = decoded cleanly, <= 0 = rejected
  unsigned final_;       // BFINAL
  unsigned pad_;
};

// Per-thread decoding tables (global memory scratch, one per candidate).
static const int PAR_LBITS = 9, PAR_DBITS = 6;
struct ParTables {
  unsigned ltab[1 << PAR_LBITS];
  unsigned dtab[1 << PAR_DBITS];
  unsigned short lsorted[288], dsorted[32];
  unsigned short lcount[16], dcount[16];
  unsigned char lens[320];
};

// ---------------------------------------------------------------------------------------------- per-thread bit reader
struct TBits {
  const unsigned* w;     // 4-byte aligned base of the stream
  unsigned sh;           // 8 * (stream address & 3)
  unsigned kmax;         // last readable word
  unsigned k;            // next raw word
  unsigned raw, lo, hi;  // raw = w[k-1]; (lo, hi) = 64 stream bits
  unsigned pos;          // cursor inside (lo, hi), < 32 after refill
  unsigned base_bit;     // stream bit offset of bit 0 of lo
  __device__ __forceinline__ unsigned next_word() {
    unsigned nw = w[min(k, kmax)];
    unsigned v = __funnelshift_r(raw, nw, sh);
    raw = nw; k++;
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
    lo = next_word();
    hi = next_word();
    pos = bit & 31;
    base_bit = word << 5;
  }
  __device__ __forceinline__ void refill() {
    if (pos >= 32) { lo = hi; hi = next_word(); pos -= 32; base_bit += 32; }
  }
  __device__ __forceinline__ unsigned window() const { return __funnelshift_r(lo, hi, pos); }
  __device__ __forceinline__ void drop(unsigned n) { pos += n; }
  __device__ __forceinline__ unsigned get(unsigned n) { refill(); unsigned v = window() & ((1u << n) - 1); pos += n; return v; }
  __device__ __forceinline__ unsigned bit_pos() const { return base_bit + pos; }
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

// Parse a dynamic block header at the reader's position.  On success the reader is at the first symbol, lens[0..nl) and
// lens[nl..nl+nd) hold the code lengths.  strict: require complete literal/length and distance codes (zlib's output).
__device__ bool par_parse_header(TBits& br, unsigned char* lens, int& nl, int& nd, unsigned& final_, bool strict) {
  final_ = br.get(1);
  if (br.get(2) != 2) return false;
  nl = (int)br.get(5) + 257;
  nd = (int)br.get(5) + 1;
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
  if (strict && left != 0) return false;
  for (int i = 0; i < 19; i++) if (cl[i]) sorted[offs[cl[i]]++] = (unsigned char)i;
  int idx = 0;
  unsigned prev = 0, kl_run = 0, kd_run = 0;
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
    for (unsigned k = 0; k < rep; k++) lens[idx + k] = (unsigned char)val;
    if (strict && val) {
      // running Kraft sums (units of 2^-15): a random bit string over-subscribes one of the codes within a few dozen
      // lengths, so false survivors are dropped long before the end of the header
      for (unsigned k = 0; k < rep; k++) { if (idx + (int)k < nl) kl_run += 32768u >> val; else kd_run += 32768u >> val; }
      if (kl_run > 32768u || kd_run > 32768u) return false;
    }
    idx += (int)rep;
    prev = val;
  }
  if (lens[256] == 0) return false;
  if (strict) {
    // Kraft sums in units of 2^-15
    unsigned kl = 0, kd = 0, ndist = 0;
    for (int i = 0; i < nl; i++) if (lens[i]) kl += 32768u >> lens[i];
    for (int i = 0; i < nd; i++) if (lens[nl + i]) { kd += 32768u >> lens[nl + i]; ndist++; }
    if (kl != 32768u) return false;
    if (kd != 32768u && !(ndist <= 1 && kd <= 16384u)) return false;   // zlib emits >= 2 distance codes; tolerate 0/1
  }
  return true;
}

// Fast table + canonical arrays for one code (per thread, serial).  KIND 1 literal/length, 2 distance.
template <int KIND>
__device__ bool par_build(const unsigned char* lens, int n, unsigned* tab, int tb, unsigned short* sorted,
                          unsigned short* count) {
  for (int i = 0; i < 16; i++) count[i] = 0;
  for (int i = 0; i < n; i++) count[lens[i]]++;
  count[0] = 0;
  unsigned first[16], offs[16];
  unsigned code = 0, o = 0;
  int left = 1;
  for (int l = 1; l <= 15; l++) {
    code = (code + (l > 1 ? count[l - 1] : 0)) << 1;
    first[l] = code;
    offs[l] = o;
    o += count[l];
    left = (left << 1) - count[l];
    if (left < 0) return false;
  }
  for (int i = 0; i < (1 << tb); i++) tab[i] = 0;
  for (int s = 0; s < n; s++) {
    const unsigned l = lens[s];
    if (!l) continue;
    sorted[offs[l]++] = (unsigned short)s;
    if ((int)l <= tb) {
      const unsigned r = __brev(first[l]++) >> (32 - l);
      const unsigned e = KIND == 1 ? ll_entry((unsigned)s, l) : d_entry((unsigned)s, l);
      for (unsigned k = r; k < (1u << tb); k += 1u << l) tab[k] = e;
    } else first[l]++;
  }
  return true;
}

template <int KIND>
__device__ __forceinline__ unsigned par_slow(unsigned win, const unsigned short* count, const unsigned short* sorted) {
  unsigned code = 0, first = 0, index = 0;
  for (int l = 1; l <= 15; l++) {
    code |= (win >> (l - 1)) & 1;
    const unsigned c = count[l];
    if (code - first < c) {
      const unsigned s = sorted[index + (code - first)];
      return KIND == 1 ? ll_entry(s, (unsigned)l) : d_entry(s, (unsigned)l);
    }
    index += c;
    first = (first + c) << 1;
    code <<= 1;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------- kernels
// 1. Survivors of the cheap header test: one thread per stream byte (8 bit offsets).  Phase 1 tests the fixed fields of
//    all 8 offsets (BTYPE = dynamic, HLIT <= 29, HDIST <= 29); phase 2 visits only the passing offsets and checks that the
//    code-length code is complete: its 3-bit lengths are summed 3 at a time through a 512-entry table of 2^(7-len).
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
  const unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
  // a dynamic header is at least 17 + 12 bits + two codes: ignore the last bytes (zlib trailer + shortest block)
  if (st.in_len < 16 || b < 2 || b + 12 > (unsigned)st.in_len) return;
  const unsigned char* p = comp + st.in_off + b;
  const unsigned mis = (unsigned)((uintptr_t)p & 3);
  const unsigned* w = (const unsigned*)(p - mis);
  const unsigned kmax = (mis + (unsigned)st.in_len - b - 1) >> 2;
  unsigned x[4];
  {
    unsigned r0 = w[0], r1 = w[min(1u, kmax)], r2 = w[min(2u, kmax)], r3 = w[min(3u, kmax)], r4 = w[min(4u, kmax)];
    x[0] = __funnelshift_r(r0, r1, mis * 8); x[1] = __funnelshift_r(r1, r2, mis * 8);
    x[2] = __funnelshift_r(r2, r3, mis * 8); x[3] = __funnelshift_r(r3, r4, mis * 8);
  }
  unsigned mask = 0;
#pragma unroll
  for (unsigned s = 0; s < 8; s++) {
    const unsigned t0 = __funnelshift_r(x[0], x[1], s);
    const bool pass = ((t0 >> 1) & 3) == 2 && ((t0 >> 3) & 31) <= 29 && ((t0 >> 8) & 31) <= 29;
    mask |= (unsigned)pass << s;
  }
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
    const unsigned kraft = k9[lo & 511] + k9[(lo >> 9) & 511] + k9[(lo >> 18) & 511] +
                           k9[((lo >> 27) | (hi << 5)) & 511] + k9[(hi >> 4) & 511] + k9[(hi >> 13) & 511] +
                           k9[(hi >> 22) & 511];
    if (kraft != 128) continue;
    const unsigned at = atomicAdd(&counters[0], 1u);
    if (at < cap) surv[at] = ((unsigned long long)blockIdx.y << 32) | (b * 8 + s);
  }
}

// 2. Full header validation: thread per survivor, valid ones are appended to the candidate list.
__global__ void __launch_bounds__(128) par_validate_kernel(const unsigned char* __restrict__ comp,
                                                           const ParStream* __restrict__ streams,
                                                           const unsigned long long* __restrict__ surv, unsigned n_surv,
                                                           ParCand* __restrict__ cand, unsigned cap,
                                                           unsigned* __restrict__ counters) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_surv) return;
  const unsigned long long sv = surv[i];
  const unsigned sidx = (unsigned)(sv >> 32), bit = (unsigned)sv;
  const ParStream st = streams[sidx];
  TBits br;
  br.init(comp + st.in_off, (unsigned)st.in_len, bit);
  unsigned char lens[320];
  int nl, nd;
  unsigned fin;
  if (!par_parse_header(br, lens, nl, nd, fin, true)) return;
  if (br.bit_pos() > (unsigned)st.in_len * 8) return;
  const unsigned at = atomicAdd(&counters[1], 1u);
  if (at < cap) {
    ParCand c;
    c.stream = sidx; c.bit = bit; c.end_bit = 0; c.out_len = 0; c.out_off = 0; c.state = 0; c.final_ = fin; c.pad_ = 0;
    cand[at] = c;
  }
}

// 3/4. Thread per block.  REAL = 0: dry decode of every candidate (end bit, output length).  REAL = 1: decode the
// chained blocks (list[] holds their candidate indices) into 16-bit cells.
template <int REAL>
__global__ void __launch_bounds__(128) par_decode_kernel(const unsigned char* __restrict__ comp,
                                                         const ParStream* __restrict__ streams,
                                                         ParCand* __restrict__ cand, const unsigned* __restrict__ list,
                                                         unsigned n, ParTables* __restrict__ tables,
                                                         unsigned short* __restrict__ cells, long long cells_base) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned ci = REAL ? list[i] : i;
  ParCand c = cand[ci];
  const ParStream st = streams[c.stream];
  ParTables& T = tables[i];
  TBits br;
  br.init(comp + st.in_off, (unsigned)st.in_len, c.bit);
  int nl, nd;
  unsigned fin;
  bool ok = par_parse_header(br, T.lens, nl, nd, fin, false);
  ok = ok && par_build<1>(T.lens, nl, T.ltab, PAR_LBITS, T.lsorted, T.lcount);
  ok = ok && par_build<2>(T.lens + nl, nd, T.dtab, PAR_DBITS, T.dsorted, T.dcount);
  const unsigned in_bits = (unsigned)st.in_len * 8;
  // output budget: a block cannot produce more than what is left of the stream
  const unsigned cap = REAL ? c.out_len : (unsigned)st.out_len;
  unsigned short* out = REAL ? cells + (st.out_off - cells_base) + c.out_off : nullptr;
  unsigned opos = 0;
  bool done = false;
  while (ok && !done) {
    br.refill();
    if (br.bit_pos() > in_bits) { ok = false; break; }
    unsigned win = br.window();
    unsigned e = T.ltab[win & ((1u << PAR_LBITS) - 1)];
    if ((e & 15) == 0) { e = par_slow<1>(win, T.lcount, T.lsorted); if (!e) { ok = false; break; } }
    const unsigned kind = e >> 24;
    unsigned val = (e >> 8) & 0xffff;
    if (kind == K_LIT) {
      br.drop(e & 15);
      if (opos >= cap) { ok = false; break; }
      if (REAL) out[opos] = (unsigned short)val;
      opos++;
      continue;
    }
    if (kind != K_LEN) {
      br.drop(e & 15);
      if (kind == K_EOB) done = true; else ok = false;
      break;
    }
    const unsigned cl = e & 15, xb = (e >> 4) & 15;
    val += (win >> cl) & ((1u << xb) - 1);
    br.drop(cl + xb);
    br.refill();
    win = br.window();
    unsigned e2 = T.dtab[win & ((1u << PAR_DBITS) - 1)];
    if ((e2 & 15) == 0) { e2 = par_slow<2>(win, T.dcount, T.dsorted); if (!e2) { ok = false; break; } }
    if ((e2 >> 24) != 0) { ok = false; break; }
    const unsigned cl2 = e2 & 15, xb2 = (e2 >> 4) & 15;
    const unsigned dist = ((e2 >> 8) & 0xffff) + ((win >> cl2) & ((1u << xb2) - 1));
    br.drop(cl2 + xb2);
    if (opos + val > cap || br.bit_pos() > in_bits) { ok = false; break; }
    if (REAL) {
      if (dist >= val && dist <= opos) {
        // non-overlapping copy inside the block: loads first, then stores, 8 cells at a time (one memory latency)
        const unsigned short* sp = out + (opos - dist);
        unsigned short* dp = out + opos;
        for (unsigned j0 = 0; j0 < val; j0 += 8) {
          unsigned short t[8];
#pragma unroll
          for (unsigned j = 0; j < 8; j++) t[j] = (j0 + j < val) ? sp[j0 + j] : (unsigned short)0;
#pragma unroll
          for (unsigned j = 0; j < 8; j++) if (j0 + j < val) dp[j0 + j] = t[j];
        }
      } else {
        for (unsigned j = 0; j < val; j++) {
          const int src = (int)(opos + j) - (int)dist;
          out[opos + j] = src >= 0 ? out[src] : (unsigned short)(0x8000u | (unsigned)(32768 + src));
        }
      }
    }
    opos += val;
  }
  if (!REAL) {
    c.end_bit = br.bit_pos();
    c.out_len = opos;
    c.state = (ok && done && c.end_bit <= in_bits) ? 1 : -1;
    cand[ci] = c;
  } else if (!(ok && done && opos == c.out_len)) {
    cand[ci].state = -2;      // cannot happen if the dry pass succeeded; the host falls back if it does
  }
}

// 5. Cells -> bytes, one CTA per stream, chained blocks in order.  chain[first[s] .. first[s+1]) = candidate indices.
__global__ void __launch_bounds__(1024) par_resolve_kernel(const ParStream* __restrict__ streams,
                                                           const ParCand* __restrict__ cand,
                                                           const unsigned* __restrict__ chain,
                                                           const unsigned* __restrict__ first,
                                                           const unsigned short* __restrict__ cells, long long cells_base,
                                                           unsigned char* __restrict__ out_base, int* __restrict__ bad) {
  const ParStream st = streams[blockIdx.x];
  unsigned char* out = out_base + st.out_off;
  const unsigned short* cl = cells + (st.out_off - cells_base);
  for (unsigned j = first[blockIdx.x]; j < first[blockIdx.x + 1]; j++) {
    const ParCand c = cand[chain[j]];
    const unsigned o = c.out_off, n = c.out_len;
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned v = cl[o + i];
      unsigned char b;
      if (v & 0x8000u) {
        const int src = (int)o - 32768 + (int)(v & 0x7fffu);
        if (src < 0) { bad[blockIdx.x] = 1; b = 0; } else b = out[src];
      } else b = (unsigned char)v;
      out[o + i] = b;
    }
    __syncthreads();   // the next block's markers read these bytes
  }
}

}  // namespace mts
