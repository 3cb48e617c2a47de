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
  unsigned tok_off;      // first token of the block inside its stream's token array (set on the host for chained blocks)
  int state;             // 1 = decoded cleanly, <= 0 = rejected
  unsigned final_;       // BFINAL
  unsigned n_tok;        // tokens (literals + matches) the block decodes to
};

struct ParLz {           // per stream, for par_lz_kernel
  long long tok_base;    // first token of the stream in the token buffer
  unsigned n_tok;        // tokens of its chained blocks
  unsigned out_len;      // bytes they must produce
};

// Decoding tables.  The fast tables live in SHARED memory, one column of 32-bit words per thread (word w of lane l at
// tabw[w * 32 + l]: every lane owns a bank, so 32 unrelated lookups never conflict); entries are 16 bits, two per word:
//   literal/length (9 index bits): code length | kind << 4 | (literal byte or length symbol - 257) << 6
//   distance       (7 index bits): code length | symbol << 4
// A zero entry means "longer code": see the limit words below; only the symbol order (sorted[]) is in global scratch.
// Longer codes: per code length l one word  limit | oend << 17  where limit = one past the last l-bit code, left-justified
// to 16 bits, and oend = number of symbols with codes of length <= l; the symbol is sorted[oend - ((limit - code16) >> (16-l))].
static const int PAR_LBITS = 9, PAR_DBITS = 7;
static const int PAR_LWORDS = (1 << PAR_LBITS) / 2, PAR_DWORDS = (1 << PAR_DBITS) / 2;
static const int PAR_LLONG = 16 - PAR_LBITS, PAR_DLONG = 16 - PAR_DBITS;      // limit words of lengths tb..15
static const int PAR_TWORDS = PAR_LWORDS + PAR_DWORDS + PAR_LLONG + PAR_DLONG;
static const int PAR_DEC_THREADS = 32;       // one warp per CTA: 42 KB of tables, 5 CTAs per SM
struct ParTables {
  unsigned short lsorted[288], dsorted[32];
  unsigned short lcount[16], dcount[16];
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

// Fast table (this thread's shared-memory column starting at word w0) + canonical arrays for one code (serial, per
// thread).  KIND 1 literal/length, 2 distance.
template <int KIND>
__device__ bool par_build(const unsigned char* lens, int n, unsigned* tabw, unsigned lane, int w0, int tb, int wlong,
                          unsigned short* sorted, unsigned short* count) {
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
    if (l >= tb) tabw[(wlong + l - tb) * 32 + lane] = ((code + count[l]) << (16 - l)) | (o << 17);
  }
  for (int w = 0; w < (1 << tb) / 2; w++) tabw[(w0 + w) * 32 + lane] = 0;
  unsigned short* col = (unsigned short*)(tabw + w0 * 32 + lane);   // entry k at col[(k >> 1) * 64 + (k & 1)]
  for (int s = 0; s < n; s++) {
    const unsigned l = lens[s];
    if (!l) continue;
    sorted[offs[l]++] = (unsigned short)s;
    if ((int)l <= tb) {
      const unsigned r = __brev(first[l]++) >> (32 - l);
      unsigned e;
      if (KIND == 1) e = l | ((s < 256 ? (unsigned)K_LIT : s == 256 ? (unsigned)K_EOB : s < 286 ? (unsigned)K_LEN : (unsigned)K_BAD) << 4) |
                         ((unsigned)(s < 256 ? s : s > 256 ? s - 257 : 0) << 6);
      else e = l | ((unsigned)s << 4);
      for (unsigned k = r; k < (1u << tb); k += 1u << l) col[(k >> 1) * 64 + (k & 1)] = (unsigned short)e;
    } else first[l]++;
  }
  return true;
}

// Decode of a code longer than the fast table (tb bits): the limit words of lengths tb+1..15 are scanned (the first
// length whose limit exceeds the left-justified code), then one global load fetches the symbol.
template <int TB>
__device__ __forceinline__ bool par_long(unsigned win, const unsigned* tabw, unsigned lane, int wlong,
                                         const unsigned short* sorted, unsigned& sym, unsigned& nbits) {
  const unsigned c16 = __brev(win) >> 16;
  unsigned word = 0, prev = 0, l = 0;
  unsigned below = tabw[wlong * 32 + lane];                 // length TB: only its symbol count matters
#pragma unroll
  for (int j = 1; j <= 15 - TB; j++) {                      // the limits grow with the length: the first that fits
    const unsigned wj = tabw[(wlong + j) * 32 + lane];
    if (!l && c16 < (wj & 0x1ffffu)) { word = wj; prev = below; l = (unsigned)(TB + j); }
    below = wj;
  }
  if (!l) return false;
  const unsigned back = ((word & 0x1ffffu) - c16 - 1) >> (16 - l);      // codes between this one and the last of length l
  const unsigned idx = (word >> 17) - 1 - back;
  // idx must lie among the symbols of length l (an incomplete code leaves holes that are not codes)
  if ((int)idx < (int)(prev >> 17)) return false;
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
  const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;          // aligned word index
  const unsigned char* in = comp + st.in_off;
  const unsigned mis = (unsigned)((uintptr_t)in & 3);
  const unsigned* w = (const unsigned*)(in - mis);
  if (st.in_len < 16) return;
  const unsigned kmax = (mis + (unsigned)st.in_len - 1) >> 2;
  if (j > kmax) return;
  // stream bit of this word's bit 0 (may be negative for the first word); valid header bits: [16, 8 * (in_len - 12)]
  // (a dynamic header is at least 17 + 12 bits + two codes: the zlib trailer and the shortest block are ignored)
  const long long bit0 = 32ll * j - 8ll * mis;
  const long long first_ok = 16, last_ok = 8ll * ((long long)st.in_len - 12) + 7;
  if (bit0 + 31 < first_ok || bit0 > last_ok) return;
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
    c.stream = sidx; c.bit = bit; c.end_bit = 0; c.out_len = 0; c.tok_off = 0; c.state = 0; c.final_ = fin; c.n_tok = 0;
    cand[at] = c;
  }
}

// 3/4. Thread per block, one warp per CTA, fast tables in shared memory.  REAL = 0: dry decode of every candidate (end
// bit, output length, token count).  REAL = 1: decode the chained blocks (list[] holds their candidate indices) into
// tokens; the back-references are resolved afterwards by par_lz_kernel, in stream order.
template <int REAL>
__global__ void __launch_bounds__(PAR_DEC_THREADS) par_decode_kernel(const unsigned char* __restrict__ comp,
                                                                     const ParStream* __restrict__ streams,
                                                                     ParCand* __restrict__ cand,
                                                                     const unsigned* __restrict__ list, unsigned n,
                                                                     ParTables* __restrict__ tables,
                                                                     unsigned* __restrict__ tokens,
                                                                     const ParLz* __restrict__ lz) {
  __shared__ unsigned tabw[PAR_TWORDS * 32];
  __shared__ unsigned short lenx[32];
  __shared__ unsigned distx[32];
  const unsigned lane = threadIdx.x;
  lenx[lane] = (unsigned short)par_len_info(lane);
  distx[lane] = par_dist_info(lane);
  __syncwarp();
  const unsigned i = blockIdx.x * PAR_DEC_THREADS + lane;
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
  const int wll = PAR_LWORDS + PAR_DWORDS, wdl = wll + PAR_LLONG;       // limit words of the two codes
  ok = ok && par_build<1>(T.lens, nl, tabw, lane, 0, PAR_LBITS, wll, T.lsorted, T.lcount);
  ok = ok && par_build<2>(T.lens + nl, nd, tabw, lane, PAR_LWORDS, PAR_DBITS, wdl, T.dsorted, T.dcount);
  const unsigned short* lcol = (const unsigned short*)(tabw + lane);
  const unsigned short* dcol = (const unsigned short*)(tabw + PAR_LWORDS * 32 + lane);
  const unsigned in_bits = (unsigned)st.in_len * 8;
  // output budget: a block cannot produce more than what is left of the stream
  const unsigned cap = REAL ? c.out_len : (unsigned)st.out_len;
  unsigned* tok = REAL ? tokens + lz[c.stream].tok_base + c.tok_off : nullptr;
  unsigned opos = 0, ntok = 0;
  bool done = false;
  while (ok && !done) {
    br.refill();
    if (br.bit_pos() > in_bits) { ok = false; break; }
    unsigned win = br.window();
    unsigned k = win & ((1u << PAR_LBITS) - 1);
    unsigned e = lcol[(k >> 1) * 64 + (k & 1)];
    unsigned cl = e & 15, kind = (e >> 4) & 3, val = e >> 6;
    if (cl == 0) {
      unsigned sym;
      if (!par_long<PAR_LBITS>(win, tabw, lane, wll, T.lsorted, sym, cl)) { ok = false; break; }
      kind = sym < 256 ? (unsigned)K_LIT : sym == 256 ? (unsigned)K_EOB : sym < 286 ? (unsigned)K_LEN : (unsigned)K_BAD;
      val = sym < 256 ? sym : sym > 256 ? sym - 257 : 0;
    }
    if (kind == K_LIT) {
      br.drop(cl);
      if (opos >= cap) { ok = false; break; }
      if (REAL) tok[ntok] = val;
      ntok++; opos++;
      continue;
    }
    if (kind != K_LEN) {
      br.drop(cl);
      if (kind == K_EOB) done = true; else ok = false;
      break;
    }
    const unsigned lx = lenx[val];
    const unsigned xb = lx >> 12;
    const unsigned len = (lx & 0xfffu) + ((win >> cl) & ((1u << xb) - 1));
    br.drop(cl + xb);
    br.refill();
    win = br.window();
    k = win & ((1u << PAR_DBITS) - 1);
    const unsigned e2 = dcol[(k >> 1) * 64 + (k & 1)];
    unsigned cl2 = e2 & 15, dsym = e2 >> 4;
    if (cl2 == 0 && !par_long<PAR_DBITS>(win, tabw, lane, wdl, T.dsorted, dsym, cl2)) { ok = false; break; }
    const unsigned dx = distx[dsym & 31];
    if (dx == 0) { ok = false; break; }
    const unsigned xb2 = dx >> 16;
    const unsigned dist = (dx & 0xffffu) + ((win >> cl2) & ((1u << xb2) - 1));
    br.drop(cl2 + xb2);
    if (opos + len > cap || br.bit_pos() > in_bits) { ok = false; break; }
    if (REAL) tok[ntok] = 0x80000000u | (len << 16) | (dist - 1);
    ntok++; opos += len;
  }
  if (!REAL) {
    c.end_bit = br.bit_pos();
    c.out_len = opos;
    c.n_tok = ntok;
    c.state = (ok && done && c.end_bit <= in_bits) ? 1 : -1;
    cand[ci] = c;
  } else if (!(ok && done && opos == c.out_len && ntok == c.n_tok)) {
    cand[ci].state = -2;      // cannot happen if the dry pass succeeded; the host falls back if it does
  }
}

// 5. Tokens -> bytes: one CTA per stream walks the stream's tokens in order, one tile of up to PAR_LZ_THREADS tokens /
//    PAR_LZ_CAP output bytes at a time (thread = token).  A block-wide scan of the token lengths gives every token its
//    position; the tile's output is assembled in shared memory and then stored coalesced.  Literals and the bytes that
//    matches copy from before the tile (global memory, written by earlier tiles) are placed at once; bytes copied from
//    inside the tile wait, without block barriers, until the PENDING bitmap (one bit per staged byte, set by the
//    matches that still have to produce it) is clear over their source range.
static const int PAR_LZ_THREADS = 256;
static const int PAR_LZ_CAP = 4096;
__device__ __forceinline__ unsigned par_bits(unsigned a, unsigned b, unsigned w) {   // bits of [a, b) that fall in word w
  const unsigned lo = max(a, w * 32), hi = min(b, w * 32 + 32);
  if (hi <= lo) return 0;
  return (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1) << (lo & 31));
}
__global__ void __launch_bounds__(PAR_LZ_THREADS) par_lz_kernel(const ParStream* __restrict__ streams,
                                                                const ParLz* __restrict__ lz,
                                                                const unsigned* __restrict__ tokens,
                                                                unsigned char* out_base, int* __restrict__ bad) {
  const int NT = PAR_LZ_THREADS;
  __shared__ unsigned ob_w[PAR_LZ_CAP / 4];
  __shared__ unsigned pend_w[PAR_LZ_CAP / 32];
  __shared__ unsigned wsum[NT / 32];
  __shared__ unsigned s_total;
  unsigned char* ob = (unsigned char*)ob_w;
  volatile unsigned* pend = pend_w;
  const ParStream st = streams[blockIdx.x];
  const ParLz z = lz[blockIdx.x];
  unsigned char* out = out_base + st.out_off;
  const unsigned* tk = tokens + z.tok_base;
  const unsigned T = z.n_tok;
  const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (unsigned i = tid; i < PAR_LZ_CAP / 32; i += NT) pend_w[i] = 0;
  unsigned obase = 0, t0 = 0;
  bool fail = false;
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
    unsigned woff = 0;
    for (int w = 0; w < NT / 32; w++) { const unsigned v = wsum[w]; if (w < (int)wid) woff += v; }
    const unsigned rel = woff + incl - L;                           // position inside the tile
    const bool fits = has && rel + L <= (unsigned)PAR_LZ_CAP;       // monotone: the tile is the longest fitting prefix
    const unsigned ncut = (unsigned)__syncthreads_count(fits);
    if (fits && tid + 1 == ncut) s_total = rel + L;
    const unsigned dist = (t & 0x7fffu) + 1;
    bool m = fits && isM;
    if (m && dist > obase + rel) { fail = true; m = false; }
    if (fits && obase + rel + L > z.out_len) { fail = true; m = false; }
    else if (fits && !isM) ob[rel] = (unsigned char)t;
    const int srel = (int)rel - (int)dist;                          // source position inside the tile (negative: before it)
    unsigned n_old = 0;
    if (m) {
      n_old = srel < 0 ? min(L, (unsigned)(-srel)) : 0u;
      if (n_old < L) {                                              // has an in-tile part: its bytes are pending
        for (unsigned w = rel >> 5; w <= (rel + L - 1) >> 5; w++) atomicOr(&pend_w[w], par_bits(rel, rel + L, w));
      }
      const unsigned char* sp = out + obase + srel;                 // bytes from before the tile
      for (unsigned j0 = 0; j0 < n_old; j0 += 8) {
        unsigned char v[8];
#pragma unroll
        for (unsigned j = 0; j < 8; j++) v[j] = (j0 + j < n_old) ? sp[j0 + j] : (unsigned char)0;
#pragma unroll
        for (unsigned j = 0; j < 8; j++) if (j0 + j < n_old) ob[rel + j0 + j] = v[j];
      }
      if (n_old == L) m = false;
    }
    __syncthreads();                                                // literals, old bytes, pending bits and s_total are visible
    const unsigned a = (unsigned)max(srel, 0), b = min((unsigned)(srel + (int)L), rel);   // in-tile source bytes outside my own output
    while (__any_sync(0xffffffffu, m)) {
      if (m) {
        bool clear = true;
        if (b > a) for (unsigned w = a >> 5; w <= (b - 1) >> 5; w++) if (pend[w] & par_bits(a, b, w)) { clear = false; break; }
        if (clear) {
          __threadfence_block();
          for (unsigned j = n_old; j < L; j++) ob[rel + j] = ob[(unsigned)(srel + (int)j)];   // in order: may read my own bytes
          __threadfence_block();
          for (unsigned w = rel >> 5; w <= (rel + L - 1) >> 5; w++) atomicAnd(&pend_w[w], ~par_bits(rel, rel + L, w));
          m = false;
        }
      }
    }
    __syncthreads();
    const unsigned total = s_total;
    for (unsigned i = tid; i < total; i += NT) out[obase + i] = ob[i];
    obase += total;
    t0 += ncut;
    if (ncut != (unsigned)NT && t0 < T) nxt = t0 + tid < T ? tk[t0 + tid] : 0;   // the tile was cut short: reload
  }
  if (__syncthreads_or(fail) || obase != z.out_len) { if (tid == 0) bad[blockIdx.x] = 1; }
}

}  // namespace mts
