// inflate.cuh — K3: DEFLATE decoder (RFC 1951 / zlib container RFC 1950), one warp per stream.
//
// Replaces `zlib.decompress(cbuffer)` at mtscomp.py:619.  It accepts what zlib accepts for a `.cbin` chunk (any mix of
// stored / fixed / dynamic blocks, any declared window, trailing bytes ignored) and reports a status per
// stream; the adler32 found in the trailer is returned so the caller can compare it with the adler32 that the K4
// pass computes over the inflated bytes (a mismatch surfaces as the reference's "Compressed chunk #i is corrupted").
//
// A "stream" is either a whole chunk (reference-written files: strictly serial, block boundaries are only known by
// decoding) or one encoder segment of a GPU-written file (byte-aligned start, fresh window, listed in the in-band index).
//
// Warp organisation: all 32 lanes run the Huffman decode loop redundantly (uniform control flow, broadcast loads), lane
// k keeps the k-th symbol of a batch of 32.  Literals are then stored in parallel, matches that only read bytes produced
// before the batch are copied one per lane, the (few) others cooperatively in stream order.  The last 32 KB of output
// plus the batch in flight live in a shared-memory ring, finished bytes are streamed to HBM as aligned 32-bit words.
// The decode tables hold fully decoded entries (code length, extra-bit count, base value, kind) so the symbol loop has
// no per-symbol arithmetic on symbol numbers.
#pragma once
#include "common.cuh"

namespace mts {

struct InflateSeg {
  long long in_off;    // byte offset of the stream in the compressed buffer
  long long out_off;   // byte offset of its output in the transformed buffer
  int in_len;          // compressed bytes available
  int out_len;         // exact number of bytes the stream must produce
  int flags;           // INF_ZLIB: 2-byte zlib header + adler32 trailer, decode to the final block
                       // otherwise raw deflate blocks, stop once out_len bytes are produced at a block end
  unsigned start_bit;  // INF_RESUME: bit offset (from in_off) of the block header at which decoding resumes ...
  unsigned opos0;      // ... and the number of output bytes already produced (by the block-parallel path)
  int pad_;
};
enum { INF_ZLIB = 1, INF_RESUME = 2, INF_NO_BLOCKS = 4 };   // INF_NO_BLOCKS: the final block is already decoded
enum {
  INF_OK = 0, INF_BAD_HEADER = 1, INF_BAD_BLOCK = 2, INF_BAD_LENGTHS = 3, INF_BAD_CODE = 4, INF_BAD_DISTANCE = 5,
  INF_BAD_SIZE = 6, INF_INPUT_OVERRUN = 7, INF_BAD_STORED = 8, INF_BAD_ADLER = 9
};

static const int INF_LBITS = 10, INF_DBITS = 8;
static const unsigned INF_RING = 36864;        // 32768 history + the batch in flight (<= INF_BATCH_BYTES + 258), % 16 == 0
static const unsigned INF_BATCH_BYTES = 3800;   // a batch stops early once it has produced this much (keeps 5 CTAs / SM)

// Table entry: bits 0-3 code length (0 = not in the fast table), 4-7 extra-bit count, 8-23 base value, 24-25 kind.
enum { K_LIT = 0, K_LEN = 1, K_EOB = 2, K_BAD = 3 };
__device__ __forceinline__ unsigned ll_entry(unsigned s, unsigned l) {
  if (s < 256) return l | (s << 8) | (K_LIT << 24);
  if (s == 256) return l | (K_EOB << 24);
  if (s < 265) return l | ((s - 254) << 8) | (K_LEN << 24);
  if (s < 285) { unsigned nb = (s - 261) >> 2; return l | (nb << 4) | ((3 + ((4 + ((s - 257) & 3)) << nb)) << 8) | (K_LEN << 24); }
  if (s == 285) return l | (258u << 8) | (K_LEN << 24);
  return l | (K_BAD << 24);
}
__device__ __forceinline__ unsigned d_entry(unsigned s, unsigned l) {
  if (s < 4) return l | ((s + 1) << 8);
  if (s < 30) { unsigned nb = (s >> 1) - 1; return l | (nb << 4) | ((1 + ((2 + (s & 1)) << nb)) << 8); }
  return l | (K_BAD << 24);
}
__device__ __forceinline__ unsigned cl_entry(unsigned s, unsigned l) { return l | (s << 8); }

struct InflateWarpSmem {
  unsigned ltab[1 << INF_LBITS];
  unsigned dtab[1 << INF_DBITS];
  unsigned short lsorted[288];           // symbols in canonical order (slow path for codes longer than the table)
  unsigned short dsorted[32];
  unsigned short lcount[16], dcount[16];
  unsigned short run[16];
  unsigned char lens[352];               // [0,19) code-length code, [24, 24+316) literal/length + distance lengths
};

// LSB-first bit reader over 32-bit words: (lo, hi) hold 64 stream bits, `pos` (< 32 after refill()) is the cursor in
// them, so 32 bits are always available right after a refill and peek() is one funnel shift.
struct BitR {
  const unsigned* w;      // 4-byte aligned base
  unsigned sh;            // 8 * (address & 3)
  unsigned kmax;          // last valid word index
  unsigned k;             // index of the next raw word to fetch
  unsigned raw;           // w[k - 1]: raw word whose upper part belongs to the next stream word
  unsigned lo, hi;        // stream words n and n+1
  unsigned pos;           // bit cursor inside (lo, hi)
  unsigned ip;            // stream bytes represented by lo and everything before it, + 4 (i.e. end of lo)
  __device__ __forceinline__ unsigned next_word() {
    unsigned nw = w[min(k, kmax)];
    unsigned v = __funnelshift_r(raw, nw, sh);
    raw = nw;
    k++;
    return v;
  }
  // start reading at `p` (= stream start + ip0) with `avail` bytes left
  __device__ __forceinline__ void init(const unsigned char* p, unsigned avail, unsigned ip0) {
    unsigned mis = (unsigned)((uintptr_t)p & 3);
    w = (const unsigned*)(p - mis);
    sh = mis * 8;
    kmax = (mis + max(avail, 1u) - 1) >> 2;
    raw = w[0];
    k = 1;
    lo = next_word();
    hi = next_word();
    pos = 0;
    ip = ip0 + 4;
  }
  __device__ __forceinline__ void refill() {
    if (pos >= 32) { lo = hi; hi = next_word(); pos -= 32; ip += 4; }
  }
  __device__ __forceinline__ unsigned window() const { return __funnelshift_r(lo, hi, pos); }   // next 32 bits
  __device__ __forceinline__ unsigned peek(unsigned n) const { return window() & ((1u << n) - 1); }
  __device__ __forceinline__ void drop(unsigned n) { pos += n; }
  __device__ __forceinline__ unsigned get(unsigned n) { refill(); unsigned v = peek(n); drop(n); return v; }   // n <= 32
  __device__ __forceinline__ void align_byte() { pos = (pos + 7) & ~7u; }
  __device__ __forceinline__ unsigned byte_pos() const { return ip - 4 + ((pos + 7) >> 3); }   // bytes consumed
};

// Build one decoding table from code lengths lens[0..n): fast table of `tb` bits + canonical arrays for longer codes.
// KIND selects the entry encoder (0 code-length code, 1 literal/length, 2 distance).  Returns false if the lengths are
// over-subscribed.  All lanes participate.
template <int KIND>
__device__ bool inflate_build(const unsigned char* lens, int n, unsigned* tab, int tb, unsigned short* sorted,
                              unsigned short* count, unsigned short* run) {
  const unsigned lane = lane_id();
  if (lane < 16) { count[lane] = 0; run[lane] = 0; }
  for (int i = lane; i < (1 << tb); i += 32) tab[i] = 0;
  __syncwarp();
  for (int s0 = 0; s0 < n; s0 += 32) {
    int s = s0 + lane;
    unsigned l = (s < n) ? lens[s] : 0;
    unsigned grp = __match_any_sync(0xffffffffu, l);
    if (l && (grp >> lane) == 1u) count[l] += (unsigned short)__popc(grp);
    __syncwarp();
  }
  // first canonical code and first sorted index of each length (every lane computes the same 15 values)
  unsigned first[16], offs[16];
  unsigned code = 0, o = 0;
  int left = 1;
  first[0] = 0; offs[0] = 0;
  for (int l = 1; l <= 15; l++) {
    code = (code + (l > 1 ? count[l - 1] : 0)) << 1;
    first[l] = code;
    offs[l] = o;
    o += count[l];
    left = (left << 1) - count[l];
    if (left < 0) return false;
  }
  for (int s0 = 0; s0 < n; s0 += 32) {
    int s = s0 + lane;
    unsigned l = (s < n) ? lens[s] : 0;
    unsigned grp = __match_any_sync(0xffffffffu, l);
    if (l) {
      unsigned rank = run[l] + __popc(grp & ((1u << lane) - 1));
      sorted[offs[l] + rank] = (unsigned short)s;
      if ((int)l <= tb) {
        unsigned r = __brev(first[l] + rank) >> (32 - l);
        unsigned e = KIND == 1 ? ll_entry((unsigned)s, l) : KIND == 2 ? d_entry((unsigned)s, l) : cl_entry((unsigned)s, l);
        for (unsigned k = r; k < (1u << tb); k += 1u << l) tab[k] = e;
      }
    }
    __syncwarp();
    if (l && (grp >> lane) == 1u) run[l] += (unsigned short)__popc(grp);
    __syncwarp();
  }
  return true;
}

// Slow path: canonical decode of a code longer than the fast table (puff-style, one bit at a time); returns the
// symbol's table entry, 0 if no code matches.
template <int KIND>
__device__ __noinline__ unsigned inflate_slow(unsigned bb, const unsigned short* count,
                                              const unsigned short* sorted) {
  unsigned code = 0, first = 0, index = 0;
  for (int l = 1; l <= 15; l++) {
    code |= (bb >> (l - 1)) & 1;
    unsigned c = count[l];
    if (code - first < c) {
      unsigned s = sorted[index + (code - first)];
      return KIND == 1 ? ll_entry(s, (unsigned)l) : d_entry(s, (unsigned)l);
    }
    index += c;
    first = (first + c) << 1;
    code <<= 1;
  }
  return 0;
}

__device__ __forceinline__ void fixed_lengths(unsigned char* lens, unsigned lane) {
  for (int s = lane; s < 288; s += 32) lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
  lens[288 + lane] = (lane < 30) ? 5 : 0;
}

// Window policies.  SMEM: the last 32 KB of output + the batch in flight live in a shared-memory ring (index of
// position p = (p + a0) mod INF_RING, a0 = output address & 3, so aligned ring words are aligned output words) and are
// streamed to HBM after every batch: right for few long streams (reference-written chunks).  GLOBAL: bytes go straight
// to the output buffer and back-references read it (L1/L2): no per-warp shared window, so 64 warps per SM can hide the
// latency: right for many short streams (the segments of GPU-written chunks).
template <bool SMEM> struct Win {
  unsigned char* base;
  __device__ __forceinline__ unsigned index(unsigned pos, unsigned a0) const { return SMEM ? (pos + a0) % INF_RING : pos; }
  __device__ __forceinline__ unsigned wrap(unsigned i) const { return (SMEM && i >= INF_RING) ? i - INF_RING : i; }
  __device__ __forceinline__ unsigned back(unsigned i, unsigned d) const {
    return (SMEM && i < d) ? i + INF_RING - d : i - d;
  }
};

__device__ __forceinline__ unsigned ring_wrap(unsigned i) { return i >= INF_RING ? i - INF_RING : i; }

// Stream bytes [from, to) of the output (ring -> global).
__device__ __forceinline__ void inflate_flush(const unsigned char* ring, unsigned char* out, unsigned a0, unsigned from,
                                              unsigned to, unsigned lane) {
  if (to <= from) return;
  unsigned g0 = from + a0, g1 = to + a0;                 // coordinates relative to the aligned word base of out
  unsigned w0 = (g0 + 3) & ~3u, w1 = g1 & ~3u;
  unsigned char* gb = out - a0;
  if (w0 >= w1) {                                        // no full word: bytes only
    for (unsigned g = g0 + lane; g < g1; g += 32) gb[g] = ring[g % INF_RING];
    return;
  }
  if (lane < w0 - g0) gb[g0 + lane] = ring[(g0 + lane) % INF_RING];
  unsigned r = (w0 + 4 * lane) % INF_RING;
  for (unsigned g = w0 + 4 * lane; g < w1; g += 128) {
    *(unsigned*)(gb + g) = *(const unsigned*)(ring + r);
    r = ring_wrap(r + 128);
  }
  if (lane < g1 - w1) gb[w1 + lane] = ring[(w1 + lane) % INF_RING];
}

template <bool SMEM, int WPC>
__global__ void __launch_bounds__(32 * WPC) inflate_kernel(const unsigned char* __restrict__ comp,
                                                           const InflateSeg* __restrict__ segs, int n_segs,
                                                           unsigned char* out_base, int* __restrict__ status,
                                                           unsigned* __restrict__ trailer_adler) {
  __shared__ InflateWarpSmem S_all[WPC];
  __shared__ __align__(16) unsigned char ring_all[SMEM ? INF_RING * WPC : 16];
  const unsigned lane = lane_id();
  const int sidx = blockIdx.x * WPC + (int)warp_id();
  if (sidx >= n_segs) return;
  InflateWarpSmem& S = S_all[warp_id()];
  const InflateSeg sg = segs[sidx];
  unsigned char* out = out_base + sg.out_off;
  const unsigned out_len = (unsigned)sg.out_len;
  if ((sg.flags & INF_RESUME) && !(sg.flags & INF_ZLIB) && sg.opos0 >= out_len) {   // an indexed segment, already complete
    if (lane == 0) status[sidx] = INF_OK;
    return;
  }
  const unsigned char* in = comp + sg.in_off;
  const unsigned in_len = (unsigned)sg.in_len;
  const unsigned a0 = SMEM ? (unsigned)((uintptr_t)out & 3) : 0;
  Win<SMEM> W;
  W.base = SMEM ? ring_all + (SMEM ? INF_RING * warp_id() : 0) : out;
  unsigned char* const win = W.base;

  BitR br;
  int err = INF_OK;
  unsigned opos = 0;            // bytes produced so far (all flushed at batch boundaries)
  unsigned adler = 0;
  bool last = false;

  if (sg.flags & INF_ZLIB) {
    if (in_len < 8) err = INF_BAD_HEADER;
    else {
      const unsigned cmf = in[0], flg = in[1];
      if ((cmf & 15) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20)) err = INF_BAD_HEADER;
    }
  }
  if (sg.flags & INF_RESUME) {
    // the block-parallel path has produced out[0, opos0): continue at the block header at start_bit
    const unsigned sb = min(sg.start_bit >> 3, in_len);
    br.init(in + sb, in_len - sb, sb);
    br.pos = sg.start_bit & 7;
    opos = min(sg.opos0, out_len);
    last = (sg.flags & INF_NO_BLOCKS) != 0;
    if (SMEM) {
      const unsigned hist = min(opos, 32768u);
      for (unsigned i = lane; i < hist; i += 32) win[W.index(opos - 1 - i, a0)] = out[opos - 1 - i];
      __syncwarp();
    }
  } else {
    br.init(in, in_len, 0);
    if (sg.flags & INF_ZLIB) br.pos = 16;
  }
  while (!err && !last) {
    if (!(sg.flags & INF_ZLIB) && opos >= out_len) break;
    br.refill();
    last = br.get(1) != 0;
    unsigned type = br.get(2);
    if (type == 0) {
      // stored: skip to the byte boundary, LEN / NLEN, raw copy through the window (later blocks may reference it)
      br.align_byte();
      br.refill();
      unsigned len = br.get(16);
      br.refill();
      unsigned nlen = br.get(16);
      if ((len ^ 0xffffu) != nlen) { err = INF_BAD_STORED; break; }
      unsigned src = br.byte_pos();
      if (src + len > in_len) { err = INF_INPUT_OVERRUN; break; }
      if (opos + len > out_len) { err = INF_BAD_SIZE; break; }
      for (unsigned b0 = 0; b0 < len; b0 += 4096) {
        unsigned m = min(4096u, len - b0);
        for (unsigned i = lane; i < m; i += 32) win[W.index(opos + i, a0)] = in[src + b0 + i];
        __syncwarp();
        if (SMEM) { inflate_flush(win, out, a0, opos, opos + m, lane); __syncwarp(); }
        opos += m;
      }
      br.init(in + src + len, in_len - (src + len), src + len);   // restart the bit reader after the stored bytes
      continue;
    }
    if (type == 3) { err = INF_BAD_BLOCK; break; }
    int nl = 288, nd = 32;
    if (type == 1) {
      fixed_lengths(S.lens, lane);
    } else {
      br.refill();
      nl = (int)br.get(5) + 257;
      nd = (int)br.get(5) + 1;
      int ncl = (int)br.get(4) + 4;
      if (nl > 286 || nd > 30) { err = INF_BAD_LENGTHS; break; }
      const unsigned char order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      if (lane < 19) S.lens[lane] = 0;
      __syncwarp();
      for (int i = 0; i < ncl; i++) {
        br.refill();
        unsigned v = br.get(3);
        if (lane == 0) S.lens[order[i]] = (unsigned char)v;
      }
      __syncwarp();
      // the code-length code is decoded with the distance-table slots (7-bit table)
      if (!inflate_build<0>(S.lens, 19, S.dtab, 7, S.dsorted, S.dcount, S.run)) { err = INF_BAD_LENGTHS; break; }
      __syncwarp();
      int idx = 0;
      unsigned prev_len = 0;
      while (idx < nl + nd) {
        br.refill();
        unsigned e = S.dtab[br.peek(7)];
        if (!e) { err = INF_BAD_LENGTHS; break; }
        br.drop(e & 15);
        unsigned sym = e >> 8;
        unsigned rep = 1, val = sym;
        if (sym == 16) {
          if (idx == 0) { err = INF_BAD_LENGTHS; break; }
          val = prev_len; rep = 3 + br.get(2);
        } else if (sym == 17) { val = 0; rep = 3 + br.get(3); }
        else if (sym == 18) { val = 0; rep = 11 + br.get(7); }
        if (idx + (int)rep > nl + nd) { err = INF_BAD_LENGTHS; break; }
        for (unsigned k = lane; k < rep; k += 32) S.lens[24 + idx + k] = (unsigned char)val;
        idx += (int)rep;
        prev_len = val;
      }
      if (err) break;
      __syncwarp();
      if (S.lens[24 + 256] == 0) { err = INF_BAD_LENGTHS; break; }   // no end-of-block code
    }
    __syncwarp();
    const unsigned char* ll = (type == 1) ? S.lens : S.lens + 24;
    const unsigned char* dl = (type == 1) ? S.lens + 288 : S.lens + 24 + nl;
    if (!inflate_build<1>(ll, nl, S.ltab, INF_LBITS, S.lsorted, S.lcount, S.run)) { err = INF_BAD_LENGTHS; break; }
    __syncwarp();
    if (!inflate_build<2>(dl, nd, S.dtab, INF_DBITS, S.dsorted, S.dcount, S.run)) { err = INF_BAD_LENGTHS; break; }
    __syncwarp();

    // ---- symbol loop: batches of up to 32 symbols / INF_BATCH_BYTES + 258 output bytes (the ring has room for that)
    bool eob = false;
    while (!eob && !err) {
      unsigned my_pos = 0, my_tok = 0;          // my_tok: literal byte, or len << 16 | dist
      unsigned bpos = opos;
      unsigned bad = 0;                          // deferred checks (distance too far / invalid symbol)
      int k = 0;
      const unsigned blimit = opos + INF_BATCH_BYTES;
      for (; k < 32 && bpos < blimit; k++) {
        br.refill();
        unsigned win32 = br.window();
        unsigned e = S.ltab[win32 & ((1u << INF_LBITS) - 1)];
        if ((e & 15) == 0) { e = inflate_slow<1>(win32, S.lcount, S.lsorted); if (!e) { err = INF_BAD_CODE; break; } }
        const unsigned kind = e >> 24;
        unsigned val = (e >> 8) & 0xffff;
        if (kind == K_LIT) {
          br.drop(e & 15);
          if (lane == (unsigned)k) { my_pos = bpos; my_tok = val; }
          bpos += 1;
          continue;
        }
        if (kind != K_LEN) {
          br.drop(e & 15);
          if (kind == K_EOB) eob = true; else err = INF_BAD_CODE;
          break;
        }
        const unsigned cl = e & 15, xb = (e >> 4) & 15;
        val += (win32 >> cl) & ((1u << xb) - 1);
        br.drop(cl + xb);
        br.refill();
        win32 = br.window();
        unsigned e2 = S.dtab[win32 & ((1u << INF_DBITS) - 1)];
        if ((e2 & 15) == 0) { e2 = inflate_slow<2>(win32, S.dcount, S.dsorted); if (!e2) { err = INF_BAD_CODE; break; } }
        const unsigned cl2 = e2 & 15, xb2 = (e2 >> 4) & 15;
        const unsigned dist = ((e2 >> 8) & 0xffff) + ((win32 >> cl2) & ((1u << xb2) - 1));
        br.drop(cl2 + xb2);
        bad |= (e2 >> 24) | (unsigned)(dist > bpos);
        if (lane == (unsigned)k) { my_pos = bpos; my_tok = (val << 16) | dist; }
        bpos += val;
      }
      if (err) break;
      if (bad) { err = (bad & 2) ? INF_BAD_CODE : INF_BAD_DISTANCE; break; }
      if (bpos > out_len) { err = INF_BAD_SIZE; break; }
      if (br.byte_pos() > in_len + 8) { err = INF_INPUT_OVERRUN; break; }
      const bool have = lane < (unsigned)k;
      const unsigned my_len = my_tok >> 16, my_dist = my_tok & 0xffff;
      const bool is_match = have && my_len != 0;
      // a match is independent of this batch's other symbols if everything it reads was produced before the batch
      const bool indep = is_match && (my_pos - my_dist + min(my_len, my_dist) <= opos);
      const unsigned rp = W.index(my_pos, a0);
      if (have && my_len == 0) win[rp] = (unsigned char)my_tok;
      if (indep) {
        unsigned rs = W.back(rp, my_dist);
        unsigned rd = rp;
        for (unsigned i = 0; i < my_len; i++) {
          win[rd] = win[rs];
          rd = W.wrap(rd + 1);
          rs = W.wrap(rs + 1);
        }
      }
      __syncwarp();
      unsigned mm = __ballot_sync(0xffffffffu, is_match && !indep);
      while (mm) {
        int src_lane = __ffs((int)mm) - 1;
        mm &= mm - 1;
        unsigned p = __shfl_sync(0xffffffffu, rp, src_lane);
        unsigned t = __shfl_sync(0xffffffffu, my_tok, src_lane);
        unsigned l = t >> 16, d = t & 0xffff;
        unsigned s = W.back(p, d);
        if (d >= l) { for (unsigned i = lane; i < l; i += 32) win[W.wrap(p + i)] = win[W.wrap(s + i)]; }
        else if (d >= 32) {
          // overlapping copy with period d >= 32: rounds of 32 bytes never read what the same round writes
          for (unsigned i0 = 0; i0 < l; i0 += 32) {
            unsigned i = i0 + lane;
            if (i < l) win[W.wrap(p + i)] = win[W.wrap(s + i)];
            __syncwarp();
          }
        } else { for (unsigned i = lane; i < l; i += 32) win[W.wrap(p + i)] = win[W.wrap(s + i % d)]; }
        __syncwarp();
      }
      if (SMEM) { inflate_flush(win, out, a0, opos, bpos, lane); __syncwarp(); }
      opos = bpos;
    }
  }
  if (!err) {
    if (opos != out_len) err = INF_BAD_SIZE;
    else if (sg.flags & INF_ZLIB) {
      br.align_byte();
      unsigned a = 0;
      for (int i = 0; i < 4; i++) { br.refill(); a = (a << 8) | br.get(8); }
      adler = a;
      if (br.byte_pos() > in_len) err = INF_INPUT_OVERRUN;
    } else if (br.byte_pos() > in_len) err = INF_INPUT_OVERRUN;
  }
  if (lane == 0) { status[sidx] = err; trailer_adler[sidx] = adler; }
}

}  // namespace mts
