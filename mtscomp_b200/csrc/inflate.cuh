// inflate.cuh — K3: DEFLATE decoder (RFC 1951 / zlib container RFC 1950), one warp per stream.
//
// Replaces `zlib.decompress(cbuffer)` at mtscomp.py:619.  It accepts what zlib accepts for a `.cbin` chunk (any mix of
// stored / fixed / dynamic blocks, any declared window, trailing bytes ignored) and reports a status per
// stream; the adler32 found in the trailer is returned so the caller can compare it with the adler32 that the K4
// pass computes over the inflated bytes (a mismatch surfaces as the reference's "Compressed chunk #i is corrupted").
//
// A "stream" is either a whole chunk (reference-written files: strictly serial, block boundaries are only known by
// decoding) or one encoder segment of a GPU-written file (byte-aligned start, fresh window, listed in the side index).
//
// Warp organisation: all 32 lanes run the Huffman decode loop redundantly (uniform control flow, broadcast loads), lane
// k keeps the k-th symbol of a batch of 32; literals are then stored in parallel and matches are copied cooperatively in
// stream order.  Back-references read the output buffer itself (the 32 KB window lives in L1/L2).
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
  int pad_;
};
enum { INF_ZLIB = 1 };
enum {
  INF_OK = 0, INF_BAD_HEADER = 1, INF_BAD_BLOCK = 2, INF_BAD_LENGTHS = 3, INF_BAD_CODE = 4, INF_BAD_DISTANCE = 5,
  INF_BAD_SIZE = 6, INF_INPUT_OVERRUN = 7, INF_BAD_STORED = 8, INF_BAD_ADLER = 9
};

static const int INF_LBITS = 10, INF_DBITS = 8;

struct InflateWarpSmem {
  unsigned short ltab[1 << INF_LBITS];   // sym << 4 | len, 0 = not in the fast table
  unsigned short dtab[1 << INF_DBITS];
  unsigned short lsorted[288];           // symbols in canonical order (slow path for codes longer than the table)
  unsigned short dsorted[32];
  unsigned short lcount[16], dcount[16];
  unsigned short run[16];
  unsigned char lens[352];             // [0,19) code-length code, [24, 24+316) literal/length + distance lengths
};

struct BitR {
  const unsigned* w;     // 4-byte aligned base
  unsigned mis;          // byte offset of the stream inside w
  unsigned kmax;         // last valid word index
  unsigned long long bb; // bit buffer
  unsigned bc;           // valid bits in bb
  unsigned ip;           // next unread byte of the stream
  __device__ __forceinline__ unsigned load32(unsigned byte_idx) const {
    unsigned a = byte_idx + mis, k = a >> 2;
    unsigned lo = w[min(k, kmax)], hi = w[min(k + 1, kmax)];
    return __funnelshift_r(lo, hi, (a & 3) * 8);
  }
  __device__ __forceinline__ void refill() {
    if (bc <= 32) { bb |= (unsigned long long)load32(ip) << bc; ip += 4; bc += 32; }
  }
  __device__ __forceinline__ unsigned peek(unsigned n) const { return (unsigned)bb & ((1u << n) - 1); }
  __device__ __forceinline__ void drop(unsigned n) { bb >>= n; bc -= n; }
  __device__ __forceinline__ unsigned get(unsigned n) { unsigned v = peek(n); drop(n); return v; }
  __device__ __forceinline__ unsigned byte_pos() const { return ip - (bc >> 3); }   // bytes fully or partly consumed
};

// Build one decoding table from code lengths lens[0..n): fast table of `tb` bits + canonical arrays for longer codes.
// Returns false if the lengths are over-subscribed.  All lanes participate.
__device__ bool inflate_build(const unsigned char* lens, int n, unsigned short* tab, int tb, unsigned short* sorted,
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
        unsigned short e = (unsigned short)((s << 4) | l);
        for (unsigned k = r; k < (1u << tb); k += 1u << l) tab[k] = e;
      }
    }
    __syncwarp();
    if (l && (grp >> lane) == 1u) run[l] += (unsigned short)__popc(grp);
    __syncwarp();
  }
  return true;
}

// Slow path: canonical decode of a code longer than the fast table (puff-style, one bit at a time).
__device__ __forceinline__ int inflate_slow(BitR& br, const unsigned short* count, const unsigned short* sorted,
                                            unsigned& sym) {
  unsigned code = 0, first = 0, index = 0;
  for (int l = 1; l <= 15; l++) {
    code |= (unsigned)(br.bb >> (l - 1)) & 1;
    unsigned c = count[l];
    if (code - first < c) { sym = sorted[index + (code - first)]; br.drop(l); return l; }
    index += c;
    first = (first + c) << 1;
    code <<= 1;
  }
  return 0;
}

__device__ __forceinline__ void fixed_lengths(unsigned char* lens, unsigned lane) {
  for (int s = lane; s < 288; s += 32) lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
  if (lane < 32) lens[288 + lane] = (lane < 30) ? 5 : 0;
}

// Output window: the last 32 KB of output plus the batch being produced live in a shared-memory ring, so that
// back-references never touch global memory; finished bytes are streamed to HBM as aligned 32-bit words.
static const unsigned INF_RING = 40960;        // 32768 history + up to 8192 bytes of the batch in flight
static const unsigned INF_BATCH_BYTES = 8192 - 258;

__device__ __forceinline__ unsigned ring_wrap(unsigned i) { return i >= INF_RING ? i - INF_RING : i; }

// Stream bytes [from, to) of the output (ring -> global).  a0 = (address of out) & 3; ring index of position p is
// (p + a0) mod INF_RING with INF_RING % 4 == 0, so aligned words of the ring are aligned words of the output.
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
  for (unsigned g = w0 + 4 * lane; g < w1; g += 128) *(unsigned*)(gb + g) = *(const unsigned*)(ring + g % INF_RING);
  if (lane < g1 - w1) gb[w1 + lane] = ring[(w1 + lane) % INF_RING];
}

__global__ void __launch_bounds__(32) inflate_kernel(const unsigned char* __restrict__ comp,
                                                     const InflateSeg* __restrict__ segs, int n_segs,
                                                     unsigned char* out_base, int* __restrict__ status,
                                                     unsigned* __restrict__ trailer_adler) {
  __shared__ InflateWarpSmem S;
  __shared__ __align__(16) unsigned char ring[INF_RING];
  const unsigned lane = lane_id();
  const int sidx = blockIdx.x;
  if (sidx >= n_segs) return;
  const InflateSeg sg = segs[sidx];
  unsigned char* out = out_base + sg.out_off;
  const unsigned out_len = (unsigned)sg.out_len;
  const unsigned char* in = comp + sg.in_off;
  const unsigned in_len = (unsigned)sg.in_len;
  const unsigned a0 = (unsigned)((uintptr_t)out & 3);

  BitR br;
  br.mis = (unsigned)((uintptr_t)in & 3);
  br.w = (const unsigned*)(in - br.mis);
  br.kmax = (br.mis + in_len - 1) >> 2;
  br.bb = 0; br.bc = 0; br.ip = 0;
  int err = INF_OK;
  unsigned opos = 0;            // bytes produced so far (all flushed at batch boundaries)
  unsigned adler = 0;

  if (sg.flags & INF_ZLIB) {
    if (in_len < 8) err = INF_BAD_HEADER;
    else {
      br.refill();
      unsigned cmf = br.get(8), flg = br.get(8);
      if ((cmf & 15) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20)) err = INF_BAD_HEADER;
    }
  }
  bool last = false;
  while (!err && !last) {
    if (!(sg.flags & INF_ZLIB) && opos >= out_len) break;
    br.refill();
    last = br.get(1) != 0;
    unsigned type = br.get(2);
    if (type == 0) {
      // stored: skip to the byte boundary, LEN / NLEN, raw copy through the ring (later blocks may reference it)
      br.drop(br.bc & 7);
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
        for (unsigned i = lane; i < m; i += 32) ring[(opos + a0 + i) % INF_RING] = in[src + b0 + i];
        __syncwarp();
        inflate_flush(ring, out, a0, opos, opos + m, lane);
        __syncwarp();
        opos += m;
      }
      br.ip = src + len; br.bb = 0; br.bc = 0;
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
      if (!inflate_build(S.lens, 19, S.dtab, 7, S.dsorted, S.dcount, S.run)) { err = INF_BAD_LENGTHS; break; }
      __syncwarp();
      int idx = 0;
      unsigned prev_len = 0;
      while (idx < nl + nd) {
        br.refill();
        unsigned e = S.dtab[br.peek(7)];
        if (!e) { err = INF_BAD_LENGTHS; break; }
        br.drop(e & 15);
        unsigned sym = e >> 4;
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
    if (!inflate_build(ll, nl, S.ltab, INF_LBITS, S.lsorted, S.lcount, S.run)) { err = INF_BAD_LENGTHS; break; }
    __syncwarp();
    if (!inflate_build(dl, nd, S.dtab, INF_DBITS, S.dsorted, S.dcount, S.run)) { err = INF_BAD_LENGTHS; break; }
    __syncwarp();

    // ---- symbol loop: batches of up to 32 symbols / INF_BATCH_BYTES output bytes
    bool eob = false;
    while (!eob && !err) {
      unsigned my_pos = 0, my_len = 0, my_dist = 0, my_lit = 0;
      unsigned bpos = opos;
      int k = 0;
      for (; k < 32 && bpos - opos < INF_BATCH_BYTES; k++) {
        br.refill();
        unsigned e = S.ltab[br.peek(INF_LBITS)];
        unsigned sym;
        if (e) { br.drop(e & 15); sym = e >> 4; }
        else if (!inflate_slow(br, S.lcount, S.lsorted, sym)) { err = INF_BAD_CODE; break; }
        if (sym < 256) {
          if (lane == (unsigned)k) { my_pos = bpos; my_lit = sym; my_len = 0; }
          bpos += 1;
        } else if (sym == 256) {
          eob = true;
          break;
        } else {
          sym -= 257;
          if (sym >= 29) { err = INF_BAD_CODE; break; }
          unsigned len;
          if (sym < 8) len = sym + 3;
          else if (sym == 28) len = 258;
          else { unsigned nb = (sym - 4) >> 2; len = 3 + ((4 + (sym & 3)) << nb) + br.get(nb); }
          br.refill();
          unsigned e2 = S.dtab[br.peek(INF_DBITS)];
          unsigned ds;
          if (e2) { br.drop(e2 & 15); ds = e2 >> 4; }
          else if (!inflate_slow(br, S.dcount, S.dsorted, ds)) { err = INF_BAD_CODE; break; }
          if (ds >= 30) { err = INF_BAD_CODE; break; }
          unsigned dist;
          if (ds < 4) dist = ds + 1;
          else { unsigned nb = (ds >> 1) - 1; br.refill(); dist = 1 + ((2 + (ds & 1)) << nb) + br.get(nb); }
          if (dist > bpos) { err = INF_BAD_DISTANCE; break; }
          if (lane == (unsigned)k) { my_pos = bpos; my_len = len; my_dist = dist; }
          bpos += len;
        }
        if (bpos > out_len) { err = INF_BAD_SIZE; break; }
      }
      if (err) break;
      if (br.byte_pos() > in_len + 8) { err = INF_INPUT_OVERRUN; break; }
      const bool have = lane < (unsigned)k;
      const bool is_match = have && my_len != 0;
      // a match is independent of this batch's other symbols if everything it reads was produced before the batch
      const bool indep = is_match && (my_pos - my_dist + min(my_len, my_dist) <= opos);
      const unsigned rp = (my_pos + a0) % INF_RING;
      if (have && my_len == 0) ring[rp] = (unsigned char)my_lit;
      if (indep) {
        unsigned rs = rp >= my_dist ? rp - my_dist : rp + INF_RING - my_dist;
        for (unsigned i = 0; i < my_len; i++) ring[ring_wrap(rp + i)] = ring[ring_wrap(rs + i)];
      }
      __syncwarp();
      unsigned mm = __ballot_sync(0xffffffffu, is_match && !indep);
      while (mm) {
        int src_lane = __ffs((int)mm) - 1;
        mm &= mm - 1;
        unsigned p = __shfl_sync(0xffffffffu, rp, src_lane);
        unsigned l = __shfl_sync(0xffffffffu, my_len, src_lane);
        unsigned d = __shfl_sync(0xffffffffu, my_dist, src_lane);
        unsigned s = p >= d ? p - d : p + INF_RING - d;
        if (d >= l) { for (unsigned i = lane; i < l; i += 32) ring[ring_wrap(p + i)] = ring[ring_wrap(s + i)]; }
        else if (d >= 32) {
          // overlapping copy with period d >= 32: rounds of 32 bytes never read what the same round writes
          for (unsigned i0 = 0; i0 < l; i0 += 32) {
            unsigned i = i0 + lane;
            if (i < l) ring[ring_wrap(p + i)] = ring[ring_wrap(s + i)];
            __syncwarp();
          }
        } else { for (unsigned i = lane; i < l; i += 32) ring[ring_wrap(p + i)] = ring[ring_wrap(s + i % d)]; }
        __syncwarp();
      }
      inflate_flush(ring, out, a0, opos, bpos, lane);
      __syncwarp();
      opos = bpos;
    }
  }
  if (!err) {
    if (opos != out_len) err = INF_BAD_SIZE;
    else if (sg.flags & INF_ZLIB) {
      br.drop(br.bc & 7);
      unsigned a = 0;
      for (int i = 0; i < 4; i++) { br.refill(); a = (a << 8) | br.get(8); }
      adler = a;
      if (br.byte_pos() > in_len) err = INF_INPUT_OVERRUN;
    } else if (br.byte_pos() > in_len) err = INF_INPUT_OVERRUN;
  }
  if (lane == 0) { status[sidx] = err; trailer_adler[sidx] = adler; }
}

}  // namespace mts
