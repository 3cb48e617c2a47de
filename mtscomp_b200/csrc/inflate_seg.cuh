// inflate_seg.cuh — decoder of the indexed segments of chunks that carry the second in-band index ("MTS2", deflate.cuh).
//
// Such a segment is ONE dynamic-Huffman block whose sub-blocks (every IDX_SUB_BYTES of output) are indexed by bit offset,
// and whose matches obey the STEP RULE: a match never reads output at or after the beginning of the step (step_bytes of
// output, counted from the segment's start) in which it starts.  That removes both serial chains of inflate:
//   seg_tokens_kernel   warp per segment, LANE per sub-block (the first one, mostly literals behind an empty window, is
//                       indexed as two halves so that the lanes have equal work): every lane decodes its codes once, from
//                       its indexed bit offset, into 32-bit tokens (runs of up to 3 literals packed into one; batches of
//                       32 token slots never straddle a step) + the output offset before every batch
//   seg_resolve_kernel  CTA per segment: step after step, ALL the tokens of a step are resolved at once (their sources
//                       are complete by the step rule) through a 32 KB shared-memory ring of the segment's output; one
//                       block barrier per step, coalesced word stores to global memory
// Every assumption is checked on the device (bit offsets and output counts of the sub-blocks must chain exactly, every
// match must obey the step rule and the ring's reach); a segment that fails any check is left to the serial decoder,
// which also decides whether the data is corrupt.  Reference-written (plain zlib) streams never come here.
#pragma once
#include "deflate.cuh"
#include "inflate_par.cuh"

namespace mts {

struct SegV2 {              // one indexed segment
  long long in_off;         // byte offset of its deflate data in the compressed buffer
  long long out_off;        // byte offset of its output in the transformed buffer
  long long tab_off;        // byte offset of its sub-block table in the compressed buffer (4 bytes per sub-block)
  int in_len, out_len;
  int sub_first;            // index of its first sub-block in the batch (token areas and tables)
  unsigned step_bytes;      // the step rule's step
};

static const int SEG_TOK_STRIDE = 4480;                  // token slots per sub-block: (8192 + 257 bytes) / 2 bytes per token
                                                         // (a run of <= 3 literals is one token) + 31 slots of padding per step
static const int SEG_BATCHES = SEG_TOK_STRIDE / 32;
static const int SEG_MAX_SPS = IDX_SUB_BYTES / 256;      // steps per sub-block at the smallest step
static const int SEG_RING = 32768;
static const int SEG_MAX_STEPS = 1024;                   // steps per segment the resolve kernel keeps in shared memory
#ifndef MTS_SEG_RES_WARPS
#define MTS_SEG_RES_WARPS 8
#endif
static const int SEG_RES_WARPS = MTS_SEG_RES_WARPS;
static const unsigned TOK_MATCH = 0x80000000u;

// ---------------------------------------------------------------------------------------------- seg_tokens_kernel
struct SubOut {             // per sub-block, written by seg_tokens_kernel
  unsigned n_slots;         // token slots used (batches are padded at step starts)
  unsigned stepb[SEG_MAX_SPS];   // first batch of each step of the sub-block
};

struct TokSink {            // a lane's token writer
  unsigned* tok;            // the sub-block's slots
  uint2* btab;              // per batch: {output offset before it, tokens in it}
  unsigned* stepb;
  unsigned n, bpos, step0, cur_step, step_shift;
  bool over;
  __device__ __forceinline__ void open(unsigned* t, uint2* b, unsigned* sb, unsigned first_pos, unsigned shift, unsigned sub_pos) {
    tok = t; btab = b; stepb = sb; n = 0; bpos = first_pos; step_shift = shift;
    step0 = sub_pos >> shift; cur_step = first_pos >> shift; over = false;
    for (unsigned q = 0; q <= cur_step - step0 && q < (unsigned)SEG_MAX_SPS; q++) stepb[q] = 0;
  }
  // token `word` whose first output byte is at `pos` (segment-relative)
  __device__ __forceinline__ void put(unsigned word, unsigned pos) {
    const unsigned st = pos >> step_shift;
    if (st != cur_step) {                                  // a new step begins: close the batch, note where the step starts
      if (n & 31u) { btab[n >> 5] = make_uint2(bpos, n & 31u); n = (n + 31u) & ~31u; }
      for (unsigned q = cur_step + 1; q <= st; q++) if (q - step0 < (unsigned)SEG_MAX_SPS) stepb[q - step0] = n >> 5;
      cur_step = st;
    }
    if ((n & 31u) == 0) bpos = pos;
    if (n < (unsigned)SEG_TOK_STRIDE) tok[n] = word; else over = true;
    n++;
    if ((n & 31u) == 0 && !over) btab[(n >> 5) - 1] = make_uint2(bpos, 32u);
  }
  __device__ __forceinline__ void close(unsigned n_steps_sub) {
    if ((n & 31u) && !over) { btab[n >> 5] = make_uint2(bpos, n & 31u); n = (n + 31u) & ~31u; }
    for (unsigned q = cur_step - step0 + 1; q < n_steps_sub && q < (unsigned)SEG_MAX_SPS; q++) stepb[q] = n >> 5;
  }
};

__device__ __forceinline__ unsigned rd32u(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((unsigned)p[3] << 24); }

__global__ void __launch_bounds__(32) seg_tokens_kernel(const unsigned char* __restrict__ comp, const SegV2* __restrict__ segs,
                                                        unsigned* __restrict__ tokens, uint2* __restrict__ btab,
                                                        SubOut* __restrict__ subs, ParRes* __restrict__ res) {
  __shared__ BlkTabs T;
  __shared__ uint4 ring[4 * 32];                                // SBits: four 16-byte slots per lane
  __shared__ unsigned short lenx[32];
  __shared__ unsigned distx[32];
  const unsigned lane = threadIdx.x;
  lenx[lane] = (unsigned short)par_len_info(lane);
  distx[lane] = par_dist_info(lane);
  const SegV2 sg = segs[blockIdx.x];
  const unsigned char* in = comp + sg.in_off;
  const unsigned in_len = (unsigned)sg.in_len, in_bits = in_len * 8, out_len = (unsigned)sg.out_len;
  ParRes r0;
  r0.tail_bit = 0; r0.tail_out = 0; r0.n_done = 0; r0.flags = 2; r0.adler = 0;   // until proven good
  int nl = 0, nd = 0, hok = 0;
  unsigned fin = 0, hdr_end = 0;
  if (lane == 0) hok = blk_parse_header(in, in_len, 0, T.lens, (unsigned char*)T.dtab, nl, nd, fin, hdr_end) ? 1 : 0;
  hok = __shfl_sync(0xffffffffu, hok, 0);
  nl = __shfl_sync(0xffffffffu, nl, 0);
  nd = __shfl_sync(0xffffffffu, nd, 0);
  fin = __shfl_sync(0xffffffffu, fin, 0);
  hdr_end = __shfl_sync(0xffffffffu, hdr_end, 0);
  __syncwarp();
  bool ok = hok && !fin && hdr_end < in_bits;
  ok = ok && blk_build<1>(T.lens, nl, T.ltab, PAR_LBITS, T.llong, T.lsorted, T.code, lane);
  ok = ok && blk_build<2>(T.lens + nl, nd, T.dtab, PAR_DBITS, T.dlong, T.dsorted, T.code, lane);
  unsigned step_shift = 0;
  while ((1u << step_shift) < sg.step_bytes) step_shift++;
  const unsigned n_sub = (unsigned)idx_n_sub(out_len);
  const unsigned char* tab = comp + sg.tab_off;
  unsigned carry = 0, end_bit = 0;
  for (unsigned g0 = 0; ok && g0 < n_sub; g0 += 32) {
    const unsigned j = g0 + lane;
    const bool act = j < n_sub;
    const unsigned e = act ? rd32u(tab + 4 * j) : 0u, e1 = (j + 1 < n_sub) ? rd32u(tab + 4 * (j + 1)) : 0u;
    unsigned incl = e & 0x1ffffu;
    for (int d = 1; d < 32; d <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += v; }
    const unsigned start = carry + incl;                                      // bit of the sub-block's first token
    carry += __shfl_sync(0xffffffffu, incl, 31);
    const bool last = j + 1 == n_sub;
    const unsigned bound = last ? in_bits : start + (e1 & 0x1ffffu);
    const unsigned pos0 = idx_bound(j) + ((e >> 17) & 511u);                  // output offset of its first token
    const unsigned pos1 = last ? out_len : idx_bound(j + 1) + ((e1 >> 17) & 511u);
    bool good = !act || (start < in_bits && bound <= in_bits && bound > start && pos0 < pos1 && pos1 <= out_len &&
                         (j > 0 || (start == hdr_end && pos0 == 0)));
    // ---- the lane's sub-block: one pass over its codes (the loop of blk_span, with a token sink)
    TokSink sink;
    const size_t sub = (size_t)sg.sub_first + j;
    SubOut* so = subs + sub;
    if (act) sink.open(tokens + sub * SEG_TOK_STRIDE, btab + sub * SEG_BATCHES, so->stepb, pos0, step_shift, idx_bound(j));
    SBits br;
    br.init(in, in_len, (act && good) ? start : 0u, ring + lane);
    // One symbol per iteration, at most one token written: the token in hand (`pend`: a run of 1..3 literals, or a
    // match) is written when the next symbol cannot join it.
    unsigned pos = pos0, pend = 0, ppos = pos0, np = 0, flags = 0;      // np: literals in hand (0: pend is a match or nothing)
    bool have = false;
    bool active = act && good;
    while (__any_sync(0xffffffffu, active)) {
      if (active) {
        br.refill();
        const bool past = br.bit_pos() >= bound;
        const unsigned win = br.window();
        const unsigned en = T.ltab[win & ((1u << PAR_LBITS) - 1)];
        unsigned cl = en & 15, kind = (en >> 4) & 3, val = en >> 6;
        if (cl == 0 && !past) {                                               // rare: a code longer than the table
          unsigned sym;
          if (!blk_long<PAR_LBITS>(win, T.llong, T.lsorted, sym, cl)) { kind = K_BAD; cl = 1; }
          else {
            kind = sym < 256 ? (unsigned)K_LIT : sym == 256 ? (unsigned)K_EOB : sym < 286 ? (unsigned)K_LEN : (unsigned)K_BAD;
            val = sym < 256 ? sym : sym > 256 ? sym - 257 : 0;
          }
        }
        const bool stop = past || kind >= (unsigned)K_EOB;
        const bool isl = kind == K_LEN;
        const unsigned lx = lenx[isl ? val : 0u];
        const unsigned xb = isl ? lx >> 12 : 0u;
        const unsigned len = (lx & 0xfffu) + ((win >> cl) & ((1u << xb) - 1));
        br.drop(past ? 0u : cl + xb);
        br.refill();
        const unsigned win2 = br.window();
        const unsigned e2 = T.dtab[win2 & ((1u << PAR_DBITS) - 1)];
        unsigned cl2 = e2 & 15, dsym = e2 >> 4;
        bool dbad = false;
        if (isl && cl2 == 0 && !stop) dbad = !blk_long<PAR_DBITS>(win2, T.dlong, T.dsorted, dsym, cl2);
        const unsigned dx = distx[dsym & 31];
        dbad = dbad || (isl && dx == 0);
        const unsigned xb2 = dx >> 16;
        const unsigned dist = (dx & 0xffffu) + ((win2 >> cl2) & ((1u << xb2) - 1));
        br.drop((isl && !stop) ? cl2 + xb2 : 0u);
        if (stop || dbad) {
          flags = (!past && kind == K_EOB) ? 1u : (past ? 0u : 2u);
          if (dbad) flags = 2;
          active = false;
        } else {
          const bool join = !isl && have && np > 0 && np < 3;                 // a literal joins the literals in hand
          if (have && !join) sink.put(pend, ppos);
          if (join) { pend |= val << (8 * np); np++; pend = (pend & 0x00ffffffu) | ((np - 1) << 24); }
          else if (isl) { pend = TOK_MATCH | (len << 16) | (dist - 1); ppos = pos; np = 0; }
          else { pend = val; ppos = pos; np = 1; }
          have = true;
          pos += isl ? len : 1u;
        }
      }
    }
    br.finish();
    if (act && good) {
      if (have) sink.put(pend, ppos);
      sink.close(((last ? out_len : idx_bound(j + 1)) - idx_bound(j) + sg.step_bytes - 1) >> step_shift);
      // the sub-block must end exactly where the next one begins (bits and bytes); the last one with the end-of-block code
      good = !(flags & 2) && !sink.over && pos == pos1 && br.bit_pos() <= in_bits &&
             (last ? (flags & 1) != 0 : (br.bit_pos() == bound && !(flags & 1)));
      so->n_slots = sink.n;
      if (last) end_bit = br.bit_pos();
    }
    ok = __all_sync(0xffffffffu, good);
    end_bit = __shfl_sync(0xffffffffu, end_bit, (n_sub - 1 - g0) & 31);       // (only meaningful in the last group)
  }
  if (lane == 0) {
    if (ok) { r0.tail_bit = end_bit; r0.tail_out = out_len; r0.n_done = 1; r0.flags = 0; }
    res[blockIdx.x] = r0;
  }
}

// ---------------------------------------------------------------------------------------------- seg_resolve_kernel
// Also leaves the segment's standalone adler32 (seg_adler, as the encoder's lz77_kernel does): the bytes pass through
// the flushing threads' registers anyway.
__global__ void __launch_bounds__(SEG_RES_WARPS * 32) seg_resolve_kernel(const SegV2* __restrict__ segs,
                                                                         const unsigned* __restrict__ tokens,
                                                                         const uint2* __restrict__ btab,
                                                                         const SubOut* __restrict__ subs,
                                                                         unsigned char* out_base, ParRes* __restrict__ res,
                                                                         unsigned* __restrict__ seg_adler) {
  MTS_DYN_SMEM(sm);                                             // [ring SEG_RING][per step: first batch | batches << 24]
  unsigned char* ring = sm;
  unsigned short* ring16 = (unsigned short*)sm;
  const unsigned* ring32 = (const unsigned*)sm;
  unsigned* s_step = (unsigned*)(sm + SEG_RING);                // first batch (relative to the segment's) | batches << 24
  __shared__ unsigned long long s_red[2 * SEG_RES_WARPS];
  const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const unsigned W = SEG_RES_WARPS, NT = SEG_RES_WARPS * 32, RM = SEG_RING - 1;
  const SegV2 sg = segs[blockIdx.x];
  if (res[blockIdx.x].flags & 2) return;                        // the token kernel gave up on this segment
  const unsigned out_len = (unsigned)sg.out_len, DS = sg.step_bytes;
  unsigned step_shift = 0;
  while ((1u << step_shift) < DS) step_shift++;
  const unsigned n_steps = (out_len + DS - 1) >> step_shift;
  unsigned char* out = out_base + sg.out_off;
  const bool al4 = ((uintptr_t)out & 3) == 0;
  if (n_steps > (unsigned)SEG_MAX_STEPS) { if (tid == 0) res[blockIdx.x].flags |= 2u; return; }
  // batches of a step: from its first batch to the next step's (the steps of a sub-block are consecutive in its token
  // area; the last step of a sub-block ends with the sub-block's slots)
  for (unsigned s = tid; s < n_steps; s += NT) {
    const unsigned j = idx_item(s << step_shift), ls = s - (idx_bound(j) >> step_shift);
    const SubOut* so = subs + (size_t)sg.sub_first + j;
    const unsigned b0 = so->stepb[ls];
    const unsigned b1 = (s + 1 < n_steps && idx_item((s + 1) << step_shift) == j) ? so->stepb[ls + 1] : (so->n_slots + 31) >> 5;
    s_step[s] = (j * SEG_BATCHES + b0) | (min(b1 > b0 ? b1 - b0 : 0u, 255u) << 24);
  }
  __syncthreads();
  const unsigned max_dist = SEG_RING - DS - 320;                // what the ring still holds while a step is being written
  const size_t bat0 = (size_t)sg.sub_first * SEG_BATCHES;
  // software pipeline: a warp's first batch of step s + 1 is fetched while step s is resolved (slots beyond a batch's
  // token count hold stale words: the count masks them)
  unsigned t_nx = 0;
  uint2 bt_nx = make_uint2(0, 0);
  auto fetch = [&](unsigned s) {
    if (s >= n_steps) return;
    const unsigned e = s_step[s];
    if (wid < (e >> 24)) { const size_t b = bat0 + (e & 0xffffffu) + wid; bt_nx = btab[b]; t_nx = tokens[b * 32 + lane]; }
  };
  fetch(0);
  bool bad = false;
  unsigned ad_a = 0;
  unsigned long long ad_b = 0;
  for (unsigned s = 0; s < n_steps; s++) {
    const unsigned e = s_step[s], nb = e >> 24, s_start = s << step_shift;
    unsigned t = t_nx;
    uint2 bt = bt_nx;
    fetch(s + 1);
    for (unsigned q = wid; q < nb; q += W) {
      if (q != wid) { const size_t b = bat0 + (e & 0xffffffu) + q; bt = btab[b]; t = tokens[b * 32 + lane]; }
      const bool has = lane < bt.y;
      const bool isM = has && (t >> 31);
      const unsigned L = has ? (isM ? (t >> 16) & 0x1ffu : ((t >> 24) & 3u) + 1u) : 0u;
      const unsigned dist = (t & 0x7fffu) + 1;
      unsigned incl = L;
      for (int d = 1; d < 32; d <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += v; }
      const unsigned p = bt.x + incl - L;
      // every token of the batch starts in this step; a match reads only what earlier steps wrote, within the ring
      if (has && (p < s_start || p >= s_start + DS || p + L > out_len ||
                  (isM && (L < 3 || dist > p || p - dist + L > s_start || dist > max_dist)))) bad = true;
      if (has && !bad) {
        if (!isM) {
          ring[p & RM] = (unsigned char)t;
          if (L > 1) ring[(p + 1) & RM] = (unsigned char)(t >> 8);
          if (L > 2) ring[(p + 2) & RM] = (unsigned char)(t >> 16);
        } else if (((p | dist | L) & 1u) == 0) {
          const unsigned p2 = p >> 1, s2 = (p - dist) >> 1, n2 = L >> 1, M2 = RM >> 1;
          // the first 8 bytes without a loop (most matches), the rest of a longer one in a loop
          unsigned short v0 = ring16[s2 & M2], v1 = ring16[(s2 + 1) & M2], v2 = 0, v3 = 0;
          if (n2 > 2) v2 = ring16[(s2 + 2) & M2];
          if (n2 > 3) v3 = ring16[(s2 + 3) & M2];
          ring16[p2 & M2] = v0;
          ring16[(p2 + 1) & M2] = v1;
          if (n2 > 2) ring16[(p2 + 2) & M2] = v2;
          if (n2 > 3) ring16[(p2 + 3) & M2] = v3;
          for (unsigned i = 4; i < n2; i++) ring16[(p2 + i) & M2] = ring16[(s2 + i) & M2];
        } else {
          for (unsigned i = 0; i < L; i++) ring[(p + i) & RM] = ring[(p - dist + i) & RM];
        }
      }
    }
    if (__syncthreads_or(bad)) { bad = true; break; }
    // the step's bytes are complete (matches of earlier steps may have spilled into it): ring -> global, coalesced
    const unsigned end = min(s_start + DS, out_len);
    if (al4) {
      for (unsigned i = s_start + 4 * tid; i + 4 <= end; i += 4 * NT) {
        const unsigned w = ring32[(i & RM) >> 2];
        *(unsigned*)(out + i) = w;
        const unsigned sum = (w & 0xffu) + ((w >> 8) & 0xffu) + ((w >> 16) & 0xffu) + (w >> 24);
        ad_a += sum;
        ad_b += (unsigned long long)(out_len - i) * sum - (((w >> 8) & 0xffu) + 2 * ((w >> 16) & 0xffu) + 3 * (w >> 24));
      }
      if (tid < (end & 3u)) {
        const unsigned i = (end & ~3u) + tid, v = ring[i & RM];
        out[i] = (unsigned char)v;
        ad_a += v; ad_b += (unsigned long long)(out_len - i) * v;
      }
    } else {
      for (unsigned i = s_start + tid; i < end; i += NT) {
        const unsigned v = ring[i & RM];
        out[i] = (unsigned char)v;
        ad_a += v; ad_b += (unsigned long long)(out_len - i) * v;
      }
    }
  }
  if (bad) { if (tid == 0) res[blockIdx.x].flags |= 2u; return; }
  {
    unsigned long long a = ad_a, b = ad_b % ADLER_BASE;
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) { s_red[2 * wid] = a; s_red[2 * wid + 1] = b; }
    __syncthreads();
    if (tid == 0) {
      unsigned long long ta = 0, tb = 0;
      for (unsigned w = 0; w < W; w++) { ta += s_red[2 * w]; tb += s_red[2 * w + 1]; }
      seg_adler[blockIdx.x] = ((unsigned)((out_len % ADLER_BASE + tb) % ADLER_BASE) << 16) | (unsigned)((1 + ta) % ADLER_BASE);
    }
  }
}

}  // namespace mts
