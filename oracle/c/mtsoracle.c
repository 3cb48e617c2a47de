/* mtsoracle.c — ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the arithmetic on mtscomp's per-chunk codec path, independent of both NumPy and libz, used
 * to cross-check oracle/codec.py and the CUDA kernels:
 *   ora_transform / ora_untransform   mtscomp.py:143-169, 381-394, 622-635 (diff / cumsum with wrap-around, 'F'/'C'
 *                                     serialisation)
 *   ora_adler32                       RFC 1950 section 8.2 (zlib's trailer checksum)
 *   ora_inflate                       RFC 1951 decoder + RFC 1950 container: what `zlib.decompress` does at
 *                                     mtscomp.py:619 (third-party libz, not in the reference tree; pinned here by the
 *                                     golden .cbin files written by the reference with zlib 1.3)
 * Parity status: pinned by tests/test_oracle_golden.py::test_c_oracle_* against the reference-written fixtures.
 * Build: oracle/Makefile -> oracle/_build/libmtsoracle.so (gcc, no dependencies).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ transforms */
/* element access by byte width; arithmetic is modulo 2^(8*w), identical to NumPy's integer wrap-around */
static uint64_t ld(const uint8_t* p, int w) { uint64_t v = 0; memcpy(&v, p, (size_t)w); return v; }
static void st(uint8_t* p, int w, uint64_t v) { memcpy(p, &v, (size_t)w); }

/* src: row-major (ns, nc); dst: bytes handed to deflate.  flags: 1 time diff, 2 spatial diff, 4 order 'C'. */
void ora_transform(const uint8_t* src, long ns, long nc, int w, int flags, uint8_t* dst) {
  for (long t = 0; t < ns; t++)
    for (long c = 0; c < nc; c++) {
      uint64_t v = ld(src + (t * nc + c) * w, w);
      if ((flags & 1) && t > 0) v -= ld(src + ((t - 1) * nc + c) * w, w);
      if ((flags & 2) && c > 0) {
        uint64_t u = ld(src + (t * nc + c - 1) * w, w);
        if ((flags & 1) && t > 0) u -= ld(src + ((t - 1) * nc + c - 1) * w, w);
        v -= u;
      }
      long o = (flags & 4) ? t * nc + c : c * ns + t;
      st(dst + o * w, w, v);
    }
}

/* inverse: spatial running sum first, then time running sum, exactly the order of mtscomp.py:631-632 */
void ora_untransform(const uint8_t* src, long ns, long nc, int w, int flags, uint8_t* dst) {
  for (long t = 0; t < ns; t++) {
    uint64_t run = 0;
    for (long c = 0; c < nc; c++) {
      long i = (flags & 4) ? t * nc + c : c * ns + t;
      uint64_t v = ld(src + i * w, w);
      if (flags & 2) { run += v; v = run; }
      st(dst + (t * nc + c) * w, w, v);
    }
  }
  if (flags & 1)
    for (long t = 1; t < ns; t++)
      for (long c = 0; c < nc; c++) {
        uint64_t v = ld(dst + (t * nc + c) * w, w) + ld(dst + ((t - 1) * nc + c) * w, w);
        st(dst + (t * nc + c) * w, w, v);
      }
}

uint32_t ora_adler32(const uint8_t* p, long n) {
  uint32_t a = 1, b = 0;
  for (long i = 0; i < n; i++) { a = (a + p[i]) % 65521u; b = (b + a) % 65521u; }
  return (b << 16) | a;
}

/* ------------------------------------------------------------------------------------------------ inflate */
typedef struct { const uint8_t* in; long n, pos; uint32_t bit, cnt; int err; } bits_t;

static uint32_t getbits(bits_t* s, int need) {
  uint32_t v = s->bit;
  while ((int)s->cnt < need) {
    if (s->pos >= s->n) { s->err = 1; return 0; }
    v |= (uint32_t)s->in[s->pos++] << s->cnt;
    s->cnt += 8;
  }
  s->bit = need < 32 ? v >> need : 0;
  s->cnt -= (uint32_t)need;
  return need < 32 ? v & ((1u << need) - 1) : v;
}

typedef struct { short count[16]; short symbol[288]; } huff_t;

static int build(huff_t* h, const short* len, int n) {
  short offs[16];
  int left = 1;
  memset(h->count, 0, sizeof h->count);
  for (int i = 0; i < n; i++) h->count[len[i]]++;
  if (h->count[0] == n) return 0;
  for (int l = 1; l < 16; l++) { left <<= 1; left -= h->count[l]; if (left < 0) return left; }
  offs[1] = 0;
  for (int l = 1; l < 15; l++) offs[l + 1] = (short)(offs[l] + h->count[l]);
  for (int i = 0; i < n; i++) if (len[i]) h->symbol[offs[len[i]]++] = (short)i;
  return left;
}

static int decode(bits_t* s, const huff_t* h) {
  int code = 0, first = 0, index = 0;
  for (int l = 1; l < 16; l++) {
    code |= (int)getbits(s, 1);
    if (s->err) return -1;
    int c = h->count[l];
    if (code - c < first) return h->symbol[index + (code - first)];
    index += c; first += c; first <<= 1; code <<= 1;
  }
  return -1;
}

static const short LBASE[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
static const short LEXT[29] = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
static const short DBASE[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
static const short DEXT[30] = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};

static int codes(bits_t* s, const huff_t* lc, const huff_t* dc, uint8_t* out, long cap, long* opos) {
  for (;;) {
    int sym = decode(s, lc);
    if (sym < 0) return -10;
    if (sym < 256) { if (*opos >= cap) return -11; out[(*opos)++] = (uint8_t)sym; }
    else if (sym == 256) return 0;
    else {
      sym -= 257;
      if (sym >= 29) return -12;
      long len = LBASE[sym] + (long)getbits(s, LEXT[sym]);
      int ds = decode(s, dc);
      if (ds < 0 || ds >= 30) return -13;
      long dist = DBASE[ds] + (long)getbits(s, DEXT[ds]);
      if (s->err) return -14;
      if (dist > *opos) return -15;
      if (*opos + len > cap) return -11;
      for (long i = 0; i < len; i++) { out[*opos] = out[*opos - dist]; (*opos)++; }
    }
  }
}

/* Decode a zlib stream.  Returns the number of bytes written (>= 0) or a negative error; trailing bytes after the
 * adler32 are ignored, a wrong adler32 is an error (-20) — the behaviour of zlib.decompress. */
long ora_inflate(const uint8_t* in, long n, uint8_t* out, long cap) {
  if (n < 6) return -1;
  if ((in[0] & 15) != 8 || (in[0] >> 4) > 7 || ((in[0] << 8) | in[1]) % 31 || (in[1] & 0x20)) return -2;
  bits_t s = {in, n, 2, 0, 0, 0};
  long opos = 0;
  int last;
  do {
    last = (int)getbits(&s, 1);
    int type = (int)getbits(&s, 2);
    if (s.err) return -3;
    if (type == 0) {
      s.bit = 0; s.cnt = 0;
      if (s.pos + 4 > n) return -4;
      unsigned len = in[s.pos] | (in[s.pos + 1] << 8), nlen = in[s.pos + 2] | (in[s.pos + 3] << 8);
      s.pos += 4;
      if ((len ^ 0xffff) != nlen) return -5;
      if (s.pos + len > n || opos + len > cap) return -6;
      memcpy(out + opos, in + s.pos, len);
      opos += len; s.pos += len;
    } else if (type == 1 || type == 2) {
      short lengths[320];
      huff_t lc, dc;
      int nl = 288, nd = 30;
      if (type == 1) {
        for (int i = 0; i < 288; i++) lengths[i] = (short)(i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
        for (int i = 0; i < 30; i++) lengths[288 + i] = 5;
      } else {
        static const short order[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};
        nl = (int)getbits(&s, 5) + 257; nd = (int)getbits(&s, 5) + 1;
        int nc = (int)getbits(&s, 4) + 4;
        if (nl > 286 || nd > 30) return -7;
        for (int i = 0; i < 19; i++) lengths[i] = 0;
        for (int i = 0; i < nc; i++) lengths[order[i]] = (short)getbits(&s, 3);
        if (build(&lc, lengths, 19) != 0) return -8;
        int idx = 0;
        while (idx < nl + nd) {
          int sym = decode(&s, &lc);
          if (sym < 0) return -9;
          if (sym < 16) lengths[idx++] = (short)sym;
          else {
            int rep, v = 0;
            if (sym == 16) { if (!idx) return -9; v = lengths[idx - 1]; rep = 3 + (int)getbits(&s, 2); }
            else if (sym == 17) rep = 3 + (int)getbits(&s, 3);
            else rep = 11 + (int)getbits(&s, 7);
            if (idx + rep > nl + nd) return -9;
            while (rep--) lengths[idx++] = (short)v;
          }
        }
        if (s.err || lengths[256] == 0) return -9;
      }
      int e = build(&lc, lengths, nl);
      if (e < 0 || (type == 2 && e > 0 && nl - lc.count[0] != 1)) return -8;
      e = build(&dc, lengths + nl, nd);
      if (e < 0 || (type == 2 && e > 0 && nd - dc.count[0] != 1)) return -8;   /* the fixed distance code is incomplete by design */
      int r = codes(&s, &lc, &dc, out, cap, &opos);
      if (r) return r;
    } else return -3;
  } while (!last);
  s.bit = 0; s.cnt = 0;
  if (s.pos + 4 > n) return -4;
  uint32_t want = ((uint32_t)in[s.pos] << 24) | (in[s.pos + 1] << 16) | (in[s.pos + 2] << 8) | in[s.pos + 3];
  if (want != ora_adler32(out, opos)) return -20;
  return opos;
}
