// Deflate (RFC 1951) building blocks of the entropy front-end.
//
// The reference hands the uint8 index planes of the transform path to zlib: np.savez_compressed
// in src/z_lib.py:19-23, tifffile's zlib codec in src/TIFF.py:23-31.  The planes are long runs of
// the bias value 128 with sparse, strongly skewed literals.  The parse is greedy over a FIXED set of
// candidate distances (MatchParams): distance 1 (runs -- zlib's Z_RLE strategy) and, when the caller
// names the row length of the planes, the bytes of the row above (row - 1, row, row + 1: where
// zlib's hash chains find most of their matches in a subband plane).  Whether a position starts a
// match is then a function of a bounded neighbourhood of the input alone -- no hash table, no
// state carried along the stream -- so every thread can parse its own piece.
//
// One *segment* (one CTA) becomes one dynamic-Huffman block that ends on a byte boundary (an
// empty stored block, like zlib's Z_SYNC_FLUSH), so the streams of the segments concatenate.
// Inside a segment every thread parses a *piece*; the pieces' bit strings are packed back to
// back at bit granularity.
//
// Everything here is __host__ __device__: tests/deflate_emul.cpp compiles the same code with g++
// and runs the CTA's phases one "thread" after the other, which is how the logic is tested
// without a GPU (the shipped library never takes that path).
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define DFL_HD __host__ __device__ __forceinline__
#else
#define DFL_HD inline
#endif

namespace vcfb {
namespace dfl {

constexpr int NLIT = 286;      // literal/length alphabet: 0..255 literals, 256 end of block, 257..285 lengths
constexpr int EOB = 256;
constexpr int NCL = 19;        // code-length alphabet
constexpr int MAX_LIT_BITS = 15;
constexpr int MAX_CL_BITS = 7;
constexpr int MAX_MATCH = 258;
constexpr int MIN_MATCH = 3;
constexpr int STORED_MAX = 65535;
constexpr int NDIST = 30;      // distance alphabet
constexpr int MAX_CAND = 8;    // candidate distances of the parse
constexpr int FAR_MIN = 4;     // shortest match at a distance other than 1 (found by comparing eight bytes at once)

// ---- match lengths ------------------------------------------------------------------------------

DFL_HD int ilog2_u32(uint32_t x) {           // x >= 1
#if defined(__CUDA_ARCH__)
  return 31 - __clz(int(x));
#else
  return 31 - __builtin_clz(x);
#endif
}

// RFC 1951 3.2.5: length L in [3, 258] -> symbol 257..285, number of extra bits, extra value
DFL_HD void length_symbol(int L, int* sym, int* ebits, int* eval) {
  if (L == MAX_MATCH) { *sym = 285; *ebits = 0; *eval = 0; return; }
  const int l = L - 3;
  if (l < 8) { *sym = 257 + l; *ebits = 0; *eval = 0; return; }
  const int e = ilog2_u32(uint32_t(l)) - 2;    // l in [8, 254] -> e in [1, 5]
  *sym = 261 + 4 * e + ((l >> e) & 3);
  *ebits = e;
  *eval = l & ((1 << e) - 1);
}

// ---- distances ------------------------------------------------------------------------------------

// RFC 1951 3.2.5: distance d in [1, 32768] -> symbol 0..29, number of extra bits, extra value
DFL_HD void distance_symbol(int d, int* sym, int* ebits, int* eval) {
  const int x = d - 1;
  if (x < 4) { *sym = x; *ebits = 0; *eval = 0; return; }
  int e = -1;                      // floor(log2(x)) - 1
  for (int t = x >> 1; t; t >>= 1) ++e;
  *sym = 2 * (e + 1) + ((x >> e) & 1);
  *ebits = e;
  *eval = x & ((1 << e) - 1);
}

// The candidate distances of the parse and what each costs.  dist[0] is always 1.
struct MatchParams {
  int nd;
  int good;                     // a run at least this long is taken without looking at the other candidates
  int patience;                 // positions in a row without such a match after which a piece stops looking for one
  int rows3;                    // the candidates are 1, px, row, row - px, row + px with px in [2, 4] (and row > 2 px)
  int dist[MAX_CAND];
  uint8_t dsym[MAX_CAND], debits[MAX_CAND];
  uint16_t deval[MAX_CAND];
  uint8_t dcost8[MAX_CAND];     // estimated size of the distance code + its extra bits, in 1/8 bit
};

inline void match_params_add(MatchParams& P, int d) {
  if (d < 1 || d > 32768 || P.nd >= MAX_CAND) return;
  for (int i = 0; i < P.nd; ++i) if (P.dist[i] == d) return;
  int s, eb, ev;
  distance_symbol(d, &s, &eb, &ev);
  const int i = P.nd++;
  P.dist[i] = d; P.dsym[i] = uint8_t(s); P.debits[i] = uint8_t(eb); P.deval[i] = uint16_t(ev);
  P.dcost8[i] = uint8_t(d == 1 ? 12 : 8 * (3 + eb));     // the run distance gets the shortest code; the others about 3 bits
}

// row = 0: runs only (the parse of zlib's Z_RLE).  Otherwise the input is an array of rows of `row`
// bytes whose samples lie `pixel` bytes apart (3 for the H x W x 3 index image the reference hands
// to its entropy stage): the previous sample of the same channel and the samples above it are
// candidates too.
inline MatchParams make_match_params(long long row, int pixel) {
  MatchParams P;
  P.nd = 0;
  P.good = 64;
  P.patience = 64;
  P.rows3 = 0;
  for (int i = 0; i < MAX_CAND; ++i) { P.dist[i] = 0; P.dsym[i] = 0; P.debits[i] = 0; P.deval[i] = 0; P.dcost8[i] = 0; }
  match_params_add(P, 1);
  if (pixel < 1) pixel = 1;
  if (row > 0) {
    if (pixel > 1) match_params_add(P, pixel);
    if (row > 2 * pixel && row + pixel <= 32768) {
      match_params_add(P, int(row));
      match_params_add(P, int(row) - pixel);
      match_params_add(P, int(row) + pixel);
      P.rows3 = P.nd == 5 && pixel >= 2 && pixel <= 4;
    }
  }
  return P;
}

#if defined(__CUDA_ARCH__)
#define DFL_CTZ64(x) (__ffsll((long long)(x)) - 1)
#else
#define DFL_CTZ64(x) __builtin_ctzll(x)
#endif

// src must be 8-byte aligned.  The input is read as aligned 64-bit words; bytes past n are never touched.
// CHK = false: the caller knows that base + 8 <= n (every piece but the last few of the input), which
// takes a 64-bit comparison and a branch off every load -- a sixth of the parse's instructions.
// I: the type positions are held in -- long long, or int where the caller has moved src close to them
// (32-bit arithmetic: half the instructions of the 64-bit one on this machine).
template <bool CHK = true, class I = long long>
DFL_HD uint64_t load_word(const uint8_t* src, long long n, I base) {
  if (!CHK || base + 8 <= n) return *reinterpret_cast<const uint64_t*>(src + base);
  uint64_t x = 0;
  for (int k = 0; base + k < n; ++k) x |= uint64_t(src[base + k]) << (8 * k);
  return x;
}

// the eight bytes at an arbitrary position (zeros past n); CHK = false: pos + 16 <= n
template <bool CHK = true, class I = long long>
DFL_HD uint64_t load_u64_at(const uint8_t* src, long long n, I pos) {
  const I base = pos & ~I(7);
  const int off = int(pos & 7);
  const uint64_t w0 = load_word<CHK>(src, n, base);
  if (!CHK)                       // both words without a branch: independent loads stay in flight together
    return (w0 >> (8 * off)) | ((load_word<false>(src, n, base + 8) << 1) << (63 - 8 * off));
  if (!off) return w0;
  const uint64_t w1 = (!CHK || base + 8 < n) ? load_word<CHK>(src, n, base + 8) : 0;
  return (w0 >> (8 * off)) | (w1 << (64 - 8 * off));
}

// ---- cost model of the parse --------------------------------------------------------------------
//
// A greedy parse that takes every match it finds is larger than the run-length parse on dense planes:
// three literals of a skewed alphabet take fewer bits than a match with a 7-bit distance.  So the
// segment is sampled first (every fourth 64-bit word): bytes that differ from their predecessor
// stand for the literals, those followed by a copy of themselves for the starts of runs; the
// estimated code lengths (in 1/8 bit, integer arithmetic only -- the host emulation gives the same
// tokens) decide whether a match is worth its bits.

constexpr int SAMPLE_EVERY = 4;

// 8 * log2(num / den), num >= den >= 1, num < 2^23
DFL_HD int cost8_ratio(uint32_t num, uint32_t den) {
  const uint32_t r = (num << 8) / den;        // >= 256
  const int e = ilog2_u32(r);
  const int f = int((r >> (e - 3)) & 7);      // the three bits below the leading one
  const int frac = (0x76654310 >> (4 * f)) & 15;   // 8 * log2(1 + f / 8), rounded: 0 1 3 4 5 6 6 7
  return 8 * (e - 8) + frac;
}

struct CostModel {
  uint8_t lit8[256];    // estimated bits * 8 of a literal
  int len8;             // of a length symbol (without extra bits)
  int far_on;           // whether the segment looks at the candidates other than the run at all
};

// bit i: byte i of w equals byte i - 1 (byte -1 = pb).  Exact zero-byte detector on w ^ (w shifted by a byte).
DFL_HD uint32_t eq_mask8(uint64_t w, uint32_t pb) {
  const uint64_t x = w ^ ((w << 8) | uint64_t(pb & 0xff));
  const uint64_t m7 = 0x7f7f7f7f7f7f7f7full;
  const uint64_t z = ~(((x & m7) + m7) | x | m7);    // 0x80 where a byte of x is zero
  return uint32_t(((z >> 7) * 0x0102040810204080ull) >> 56);
}

// Counts of one sampled word w at byte position pos: wp / wn are the words before and after it
// (zero where there is none).  A byte belongs to a run when three consecutive bytes around it equal
// their predecessors; every other byte counts as a literal, every start of a run as a match.
template <class Add>
DFL_HD void sample_word(uint64_t wp, uint64_t w, uint64_t wn, long long pos, long long n, Add& add) {
  uint32_t eq;                                       // bit i: byte i of (wp, w, wn) equals byte i - 1, i in [1, 24)
  if (pos >= 16 && pos + 16 <= n) {
    eq = (eq_mask8(wp, 0) & 0xfeu) | (eq_mask8(w, uint32_t(wp >> 56)) << 8) | (eq_mask8(wn, uint32_t(w >> 56)) << 16);
  } else {                                           // the ends of the input: positions outside [1, n) equal nothing
    eq = 0;
    int prev = int(wp & 0xff);
    for (int i = 1; i < 24; ++i) {
      const uint64_t x = i < 8 ? wp : i < 16 ? w : wn;
      const int b = int((x >> (8 * (i & 7))) & 0xff);
      const long long at = pos - 8 + i;
      if (b == prev && at >= 1 && at < n) eq |= 1u << i;
      prev = b;
    }
  }
  const uint32_t t = eq & (eq >> 1) & (eq >> 2);     // bit i: bytes i, i + 1, i + 2 all equal their predecessors
  const uint32_t inrun = t | (t << 1) | (t << 2);
  for (int i = 8; i < 16; ++i) {
    if (pos - 8 + i >= n) break;
    if (!((inrun >> i) & 1)) add.lit(int((w >> (8 * (i & 7))) & 0xff));
    else if (!((inrun >> (i - 1)) & 1)) add.run();
  }
}

// the sampled words of the segment [s0, s0 + nseg) that thread tid of NT looks at
// Also counted: the sampled words that are not one repeated byte (add.dense()) and those of them that
// stand, all eight bytes, at one of the other candidate distances as well (add.hit()): where the
// content is dense and almost none do -- noise, fine quantisation steps -- looking for such matches
// costs more than half of the parse and finds nothing (model_far_on()).
template <class Add>
DFL_HD void sample_segment(const uint8_t* src, long long n, long long s0, long long nseg, int tid, int NT,
                           const MatchParams& P, Add& add) {
  const long long first = ((s0 + 7) / 8 + SAMPLE_EVERY - 1) / SAMPLE_EVERY * SAMPLE_EVERY;
  for (long long aw = first + (long long)SAMPLE_EVERY * tid; aw * 8 < s0 + nseg; aw += (long long)SAMPLE_EVERY * NT) {
    const long long pos = aw * 8;
    const uint64_t w = load_word(src, n, pos);
    sample_word(pos >= 8 ? load_word(src, n, pos - 8) : 0, w, pos + 8 < n ? load_word(src, n, pos + 8) : 0, pos, n, add);
    if (P.nd > 1 && pos + 8 <= n && w != 0x0101010101010101ull * (w & 0xff)) {
      add.dense();
      for (int c = 1; c < P.nd; ++c)
        if (pos >= P.dist[c] && load_u64_at(src, n, pos - P.dist[c]) == w) { add.hit(); break; }
    }
  }
}

// dense, hits: the counts of sample_segment() over the segment
DFL_HD int model_far_on(uint32_t dense, uint32_t hits) { return dense < 256 || hits * 16 >= dense; }

// entry b of the cost table from the sampled counts (cnt[256] literals, nlit their sum, nrun run starts)
DFL_HD uint8_t model_lit8(const uint32_t* cnt, uint32_t nlit, uint32_t nrun, int b) {
  const uint32_t ntok = nlit + nrun + 1;
  const uint32_t c = cnt[b];
  int v = c ? cost8_ratio(ntok, c) : cost8_ratio(2 * ntok, 1);
  if (v < 8) v = 8;
  if (v > 120) v = 120;
  return uint8_t(v);
}

DFL_HD int model_len8(uint32_t nlit, uint32_t nrun) {
  const uint32_t ntok = nlit + nrun + 1;
  int v = cost8_ratio(ntok, nrun ? nrun : 1) + 16;     // the matches spread over several length symbols
  if (v < 16) v = 16;
  if (v > 96) v = 96;
  return v;
}


// ---- parse ----------------------------------------------------------------------------------------

// the eight bytes at byte offset off (0 <= off <= 15) of three consecutive words
DFL_HD uint64_t window3(uint64_t w0, uint64_t w1, uint64_t w2, int off) {
  const uint64_t lo = off >= 8 ? w1 : w0, hi = off >= 8 ? w2 : w1;
  const int sh = 8 * (off & 7);
  return (lo >> sh) | ((hi << 1) << (63 - sh));
}

// number of bytes (at most lim) for which src[p + k] == src[p - d + k]; CHK = false: p + lim + 16 <= n
template <bool CHK = true, class I = long long>
DFL_HD int match_length(const uint8_t* src, long long n, I p, int d, int lim) {
  int L = 0;
  while (L < lim) {
    const uint64_t x = load_u64_at<CHK>(src, n, p + L) ^ load_u64_at<CHK>(src, n, p + L - d);
    if (x) { L += DFL_CTZ64(x) >> 3; break; }
    L += 8;
  }
  return L < lim ? L : lim;
}

// estimated size (1/8 bit) of a run of r bytes equal to their predecessor: literals, or a match of distance 1
DFL_HD int run_cost8(int r, int lit8, const MatchParams& P, const CostModel& M) {
  const int as_lit = r * lit8;
  if (r < MIN_MATCH) return as_lit;
  int sym, eb, ev;
  length_symbol(r, &sym, &eb, &ev);
  const int as_match = M.len8 + 8 * eb + int(P.dcost8[0]);
  return as_match < as_lit ? as_match : as_lit;
}

// estimated size of src[p, p + L) under the run-length parse (prev = the byte before p): what a
// match at another distance has to beat.  Eight bytes at a time where they continue a run; gives
// up (returning what it has) once the estimate is above `enough`.
template <bool CHK = true, class I = long long>
DFL_HD int span_cost8(const uint8_t* src, long long n, I p, int L, int prev, const MatchParams& P,
                      const CostModel& M, int enough) {
  int cost = 0, last = -1, cur = prev;                  // last: the latest byte that differs from its predecessor
  uint32_t pb = uint32_t(prev);
  for (int k = 0; k < L; k += 8) {
    const uint64_t w = load_u64_at<CHK>(src, n, p + k);
    const int nb = L - k < 8 ? L - k : 8;
    uint32_t m = ~eq_mask8(w, pb) & ((1u << nb) - 1u);
    if (prev < 0 && k == 0) m |= 1u;                    // the first byte of the input has no predecessor
    while (m) {
      const int i = DFL_CTZ64(uint64_t(m));
      m &= m - 1;
      const int run = k + i - last - 1;
      if (run) cost += run_cost8(run, int(M.lit8[cur & 0xff]), P, M);
      cur = int((w >> (8 * i)) & 0xff);
      cost += M.lit8[cur];
      last = k + i;
    }
    if (cost > enough) return cost;
    pb = uint32_t(w >> 56);
  }
  const int run = L - last - 1;
  if (run) cost += run_cost8(run, int(M.lit8[cur & 0xff]), P, M);
  return cost;
}

// Tokens of the piece [s, e) of src[0, n), greedy.  At every position: the run (distance 1) when it
// is cheaper than its bytes as literals; the longest match over the other candidate distances when
// it is estimated to take fewer bits than the run-length parse of the bytes it covers; else a
// literal.  A match never leaves the piece; its source may lie in an earlier piece or segment (the
// decoder has produced it by then).  V::lit(byte) / V::match(length, candidate).  M == nullptr:
// every run of MIN_MATCH bytes or more is a match (the plain run-length parse; one candidate only).
// CHK = false: e + 16 <= n, no load looks at n.
// `lowest`: the position of the first byte of the input (0 unless src has been moved).
// NC > 0: exactly NC candidates besides the run -- their first eight bytes are fetched and compared without a
// branch in between, so the loads are in flight together (the walk is latency-bound); NC = -1: any number.
template <bool CHK, class I, int NC, class V>
DFL_HD void parse_piece_impl(const uint8_t* src, long long n, I s, I e, I lowest, const MatchParams& P,
                             const CostModel* M, V& v) {
  I p = s;
  int prev = p > lowest ? int(src[p - 1]) : -1;
  I wbase = (s & ~I(7)) - 8;
  uint64_t w = 0;
  bool far_on = M && P.nd > 1 && M->far_on;
  int fails = 0;
  while (p < e) {
    const I base = p & ~I(7);
    if (base != wbase) { w = load_word<CHK>(src, n, base); wbase = base; }
    const int b = int((w >> (8 * int(p & 7))) & 0xff);
    const I rem = e - p;
    const int lim = rem < MAX_MATCH ? int(rem) : MAX_MATCH;
    int best = 0, bc = 0;
    if (b == prev) {
      const uint64_t splat = 0x0101010101010101ull * uint64_t(prev);
      int L = 0;
      I rb = wbase;
      uint64_t rw = w;
      for (;;) {
        const int off = int((p + L) & 7);
        const uint64_t x = (rw ^ splat) >> (8 * off);
        const int avail = 8 - off;
        const int z = x ? (DFL_CTZ64(x) >> 3) : avail;    // a differing byte lies inside the `avail` valid ones
        L += z;
        if (z < avail || L >= lim) break;
        rb += 8;
        rw = load_word<CHK>(src, n, rb);
      }
      if (L > lim) L = lim;
      if (L >= MIN_MATCH) {
        if (!M) best = L;
        else {
          int sym, eb, ev;
          length_symbol(L, &sym, &eb, &ev);
          if (L * int(M->lit8[b]) > M->len8 + 8 * eb + int(P.dcost8[0])) best = L;
        }
      }
    }
    if (far_on && best < lim && best < P.good && lim >= FAR_MIN) {
      const uint64_t win = load_u64_at<CHK>(src, n, p);
      int fl = 0, fc = 0;                                  // the longest match at another distance
      if constexpr (NC > 0) {
        int first[NC];
        bool done = false;
        if constexpr (NC == 4) {
          // make_match_params() of an image with samples of 2-4 bytes: the three positions above lie within 24
          // bytes of each other -- three aligned words serve all of them (five loads per position instead of eight)
          const int px = P.dist[1];
          const I top = p - P.dist[4];                     // the lowest of them: row + px back
          if (top >= lowest) {
            const I wb = top & ~I(7);
            const int o = int(top & 7);
            const uint64_t w0 = load_word<CHK>(src, n, wb), w1 = load_word<CHK>(src, n, I(wb + 8)),
                           w2 = load_word<CHK>(src, n, I(wb + 16));
            const uint64_t x1 = win ^ load_u64_at<CHK>(src, n, I(p - px));
            const uint64_t x4 = win ^ window3(w0, w1, w2, o);
            const uint64_t x2 = win ^ window3(w0, w1, w2, o + px);
            const uint64_t x3 = win ^ window3(w0, w1, w2, o + 2 * px);
            first[0] = x1 ? (DFL_CTZ64(x1) >> 3) : 8;
            first[1] = x2 ? (DFL_CTZ64(x2) >> 3) : 8;
            first[2] = x3 ? (DFL_CTZ64(x3) >> 3) : 8;
            first[3] = x4 ? (DFL_CTZ64(x4) >> 3) : 8;
            done = true;
          }
        }
        if (!done) {
#pragma unroll
          for (int k = 0; k < NC; ++k) {
            const int d = P.dist[k + 1];
            const bool ok = p - d >= lowest;
            const uint64_t x = win ^ load_u64_at<CHK>(src, n, ok ? I(p - d) : p);
            const int L = x ? (DFL_CTZ64(x) >> 3) : 8;
            first[k] = ok ? L : 0;
          }
        }
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          int L = first[k];
          if (L < FAR_MIN) continue;
          if (L == 8 && lim > 8) L += match_length<CHK>(src, n, I(p + 8), P.dist[k + 1], lim - 8);
          if (L > lim) L = lim;
          if (L > fl) { fl = L; fc = k + 1; }
        }
      } else {
        for (int c = 1; c < P.nd; ++c) {
          const int d = P.dist[c];
          if (p - d < lowest) continue;
          const uint64_t x = win ^ load_u64_at<CHK>(src, n, I(p - d));
          int L = x ? (DFL_CTZ64(x) >> 3) : 8;
          if (L < FAR_MIN) continue;
          if (L == 8 && lim > 8) L += match_length<CHK>(src, n, I(p + 8), d, lim - 8);
          if (L > lim) L = lim;
          if (L > fl) { fl = L; fc = c; }
        }
      }
      bool taken = false;
      if (fl > best) {
        int sym, eb, ev;
        length_symbol(fl, &sym, &eb, &ev);
        const int cost = M->len8 + 8 * eb + int(P.dcost8[fc]);
        if (span_cost8<CHK>(src, n, p, fl, prev, P, *M, cost) > cost) { best = fl; bc = fc; taken = true; }
      }
      if (taken) fails = 0;
      else if (++fails >= P.patience) far_on = false;      // dense content: nothing to find above, stop looking
    }
    if (best) {
      v.match(best, bc);
      p += best;
      if (bc) prev = int(src[p - 1]);
      continue;
    }
    v.lit(b);
    prev = b;
    ++p;
  }
}

template <class V>
DFL_HD void parse_piece(const uint8_t* src, long long n, long long s, long long e, const MatchParams& P,
                        const CostModel* M, V& v) {
  if (e + 16 <= n) {            // all but the last pieces of the input: no bound checks, positions relative to the piece
    const long long s8 = s & ~7ll;
    const int lowest = s8 > (1ll << 30) ? -(1 << 30) : -int(s8);
    if (P.rows3) parse_piece_impl<false, int, 4>(src + s8, n, int(s - s8), int(e - s8), lowest, P, M, v);    // make_match_params() of an image
    else parse_piece_impl<false, int, -1>(src + s8, n, int(s - s8), int(e - s8), lowest, P, M, v);
  } else {
    parse_piece_impl<true, long long, -1>(src, n, s, e, 0ll, P, M, v);
  }
}

// A token in 16 bits: a literal is its byte; a match is 256 + (length - 3) in bits 0..8 and the
// candidate in bits 9..11.  The kernel parses once, keeps the tokens and walks them twice more.
DFL_HD uint16_t token_lit(int b) { return uint16_t(b); }
DFL_HD uint16_t token_match(int L, int c) { return uint16_t(256 + (L - MIN_MATCH) + (c << 9)); }

template <class V>
DFL_HD void visit_token(uint16_t t, V& v) {
  const int x = t & 511;
  if (x < 256) v.lit(x);
  else v.match(x - 256 + MIN_MATCH, t >> 9);
}

// ---- bit output ---------------------------------------------------------------------------------

#if defined(__CUDA_ARCH__)
#define DFL_ATOMIC_OR(ptr, val) atomicOr((ptr), (val))
#else
#define DFL_ATOMIC_OR(ptr, val) (*(ptr) |= (val))
#endif

// LSB-first bit writer starting at an arbitrary bit offset of a zero-initialised word array.  The
// first and the last word it touches may be shared with the neighbouring writers (atomic OR); the
// words in between are covered by this writer alone (plain stores).
struct BitWriter {
  uint32_t* base;
  uint64_t acc;
  long long w;
  int nb;
  bool first;
  DFL_HD void init(uint32_t* b, long long bitoff) {
    base = b; acc = 0; w = bitoff >> 5; nb = int(bitoff & 31); first = true;
  }
  DFL_HD void put(uint32_t v, int n) {       // n <= 32, v < 2^n
    acc |= uint64_t(v) << nb;
    nb += n;
    if (nb >= 32) {
      const uint32_t word = uint32_t(acc);
      if (first) { DFL_ATOMIC_OR(base + w, word); first = false; }
      else base[w] = word;
      ++w; acc >>= 32; nb -= 32;
    }
  }
  DFL_HD long long bitpos() const { return (w << 5) + nb; }
  DFL_HD void align_byte() { const int r = (-nb) & 7; if (r) put(0, r); }
  DFL_HD void finish() { if (nb > 0) { DFL_ATOMIC_OR(base + w, uint32_t(acc)); } nb = 0; acc = 0; }
};

// ---- visitors -----------------------------------------------------------------------------------

// Code tables of one segment
struct Codes {
  uint16_t code[NLIT];   // bit-reversed, ready for LSB-first output
  uint8_t len[NLIT];
  uint16_t dcode[NDIST + 2];
  uint8_t dlen[NDIST + 2];
};

#if defined(__CUDA_ARCH__)
#define DFL_ATOMIC_ADD(ptr, val) atomicAdd((ptr), (val))
#else
#define DFL_ATOMIC_ADD(ptr, val) (*(ptr) += (val))
#endif

struct CountVisitor {      // symbol frequencies; equal consecutive symbols (runs of 258-byte matches) are added at once
  uint32_t* hist;
  uint32_t* dhist;
  const MatchParams* P;
  int cur;
  uint32_t cnt, d0;
  DFL_HD void init(uint32_t* h, uint32_t* dh, const MatchParams* p) { hist = h; dhist = dh; P = p; cur = 0; cnt = 0; d0 = 0; }
  DFL_HD void flush() {
    if (cnt) { DFL_ATOMIC_ADD(hist + cur, cnt); }
    if (d0) { DFL_ATOMIC_ADD(dhist + P->dsym[0], d0); }
    cnt = 0; d0 = 0;
  }
  DFL_HD void add(int sym) {
    if (sym != cur) { if (cnt) { DFL_ATOMIC_ADD(hist + cur, cnt); } cnt = 0; cur = sym; }
    ++cnt;
  }
  DFL_HD void lit(int b) { add(b); }
  DFL_HD void match(int L, int c) {
    int sym, eb, ev;
    length_symbol(L, &sym, &eb, &ev);
    add(sym);
    if (c == 0) ++d0;
    else DFL_ATOMIC_ADD(dhist + P->dsym[c], 1u);
  }
};

struct SizeVisitor {       // bits the tokens take under the codes
  const Codes* c;
  const MatchParams* P;
  unsigned bits;
  DFL_HD void lit(int b) { bits += c->len[b]; }
  DFL_HD void match(int L, int k) {
    int sym, eb, ev;
    length_symbol(L, &sym, &eb, &ev);
    bits += c->len[sym] + eb + c->dlen[P->dsym[k]] + P->debits[k];
  }
};

struct EmitVisitor {
  const Codes* c;
  const MatchParams* P;
  BitWriter* bw;
  DFL_HD void lit(int b) { bw->put(c->code[b], c->len[b]); }
  DFL_HD void match(int L, int k) {
    int sym, eb, ev;
    length_symbol(L, &sym, &eb, &ev);
    const int n = c->len[sym];
    bw->put(uint32_t(c->code[sym]) | (uint32_t(ev) << n), n + eb);               // <= 15 + 5 bits
    const int ds = P->dsym[k], dn = c->dlen[ds];
    bw->put(uint32_t(c->dcode[ds]) | (uint32_t(P->deval[k]) << dn), dn + P->debits[k]);   // <= 15 + 13 bits
  }
};

// parse visitor of the kernel's first pass: keeps the token (strided store) and counts it
template <class Store>
struct TokenVisitor {
  Store st;
  CountVisitor cv;
  int n;
  DFL_HD void lit(int b) { st(n++, token_lit(b)); cv.lit(b); }
  DFL_HD void match(int L, int c) { st(n++, token_match(L, c)); cv.match(L, c); }
};

// ---- Huffman code construction ------------------------------------------------------------------

// rank of symbol i among the symbols with non-zero frequency, ordered by (frequency, symbol)
DFL_HD int rank_of(const uint32_t* freq, int n, int i) {
  const uint32_t fi = freq[i];
  int r = 0;
  for (int j = 0; j < n; ++j) {
    const uint32_t fj = freq[j];
    r += (fj != 0) && (fj < fi || (fj == fi && j < i));
  }
  return r;
}

struct BuildScratch {          // n <= NLIT
  uint16_t sorted[NLIT];       // symbols by ascending (frequency, symbol), filled through rank_of()
  uint32_t sw[NLIT];           // their frequencies in that order (parallel construction only)
  uint32_t cnt[MAX_LIT_BITS + 1];    // codes per length           "
  uint32_t next[MAX_LIT_BITS + 2];   // first code of each length  "
  uint32_t cum[MAX_LIT_BITS + 1];    // codes of length <= i       "
  uint32_t m;                        // used symbols               "
  uint32_t hi;                       // largest used symbol        "
  uint32_t iw[NLIT];           // weights of the internal nodes in creation order
  uint16_t ipar[NLIT];         // parent (internal node index) of internal node
  uint16_t lpar[NLIT];         // parent of leaf sorted[k]
  uint8_t idepth[NLIT];
};

DFL_HD uint32_t bit_reverse(uint32_t v, int n) {
#if defined(__CUDA_ARCH__)
  return n ? __brev(v) >> (32 - n) : 0u;
#endif
  uint32_t r = 0;
  for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1); v >>= 1; }
  return r;
}

// Code lengths (<= maxbits, complete code) and canonical codes for the m >= 2 used symbols listed
// in S.sorted[0, m); unused symbols get length 0.  Two-queue Huffman construction, then the
// length limit by moving codes between levels until the Kraft sum is exactly one.
DFL_HD void build_code(const uint32_t* freq, int n, int m, int maxbits, BuildScratch& S, uint16_t* code, uint8_t* len) {
  for (int i = 0; i < n; ++i) { len[i] = 0; code[i] = 0; }
  if (m < 2) {                       // callers force two used symbols; kept for safety
    if (m == 1) len[S.sorted[0]] = 1;
    return;
  }
  int li = 0, ii = 0, ni = 0;
  for (int k = 0; k < m - 1; ++k) {
    uint32_t wsum = 0;
    for (int pick = 0; pick < 2; ++pick) {
      const bool leaf = li < m && (ii >= ni || freq[S.sorted[li]] <= S.iw[ii]);
      if (leaf) { wsum += freq[S.sorted[li]]; S.lpar[li] = uint16_t(ni); ++li; }
      else { wsum += S.iw[ii]; S.ipar[ii] = uint16_t(ni); ++ii; }
    }
    S.iw[ni++] = wsum;
  }
  int cnt[64];
  for (int i = 0; i < 64; ++i) cnt[i] = 0;
  S.idepth[m - 2] = 0;                                  // the root
  for (int j = m - 3; j >= 0; --j) {
    const int d = S.idepth[S.ipar[j]] + 1;
    S.idepth[j] = uint8_t(d > 62 ? 62 : d);
  }
  for (int k = 0; k < m; ++k) {
    int d = S.idepth[S.lpar[k]] + 1;
    if (d > 63) d = 63;
    ++cnt[d];
  }
  for (int i = maxbits + 1; i < 64; ++i) cnt[maxbits] += cnt[i];
  uint32_t total = 0;
  for (int i = maxbits; i > 0; --i) total += uint32_t(cnt[i]) << (maxbits - i);
  while (total != (1u << maxbits)) {
    --cnt[maxbits];
    for (int i = maxbits - 1; i > 0; --i)
      if (cnt[i]) { --cnt[i]; cnt[i + 1] += 2; break; }
    --total;
  }
  {
    int j = m;
    for (int i = 1; i <= maxbits; ++i)
      for (int l = cnt[i]; l > 0; --l) len[S.sorted[--j]] = uint8_t(i);   // most frequent first
  }
  uint32_t next[MAX_LIT_BITS + 2];
  next[0] = 0; next[1] = 0;
  for (int i = 2; i <= maxbits; ++i) next[i] = (next[i - 1] + uint32_t(cnt[i - 1])) << 1;
  for (int i = 0; i < n; ++i) {
    const int l = len[i];
    code[i] = l ? uint16_t(bit_reverse(next[l]++, l)) : uint16_t(0);
  }
}

// ---- the same construction in CTA-parallel steps ---------------------------------------------------
//
// One thread executes an instruction every ~10 cycles when each depends on the last, so a serial
// build_code() of the 286-symbol literal/length code holds the other 511 threads at a barrier for
// longer than they take to parse the segment.  The steps below leave only the two-queue merge and
// the walk over the internal nodes to one thread; every function is called by all `nt` threads with
// their `tid`, with a CTA barrier between consecutive steps.  Results equal build_code()'s.
// S.m, S.hi and S.cnt[] must be zero before par_rank().

#if defined(__CUDA_ARCH__)
#define DFL_ATOMIC_MAX(ptr, val) atomicMax((ptr), (val))
#else
#define DFL_ATOMIC_MAX(ptr, val) (*(ptr) = *(ptr) > (val) ? *(ptr) : (val))
#endif

DFL_HD void par_rank(const uint32_t* freq, int n, BuildScratch& S, uint16_t* code, uint8_t* len, int tid, int nt) {
  for (int i = tid; i < n; i += nt) {
    len[i] = 0;
    code[i] = 0;
    const uint32_t f = freq[i];
    if (f) {
      const int r = rank_of(freq, n, i);
      S.sorted[r] = uint16_t(i);
      S.sw[r] = f;
      DFL_ATOMIC_ADD(&S.m, 1u);
      DFL_ATOMIC_MAX(&S.hi, uint32_t(i));
    }
  }
}

DFL_HD void par_tree(BuildScratch& S, int tid) {         // two-queue merge, depths of the internal nodes
  if (tid != 0) return;
  const int m = int(S.m);
  if (m < 2) return;
  const uint32_t INF = 0xffffffffu;                      // above any weight: a segment has < 2^32 symbols
  int li = 0, ii = 0, ni = 0;
  uint32_t lw = S.sw[0], ih = INF;                       // heads of the leaf and the internal-node queue
  for (int k = 0; k < m - 1; ++k) {
    uint32_t wsum = 0;
    for (int pick = 0; pick < 2; ++pick) {
      if (li < m && lw <= ih) {
        wsum += lw; S.lpar[li] = uint16_t(ni); ++li;
        lw = li < m ? S.sw[li] : INF;
      } else {
        wsum += ih; S.ipar[ii] = uint16_t(ni); ++ii;
        ih = ii < ni ? S.iw[ii] : INF;
      }
    }
    S.iw[ni] = wsum;
    if (ii == ni) ih = wsum;                              // the queue was empty: the new node is its head
    ++ni;
  }
  S.idepth[m - 2] = 0;                                    // the root
  for (int j = m - 3; j >= 0; --j) {
    const int d = S.idepth[S.ipar[j]] + 1;
    S.idepth[j] = uint8_t(d > 62 ? 62 : d);
  }
}

DFL_HD void par_count(BuildScratch& S, int maxbits, int tid, int nt) {
  const int m = int(S.m);
  if (m < 2) return;
  for (int k = tid; k < m; k += nt) {
    int d = S.idepth[S.lpar[k]] + 1;
    if (d > maxbits) d = maxbits;
    DFL_ATOMIC_ADD(&S.cnt[d], 1u);
  }
}

DFL_HD void par_limit(BuildScratch& S, int maxbits, int tid) {   // Kraft sum to exactly one, first codes, prefix counts
  if (tid != 0) return;
  const int m = int(S.m);
  if (m < 2) {
    for (int i = 0; i <= maxbits; ++i) S.cnt[i] = 0;
    S.cnt[1] = uint32_t(m);
  } else {
    uint32_t total = 0;
    for (int i = maxbits; i > 0; --i) total += S.cnt[i] << (maxbits - i);
    while (total != (1u << maxbits)) {
      --S.cnt[maxbits];
      for (int i = maxbits - 1; i > 0; --i)
        if (S.cnt[i]) { --S.cnt[i]; S.cnt[i + 1] += 2; break; }
      --total;
    }
  }
  S.next[0] = 0; S.next[1] = 0;
  for (int i = 2; i <= maxbits; ++i) S.next[i] = (S.next[i - 1] + S.cnt[i - 1]) << 1;
  S.cum[0] = 0;
  for (int i = 1; i <= maxbits; ++i) S.cum[i] = S.cum[i - 1] + S.cnt[i];
}

DFL_HD void par_lengths(const BuildScratch& S, int maxbits, uint8_t* len, int tid, int nt) {
  const int m = int(S.m);
  for (int k = tid; k < m; k += nt) {
    const uint32_t r = uint32_t(m - 1 - k);              // rank by descending frequency: the shortest codes first
    int i = 1;
    while (i < maxbits && r >= S.cum[i]) ++i;
    len[S.sorted[k]] = uint8_t(i);
  }
}

DFL_HD void par_codes(const BuildScratch& S, int n, const uint8_t* len, uint16_t* code, int tid, int nt) {
  for (int i = tid; i < n; i += nt) {
    const int l = len[i];
    if (!l) continue;
    uint32_t c = S.next[l];
    for (int j = 0; j < i; ++j) c += len[j] == l;         // canonical order: by symbol inside a length
    code[i] = uint16_t(bit_reverse(c, l));
  }
}

// ---- block header -------------------------------------------------------------------------------

struct Header {
  uint8_t tok_sym[NLIT + NDIST + 8];   // code-length tokens of the concatenated length arrays
  uint8_t tok_ext[NLIT + NDIST + 8];
  int ntok;
  int hlit;                         // literal/length codes sent (>= 257)
  int hdist;                        // distance codes sent (>= 2 here)
  int hclen;                        // code-length codes sent (>= 4)
  uint32_t clfreq[NCL];
  uint16_t clcode[NCL];
  uint8_t cllen[NCL];
  int bits;                         // size of the whole header including BFINAL/BTYPE
  // the CTA-parallel construction (hpar_*) only
  uint8_t v[NLIT + NDIST + 4];      // the two length arrays one after the other
  uint16_t tcnt[NLIT + NDIST + 4];  // tokens of the run that starts at i (0 elsewhere)
  uint16_t rlen[NLIT + NDIST + 4];  // its length
  uint16_t bsum[(NLIT + NDIST + 4) / 16 + 1];   // tokens per 16 positions
  int clused;                       // code-length symbols in use
  uint32_t tokbits;                 // sum of the tokens' sizes under the code-length code
};

DFL_HD void header_tok(Header& h, int sym, int ext) {
  h.tok_sym[h.ntok] = uint8_t(sym);
  h.tok_ext[h.ntok] = uint8_t(ext);
  ++h.ntok;
  ++h.clfreq[sym];
}

// Run-length tokens (symbols 16/17/18 of RFC 1951 3.2.7) for the lengths of the literal/length
// code followed by those of the distance code.
DFL_HD void header_tokens(const uint8_t* litlen, const uint8_t* dlen, Header& h) {
  int hlit = NLIT;
  while (hlit > 257 && litlen[hlit - 1] == 0) --hlit;
  h.hlit = hlit;
  int hdist = NDIST;
  while (hdist > 1 && dlen[hdist - 1] == 0) --hdist;
  h.hdist = hdist;
  h.ntok = 0;
  for (int i = 0; i < NCL; ++i) h.clfreq[i] = 0;
  const int total = hlit + hdist;
  int i = 0;
  while (i < total) {
    const int v = i < hlit ? litlen[i] : dlen[i - hlit];
    int r = 1;
    while (i + r < total && (i + r < hlit ? litlen[i + r] : dlen[i + r - hlit]) == v) ++r;
    i += r;
    if (v == 0) {
      while (r >= 11) { const int c = r < 138 ? r : 138; header_tok(h, 18, c - 11); r -= c; }
      if (r >= 3) { header_tok(h, 17, r - 3); r = 0; }
      while (r-- > 0) header_tok(h, 0, 0);
    } else {
      header_tok(h, v, 0);
      --r;
      while (r >= 3) { const int c = r < 6 ? r : 6; header_tok(h, 16, c - 3); r -= c; }
      while (r-- > 0) header_tok(h, v, 0);
    }
  }
}

DFL_HD int cl_order(int i) {
  const uint8_t order[NCL] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  return order[i];
}

// After the code-length code (h.cllen / h.clcode) has been built: HCLEN and the header size.
DFL_HD void header_finish(Header& h) {
  int hclen = NCL;
  while (hclen > 4 && h.cllen[cl_order(hclen - 1)] == 0) --hclen;
  h.hclen = hclen;
  int bits = 3 + 5 + 5 + 4 + 3 * hclen;
  for (int t = 0; t < h.ntok; ++t) {
    const int s = h.tok_sym[t];
    bits += h.cllen[s] + (s == 16 ? 2 : s == 17 ? 3 : s == 18 ? 7 : 0);
  }
  h.bits = bits;
}

// ---- the header in CTA-parallel steps ------------------------------------------------------------
//
// header_tokens() + header_finish() by one thread were ~10-20 % of a segment's time (every other
// thread waits for the header's size before it can place its tokens).  The steps below are called by
// all `nt` threads with their `tid`, a CTA barrier between consecutive steps; tokens, frequencies
// and sizes equal the serial functions' (checked by the host emulation).

DFL_HD void hpar_prepare(const Codes& c, Header& h, int tid, int nt) {
  for (int i = tid; i < NCL; i += nt) h.clfreq[i] = 0;
  if (tid != 0) return;
  int hlit = NLIT;
  while (hlit > 257 && c.len[hlit - 1] == 0) --hlit;
  int hdist = NDIST;
  while (hdist > 1 && c.dlen[hdist - 1] == 0) --hdist;
  h.hlit = hlit;
  h.hdist = hdist;
  h.ntok = 0;
  h.tokbits = 0;
}

DFL_HD void hpar_fill(const Codes& c, Header& h, int tid, int nt) {
  const int total = h.hlit + h.hdist;
  for (int i = tid; i < total; i += nt) h.v[i] = i < h.hlit ? c.len[i] : c.dlen[i - h.hlit];
}

// tokens of a run of r equal lengths v, in the order header_tokens() sends them; PUT(symbol, extra)
template <class Put>
DFL_HD int run_tokens(int v, int r, Put& put) {
  int k = 0;
  if (v == 0) {
    while (r >= 11) { const int c = r < 138 ? r : 138; put(k++, 18, c - 11); r -= c; }
    if (r >= 3) { put(k++, 17, r - 3); r = 0; }
    while (r-- > 0) put(k++, 0, 0);
  } else {
    put(k++, v, 0);
    --r;
    while (r >= 3) { const int c = r < 6 ? r : 6; put(k++, 16, c - 3); r -= c; }
    while (r-- > 0) put(k++, v, 0);
  }
  return k;
}

struct NoPut { DFL_HD void operator()(int, int, int) const {} };

DFL_HD int run_length_at(const Header& h, int total, int i) {
  const int v = h.v[i];
  int r = 1;
  while (i + r < total && h.v[i + r] == v) ++r;
  return r;
}

DFL_HD void hpar_count(Header& h, int tid, int nt) {
  const int total = h.hlit + h.hdist;
  for (int i = tid; i < total; i += nt) {
    int k = 0, r = 0;
    if (i == 0 || h.v[i] != h.v[i - 1]) {
      NoPut np;
      r = run_length_at(h, total, i);
      k = run_tokens(h.v[i], r, np);
    }
    h.tcnt[i] = uint16_t(k);
    h.rlen[i] = uint16_t(r);
  }
}

DFL_HD void hpar_blocks(Header& h, int tid, int nt) {
  const int total = h.hlit + h.hdist;
  for (int b = tid; b * 16 < total; b += nt) {
    int sum = 0;
    for (int i = b * 16; i < b * 16 + 16 && i < total; ++i) sum += h.tcnt[i];
    h.bsum[b] = uint16_t(sum);
  }
}

struct HeaderPut {
  Header* h;
  int base;
  DFL_HD void operator()(int k, int sym, int ext) const {
    h->tok_sym[base + k] = uint8_t(sym);
    h->tok_ext[base + k] = uint8_t(ext);
    DFL_ATOMIC_ADD(&h->clfreq[sym], 1u);
  }
};

DFL_HD void hpar_tokens(Header& h, int tid, int nt) {
  const int total = h.hlit + h.hdist;
  for (int i = tid; i < total; i += nt) {
    if (!h.tcnt[i]) continue;
    int base = 0;
    for (int b = 0; b < (i >> 4); ++b) base += h.bsum[b];
    for (int j = i & ~15; j < i; ++j) base += h.tcnt[j];
    HeaderPut put{&h, base};
    const int r = h.rlen[i];
    const int k = run_tokens(h.v[i], r, put);
    if (i + r == total) h.ntok = base + k;
  }
}

// the code-length code from h.clfreq (S is free by now): a complete code needs two symbols (one thread),
// ranks by all threads, construction by one thread
DFL_HD void hpar_clprepare(Header& h, int tid) {
  if (tid != 0) return;
  int used = 0;
  for (int i = 0; i < NCL; ++i) used += h.clfreq[i] != 0;
  for (int i = 0; used < 2 && i < NCL; ++i)
    if (h.clfreq[i] == 0) { h.clfreq[i] = 1; ++used; }
  h.clused = used;
}

DFL_HD void hpar_clrank(BuildScratch& S, Header& h, int tid, int nt) {
  for (int i = tid; i < NCL; i += nt)
    if (h.clfreq[i]) S.sorted[rank_of(h.clfreq, NCL, i)] = uint16_t(i);
}

DFL_HD void hpar_clcode(BuildScratch& S, Header& h, int tid) {
  if (tid != 0) return;
  build_code(h.clfreq, NCL, h.clused, MAX_CL_BITS, S, h.clcode, h.cllen);
  int hclen = NCL;
  while (hclen > 4 && h.cllen[cl_order(hclen - 1)] == 0) --hclen;
  h.hclen = hclen;
}

DFL_HD void hpar_size(Header& h, int tid, int nt) {
  uint32_t bits = 0;
  for (int t = tid; t < h.ntok; t += nt) {
    const int s = h.tok_sym[t];
    bits += h.cllen[s] + (s == 16 ? 2 : s == 17 ? 3 : s == 18 ? 7 : 0);
  }
  if (bits) { DFL_ATOMIC_ADD(&h.tokbits, bits); }
}

DFL_HD void hpar_finish(Header& h, int tid) {
  if (tid == 0) h.bits = 3 + 5 + 5 + 4 + 3 * h.hclen + int(h.tokbits);
}

DFL_HD void header_emit(const Header& h, BitWriter& bw) {
  bw.put(0, 1);                 // BFINAL = 0: the stream is closed by the caller
  bw.put(2, 2);                 // BTYPE = 10, dynamic Huffman
  bw.put(uint32_t(h.hlit - 257), 5);
  bw.put(uint32_t(h.hdist - 1), 5);
  bw.put(uint32_t(h.hclen - 4), 4);
  for (int i = 0; i < h.hclen; ++i) bw.put(h.cllen[cl_order(i)], 3);
  for (int t = 0; t < h.ntok; ++t) {
    const int s = h.tok_sym[t];
    bw.put(h.clcode[s], h.cllen[s]);
    if (s == 16) bw.put(h.tok_ext[t], 2);
    else if (s == 17) bw.put(h.tok_ext[t], 3);
    else if (s == 18) bw.put(h.tok_ext[t], 7);
  }
}

// ---- stored blocks (fallback when the Huffman block would be larger) ------------------------------

DFL_HD long long stored_size(long long n) { return n + 5 * ((n + STORED_MAX - 1) / STORED_MAX); }

// bytes [s, e) of a segment of n bytes -> their place in a sequence of stored blocks
DFL_HD void stored_copy(const uint8_t* seg, long long n, long long s, long long e, uint8_t* out) {
  for (long long i = s; i < e; ++i) {
    const long long k = i / STORED_MAX;
    if (i == k * STORED_MAX) {
      const long long rem = n - i;
      const unsigned L = unsigned(rem < STORED_MAX ? rem : STORED_MAX);
      uint8_t* h = out + i + 5 * k;
      h[0] = 0;                               // BFINAL = 0, BTYPE = 00, padding
      h[1] = uint8_t(L & 255); h[2] = uint8_t(L >> 8);
      h[3] = uint8_t(~L & 255); h[4] = uint8_t((~L >> 8) & 255);
    }
    out[i + 5 * (k + 1)] = seg[i];
  }
}

// ---- the serial part of a segment (one thread) ------------------------------------------------------

// The distance code of a segment from the frequencies of its distance symbols (one thread; at most
// MAX_CAND symbols are in use).  A complete code needs two symbols: like zlib, unused ones are
// sent to fill up.  dhist is changed.
DFL_HD void distance_code(uint32_t* dhist, BuildScratch& S, Codes& c) {
  int used = 0;
  for (int i = 0; i < NDIST; ++i) used += dhist[i] != 0;
  for (int i = 0; used < 2 && i < NDIST; ++i)
    if (dhist[i] == 0) { dhist[i] = 1; ++used; }
  for (int i = 0; i < NDIST; ++i)
    if (dhist[i]) S.sorted[rank_of(dhist, NDIST, i)] = uint16_t(i);
  build_code(dhist, NDIST, used, MAX_LIT_BITS, S, c.dcode, c.dlen);
}

// The same in steps for the kernel (one thread / all threads / one thread, a CTA barrier between them; they
// run beside par_rank, par_tree and par_count): the 900 comparisons of the ranks by one thread were a tenth
// of a segment's time.  S.m carries the number of symbols in use.
DFL_HD void dpar_prepare(uint32_t* dhist, BuildScratch& S, int tid) {
  if (tid != 0) return;
  int used = 0;
  for (int i = 0; i < NDIST; ++i) used += dhist[i] != 0;
  for (int i = 0; used < 2 && i < NDIST; ++i)
    if (dhist[i] == 0) { dhist[i] = 1; ++used; }
  S.m = uint32_t(used);
}

DFL_HD void dpar_rank(const uint32_t* dhist, BuildScratch& S, int tid, int nt) {
  for (int i = tid; i < NDIST; i += nt)
    if (dhist[i]) S.sorted[rank_of(dhist, NDIST, i)] = uint16_t(i);
}

DFL_HD void dpar_build(const uint32_t* dhist, BuildScratch& S, Codes& c, int tid) {
  if (tid != 0) return;
  if (S.m == 2) {                        // runs only: two codes of one bit, the lower symbol gets 0
    int k = 0;
    for (int i = 0; i < NDIST; ++i) {
      c.dlen[i] = dhist[i] ? 1 : 0;
      c.dcode[i] = dhist[i] ? uint16_t(k++) : uint16_t(0);
    }
    return;
  }
  build_code(dhist, NDIST, int(S.m), MAX_LIT_BITS, S, c.dcode, c.dlen);
}

// hist: frequencies of the literal/length symbols of the segment (EOB counted once, so at least
// two symbols are used); S.sorted: the used symbols by ascending (frequency, symbol).  Builds the literal/length
// code, the header tokens and the code-length code; c.dlen is the finished distance code.
DFL_HD void segment_header(BuildScratch& S, const Codes& c, Header& h);

DFL_HD void segment_build(const uint32_t* hist, BuildScratch& S, Codes& c, Header& h) {
  int m = 0;
  for (int i = 0; i < NLIT; ++i) m += hist[i] != 0;
  build_code(hist, NLIT, m, MAX_LIT_BITS, S, c.code, c.len);
  segment_header(S, c, h);
}

// the block header for the finished codes (one thread)
DFL_HD void segment_header(BuildScratch& S, const Codes& c, Header& h) {
  header_tokens(c.len, c.dlen, h);
  int used = 0;
  for (int i = 0; i < NCL; ++i) used += h.clfreq[i] != 0;
  for (int i = 0; used < 2 && i < NCL; ++i)            // a complete code needs two symbols
    if (h.clfreq[i] == 0) { h.clfreq[i] = 1; ++used; }
  for (int i = 0; i < NCL; ++i)
    if (h.clfreq[i]) S.sorted[rank_of(h.clfreq, NCL, i)] = uint16_t(i);
  build_code(h.clfreq, NCL, used, MAX_CL_BITS, S, h.clcode, h.cllen);
  header_finish(h);
}

// size in bytes of the segment's dynamic block: header, tokens (token_bits), end of block, then
// the empty stored block that realigns the stream (3 bits, padding, 00 00 FF FF)
DFL_HD long long dynamic_size(const Header& h, const Codes& c, long long token_bits) {
  const long long bits = h.bits + token_bits + c.len[EOB] + 3;
  return (bits + 7) / 8 + 4;
}

// what the last writer of a segment appends after its tokens
DFL_HD void segment_close(const Codes& c, BitWriter& bw) {
  bw.put(c.code[EOB], c.len[EOB]);
  bw.put(0, 3);                   // BFINAL = 0, BTYPE = 00
  bw.align_byte();
  bw.put(0x0000u, 16);            // LEN = 0
  bw.put(0xffffu, 16);            // NLEN
}

}  // namespace dfl
}  // namespace vcfb
