// Deflate (RFC 1951) building blocks of the entropy front-end.
//
// The reference hands the uint8 index planes of the transform path to zlib: np.savez_compressed
// in src/z_lib.py:19-23, tifffile's zlib codec in src/TIFF.py:23-31.  The planes are long runs of
// the bias value 128 with sparse, strongly skewed literals, so a run-length parse (matches of
// distance 1 only, zlib's Z_RLE strategy) under a dynamic Huffman code is within a few per cent of
// zlib's default parse -- and, unlike hash-chain matching, it is a pure function of each byte's
// neighbours, so every thread can parse its own piece of the input.
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

// ---- match lengths ------------------------------------------------------------------------------

// RFC 1951 3.2.5: length L in [3, 258] -> symbol 257..285, number of extra bits, extra value
DFL_HD void length_symbol(int L, int* sym, int* ebits, int* eval) {
  if (L == MAX_MATCH) { *sym = 285; *ebits = 0; *eval = 0; return; }
  const int l = L - 3;
  if (l < 8) { *sym = 257 + l; *ebits = 0; *eval = 0; return; }
  int e = 0;                       // floor(log2(l)) - 2, l in [8, 254] -> e in [1, 5]
  for (int t = l >> 3; t; t >>= 1) ++e;
  *sym = 261 + 4 * e + ((l >> e) & 3);
  *ebits = e;
  *eval = l & ((1 << e) - 1);
}

// ---- run-length parse ---------------------------------------------------------------------------

// Tokens of the piece [s, e) of src[0, n): a byte equal to its predecessor starts a match of
// distance 1 when at least MIN_MATCH bytes repeat (the predecessor may lie in the previous piece
// or segment: the decoder has produced it by then).  V::lit(byte) / V::match(length).
// src must be 8-byte aligned.  The input is read as aligned 64-bit words kept in a register (one
// load per 8 bytes: the pieces of a warp's threads lie `piece` bytes apart, so every load costs one
// L1 wavefront per thread -- measured neutral against byte loads); bytes past n are
// never touched.
#if defined(__CUDA_ARCH__)
#define DFL_CTZ64(x) (__ffsll((long long)(x)) - 1)
#else
#define DFL_CTZ64(x) __builtin_ctzll(x)
#endif

DFL_HD uint64_t load_word(const uint8_t* src, long long n, long long base) {
  if (base + 8 <= n) return *reinterpret_cast<const uint64_t*>(src + base);
  uint64_t x = 0;
  for (int k = 0; base + k < n; ++k) x |= uint64_t(src[base + k]) << (8 * k);
  return x;
}

template <class V>
DFL_HD void parse_piece(const uint8_t* src, long long n, long long s, long long e, V& v) {
  long long p = s;
  int prev = p > 0 ? int(src[p - 1]) : -1;
  long long wbase = -8;
  uint64_t w = 0;
  while (p < e) {
    const long long base = p & ~7ll;
    if (base != wbase) { w = load_word(src, n, base); wbase = base; }
    const int b = int((w >> (8 * int(p & 7))) & 0xff);
    if (b == prev) {
      const long long rem = e - p;
      const int lim = rem < MAX_MATCH ? int(rem) : MAX_MATCH;
      const uint64_t splat = 0x0101010101010101ull * uint64_t(prev);
      int L = 0;
      for (;;) {
        const int off = int((p + L) & 7);
        const uint64_t x = (w ^ splat) >> (8 * off);
        const int avail = 8 - off;
        const int z = x ? (DFL_CTZ64(x) >> 3) : avail;    // a differing byte lies inside the `avail` valid ones
        L += z;
        if (z < avail || L >= lim) break;
        wbase += 8;
        w = load_word(src, n, wbase);
      }
      if (L > lim) L = lim;
      if (L >= MIN_MATCH) { v.match(L); p += L; continue; }
    }
    v.lit(b);
    prev = b;
    ++p;
  }
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
};

#if defined(__CUDA_ARCH__)
#define DFL_ATOMIC_ADD(ptr, val) atomicAdd((ptr), (val))
#else
#define DFL_ATOMIC_ADD(ptr, val) (*(ptr) += (val))
#endif

struct CountVisitor {      // symbol frequencies; equal consecutive symbols (runs of 258-byte matches) are added at once
  uint32_t* hist;
  int cur;
  uint32_t cnt;
  DFL_HD void init(uint32_t* h) { hist = h; cur = 0; cnt = 0; }
  DFL_HD void flush() { if (cnt) { DFL_ATOMIC_ADD(hist + cur, cnt); } cnt = 0; }
  DFL_HD void add(int sym) {
    if (sym != cur) { flush(); cur = sym; }
    ++cnt;
  }
  DFL_HD void lit(int b) { add(b); }
  DFL_HD void match(int L) {
    int sym, eb, ev;
    length_symbol(L, &sym, &eb, &ev);
    add(sym);
  }
};

struct SizeVisitor {       // bits the tokens take under `len`
  const uint8_t* len;
  unsigned bits;
  DFL_HD void lit(int b) { bits += len[b]; }
  DFL_HD void match(int L) {
    int sym, eb, ev;
    length_symbol(L, &sym, &eb, &ev);
    bits += len[sym] + eb + 1;      // + the 1-bit distance code
  }
};

struct EmitVisitor {
  const Codes* c;
  BitWriter* bw;
  DFL_HD void lit(int b) { bw->put(c->code[b], c->len[b]); }
  DFL_HD void match(int L) {
    int sym, eb, ev;
    length_symbol(L, &sym, &eb, &ev);
    const int n = c->len[sym];
    // length code, extra bits, distance symbol 0 = code "0" of length 1
    bw->put(uint32_t(c->code[sym]) | (uint32_t(ev) << n), n + eb + 1);
  }
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
  uint8_t tok_sym[NLIT + 2 + 8];   // code-length tokens of the concatenated length arrays
  uint8_t tok_ext[NLIT + 2 + 8];
  int ntok;
  int hlit;                         // literal/length codes sent (>= 257)
  int hclen;                        // code-length codes sent (>= 4)
  uint32_t clfreq[NCL];
  uint16_t clcode[NCL];
  uint8_t cllen[NCL];
  int bits;                         // size of the whole header including BFINAL/BTYPE
};

DFL_HD void header_tok(Header& h, int sym, int ext) {
  h.tok_sym[h.ntok] = uint8_t(sym);
  h.tok_ext[h.ntok] = uint8_t(ext);
  ++h.ntok;
  ++h.clfreq[sym];
}

// Run-length tokens (symbols 16/17/18 of RFC 1951 3.2.7) for the lengths of the literal/length
// code followed by the two distance codes of length 1 (one distance is ever used; like zlib we
// send two so the distance code is complete).
DFL_HD void header_tokens(const uint8_t* litlen, Header& h) {
  int hlit = NLIT;
  while (hlit > 257 && litlen[hlit - 1] == 0) --hlit;
  h.hlit = hlit;
  h.ntok = 0;
  for (int i = 0; i < NCL; ++i) h.clfreq[i] = 0;
  const int total = hlit + 2;
  int i = 0;
  while (i < total) {
    const int v = i < hlit ? litlen[i] : 1;
    int r = 1;
    while (i + r < total && (i + r < hlit ? litlen[i + r] : 1) == v) ++r;
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

DFL_HD void header_emit(const Header& h, BitWriter& bw) {
  bw.put(0, 1);                 // BFINAL = 0: the stream is closed by the caller
  bw.put(2, 2);                 // BTYPE = 10, dynamic Huffman
  bw.put(uint32_t(h.hlit - 257), 5);
  bw.put(1, 5);                 // HDIST = 2 codes
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

// hist: frequencies of the literal/length symbols of the segment (EOB counted once, so at least
// two symbols are used); S.sorted: the used symbols by ascending (frequency, symbol).  Builds the literal/length
// code, the header tokens and the code-length code.
DFL_HD void segment_header(BuildScratch& S, const Codes& c, Header& h);

DFL_HD void segment_build(const uint32_t* hist, BuildScratch& S, Codes& c, Header& h) {
  int m = 0;
  for (int i = 0; i < NLIT; ++i) m += hist[i] != 0;
  build_code(hist, NLIT, m, MAX_LIT_BITS, S, c.code, c.len);
  segment_header(S, c, h);
}

// the block header for the finished literal/length code (one thread)
DFL_HD void segment_header(BuildScratch& S, const Codes& c, Header& h) {
  header_tokens(c.len, h);
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
