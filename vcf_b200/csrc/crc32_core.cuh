// CRC-32 (the zip / zlib / PNG checksum: reflected polynomial 0xEDB88320, initial value and final
// XOR 0xFFFFFFFF) in pieces that combine -- the checksum a zip member needs next to the deflate
// stream of kernels_deflate.cu (np.savez_compressed, /root/reference src/z_lib.py:19-23).
//
//   crc(A || B) = crc(A) * x^(8 |B|) mod p  XOR  crc(B)        (polynomials over GF(2))
//
// holds for the finished CRC values, so with the input cut into chunks
//   crc(M) = XOR_i  crc(chunk_i) * x^(8 * bytes after chunk i) mod p :
// every thread checksums its own chunk, shifts the result by its own distance to the end and the
// CTA XORs everything into one word.  __host__ __device__ like deflate_core.cuh: the CPU suite runs
// the same functions through tests/emul/deflate_emul.cpp against zlib.crc32.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define CRC_HD __host__ __device__ __forceinline__
#else
#define CRC_HD inline
#endif

namespace vcfb {
namespace crc {

constexpr uint32_t POLY = 0xEDB88320u;

// a(x) * b(x) mod p(x); bit 31 is the coefficient of x^0
CRC_HD uint32_t multmodp(uint32_t a, uint32_t b) {
  uint32_t p = 0;
  for (uint32_t m = 1u << 31; m; m >>= 1) {
    if (a & m) p ^= b;
    b = (b & 1u) ? (b >> 1) ^ POLY : b >> 1;
  }
  return p;
}

struct Powers {            // x2n[k] = x^(2^k) mod p
  uint32_t x2n[32];
};

CRC_HD void make_powers(Powers& P) {
  uint32_t p = 1u << 30;   // x^1
  P.x2n[0] = p;
  for (int k = 1; k < 32; ++k) { p = multmodp(p, p); P.x2n[k] = p; }
}

// x^(8 n) mod p
CRC_HD uint32_t x8nmodp(const Powers& P, unsigned long long n) {
  uint32_t p = 1u << 31;   // x^0
  for (int k = 3; n; n >>= 1, ++k)
    if (n & 1ull) p = multmodp(P.x2n[k & 31], p);
  return p;
}

CRC_HD uint32_t table_entry(uint32_t i) {          // the byte-wise table
  uint32_t c = i;
  for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ POLY : c >> 1;
  return c;
}

// finished CRC of src[s, e) with the 256-entry table `tab`
CRC_HD uint32_t chunk_crc(const uint8_t* src, long long s, long long e, const uint32_t* tab) {
  uint32_t c = 0xFFFFFFFFu;
  long long p = s;
  for (; p < e && (p & 7); ++p) c = tab[(c ^ src[p]) & 0xff] ^ (c >> 8);
  for (; p + 8 <= e; p += 8) {                     // one aligned 64-bit load per 8 bytes
    uint64_t w = *reinterpret_cast<const uint64_t*>(src + p);
    for (int k = 0; k < 8; ++k) { c = tab[(c ^ uint32_t(w)) & 0xff] ^ (c >> 8); w >>= 8; }
  }
  for (; p < e; ++p) c = tab[(c ^ src[p]) & 0xff] ^ (c >> 8);
  return ~c;
}

// what thread-chunk [s, e) of an input of n bytes contributes to the CRC of the whole input
CRC_HD uint32_t chunk_term(const uint8_t* src, long long n, long long s, long long e, const uint32_t* tab, const Powers& P) {
  if (s >= e) return 0u;
  return multmodp(x8nmodp(P, (unsigned long long)(n - e)), chunk_crc(src, s, e, tab));
}

}  // namespace crc
}  // namespace vcfb
