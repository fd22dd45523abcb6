// Exact evaluation of blocks without AC indices (float64 decoders of the B=8 fast path).
#pragma once

#include "fast_common.cuh"

namespace vcfb {
namespace fast {

__device__ __forceinline__ int clamp255(int v) { return min(max(v, 0), 255); }

// per-lane constants of the kernel
struct Lane {
  int i1, G1, G2, y2;     // pass 1: (coefficient column, block pair); pass 2: (block pair, pixel row)
  int sh0;                // bit offset of the lane's 6-byte run inside its first word (0 or 16)
  int q;
};

// ---- tier 2a: blocks without AC indices -------------------------------------------------
// With X[u][i] = 0 for (u,i) != (0,0), every sum in dct8_inv adds an exact zero and every
// other product is a zero, so all 8 outputs equal  X * c0  (c0 = pocketfft's sqrt(2) constant,
// first multiplication of dct8_inv) -- once per axis.  Pinned against the real scipy by
// tests/test_oracle.py::test_dc_only_block_chain.
__device__ __forceinline__ double dc_chain(int byte, int q) {
  constexpr double C0 = 0x1.6a09e667f3bcdp+0;
  const double X = __int2double_rn(byte * q - 128 * q);
  return __dmul_rn(__dmul_rn(X, C0), C0);
}
__device__ __forceinline__ unsigned dc_rgb(unsigned ycc, int q) {
  constexpr double SCALE = p2(2 * M8I::exp(0));
  const double Y = dc_chain(int(ycc & 255u), q), Co = dc_chain(int((ycc >> 8) & 255u), q),
               Cg = dc_chain(int((ycc >> 16) & 255u), q);
  const double R = __fma_rn(__dsub_rn(__dadd_rn(Y, Co), Cg), SCALE, 128.0);
  const double G = __fma_rn(__dadd_rn(Y, Cg), SCALE, 128.0);
  const double B = __fma_rn(__dsub_rn(__dsub_rn(Y, Co), Cg), SCALE, 128.0);
  return unsigned(clamp255(__double2int_rz(R))) | unsigned(clamp255(__double2int_rz(G))) << 8 |
         unsigned(clamp255(__double2int_rz(B))) << 16;
}

// Writes the constant colour of the blocks selected by mask8 (bit 2*pair + block of the pair)
// into row y2.  w0, w1: the lane's words of coefficient row u = 0 (lanes with i1 == 0 hold the DC).
__device__ __forceinline__ void dc_blocks(const Lane& L, uint32_t w0, uint32_t w1, unsigned char* tb, int h,
                                          unsigned mask8) {
  const uint32_t lo = __funnelshift_r(w0, w1, L.sh0), hi = w1 >> L.sh0;
  unsigned rgbA = 0, rgbB = 0;
  if (L.i1 == 0) {
    rgbA = dc_rgb(lo, L.q);
    rgbB = dc_rgb(__byte_perm(lo, hi, 0x0543), L.q);
  }
  rgbA = __shfl_sync(0xffffffffu, rgbA, 8 * L.G2);
  rgbB = __shfl_sync(0xffffffffu, rgbB, 8 * L.G2);
  uint2* o = reinterpret_cast<uint2*>(tb + L.y2 * (WT * 3) + 192 * h + 48 * L.G2);
  const unsigned m = mask8 >> (2 * L.G2);
  if (m & 1u) {
    const uint32_t a = __byte_perm(rgbA, 0, 0x0210), b = __byte_perm(rgbA, 0, 0x1021), c = __byte_perm(rgbA, 0, 0x2102);
    o[0] = make_uint2(a, b);
    o[1] = make_uint2(c, a);
    o[2] = make_uint2(b, c);
  }
  if (m & 2u) {
    const uint32_t a = __byte_perm(rgbB, 0, 0x0210), b = __byte_perm(rgbB, 0, 0x1021), c = __byte_perm(rgbB, 0, 0x2102);
    o[3] = make_uint2(a, b);
    o[4] = make_uint2(c, a);
    o[5] = make_uint2(b, c);
  }
}


// ac24 of one half-tile: bits 6*pair + 3*block + {0,1,2} all set when the block carries any AC
// index.  nz0, nz1: OR over the coefficient rows of (word ^ 0x80808080) of the lane's two words,
// with the DC position (lane i1 == 0, row 0) masked out.
__device__ __forceinline__ unsigned half_ac24(const Lane& L, uint32_t nz0, uint32_t nz1) {
  const uint32_t lo = __funnelshift_r(nz0, nz1, L.sh0), hi = nz1 >> L.sh0;
  const unsigned f = ((lo & 0x00ffffffu) ? 7u : 0u) | (((lo >> 24) | (hi << 8 & 0x00ffff00u)) ? 56u : 0u);
  return __reduce_or_sync(0xffffffffu, f << (6 * L.G1));
}

}  // namespace fast
}  // namespace vcfb
