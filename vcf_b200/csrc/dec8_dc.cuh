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


// ---- distortion statistics fused into the float64 decoders (src/RDE.py:41-49) --------------
// Each lane compares the 48 bytes it owns in a finished half-tile row (16 pixels of two blocks,
// starting on a pixel boundary, so byte j of word k belongs to channel (4k + j) mod 3) with the
// same bytes of the original frame: per-channel sum of squared differences through a masked
// dp4a of the byte-wise absolute difference with itself, and the signed sum of differences.
struct SseAcc {
  unsigned ph[3];
  int sx, sy, it;
  unsigned long long tot[3];
  long long sdiff;
};
__device__ __forceinline__ void sse_reset(SseAcc& A) {
#pragma unroll
  for (int p = 0; p < 3; ++p) { A.ph[p] = 0; A.tot[p] = 0; }
  A.sx = A.sy = A.it = 0;
  A.sdiff = 0;
}
__device__ __forceinline__ void sse_flush(SseAcc& A) {
#pragma unroll
  for (int p = 0; p < 3; ++p) { A.tot[p] += A.ph[p]; A.ph[p] = 0; }
  A.sdiff += (long long)A.sx - A.sy;
  A.sx = A.sy = A.it = 0;
}
__device__ __forceinline__ void sse_row48(SseAcc& A, const uint4 (&orig)[3], const unsigned char* mine) {
  constexpr unsigned M[3][3] = {{0xFF0000FFu, 0x0000FF00u, 0x00FF0000u},
                                {0x00FF0000u, 0xFF0000FFu, 0x0000FF00u},
                                {0x0000FF00u, 0x00FF0000u, 0xFF0000FFu}};
  const uint4* p = reinterpret_cast<const uint4*>(mine);
  const uint4 y0 = p[0], y1 = p[1], y2 = p[2];
  const unsigned xa[12] = {orig[0].x, orig[0].y, orig[0].z, orig[0].w, orig[1].x, orig[1].y,
                           orig[1].z, orig[1].w, orig[2].x, orig[2].y, orig[2].z, orig[2].w};
  const unsigned ya[12] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w, y2.x, y2.y, y2.z, y2.w};
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const unsigned d = __vabsdiffu4(xa[k], ya[k]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const unsigned dm = d & M[(4 * k) % 3][c];
      A.ph[c] = __dp4a(dm, dm, A.ph[c]);
    }
    A.sx = __dp4a(xa[k], 0x01010101u, unsigned(A.sx));
    A.sy = __dp4a(ya[k], 0x01010101u, unsigned(A.sy));
  }
  if (++A.it == 1024) sse_flush(A);        // 1024 * 48 * 65025 < 2^32
}
// end of the kernel: one set of atomics per warp
__device__ __forceinline__ void sse_finish(SseAcc& A, unsigned long long* stats, int lane) {
  sse_flush(A);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int p = 0; p < 3; ++p) A.tot[p] += __shfl_xor_sync(0xffffffffu, A.tot[p], o);
    A.sdiff += __shfl_xor_sync(0xffffffffu, A.sdiff, o);
  }
  if (lane == 0) {
#pragma unroll
    for (int p = 0; p < 3; ++p)
      if (A.tot[p]) atomicAdd(stats + VCFB_STAT_SSE_R + p, A.tot[p]);
    if (A.sdiff) atomicAdd(stats + VCFB_STAT_SUMDIFF, (unsigned long long)A.sdiff);
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
