// Shared pieces of the B=8 fast-path translation units (kernels_fast.cu, kernels_packed.cu).
#pragma once

#include <stdlib.h>

#include "common.cuh"
#include "dct_codelets.cuh"
#include "tma.cuh"

namespace vcfb {
namespace fast {

// ---- geometry of the warp-autonomous pipelines --------------------------------------
constexpr int WT = 128;               // pixels per warp tile (16 blocks)
constexpr int ROWW = WT * 3 / 4;      // 96 words per raw RGB row
constexpr int TILE = 8 * WT * 3;      // 3072 bytes: raw tile == index tile
constexpr int NSTAGE = 3;

constexpr int ENC_FP = 132;           // floats per intermediate row (132 = 4 mod 32: conflict-free)
constexpr int ENC_F_BYTES = 3 * 8 * ENC_FP * 4;
constexpr int ENC_WARP_SMEM = 22016;  // ring 9216 + F 12672 + 3 mbarriers, rounded to 128
__host__ __device__ constexpr int enc_warp_smem(int nst) { return (nst * TILE + ENC_F_BYTES + 8 * nst + 127) / 128 * 128; }

template <typename T> struct DecL;
// Intermediate X[c][y][i][16 blocks]: row pitch P and plane pitch PP (elements) are
// chosen so that pass-1 stores (lanes = 8 values of i) and pass-2 loads (lanes = 4
// block groups x 2 values of y) both touch 32 distinct banks per wavefront.
template <> struct DecL<float> {
  static constexpr int P = 20;              // 20 words = 4*odd  (mod 32)
  static constexpr int PP = 8 * P + 16;     // 176 = 16 (mod 32): odd y lands on the other 16 banks
  static constexpr int F_BYTES = 3 * 8 * PP * 4;
  static constexpr int WARP_SMEM = 26240;
};
template <> struct DecL<double> {
  static constexpr int P = 18;              // 36 words = 4*odd (mod 32)
  static constexpr int PP = 8 * P + 2;      // 292 words = 4 (mod 8): odd y fills the gaps
  static constexpr int F_BYTES = 3 * 8 * PP * 8;
  static constexpr int WARP_SMEM = 37376;
};
static_assert(NSTAGE * TILE + ENC_F_BYTES + 8 * NSTAGE <= ENC_WARP_SMEM, "encode smem");
static_assert(NSTAGE * TILE + DecL<float>::F_BYTES + 8 * NSTAGE <= DecL<float>::WARP_SMEM, "decode f32 smem");
static_assert(NSTAGE * TILE + DecL<double>::F_BYTES + 8 * NSTAGE <= DecL<double>::WARP_SMEM, "decode f64 smem");

struct FastArgs {
  int ntiles, tiles_x, ny, top;
  float q;
  float qtab[8][3];   // [u][c]: sgn_u * 2^(exp_u + colour exp_c + min_i exp_i) (/ q when q is 2^k)
  unsigned long long* stats;   // packed encoder: accumulates NONZERO and SUMABS when set
};

struct FastDecArgs {
  int ntiles, tiles_x, ny, top;
  int q;
  // Device-side dispatch between the three float64 decoders (kernels_fast.cu "probe"): when
  // `choice` is set, the kernel runs only if *choice == kind and exits at once otherwise.
  const int* choice;
  int kind;
  // distortion statistics fused into the float64 decoders (kernels instantiated with SSE = true)
  const uint8_t* original;
  unsigned long long* stats;
  long long frame_bytes;       // H * W * 3
  int row_bytes;               // W * 3
};
enum { DEC_EXACT = 0, DEC_EXACT_DCSKIP = 1, DEC_TWO_TIER = 2 };
__device__ __forceinline__ bool not_chosen(const FastDecArgs& a) {
  return a.choice != nullptr && *reinterpret_cast<const volatile int*>(a.choice) != a.kind;
}

__host__ __device__ constexpr double p2(int e) {
  double r = 1.0;
  for (int i = 0; i < (e < 0 ? -e : e); ++i) r = e < 0 ? r * 0.5 : r * 2.0;
  return r;
}

using M8F = dct8_fwd_meta;
using M8I = dct8_inv_meta;
__host__ __device__ constexpr int min_exp8() {
  int m = M8F::exp(0);
  for (int i = 1; i < 8; ++i) m = M8F::exp(i) < m ? M8F::exp(i) : m;
  return m;
}
__host__ __device__ constexpr bool inv8_uniform() {
  for (int i = 0; i < 8; ++i)
    if (M8I::exp(i) != M8I::exp(0) || M8I::sgn(i) != 1) return false;
  return true;
}
static_assert(inv8_uniform(), "decode fast path assumes a uniform lazy scale of dct8_inv");

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

constexpr int MAGIC_I = 0x4B400000;     // bits of 12582912.0f = 1.5 * 2^23
constexpr float MAGIC_F = 12582912.0f;

// (float) of the exact integer dot product of the pixel bytes with small signed
// coefficients: the accumulator starts at the bit pattern of 1.5*2^23, so the dp4a
// result *is* the float 1.5*2^23 + n; one subtraction removes the bias exactly.
__device__ __forceinline__ float dotf(unsigned px, int coef, int bias) {
  return __int_as_float(dp4a_us(px, coef, MAGIC_I + bias)) - MAGIC_F;
}

__device__ __forceinline__ unsigned pack4(int k0, int k1, int k2, int k3) {
  const unsigned lo = __byte_perm(unsigned(k0), unsigned(k1), 0x0040);
  const unsigned hi = __byte_perm(unsigned(k2), unsigned(k3), 0x0040);
  return __byte_perm(lo, hi, 0x5410);
}

// Per-warp tile walker: which tiles this warp owns, and the TMA issue for them.
struct Walker {
  int tile, stride, ntiles, per_frame, tiles_x, top;
  __device__ __forceinline__ void coords(int t, int& f, int& by, int& tx) const {
    f = t / per_frame;
    const int rem = t - f * per_frame;
    by = rem / tiles_x;
    tx = rem - by * tiles_x;
  }
};


int sm_count();

// Development knob "<warps per CTA>x<CTAs per SM>" from the environment (e.g. VCFB_ENC_CFG=5x2)
// as the integer 10*warps + ctas; 0 when unset or malformed.
inline int dev_cfg(const char* name) {
  const char* e = getenv(name);
  if (!e || e[0] < '1' || e[0] > '9' || e[1] != 'x' || e[2] < '1' || e[2] > '9' || e[3] != 0) return 0;
  return (e[0] - '0') * 10 + (e[2] - '0');
}

}  // namespace fast

// packed (f32x2) exact encoder, compiled in its own translation unit with -fmad=false
int launch_encode_packed(int nwarps_cfg, bool qpow2, const CUtensorMap& in_map, const CUtensorMap& out_map,
                         const fast::FastArgs& fa, cudaStream_t s);

// two-tier float64 decoder (kernels_dec2t.cu)
int launch_decode_2t(int cfg, const CUtensorMap& in_map, const CUtensorMap& out_map, const fast::FastDecArgs& fa,
                     cudaStream_t s);

// float32 fast-mode decoder (kernels_dec32.cu)
int launch_decode_f32a(int cfg, const CUtensorMap& in_map, const CUtensorMap& out_map, const fast::FastDecArgs& fa,
                       cudaStream_t s);

}  // namespace vcfb
