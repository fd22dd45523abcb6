// Individually rounded IEEE operations (EXACT=true) vs. contractible ones.
//
// The exact set goes through the CUDA rounding intrinsics, which nvcc never
// merges into a fused multiply-add; this is what makes the generated DCT
// codelets (dct_codelets.cuh) bit-identical to scipy/pocketfft, whose x86-64
// build rounds every product and sum separately.
#pragma once

namespace vcfb {

template <typename T, bool EXACT> struct Ops;

// constant of the working type from its float32 / float64 literal
template <typename T> __device__ __forceinline__ T konst(float f32, double f64);
template <> __device__ __forceinline__ float konst<float>(float f32, double) { return f32; }
template <> __device__ __forceinline__ double konst<double>(float, double f64) { return f64; }
template <> __device__ __forceinline__ float2 konst<float2>(float f32, double) { return make_float2(f32, f32); }

template <> struct Ops<float, true> {
  __device__ __forceinline__ static float add(float a, float b) { return __fadd_rn(a, b); }
  __device__ __forceinline__ static float sub(float a, float b) { return __fsub_rn(a, b); }
  __device__ __forceinline__ static float mul(float a, float b) { return __fmul_rn(a, b); }
  __device__ __forceinline__ static float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  __device__ __forceinline__ static float div(float a, float b) { return __fdiv_rn(a, b); }
  __device__ __forceinline__ static float neg(float a) { return -a; }
};
// Two independent float32 lanes per instruction (sm_100 FADD2 / FMUL2 / FFMA2): each lane
// is rounded exactly like the scalar instruction, so the codelets stay bit-exact while
// the floating-point work takes half the issue slots.  a - b is fma(b, -1, a): the
// product is exact, so the single rounding is that of the subtraction.
// Contraction hazard: unlike the scalar __fadd_rn / __fmul_rn, neither the float2
// intrinsics of CUDA 12.9 (__fadd2_rn, __fmul2_rn) nor PTX mul.rn.f32x2 + add.rn.f32x2 (nor
// fma.rn.f32x2 with a factor of 1) are safe: ptxas 12.9 fuses them into one FFMA2 (seen in
// SASS; 45 of 24.9 M indices moved at q = 1).  Code that relies on Ops<float2> for
// bit-exactness must therefore be compiled with -fmad=false (kernels_packed.cu is).
__device__ __forceinline__ unsigned long long f2_bits(float2 v) {
  return (unsigned long long)__float_as_uint(v.x) | ((unsigned long long)__float_as_uint(v.y) << 32);
}
__device__ __forceinline__ float2 f2_from(unsigned long long b) {
  return make_float2(__uint_as_float(unsigned(b)), __uint_as_float(unsigned(b >> 32)));
}
template <bool EXACT> struct Ops<float2, EXACT> {
  __device__ __forceinline__ static float2 add(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(d);
  }
  __device__ __forceinline__ static float2 mul(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(d);
  }
  __device__ __forceinline__ static float2 fma(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
    return f2_from(d);
  }
  __device__ __forceinline__ static float2 sub(float2 a, float2 b) { return fma(b, make_float2(-1.0f, -1.0f), a); }
  __device__ __forceinline__ static float2 neg(float2 a) { return make_float2(-a.x, -a.y); }
};
template <> struct Ops<float, false> {
  __device__ __forceinline__ static float add(float a, float b) { return a + b; }
  __device__ __forceinline__ static float sub(float a, float b) { return a - b; }
  __device__ __forceinline__ static float mul(float a, float b) { return a * b; }
  __device__ __forceinline__ static float fma(float a, float b, float c) { return fmaf(a, b, c); }
  __device__ __forceinline__ static float div(float a, float b) { return __fdiv_rn(a, b); }
  __device__ __forceinline__ static float neg(float a) { return -a; }
};
template <> struct Ops<double, true> {
  __device__ __forceinline__ static double add(double a, double b) { return __dadd_rn(a, b); }
  __device__ __forceinline__ static double sub(double a, double b) { return __dsub_rn(a, b); }
  __device__ __forceinline__ static double mul(double a, double b) { return __dmul_rn(a, b); }
  __device__ __forceinline__ static double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  __device__ __forceinline__ static double div(double a, double b) { return __ddiv_rn(a, b); }
  __device__ __forceinline__ static double neg(double a) { return -a; }
};
template <> struct Ops<double, false> {
  __device__ __forceinline__ static double add(double a, double b) { return a + b; }
  __device__ __forceinline__ static double sub(double a, double b) { return a - b; }
  __device__ __forceinline__ static double mul(double a, double b) { return a * b; }
  __device__ __forceinline__ static double fma(double a, double b, double c) { return ::fma(a, b, c); }
  __device__ __forceinline__ static double div(double a, double b) { return __ddiv_rn(a, b); }
  __device__ __forceinline__ static double neg(double a) { return -a; }
};

}  // namespace vcfb
