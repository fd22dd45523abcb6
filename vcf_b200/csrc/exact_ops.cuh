// Individually rounded IEEE operations (EXACT=true) vs. contractible ones.
//
// The exact set goes through the CUDA rounding intrinsics, which nvcc never
// merges into a fused multiply-add; this is what makes the generated DCT
// codelets (dct_codelets.cuh) bit-identical to scipy/pocketfft, whose x86-64
// build rounds every product and sum separately.
#pragma once

namespace vcfb {

template <typename T, bool EXACT> struct Ops;

template <> struct Ops<float, true> {
  __device__ __forceinline__ static float add(float a, float b) { return __fadd_rn(a, b); }
  __device__ __forceinline__ static float sub(float a, float b) { return __fsub_rn(a, b); }
  __device__ __forceinline__ static float mul(float a, float b) { return __fmul_rn(a, b); }
  __device__ __forceinline__ static float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  __device__ __forceinline__ static float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Ops<float, false> {
  __device__ __forceinline__ static float add(float a, float b) { return a + b; }
  __device__ __forceinline__ static float sub(float a, float b) { return a - b; }
  __device__ __forceinline__ static float mul(float a, float b) { return a * b; }
  __device__ __forceinline__ static float fma(float a, float b, float c) { return fmaf(a, b, c); }
  __device__ __forceinline__ static float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Ops<double, true> {
  __device__ __forceinline__ static double add(double a, double b) { return __dadd_rn(a, b); }
  __device__ __forceinline__ static double sub(double a, double b) { return __dsub_rn(a, b); }
  __device__ __forceinline__ static double mul(double a, double b) { return __dmul_rn(a, b); }
  __device__ __forceinline__ static double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  __device__ __forceinline__ static double div(double a, double b) { return __ddiv_rn(a, b); }
};
template <> struct Ops<double, false> {
  __device__ __forceinline__ static double add(double a, double b) { return a + b; }
  __device__ __forceinline__ static double sub(double a, double b) { return a - b; }
  __device__ __forceinline__ static double mul(double a, double b) { return a * b; }
  __device__ __forceinline__ static double fma(double a, double b, double c) { return ::fma(a, b, c); }
  __device__ __forceinline__ static double div(double a, double b) { return __ddiv_rn(a, b); }
};

}  // namespace vcfb
