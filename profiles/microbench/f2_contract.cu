#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long bits(float2 v) { return (unsigned long long)__float_as_uint(v.x) | ((unsigned long long)__float_as_uint(v.y) << 32); }
__device__ __forceinline__ float2 from(unsigned long long b) { return make_float2(__uint_as_float(unsigned(b)), __uint_as_float(unsigned(b >> 32))); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(bits(a)), "l"(bits(b))); return from(d); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(bits(a)), "l"(bits(b))); return from(d); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(bits(a)), "l"(bits(b)), "l"(bits(c))); return from(d); }
extern "C" __global__ void k_add(const float2* x, const float2* y, float2* o) {
  int i = threadIdx.x; o[i] = add2(mul2(x[i], make_float2(0.7071f, 0.7071f)), mul2(y[i], make_float2(0.3827f, 0.3827f)));
}
extern "C" __global__ void k_sub(const float2* x, const float2* y, float2* o) {
  int i = threadIdx.x; o[i] = fma2(mul2(y[i], make_float2(0.3827f, 0.3827f)), make_float2(-1.f, -1.f), mul2(x[i], make_float2(0.7071f, 0.7071f)));
}
extern "C" __global__ void k_add1(const float2* x, const float2* y, float2* o) {
  int i = threadIdx.x; o[i] = fma2(mul2(x[i], make_float2(0.7071f, 0.7071f)), make_float2(1.f, 1.f), mul2(y[i], make_float2(0.3827f, 0.3827f)));
}
