// Throughput of scalar FADD/FMUL vs packed add/mul.f32x2 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int MODE> __global__ void k(float* out, int iters, float s) {
  float a[16]; unsigned long long p[8];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) p[i] = (unsigned long long)__float_as_uint(a[2*i]) | ((unsigned long long)__float_as_uint(a[2*i+1]) << 32);
  unsigned long long ps = (unsigned long long)__float_as_uint(s) | ((unsigned long long)__float_as_uint(s) << 32);
  int acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {           // 16 scalar ops (8 add + 8 mul)
#pragma unroll
      for (int i = 0; i < 16; i += 2) { a[i] = __fadd_rn(a[i], s); a[i+1] = __fmul_rn(a[i+1], s); }
    } else if (MODE == 1) {    // 8 packed ops = 16 lane-ops
#pragma unroll
      for (int i = 0; i < 8; i += 2) { p[i] = add2(p[i], ps); p[i+1] = mul2(p[i+1], ps); }
    } else if (MODE == 2) {    // 16 scalar FP + 8 integer ops
#pragma unroll
      for (int i = 0; i < 16; i += 2) { a[i] = __fadd_rn(a[i], s); a[i+1] = __fmul_rn(a[i+1], s); acc = __byte_perm(acc, it, 0x3214 + i); }
    } else {                   // 8 packed + 8 integer ops
#pragma unroll
      for (int i = 0; i < 8; i += 2) { p[i] = add2(p[i], ps); p[i+1] = mul2(p[i+1], ps); acc = __byte_perm(acc, it, 0x3214 + i); acc = __byte_perm(acc, it, 0x1230 + i); }
    }
  }
  float r = acc;
  for (int i = 0; i < 16; ++i) r += a[i];
  for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, double lane_ops_per_iter) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  int iters = 20000;
  k<MODE><<<148 * 8, 256>>>(out, 100, 1.0001f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * 8 * 256 * iters * lane_ops_per_iter;
  printf("%-28s %.3f ms  %.1f T lane-ops/s (FP)  err=%s\n", name, ms, ops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() {
  run<0>("scalar fadd/fmul", 16);
  run<1>("packed f32x2", 16);
  run<2>("scalar + 8 int", 16);
  run<3>("packed + 8 int", 16);
  return 0;
}
