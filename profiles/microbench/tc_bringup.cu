// tcgen05 bring-up for the tensor-core tier of the B=8 transform:
//   D[128 x 64] (fp32, TMEM) = A[128 x 64] (fp16: exact small integers) * B[64 x 64]^T (bf16 limb of the
//   64 x 64 Kronecker DCT matrix), the shape of "128 blocks x 64 samples -> 64 coefficients".
// Variants (argv[1]): 0 = A in shared memory, K-major, no swizzle;  1 = A in shared memory, MN-major;
//   2 = A in tensor memory (written with tcgen05.st);  add 10 for B in bf16 with A in fp16 (mixed formats),
//   otherwise both are fp16.
// Prints the maximum deviation from the host product and the cycles of the MMA issue loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_bringup tc_bringup.cu && ./tc_bringup 2
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical no-swizzle layouts (cute/atom/mma_traits_sm100.hpp, make_umma_desc), byte offsets
//   K-major : ((8,n),2):((1,SBO),LBO) in 16-byte units -> row r of a core matrix at +16 r, next 8 rows at +SBO,
//             next 8 elements of K at +LBO
//   MN-major: ((1,n),(8,k)):((X,SBO),(1,LBO))          -> 8 elements of MN contiguous (16 B), k at +16 (k % 8),
//             next 8 of MN at +SBO, next 8 of K at +LBO
__host__ __device__ inline int off_kmajor(int mn, int k, int sbo, int lbo) { return (mn % 8) * 16 + (mn / 8) * sbo + (k / 8) * lbo + (k % 8) * 2; }
__host__ __device__ inline int off_mnmajor(int mn, int k, int sbo, int lbo) { return (mn % 8) * 2 + (k % 8) * 16 + (mn / 8) * sbo + (k / 8) * lbo; }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int lbo, int sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  return d;                          // base offset 0, layout type 0 = no swizzle
}

__global__ void __launch_bounds__(128, 1)
bringup(const uint16_t* __restrict__ Ag, const uint16_t* __restrict__ Bg, float* __restrict__ Dg, int variant,
        int b_bf16, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* As = smem;                    // 16 KB
  unsigned char* Bs = smem + 16384;            // 8 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 8192);
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + 16384 + 8192 + 8);
  const int tid = threadIdx.x, warp = tid >> 5;

  // A: SBO = 128 (next 8 rows follow), LBO = 16 groups * 128 = 2048
  constexpr int A_SBO = 128, A_LBO = 2048;
  // B (N = 64 rows, K-major): SBO = 128, LBO = 8 groups * 128 = 1024
  constexpr int B_SBO = 128, B_LBO = 1024;
  for (int e = tid; e < M * K; e += 128) {
    const int m = e / K, k = e % K;
    const int o = (variant == 1) ? off_mnmajor(m, k, A_SBO, A_LBO) : off_kmajor(m, k, A_SBO, A_LBO);
    *reinterpret_cast<uint16_t*>(As + o) = Ag[e];
  }
  for (int e = tid; e < N * K; e += 128) {
    const int n = e / K, k = e % K;
    *reinterpret_cast<uint16_t*>(Bs + off_kmajor(n, k, B_SBO, B_LBO)) = Bg[e];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> async proxy (tensor core)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_holder)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *tmem_holder;
  const uint32_t d_tmem = tbase;               // columns [0, 64): D
  const uint32_t a_tmem = tbase + 64;          // columns [64, 96): A, two fp16 per column
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  if (variant == 2 || variant == 3) {
    // thread = row of A: 64 fp16 = 32 packed words
    uint32_t w[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) w[j] = (uint32_t)Ag[tid * K + 2 * j] | ((uint32_t)Ag[tid * K + 2 * j + 1] << 16);
#pragma unroll
    for (int c = 0; c < 4; ++c)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a_tmem + lane_base + 8 * c),
                   "r"(w[8 * c + 0]), "r"(w[8 * c + 1]), "r"(w[8 * c + 2]), "r"(w[8 * c + 3]), "r"(w[8 * c + 4]),
                   "r"(w[8 * c + 5]), "r"(w[8 * c + 6]), "r"(w[8 * c + 7])
                   : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();

  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D = F32, A = F16, B = F16 | BF16,
    // K-major B, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (0u << 7) | ((b_bf16 ? 1u : 0u) << 10) | ((variant == 1 ? 1u : 0u) << 15) |
                           (0u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const long long t0 = clock64();
#pragma unroll
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t bdesc = make_desc(smem_u32(Bs) + ks * 2 * B_LBO, B_LBO, B_SBO);
      const uint32_t acc = ks > 0;
      if (variant == 2 || variant == 3) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
            "r"(a_tmem + 8 * ks), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
      } else {
        const uint64_t adesc = make_desc(smem_u32(As) + ks * 2 * A_LBO, A_LBO, A_SBO);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    cycles[0] = clock64() - t0;
  }
  // everyone waits for the MMAs
  {
    asm volatile(
        "{\n\t.reg .pred p;\n\tW1:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D1;\n\tbra W1;\n\tD1:\n\t}" ::"r"(smem_u32(bar))
        : "memory");
  }
  if (tid == 0) cycles[1] = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
  for (int c = 0; c < N / 8; ++c) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(d_tmem + lane_base + 8 * c)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 8; ++j) Dg[tid * N + 8 * c + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tbase) : "memory");
}


// ---- issue / execution rate of back-to-back MMAs (argv[1] = 100 + variant, argv[2] = N) ---------------
__global__ void __launch_bounds__(128, 1) mma_rate(int variant, int n_dim, int reps, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* As = smem;
  unsigned char* Bs = smem + 16384;            // up to 256 rows x 64 k x 2 B = 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + 16384 + 32768 + 8);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_holder)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *tmem_holder;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(n_dim >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const int b_lbo = n_dim / 8 * 128;
    uint64_t bdesc[4], adesc[4];
    for (int ks = 0; ks < 4; ++ks) {
      bdesc[ks] = make_desc(smem_u32(Bs) + ks * 2 * b_lbo, b_lbo, 128);
      adesc[ks] = make_desc(smem_u32(As) + ks * 2 * 2048, 2048, 128);
    }
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (variant == 2)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tbase),
                       "r"(tbase + 256 + 8 * ks), "l"(bdesc[ks]), "r"(idesc), "r"(1u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase),
                       "l"(adesc[ks]), "l"(bdesc[ks]), "r"(idesc), "r"(1u) : "memory");
      }
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("{\n\t.reg .pred p;\n\tW2:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D2;\n\tbra W2;\n\tD2:\n\t}" ::"r"(smem_u32(bar)) : "memory");
    const long long t2 = clock64();
    cycles[0] = t1 - t0;
    cycles[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

int rate_main(int variant, int n_dim) {
  long long* dC;
  CK(cudaMalloc(&dC, 16));
  const int smem_bytes = 16384 + 32768 + 64;
  CK(cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  const int reps = 500;
  for (int it = 0; it < 2; ++it) {
    mma_rate<<<1, 128, smem_bytes>>>(variant, n_dim, reps, dC);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
  }
  long long cyc[2];
  CK(cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost));
  printf("rate: A %s, M 128 N %d K 16: %d MMAs, issue %.1f cycles each, issue + drain %.1f cycles each\n",
         variant == 2 ? "in TMEM" : "in smem", n_dim, reps * 4, double(cyc[0]) / (reps * 4), double(cyc[1]) / (reps * 4));
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && atoi(argv[1]) >= 100) return rate_main(atoi(argv[1]) - 100, argc > 2 ? atoi(argv[2]) : 64);
  const int arg = argc > 1 ? atoi(argv[1]) : 0;
  const int variant = arg % 10, b_bf16 = arg >= 10;
  std::vector<uint16_t> A(M * K), B(N * K);
  std::vector<float> Af(M * K), Bf(N * K);
  srand(1);
  for (int i = 0; i < M * K; ++i) {
    const int v = rand() % 256 - 128;
    Af[i] = (float)v;
    __half h = __float2half((float)v);
    A[i] = *reinterpret_cast<uint16_t*>(&h);
    if (variant == 3) {                 // fp16 subnormal with the bits of the byte: (v + 128) * 2^-24
      A[i] = (uint16_t)(v + 128);
      Af[i] = (float)(v + 128);         // compared after scaling D by 2^24
    }
  }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      // 2-D DCT-II basis: coefficient n = (u, v), sample k = (r, x)
      const int u = n / 8, v = n % 8, r = k / 8, x = k % 8;
      const double cu = u ? 0.5 : sqrt(0.125), cv = v ? 0.5 : sqrt(0.125);
      const double val = cu * cv * cos((2 * r + 1) * u * M_PI / 16) * cos((2 * x + 1) * v * M_PI / 16);
      if (b_bf16) {
        __nv_bfloat16 h = __float2bfloat16((float)val);
        B[n * K + k] = *reinterpret_cast<uint16_t*>(&h);
        Bf[n * K + k] = __bfloat162float(h);
      } else {
        __half h = __float2half((float)val);
        B[n * K + k] = *reinterpret_cast<uint16_t*>(&h);
        Bf[n * K + k] = __half2float(h);
      }
    }
  uint16_t *dA, *dB;
  float* dD;
  long long* dC;
  CK(cudaMalloc(&dA, A.size() * 2));
  CK(cudaMalloc(&dB, B.size() * 2));
  CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMalloc(&dC, 16));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, M * N * 4));
  const int smem_bytes = 16384 + 8192 + 64;
  CK(cudaFuncSetAttribute(bringup, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  bringup<<<1, 128, smem_bytes>>>(dA, dB, dD, variant, b_bf16, dC);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D(M * N);
  long long cyc[2];
  CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cyc, dC, 16, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)Af[m * K + k] * Bf[n * K + k];
      const double e = fabs(ref - (variant == 3 ? 16777216.0 * D[m * N + n] : D[m * N + n]));
      if (!(e < 1e-2)) ++bad;
      if (e > maxerr || e != e) maxerr = e;
      if (fabs(ref) > maxref) maxref = fabs(ref);
    }
  printf("variant %d b_bf16 %d: max |D - ref| = %.3e (max |ref| %.1f), entries off by > 1e-2: %d of %d, issue loop %lld cycles, D[0][0..3] = %g %g %g %g\n",
         variant, b_bf16, maxerr, maxref, bad, M * N, cyc[0], D[0], D[1], D[2], D[3]);
  return bad ? 1 : 0;
}
