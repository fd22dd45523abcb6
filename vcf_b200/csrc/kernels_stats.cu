// Statistics as stand-alone streaming kernels, used behind the fast-path kernels (which do
// not accumulate statistics themselves): one pass over the index array for the rate side,
// one pass over (original, decoded) for the distortion side.  Both are HBM-bound reads with
// 128-bit loads; sums are integer, accumulated per warp, then one atomic per CTA.
//
//   index statistics  VCFB_STAT_NONZERO, _SUMABS, _NINDICES (+ _HIST with VCFB_F_HIST):
//                     what src/IPP_DCT.py:273-292 (get_rate) estimates bits from, and the
//                     zero-order histogram
//   SSE               VCFB_STAT_SSE_R/G/B, _SUMDIFF, _NSAMPLES: src/RDE.py:41-49
#include "common.cuh"

namespace vcfb {
namespace {

__device__ __forceinline__ unsigned warp_sum_u(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// bytes: n_vec 16-byte vectors (the caller guarantees alignment and n_bytes % 16 == 0).
// Channel of byte j of vector v is (v + j) % 3 because 16 = 1 (mod 3).
template <bool HIST>
__global__ void __launch_bounds__(256) index_stats_kernel(const uint4* __restrict__ idx, long long n_vec,
                                                          unsigned long long* __restrict__ stats) {
  __shared__ unsigned sh[3 * 256 + 2];
  if (HIST)
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) sh[i] = 0;
  if (threadIdx.x < 2) sh[768 + threadIdx.x] = 0;
  __syncthreads();
  unsigned nz = 0, sabs = 0;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n_vec;
       v += (long long)gridDim.x * blockDim.x) {
    const uint4 q = __ldg(idx + v);
    const unsigned w[4] = {q.x, q.y, q.z, q.w};
    const int c0 = int(v % 3);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned eq = __vcmpeq4(w[k], 0x80808080u);          // 0xff where the index is zero
      nz += 4 - (__popc(eq) >> 3);
      sabs = __dp4a(__vabsdiffu4(w[k], 0x80808080u), 0x01010101u, sabs);
      if (HIST) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const unsigned b = (w[k] >> (8 * j)) & 255u;
          const int c = (c0 + 4 * k + j) % 3;
          atomicAdd(&sh[c * 256 + b], 1u);
        }
      }
    }
  }
  nz = warp_sum_u(nz);
  sabs = warp_sum_u(sabs);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&sh[768], nz);
    atomicAdd(&sh[769], sabs);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(stats + VCFB_STAT_NONZERO, (unsigned long long)sh[768]);
    atomicAdd(stats + VCFB_STAT_SUMABS, (unsigned long long)sh[769]);
  }
  if (HIST)
    for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x)
      if (sh[i]) atomicAdd(stats + VCFB_STAT_HIST + i, (unsigned long long)sh[i]);
}

// |a-b| per byte with one SIMD instruction, squares summed per channel with dp4a on the
// difference masked to one channel.  The grid stride is a multiple of 3 vectors, so the
// channel phase (v mod 3) of a thread never changes and the byte -> channel map of each of
// the four words of a vector is a compile-time pattern relative to it.
__global__ void __launch_bounds__(256) sse_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                                  long long n_vec, unsigned long long* __restrict__ stats) {
  __shared__ unsigned long long sh[4];
  if (threadIdx.x < 4) sh[threadIdx.x] = 0;
  __syncthreads();
  // phase p of a byte = (4k + j) mod 3 for byte j of word k; masks select the bytes of one phase
  constexpr unsigned M[3][3] = {{0xFF0000FFu, 0x0000FF00u, 0x00FF0000u},    // k % 3 == 0: j -> (j) % 3
                                {0x00FF0000u, 0xFF0000FFu, 0x0000FF00u},    // k % 3 == 1: j -> (j + 1) % 3
                                {0x0000FF00u, 0x00FF0000u, 0xFF0000FFu}};   // k % 3 == 2: j -> (j + 2) % 3
  const long long stride = (long long)gridDim.x * blockDim.x;                 // multiple of 3 (see launcher)
  const long long v0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  unsigned ph[3] = {0, 0, 0};
  unsigned long long tot[3] = {0, 0, 0};
  int sx = 0, sy = 0;
  long long sdiff = 0;
  int it = 0;
  for (long long v = v0; v < n_vec; v += stride) {
    const uint4 x = __ldg(a + v), y = __ldg(b + v);
    const unsigned xa[4] = {x.x, x.y, x.z, x.w}, ya[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned d = __vabsdiffu4(xa[k], ya[k]);
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        const unsigned dm = d & M[(4 * k) % 3][p];
        ph[p] = __dp4a(dm, dm, ph[p]);
      }
      sx = __dp4a(xa[k], 0x01010101u, unsigned(sx));
      sy = __dp4a(ya[k], 0x01010101u, unsigned(sy));
    }
    if (++it == 4096) {            // 4096 * 16 * 65025 < 2^32; 4096 * 16 * 255 < 2^31
#pragma unroll
      for (int p = 0; p < 3; ++p) { tot[p] += ph[p]; ph[p] = 0; }
      sdiff += (long long)sx - sy;
      sx = sy = 0;
      it = 0;
    }
  }
#pragma unroll
  for (int p = 0; p < 3; ++p) tot[p] += ph[p];
  sdiff += (long long)sx - sy;
  const int c0 = int(v0 % 3);
  // per-lane accumulation into the three channel slots (lanes of a warp have different phases)
#pragma unroll
  for (int p = 0; p < 3; ++p)
    if (tot[p]) atomicAdd(&sh[(c0 + p) % 3], tot[p]);
  if (sdiff) atomicAdd(&sh[3], (unsigned long long)sdiff);
  __syncthreads();
  if (threadIdx.x < 3 && sh[threadIdx.x]) atomicAdd(stats + VCFB_STAT_SSE_R + threadIdx.x, sh[threadIdx.x]);
  if (threadIdx.x == 3 && sh[3]) atomicAdd(stats + VCFB_STAT_SUMDIFF, sh[3]);
}

int stats_grid(long long n_vec) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = (n_vec + 255) / 256;
  const long long cap = (long long)sms * 8;
  return int(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

__global__ void add_counts_kernel(unsigned long long* stats, int slot, unsigned long long n) {
  atomicAdd(stats + slot, n);
}

}  // namespace

int launch_add_count(unsigned long long* stats, int slot, unsigned long long n, cudaStream_t s) {
  add_counts_kernel<<<1, 1, 0, s>>>(stats, slot, n);
  note_extra_launches(1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "add_counts_kernel launch");
  return VCFB_OK;
}

// idx: n_bytes uint8 indices, 16-byte aligned, n_bytes % 16 == 0 (true for whole fast-path frames)
int launch_index_stats(const uint8_t* idx, long long n_bytes, bool hist, unsigned long long* stats, cudaStream_t s) {
  if ((reinterpret_cast<uintptr_t>(idx) & 15) || (n_bytes & 15)) return VCFB_E_UNSUPP;
  const long long n_vec = n_bytes / 16;
  if (hist) index_stats_kernel<true><<<stats_grid(n_vec), 256, 0, s>>>(reinterpret_cast<const uint4*>(idx), n_vec, stats);
  else index_stats_kernel<false><<<stats_grid(n_vec), 256, 0, s>>>(reinterpret_cast<const uint4*>(idx), n_vec, stats);
  add_counts_kernel<<<1, 1, 0, s>>>(stats, VCFB_STAT_NINDICES, (unsigned long long)n_bytes);
  note_extra_launches(2);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? VCFB_OK : cuda_fail(e, "index_stats_kernel launch");
}

int launch_sse(const uint8_t* a, const uint8_t* b, long long n_bytes, unsigned long long* stats, cudaStream_t s) {
  if (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) || (n_bytes & 15)) return VCFB_E_UNSUPP;
  const long long n_vec = n_bytes / 16;
  int grid = stats_grid(n_vec);
  grid = (grid + 2) / 3 * 3;      // grid * 256 threads must be a multiple of 3 vectors
  sse_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), n_vec,
                                              stats);
  add_counts_kernel<<<1, 1, 0, s>>>(stats, VCFB_STAT_NSAMPLES, (unsigned long long)n_bytes);
  note_extra_launches(2);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? VCFB_OK : cuda_fail(e, "sse_kernel launch");
}

}  // namespace vcfb
