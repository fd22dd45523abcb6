// Fast path for B = 32 (BASELINE configs[2]), the block size without a dedicated kernel family that needs one:
// measured at B = 4 the general kernels are the faster ones (252 / 148 against 145 / 120 Gpixel/s), so B = 4 stays there.
//
// The general kernels (kernels_general.cu) take every flag, shape and alignment, and pay for it at
// B = 32: a 32 x 256 tile with a three-channel intermediate fills the shared memory of an SM, so one
// CTA per SM loads, transforms and stores with nothing to overlap with (ncu: issue slots 30-36 % busy,
// half of the shared-memory wavefronts bank conflicts, long-scoreboard stalls on the loads).  The
// kernels here are the two halves of the fused rate/distortion kernel (kernels_rd.cu), whose layout was
// tuned on the profiler: one block row x 128 pixels (64 for the float64 decoder), one work item per
// thread in every phase, pitches of TW + 1 elements with row-major lanes (bank-conflict free), three
// CTAs per SM so that one CTA's loads and stores hide behind the others' arithmetic.
//
//   encode (float32, pocketfft-exact: the reference's own precision, src/2D-DCT.py:276-361)
//     load     B rows x TW pixels of RGB, zero above / below the frame (vertical padding, :216-227)
//     forward  (channel, pixel column): colour, DCT down the column -> F;  (channel, block, row u): DCT along
//              the row, quantise, +128, wrap -> stage[(u, i) run][block][channel]
//     store    runs of TBX x 3 contiguous bytes of the subband layout, 2 bytes per store
//   decode (float64 exact = the reference's chain :398-466, or float32 with fused multiply-adds = fast mode)
//     gather   the runs -> stage;  (channel, block, row u): dequantise -> G
//     inverse  columns, then rows, in place in G;  pixels: to_RGB, +128, truncate, clip -> RGB tile (+ SSE)
//     store    B rows x TW x 3 bytes, 4 bytes per store
// Preconditions (else VCFB_E_UNSUPP -> general kernels): subband layout, no perceptual weights, W a multiple
// of 128, an even number of blocks per row, 4-byte aligned pointers; encode float32 exact only.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "dct_codelets.cuh"

namespace vcfb {
namespace {

__host__ __device__ constexpr double p2(int e) {
  double r = 1.0;
  for (int i = 0; i < (e < 0 ? -e : e); ++i) r = e < 0 ? r * 0.5 : r * 2.0;
  return r;
}

template <int B, typename T> struct TileL {
  static constexpr int TW = (sizeof(T) == 8 && B >= 32) ? 64 : 128;
  static constexpr int NT = 3 * TW;
  static constexpr int TBX = TW / B;
  static constexpr int RUNB = TBX * 3;                               // bytes of one (u, i) run
  static constexpr int RPITCH = ((RUNB / 4) % 2 == 0 && RUNB >= 8) ? RUNB + 4 : RUNB;   // odd number of words where it matters
  static constexpr int ROWB = TW * 3;                                // bytes of one pixel row of the tile
  static constexpr int GP = TW + 1;
  static constexpr int PIX_BYTES = B * ROWB;                         // RGB tile (input of the encoder, output of the decoder)
  static constexpr int G_BYTES = ((3 * B * GP * int(sizeof(T)) + 15) / 16) * 16;
  static constexpr int STAGE_BYTES = ((B * B * RPITCH + 15) / 16) * 16;
  static constexpr int SMEM = PIX_BYTES + G_BYTES + STAGE_BYTES + 64;
};

__device__ __forceinline__ unsigned warp_sum(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }

template <typename T> __device__ __forceinline__ int to_int_rz(T x);
template <> __device__ __forceinline__ int to_int_rz<float>(float x) { return __float2int_rz(x); }
template <> __device__ __forceinline__ int to_int_rz<double>(double x) { return __double2int_rz(x); }

// ============================================================================================
// encode
// ============================================================================================
template <int B, bool STATS>
__global__ void __launch_bounds__(TileL<B, float>::NT, 3) enc_tile_kernel(const EncArgs a) {
  using L = TileL<B, float>;
  using OF = Ops<float, true>;
  using DF = Dct<B, false>;
  using MF = typename DF::meta;
  constexpr int TW = L::TW, NT = L::NT, ROWB = L::ROWB, GP = L::GP, TBX = L::TBX, RP = L::RPITCH, RUNB = L::RUNB;

  extern __shared__ __align__(16) unsigned char smem[];
  uint8_t* raw = smem;
  float* F = reinterpret_cast<float*>(smem + L::PIX_BYTES);
  uint8_t* stage = smem + L::PIX_BYTES + L::G_BYTES;

  const int tid = threadIdx.x;
  const int tile = blockIdx.x, by = blockIdx.y, f = blockIdx.z;
  const Geom g = a.g;
  const int x0 = tile * TW;
  const int bx0 = tile * TBX;

  // ---- load: B rows x TW pixels, zero outside the frame; rows are 4-byte aligned (W % 128 == 0) ----
  {
    constexpr int WPR = ROWB / 4;
    const uint32_t* base = reinterpret_cast<const uint32_t*>(a.rgb + (size_t(f) * g.H * g.W + x0) * 3);
    for (int t = tid; t < B * WPR; t += NT) {
      const int r = t / WPR, wd = t - r * WPR;
      const int gy = by * B + r - g.top;
      uint32_t v = 0;
      if (gy >= 0 && gy < g.H) v = __ldg(base + (size_t(gy) * g.W * 3) / 4 + wd);
      reinterpret_cast<uint32_t*>(raw + r * ROWB)[wd] = v;
    }
  }
  __syncthreads();

  const int c = tid / TW;
  const int t = tid - c * TW;
  const int u = t % B, bx = t / B;

  // ---- forward 1: colour + DCT down each pixel column (axis 0) ----
  {
    const int x = t;
    const float cs = a.color == VCFB_COLOR_YCOCG ? (c == 1 ? 0.5f : 0.25f) : 1.0f;
    float v[B];
#pragma unroll
    for (int r = 0; r < B; ++r) {
      const uint8_t* px = raw + r * ROWB + x * 3;
      const int R = px[0], Gc = px[1], Bc = px[2];
      if (a.color == VCFB_COLOR_YCOCG) {
        v[r] = float((c == 0) ? (R + 2 * Gc + Bc - 512) : (c == 1) ? (R - Bc) : (2 * Gc - R - Bc));
      } else {
        const float r_ = float(R - 128), g_ = float(Gc - 128), b_ = float(Bc - 128);
        const float y = OF::add(OF::add(OF::mul(r_, 0.299f), OF::mul(g_, 0.587f)), OF::mul(b_, 0.114f));
        v[r] = c == 0 ? y : c == 1 ? OF::mul(OF::sub(r_, y), 0.713f) : OF::mul(OF::sub(b_, y), 0.564f);
      }
    }
    DF::template run<float, true>(v);
#pragma unroll
    for (int k = 0; k < B; ++k) F[(c * B + k) * GP + x] = OF::mul(v[k], float(MF::sgn(k) * p2(MF::exp(k))) * cs);
  }
  __syncthreads();

  // ---- forward 2: DCT along each block row (axis 1), quantise (src/deadzone.py:98), +128, wrap (:348,:361) ----
  unsigned nz = 0, sabs = 0;
  {
    float v[B];
    const float* src = F + (c * B + u) * GP + bx * B;
#pragma unroll
    for (int i = 0; i < B; ++i) v[i] = src[i];
    DF::template run<float, true>(v);
    const float q = float(a.q), inv_q = float(a.inv_q);
    uint8_t* dst = stage + u * RP + bx * 3 + c;
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const float coef = OF::mul(v[i], float(MF::sgn(i) * p2(MF::exp(i))));     // the coefficient scipy returns
      const float tq = a.q_pow2 ? OF::mul(coef, inv_q) : OF::div(coef, q);
      const int k = __float2int_rz(tq);
      const unsigned byte = unsigned(k + 128) & 255u;
      dst[i * B * RP] = uint8_t(byte);                                          // stage[(i, u) run][block][channel]
      if (STATS) {
        const int k8 = int(byte) - 128;
        nz += (k8 != 0);
        sabs += unsigned(abs(k8));
      }
    }
  }
  __syncthreads();

  // ---- store the runs: sub[u * ny + by, i * nx + bx0 ..][channel], 2 bytes per store ----
  {
    constexpr int HPR = RUNB / 2;                                  // half-words per run
    for (int w = tid; w < B * B * HPR; w += NT) {
      const int run = w / HPR, hw = w - run * HPR;
      const int i = run / B, uu = run - i * B;                     // stage order: i major, u minor
      const size_t row = size_t(uu) * g.ny + by, col = size_t(i) * g.nx + bx0;
      uint16_t* gp = reinterpret_cast<uint16_t*>(a.idx + ((size_t(f) * g.Hp + row) * g.Wp + col) * 3);
      gp[hw] = reinterpret_cast<const uint16_t*>(stage + run * RP)[hw];
    }
  }
  if (STATS) {
    const unsigned packed = warp_sum((nz << 20) | sabs);          // nz <= 32, sabs <= 128 * 32 per lane
    if ((tid & 31) == 0) {
      if (packed >> 20) atomicAdd(a.stats + VCFB_STAT_NONZERO, (unsigned long long)(packed >> 20));
      if (packed & 0xFFFFFu) atomicAdd(a.stats + VCFB_STAT_SUMABS, (unsigned long long)(packed & 0xFFFFFu));
    }
    if (tid == 0) atomicAdd(a.stats + VCFB_STAT_NINDICES, (unsigned long long)(TW * B * 3));
  }
}

// ============================================================================================
// decode
// ============================================================================================
template <typename T, bool EXACT, int B, bool SSE>
__global__ void __launch_bounds__(TileL<B, T>::NT, 3) dec_tile_kernel(const DecArgs a) {
  using L = TileL<B, T>;
  using O = Ops<T, EXACT>;
  using DI = Dct<B, true>;
  using MI = typename DI::meta;
  constexpr int TW = L::TW, NT = L::NT, ROWB = L::ROWB, GP = L::GP, TBX = L::TBX, RP = L::RPITCH, RUNB = L::RUNB;

  extern __shared__ __align__(16) unsigned char smem[];
  uint8_t* outb = smem;
  T* G = reinterpret_cast<T*>(smem + L::PIX_BYTES);
  uint8_t* stage = smem + L::PIX_BYTES + L::G_BYTES;

  const int tid = threadIdx.x;
  const int tile = blockIdx.x, by = blockIdx.y, f = blockIdx.z;
  const Geom g = a.g;
  const int x0 = tile * TW;
  const int bx0 = tile * TBX;

  // ---- gather the runs of this tile ----
  {
    constexpr int HPR = RUNB / 2;
    for (int w = tid; w < B * B * HPR; w += NT) {
      const int run = w / HPR, hw = w - run * HPR;
      const int i = run / B, uu = run - i * B;
      const size_t row = size_t(uu) * g.ny + by, col = size_t(i) * g.nx + bx0;
      const uint16_t* gp = reinterpret_cast<const uint16_t*>(a.idx + ((size_t(f) * g.Hp + row) * g.Wp + col) * 3);
      reinterpret_cast<uint16_t*>(stage + run * RP)[hw] = __ldg(gp + hw);
    }
  }
  __syncthreads();

  const int c = tid / TW;
  const int t = tid - c * TW;
  const int u = t % B, bx = t / B;

  // ---- dequantise (:398-410): int16 * python int stays int16 (wraps); a non-integral step promotes to float64 ----
  {
    const uint8_t* src = stage + u * RP + bx * 3 + c;
    T* dst = G + (c * B + u) * GP + bx * B;
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const int k = int(src[i * B * RP]) - 128;
      dst[i] = a.q_int ? T(int(short(k * a.q_int))) : T(double(k) * a.q);
    }
  }
  __syncthreads();

  // ---- inverse DCT along axis 0 (coefficient columns), in place ----
  {
    T v[B];
    T* col = G + (c * B) * GP + t;
#pragma unroll
    for (int k = 0; k < B; ++k) v[k] = col[k * GP];
    DI::template run<T, EXACT>(v);
#pragma unroll
    for (int r = 0; r < B; ++r) col[r * GP] = O::mul(v[r], T(MI::sgn(r) * p2(MI::exp(r))));
  }
  __syncthreads();

  // ---- inverse DCT along axis 1 (pixel rows of each block), in place ----
  {
    T v[B];
    T* row = G + (c * B + u) * GP + bx * B;
#pragma unroll
    for (int i = 0; i < B; ++i) v[i] = row[i];
    DI::template run<T, EXACT>(v);
#pragma unroll
    for (int i = 0; i < B; ++i) row[i] = O::mul(v[i], T(MI::sgn(i) * p2(MI::exp(i))));
  }
  __syncthreads();

  // ---- pixels: to_RGB, +128, truncate, clip (:449-466); thread = pixel column t, rows c, c + 3, ... ----
  unsigned sse[3] = {0, 0, 0};
  int sdiff = 0;
  {
    const int x = t;
#pragma unroll
    for (int j = 0; j < (B + 2) / 3; ++j) {
      const int r = c + 3 * j;
      const int gy = by * B + r - g.top;
      if (r >= B) continue;
      const T c0 = G[(0 * B + r) * GP + x], c1 = G[(1 * B + r) * GP + x], c2 = G[(2 * B + r) * GP + x];
      T R, Gv, Bv;
      if (a.color == VCFB_COLOR_YCOCG) {   // Y + Co - Cg ; Y + Cg ; Y - Co - Cg, left to right
        R = O::sub(O::add(c0, c1), c2);
        Gv = O::add(c0, c2);
        Bv = O::sub(O::sub(c0, c1), c2);
      } else {                             // oracle ycrcb_to_rgb_float
        R = O::add(c0, O::mul(c1, T(1.403)));
        Gv = O::add(O::add(c0, O::mul(c1, T(-0.714))), O::mul(c2, T(-0.344)));
        Bv = O::add(c0, O::mul(c2, T(1.773)));
      }
      const int v[3] = {min(max(to_int_rz<T>(O::add(R, T(128))), 0), 255), min(max(to_int_rz<T>(O::add(Gv, T(128))), 0), 255),
                        min(max(to_int_rz<T>(O::add(Bv, T(128))), 0), 255)};
      uint8_t* o = outb + r * ROWB + x * 3;
      o[0] = uint8_t(v[0]); o[1] = uint8_t(v[1]); o[2] = uint8_t(v[2]);
      if (SSE && gy >= 0 && gy < g.H) {
        const uint8_t* px = a.original + ((size_t(f) * g.H + gy) * g.W + x0 + x) * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int d = int(__ldg(px + k)) - v[k];
          sse[k] += unsigned(d * d);
          sdiff += d;
        }
      }
    }
  }
  __syncthreads();

  // ---- store the rows inside the frame (crop, :444), 4 bytes per store ----
  if (a.rgb) {
    constexpr int WPR = ROWB / 4;
    uint32_t* base = reinterpret_cast<uint32_t*>(a.rgb + (size_t(f) * g.H * g.W + x0) * 3);
    for (int w = tid; w < B * WPR; w += NT) {
      const int r = w / WPR, wd = w - r * WPR;
      const int gy = by * B + r - g.top;
      if (gy >= 0 && gy < g.H) base[(size_t(gy) * g.W * 3) / 4 + wd] = reinterpret_cast<const uint32_t*>(outb + r * ROWB)[wd];
    }
  }
  if (SSE) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const unsigned sv = warp_sum(sse[k]);
      if ((tid & 31) == 0 && sv) atomicAdd(a.stats + VCFB_STAT_SSE_R + k, (unsigned long long)sv);
    }
    const unsigned sd = warp_sum(unsigned(sdiff));             // two's complement sum
    if ((tid & 31) == 0 && sd) atomicAdd(a.stats + VCFB_STAT_SUMDIFF, (unsigned long long)(long long)(int)sd);
    if (tid == 0) {
      int rows = 0;
      for (int r = 0; r < B; ++r) {
        const int gy = by * B + r - g.top;
        rows += (gy >= 0 && gy < g.H);
      }
      if (rows) atomicAdd(a.stats + VCFB_STAT_NSAMPLES, (unsigned long long)rows * TW * 3);
    }
  }
}

// ---- host ---------------------------------------------------------------------------------------

template <int B, typename T>
bool shape_ok(const Geom& g) {
  using L = TileL<B, T>;
  return g.W % 128 == 0 && g.left == 0 && g.Wp == g.W && g.nx % 2 == 0 && g.W % L::TW == 0;
}

template <int B>
int run_enc(const EncArgs& a, cudaStream_t s, const char* name) {
  using L = TileL<B, float>;
  if (!shape_ok<B, float>(a.g)) return VCFB_E_UNSUPP;
  void (*kern)(const EncArgs) = a.stats ? enc_tile_kernel<B, true> : enc_tile_kernel<B, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(enc_tile)");
  dim3 grid(a.g.W / L::TW, a.g.ny, a.n_frames);
  note_kernel(name);
  kern<<<grid, L::NT, L::SMEM, s>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "enc_tile_kernel launch");
  return VCFB_OK;
}

template <typename T, bool EXACT, int B>
int run_dec(const DecArgs& a, cudaStream_t s, const char* name) {
  using L = TileL<B, T>;
  if (!shape_ok<B, T>(a.g)) return VCFB_E_UNSUPP;
  void (*kern)(const DecArgs) = a.stats ? dec_tile_kernel<T, EXACT, B, true> : dec_tile_kernel<T, EXACT, B, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dec_tile)");
  dim3 grid(a.g.W / L::TW, a.g.ny, a.n_frames);
  note_kernel(name);
  kern<<<grid, L::NT, L::SMEM, s>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "dec_tile_kernel launch");
  return VCFB_OK;
}

}  // namespace

// VCFB_E_UNSUPP = outside this fast path (the caller falls back to the general kernels)
int launch_encode_tile(const EncArgs& a, int B, cudaStream_t s) {
  if (B != 32) return VCFB_E_UNSUPP;
  if (a.flags & (VCFB_F_NO_SUBBANDS | VCFB_F_PERCEPTUAL | VCFB_F_FP64 | VCFB_F_CONTRACT)) return VCFB_E_UNSUPP;
  if (getenv("VCFB_NO_TILE")) return VCFB_E_UNSUPP;            // development knob
  if ((reinterpret_cast<uintptr_t>(a.rgb) & 3) || (reinterpret_cast<uintptr_t>(a.idx) & 3)) return VCFB_E_UNSUPP;
  if (a.stats && (a.flags & VCFB_F_HIST)) {
    // the histogram takes the streaming pass over the stored indices (kernels_stats.cu)
    if (reinterpret_cast<uintptr_t>(a.idx) & 15) return VCFB_E_UNSUPP;
    EncArgs b = a;
    b.stats = nullptr;
    int rc = launch_encode_tile(b, B, s);
    if (rc) return rc;
    return launch_index_stats(a.idx, (long long)a.n_frames * a.g.Hp * a.g.Wp * 3, true, a.stats, s);
  }
  return run_enc<32>(a, s, "enc32_tile");
}

int launch_decode_tile(const DecArgs& a, int B, cudaStream_t s) {
  if (B != 32) return VCFB_E_UNSUPP;
  if (a.flags & (VCFB_F_NO_SUBBANDS | VCFB_F_PERCEPTUAL | VCFB_F_CONTRACT | VCFB_F_SYNTH_F32)) return VCFB_E_UNSUPP;
  if (getenv("VCFB_NO_TILE")) return VCFB_E_UNSUPP;
  if (a.y_out || (!a.rgb && !(a.stats && a.original))) return VCFB_E_UNSUPP;
  if ((a.stats != nullptr) != (a.original != nullptr)) return VCFB_E_UNSUPP;
  if ((reinterpret_cast<uintptr_t>(a.rgb) & 3) || (reinterpret_cast<uintptr_t>(a.idx) & 3)) return VCFB_E_UNSUPP;
  if (a.flags & VCFB_F_FP64) return run_dec<double, true, 32>(a, s, "dec32_tile");
  return run_dec<float, false, 32>(a, s, "dec32_tile_f32");
}

}  // namespace vcfb
