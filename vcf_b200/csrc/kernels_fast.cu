// Fast path: B = 8, float32, YCoCg, subband layout -- the configuration BASELINE.json
// quotes its metric on.  (Every other configuration runs the general kernels.)
//
// Design (DESIGN.md "fast path"):
//   * persistent grid; every WARP is an autonomous pipeline over tiles of 8 rows x 128
//     pixels (16 blocks): its own 3-deep ring of TMA loads and its own mbarriers, no
//     CTA-wide barrier anywhere (only __syncwarp), so global memory is never
//     touched by a thread: cp.async.bulk.tensor in, cp.async.bulk.tensor out;
//   * zero padding of src/2D-DCT.py:216-227 is TMA's out-of-bounds fill (signed row
//     coordinate = block row * 8 - top);
//   * pass 1: a thread owns 4 adjacent pixel columns x 8 rows x 3 channels.  The
//     colour transform is 3 dp4a per pixel on the packed bytes (exact integers
//     4Y, 2Co, 4Cg of the centred pixel; the 1/4, 1/2 are lazy exponents), then 12
//     pocketfft-exact length-8 DCTs in registers, written as float4 to an
//     XOR-swizzled intermediate in shared memory;
//   * pass 2: a thread owns one coefficient row u of 4 adjacent blocks x 3 channels:
//     12 DCTs, one multiply per coefficient by a per-thread constant that folds all
//     lazy power-of-two scales and 1/q (exact for power-of-two q; true division
//     otherwise), truncation, bias + wrap, and 12 bytes per (u,i) go out as three
//     32-bit words into the dense TMA store box [j][i][block*3+c];
//   * every shared-memory offset is a compile-time immediate on a per-thread base.
#include <mutex>
#include <type_traits>

#include "fast_common.cuh"
#include "dec8_dc.cuh"

namespace vcfb {

namespace tma {

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

bool make_map(CUtensorMap* out, CUtensorMapDataType dt, int rank, void* base, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i) gs[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(out, dt, cuuint32_t(rank), base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace tma

using namespace fast;

namespace {

// ============================================================================
// encode: one warp = one autonomous pipeline over tiles of 8 rows x 128 pixels
// ============================================================================
template <bool EXACT, bool QPOW2, int NWARPS, int CTAS>
__global__ void __launch_bounds__(NWARPS * 32, CTAS)
enc8_fast_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                 const FastArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* ring = smem + warp * ENC_WARP_SMEM;
  float* F = reinterpret_cast<float*>(ring + NSTAGE * TILE);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + NSTAGE * TILE + ENC_F_BYTES);

  if (lane == 0) {
    tma::prefetch_map(&in_map);
    tma::prefetch_map(&out_map);
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) tma::mbar_init(&full[s], 1);
    tma::fence_mbar_init();
  }
  __syncwarp();

  Walker w;
  w.tile = blockIdx.x * NWARPS + warp;
  w.stride = gridDim.x * NWARPS;
  w.ntiles = a.ntiles;
  w.tiles_x = a.tiles_x;
  w.per_frame = a.ny * a.tiles_x;
  w.top = a.top;

  auto issue_load = [&](int s, int t) {
    int f, by, tx;
    w.coords(t, f, by, tx);
    tma::mbar_expect_tx(&full[s], TILE);
    tma::load_3d(ring + s * TILE, &in_map, &full[s], tx * (WT * 3 / 8), by * 8 - w.top, f);
  };
  auto issue_store = [&](int s, int t) {
    int f, by, tx;
    w.coords(t, f, by, tx);
    // out_map dims: (x bytes, j, block row, i, frame); smem box is [i][j][48 B]
    tma::store_5d(&out_map, ring + s * TILE, tx * (WT / 8) * 3, 0, by, 0, f);
    tma::commit_group();
  };

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      const int t = w.tile + s * w.stride;
      if (t < w.ntiles) issue_load(s, t);
    }
  }

  // pass 1: lane = group of 4 pixel columns.  pass 2: lane = (coefficient row u, group of 4 blocks G)
  const int u = lane & 7, G = lane >> 3;
  float qs[3][2];   // [c][exp_i - min_exp]
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    qs[c][0] = a.qtab[u][c];
    qs[c][1] = a.qtab[u][c] * 2.0f;
  }
  const float qf = a.q;

  int k = 0;
  for (int tile = w.tile; tile < w.ntiles; tile += w.stride, ++k) {
    const int s = k % NSTAGE;
    unsigned char* tb = ring + s * TILE;
    tma::mbar_wait(&full[s], (k / NSTAGE) & 1);

    // ---- pass 1: colour transform + DCT down the columns (axis 0) -------------------
    {
      float v[3][4][8];
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(tb) + 3 * lane;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const uint32_t w0 = rw[r * ROWW + 0];
        const uint32_t w1 = rw[r * ROWW + 1];
        const uint32_t w2 = rw[r * ROWW + 2];
        const uint32_t p1 = __byte_perm(w0, w1, 0x0543);   // R1 G1 B1 .
        const uint32_t p2v = __byte_perm(w1, w2, 0x0432);  // R2 G2 B2 .
        // 4*Y = R + 2G + B - 512 ; 2*Co = R - B ; 4*Cg = -R + 2G - B   (centred pixel)
        v[0][0][r] = dotf(w0, 0x00010201, -512);
        v[1][0][r] = dotf(w0, 0x00FF0001, 0);
        v[2][0][r] = dotf(w0, 0x00FF02FF, 0);
        v[0][1][r] = dotf(p1, 0x00010201, -512);
        v[1][1][r] = dotf(p1, 0x00FF0001, 0);
        v[2][1][r] = dotf(p1, 0x00FF02FF, 0);
        v[0][2][r] = dotf(p2v, 0x00010201, -512);
        v[1][2][r] = dotf(p2v, 0x00FF0001, 0);
        v[2][2][r] = dotf(p2v, 0x00FF02FF, 0);
        v[0][3][r] = dotf(w2, 0x01020100, -512);            // . R3 G3 B3
        v[1][3][r] = dotf(w2, int(0xFF000100), 0);
        v[2][3][r] = dotf(w2, int(0xFF02FF00), 0);
      }
      float* fw = F + 4 * lane;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int col = 0; col < 4; ++col) dct8_fwd<float, EXACT>(v[c][col]);
#pragma unroll
        for (int uu = 0; uu < 8; ++uu)
          *reinterpret_cast<float4*>(fw + (c * 8 + uu) * ENC_FP) =
              make_float4(v[c][0][uu], v[c][1][uu], v[c][2][uu], v[c][3][uu]);
      }
    }
    __syncwarp();

    // ---- pass 2: DCT along the rows (axis 1), quantise, pack -------------------------
    {
      float v[3][4][8];
      const float* fr = F + u * ENC_FP + 32 * G;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const float4 t4 = *reinterpret_cast<const float4*>(fr + c * 8 * ENC_FP + 4 * m);
          v[c][m >> 1][(m & 1) * 4 + 0] = t4.x;
          v[c][m >> 1][(m & 1) * 4 + 1] = t4.y;
          v[c][m >> 1][(m & 1) * 4 + 2] = t4.z;
          v[c][m >> 1][(m & 1) * 4 + 3] = t4.w;
        }
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int b = 0; b < 4; ++b) dct8_fwd<float, EXACT>(v[c][b]);
      // index box in smem: [i][j = u][16 blocks * 3 bytes]
      uint32_t* ow = reinterpret_cast<uint32_t*>(tb) + u * 12 + 3 * G;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int kk[4][3];
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float sc = M8F::sgn(i) > 0 ? qs[c][M8F::exp(i) - min_exp8()] : -qs[c][M8F::exp(i) - min_exp8()];
            float t = __fmul_rn(v[c][b][i], sc);              // exact: sc is a power of two
            if (!QPOW2) t = __fdiv_rn(t, qf);                 // src/deadzone.py:98  x / Q_step
            kk[b][c] = __float2int_rz(t);                     // truncation = dead zone
          }
        uint32_t* o = ow + i * 96;
        o[0] = pack4(kk[0][0], kk[0][1], kk[0][2], kk[1][0]) ^ 0x80808080u;   // +128, wraps
        o[1] = pack4(kk[1][1], kk[1][2], kk[2][0], kk[2][1]) ^ 0x80808080u;
        o[2] = pack4(kk[2][2], kk[3][0], kk[3][1], kk[3][2]) ^ 0x80808080u;
      }
    }
    tma::fence_proxy_async();
    __syncwarp();

    if (lane == 0) {
      issue_store(s, tile);
      tma::wait_group_read<1>();          // the store issued one tile ago has drained its buffer
      if (k >= 1) {
        const int nt = tile + (NSTAGE - 1) * w.stride;
        if (nt < w.ntiles) issue_load((k - 1) % NSTAGE, nt);
      }
    }
    __syncwarp();
  }
  if (lane == 0) tma::wait_group<0>();
}

// ============================================================================
// decode: same pipeline, inverse arithmetic.
//
//   index box [j][i][48 B]  ->  pass 1: lane = (coefficient column i, group of 4
//   blocks G); per channel: dequantise 32 indices, 4 inverse DCTs over u (axis 0),
//   X[c][y][i][block] written as 4 contiguous elements  ->  pass 2: lane = (pixel row
//   y, G); inverse DCTs over i (axis 1), to_RGB, one fma for (lazy 2^-4, +128), clip +
//   truncate, 96 bytes of RGB out  ->  TMA store of the 8 x 384-byte tile.
//
// T = double, EXACT = true is the reference's float64 chain operation for operation.
// The double kernel keeps its channel / block-pair loops rolled so that the code stays
// inside the instruction cache.
// ============================================================================

template <typename T> __device__ __forceinline__ T from_biased_int(unsigned biased);
template <> __device__ __forceinline__ double from_biased_int<double>(unsigned biased) {
  // 2^52 + 2^31 + k  minus  (2^52 + 2^31): both exact
  return __dsub_rn(__hiloint2double(0x43300000, int(biased)), 4503601774854144.0);
}
template <> __device__ __forceinline__ float from_biased_int<float>(unsigned biased) {
  return __int2float_rn(int(biased ^ 0x80000000u));
}

template <typename T> __device__ __forceinline__ int trunc_to_int(T x);
template <> __device__ __forceinline__ int trunc_to_int<float>(float x) { return __float2int_rz(x); }
template <> __device__ __forceinline__ int trunc_to_int<double>(double x) { return __double2int_rz(x); }


template <typename T, bool EXACT, int NWARPS, int CTAS>
__global__ void __launch_bounds__(NWARPS * 32, CTAS)
dec8_fast_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                 const FastDecArgs a) {
  using O = Ops<T, EXACT>;
  using L = DecL<T>;
  constexpr int P = L::P;
  constexpr bool F64 = sizeof(T) == 8;
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* ring = smem + warp * L::WARP_SMEM;
  T* F = reinterpret_cast<T*>(ring + NSTAGE * TILE);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + NSTAGE * TILE + L::F_BYTES);

  if (lane == 0) {
    tma::prefetch_map(&in_map);
    tma::prefetch_map(&out_map);
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) tma::mbar_init(&full[s], 1);
    tma::fence_mbar_init();
  }
  __syncwarp();

  Walker w;
  w.tile = blockIdx.x * NWARPS + warp;
  w.stride = gridDim.x * NWARPS;
  w.ntiles = a.ntiles;
  w.tiles_x = a.tiles_x;
  w.per_frame = a.ny * a.tiles_x;
  w.top = a.top;
  auto issue_load = [&](int s, int t) {
    int f, by, tx;
    w.coords(t, f, by, tx);
    tma::mbar_expect_tx(&full[s], TILE);
    // in_map dims: (x bytes, i, block row, j, frame); smem box is [j][i][48 B]
    tma::load_5d(ring + s * TILE, &in_map, &full[s], tx * (WT / 8) * 3, 0, by, 0, f);
  };
  auto issue_store = [&](int s, int t) {
    int f, by, tx;
    w.coords(t, f, by, tx);
    tma::store_3d(&out_map, ring + s * TILE, tx * (WT * 3 / 8), by * 8 - w.top, f);
    tma::commit_group();
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      const int t = w.tile + s * w.stride;
      if (t < w.ntiles) issue_load(s, t);
    }
  }

  const int i1 = lane & 7, G1 = lane >> 3;     // pass 1: coefficient column i fastest
  const int G2 = lane & 3, y2 = lane >> 2;     // pass 2: block group fastest
  const int q = a.q;
  const int qbias = int(0x80000000u) - 128 * q;      // byte*q + qbias = (byte-128)*q + 2^31
  constexpr T SCALE = T(p2(2 * M8I::exp(0)));

  int k = 0;
  for (int tile = w.tile; tile < w.ntiles; tile += w.stride, ++k) {
    const int s = k % NSTAGE;
    unsigned char* tb = ring + s * TILE;
    tma::mbar_wait(&full[s], (k / NSTAGE) & 1);

    // ---- pass 1: dequantise + inverse DCT over u (axis 0) -------------------------
    {
      uint32_t wd[8][3];
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(tb) + i1 * 12 + 3 * G1;
#pragma unroll
      for (int uu = 0; uu < 8; ++uu) {
        wd[uu][0] = rw[uu * 96 + 0];
        wd[uu][1] = rw[uu * 96 + 1];
        wd[uu][2] = rw[uu * 96 + 2];
      }
      T* fw = F + i1 * P + 4 * G1;
#pragma unroll(F64 ? 1 : 3)
      for (int c = 0; c < 3; ++c) {
        T v[4][8];
#pragma unroll
        for (int uu = 0; uu < 8; ++uu) {
          // bytes 3b + c of the 12-byte run, b = 0..3: shift the 96-bit run right by c bytes
          const uint32_t s0 = __funnelshift_r(wd[uu][0], wd[uu][1], 8 * c);
          const uint32_t s1 = __funnelshift_r(wd[uu][1], wd[uu][2], 8 * c);
          const uint32_t s2 = wd[uu][2] >> (8 * c);
          const int b0 = int(s0 & 255u), b1 = int(s0 >> 24), b2 = int((s1 >> 16) & 255u), b3 = int((s2 >> 8) & 255u);
          // (byte - 128) * q: int16 * int of src/2D-DCT.py:398-410 (cannot wrap for q <= 255)
          v[0][uu] = from_biased_int<T>(unsigned(b0 * q + qbias));
          v[1][uu] = from_biased_int<T>(unsigned(b1 * q + qbias));
          v[2][uu] = from_biased_int<T>(unsigned(b2 * q + qbias));
          v[3][uu] = from_biased_int<T>(unsigned(b3 * q + qbias));
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) dct8_inv<T, EXACT>(v[b]);
#pragma unroll
        for (int yy = 0; yy < 8; ++yy) {
          T* d = fw + (c * 8 + yy) * L::PP;
          if (!F64) {
            *reinterpret_cast<float4*>(d) = make_float4(float(v[0][yy]), float(v[1][yy]), float(v[2][yy]), float(v[3][yy]));
          } else {
            *reinterpret_cast<double2*>(d) = make_double2(double(v[0][yy]), double(v[1][yy]));
            *reinterpret_cast<double2*>(d + 2) = make_double2(double(v[2][yy]), double(v[3][yy]));
          }
        }
      }
    }
    __syncwarp();

    // ---- pass 2: inverse DCT over i (axis 1), to_RGB, +128, clip, truncate -----------
    {
      constexpr int NB = F64 ? 2 : 4;            // blocks per step
      const T* fr = F + y2 * L::PP + 4 * G2;
      uint4* orow = reinterpret_cast<uint4*>(tb + y2 * (WT * 3) + 96 * G2);
#pragma unroll(F64 ? 1 : 4)
      for (int st = 0; st < 4 / NB; ++st) {
        T v[3][NB][8];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const T* sp = fr + c * 8 * L::PP + i * P + NB * st;
            if (!F64) {
              const float4 t4 = *reinterpret_cast<const float4*>(sp);
              v[c][0][i] = T(t4.x);
              v[c][1][i] = T(t4.y);
              v[c][NB - 2][i] = T(t4.z);
              v[c][NB - 1][i] = T(t4.w);
            } else {
              const double2 t2 = *reinterpret_cast<const double2*>(sp);
              v[c][0][i] = T(t2.x);
              v[c][1][i] = T(t2.y);
            }
          }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int b = 0; b < NB; ++b) dct8_inv<T, EXACT>(v[c][b]);
        int px[NB][8][3];
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
          for (int x = 0; x < 8; ++x) {
            const T Y = v[0][b][x], Co = v[1][b][x], Cg = v[2][b][x];
            // to_RGB: Y + Co - Cg ; Y + Cg ; Y - Co - Cg (left to right), then += 128;
            // the lazy 2^-4 of the two codelets rides on the fma (exact product).
            const T R = O::fma(O::sub(O::add(Y, Co), Cg), SCALE, T(128));
            const T Gc = O::fma(O::add(Y, Cg), SCALE, T(128));
            const T Bc = O::fma(O::sub(O::sub(Y, Co), Cg), SCALE, T(128));
            px[b][x][0] = clamp255(trunc_to_int<T>(R));      // np.clip(y,0,255).astype(uint8)
            px[b][x][1] = clamp255(trunc_to_int<T>(Gc));
            px[b][x][2] = clamp255(trunc_to_int<T>(Bc));
          }
        // NB blocks x 8 px x 3 B = NB * 24 bytes
        const int* p = &px[0][0][0];
        uint32_t ww[NB * 6];
#pragma unroll
        for (int j = 0; j < NB * 6; ++j) ww[j] = pack4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
#pragma unroll
        for (int j = 0; j < NB * 6 / 4; ++j)
          orow[(NB * 6 / 4) * st + j] = make_uint4(ww[4 * j], ww[4 * j + 1], ww[4 * j + 2], ww[4 * j + 3]);
      }
    }
    tma::fence_proxy_async();
    __syncwarp();

    if (lane == 0) {
      issue_store(s, tile);
      tma::wait_group_read<1>();
      if (k >= 1) {
        const int nt = tile + (NSTAGE - 1) * w.stride;
        if (nt < w.ntiles) issue_load((k - 1) % NSTAGE, nt);
      }
    }
    __syncwarp();
  }
  if (lane == 0) tma::wait_group<0>();
}

// ---- float64 decode, half-tile variant ------------------------------------------------
// Same pipeline, but the 16-block tile is processed as two halves of 8 blocks so that the
// float64 intermediate is 16.9 KB instead of 28 KB per warp: 8 warps per SM (two per
// scheduler) instead of 4, which is what keeps the FP64 pipe fed.  Lane mapping: pass 1 =
// (coefficient column i, pair of blocks), pass 2 = (pixel row y, pair of blocks).  The index
// words of both halves are read into registers before the first half's RGB bytes overwrite
// the (aliased) tile buffer.
constexpr int H64_P = 10;                // doubles per (y,i) row of 8 blocks: 20 words = 4*odd
constexpr int H64_PP = 8 * H64_P + 8;    // 88 doubles: odd y lands on the other 16 banks
constexpr int H64_F_BYTES = 3 * 8 * H64_PP * 8;
constexpr int H64_WARP_SMEM = 26240;
static_assert(NSTAGE * TILE + H64_F_BYTES + 8 * NSTAGE <= H64_WARP_SMEM, "decode f64 half-tile smem");

template <bool EXACT, int NWARPS, int CTAS, bool DCSKIP, bool SSE>
__global__ void __launch_bounds__(NWARPS * 32, CTAS)
dec8_f64h_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                 const FastDecArgs a) {
  using T = double;
  using O = Ops<double, EXACT>;
  extern __shared__ __align__(128) unsigned char smem[];
  if (not_chosen(a)) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* ring = smem + warp * H64_WARP_SMEM;
  T* F = reinterpret_cast<T*>(ring + NSTAGE * TILE);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + NSTAGE * TILE + H64_F_BYTES);

  if (lane == 0) {
    tma::prefetch_map(&in_map);
    tma::prefetch_map(&out_map);
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) tma::mbar_init(&full[s], 1);
    tma::fence_mbar_init();
  }
  __syncwarp();

  Walker w;
  w.tile = blockIdx.x * NWARPS + warp;
  w.stride = gridDim.x * NWARPS;
  w.ntiles = a.ntiles;
  w.tiles_x = a.tiles_x;
  w.per_frame = a.ny * a.tiles_x;
  w.top = a.top;
  auto issue_load = [&](int s, int t) {
    int f, by, tx;
    w.coords(t, f, by, tx);
    tma::mbar_expect_tx(&full[s], TILE);
    tma::load_5d(ring + s * TILE, &in_map, &full[s], tx * (WT / 8) * 3, 0, by, 0, f);
  };
  auto issue_store = [&](int s, int t) {
    int f, by, tx;
    w.coords(t, f, by, tx);
    tma::store_3d(&out_map, ring + s * TILE, tx * (WT * 3 / 8), by * 8 - w.top, f);
    tma::commit_group();
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) {
      const int t = w.tile + s * w.stride;
      if (t < w.ntiles) issue_load(s, t);
    }
  }

  const int i1 = lane & 7, G1 = lane >> 3;     // pass 1
  const int G2 = lane & 3, y2 = lane >> 2;     // pass 2
  const int q = a.q;
  const int woff = (6 * G1) >> 2;              // word of the first byte of the lane's 6-byte run
  const int sh0 = ((6 * G1) & 3) * 8;          // bit offset inside that word (0 or 16)
  constexpr T SCALE = T(p2(2 * M8I::exp(0)));
  Lane L;
  L.i1 = i1; L.G1 = G1; L.G2 = G2; L.y2 = y2; L.sh0 = sh0; L.q = q;
  const uint32_t dcm = i1 == 0 ? 0u : 0xffffffffu;     // row 0 of column 0 holds the DC indices
  SseAcc acc;
  sse_reset(acc);
  // the 48 bytes of the original frame under this lane's bytes of half h of a tile
  auto load_orig = [&](int t, int h, uint4 (&o)[3]) {
    int f, by, tx;
    w.coords(t, f, by, tx);
    const uint4* p = reinterpret_cast<const uint4*>(a.original + f * a.frame_bytes + (long long)(by * 8 + y2) * a.row_bytes +
                                                    tx * (WT * 3) + 192 * h + 48 * G2);
    o[0] = __ldg(p);
    o[1] = __ldg(p + 1);
    o[2] = __ldg(p + 2);
  };

  int k = 0;
  for (int tile = w.tile; tile < w.ntiles; tile += w.stride, ++k) {
    const int s = k % NSTAGE;
    unsigned char* tb = ring + s * TILE;
    tma::mbar_wait(&full[s], (k / NSTAGE) & 1);

    uint32_t wd[2][8][2];
    {
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(tb) + i1 * 12 + woff;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int uu = 0; uu < 8; ++uu) {
          wd[h][uu][0] = rw[uu * 96 + 6 * h];
          wd[h][uu][1] = rw[uu * 96 + 6 * h + 1];
        }
    }
    __syncwarp();
    // Per half-tile: which channels carry AC indices (bits 16..18), which coefficient rows u
    // (bits 0..7) and columns i (bits 8..15) hold any non-zero index.  A channel without AC indices
    // is not transformed -- every sample of block b equals (q k_b c0) c0 exactly (dec8_dc.cuh): smooth
    // content at a coarse step, the chroma planes of most natural content from q ~ 24 up.  Rows /
    // columns beyond the first 2 or 4 that are zero throughout select the pruned codelets
    // (dct8_inv_low2 / _low4: the same operations minus those on exact zeros).
    unsigned cls[2] = {0x7ffffu, 0x7ffffu};
    if (DCSKIP) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t nz0 = (wd[h][0][0] ^ 0x80808080u) & dcm, nz1 = (wd[h][0][1] ^ 0x80808080u) & dcm;
        unsigned rows = 1u;                       // row 0 holds the DC indices
#pragma unroll
        for (int uu = 1; uu < 8; ++uu) {
          const uint32_t t0 = wd[h][uu][0] ^ 0x80808080u, t1 = wd[h][uu][1] ^ 0x80808080u;
          nz0 |= t0;
          nz1 |= t1;
          rows |= ((t0 | t1) ? 1u : 0u) << uu;    // whole words: neighbours of the same half-tile included
        }
        const unsigned col = (i1 == 0 || ((nz0 | nz1) | (wd[h][0][0] ^ 0x80808080u) | (wd[h][0][1] ^ 0x80808080u))) ? (1u << i1) : 0u;
        // the lane's 6-byte run: Y Co Cg Y | Co Cg
        const uint32_t lo = __funnelshift_r(nz0, nz1, sh0), hi = nz1 >> sh0;
        const unsigned ch = ((lo & 0xff0000ffu) ? 1u : 0u) | (((lo & 0x0000ff00u) | (hi & 0x000000ffu)) ? 2u : 0u) |
                            (((lo & 0x00ff0000u) | (hi & 0x0000ff00u)) ? 4u : 0u);
        cls[h] = __reduce_or_sync(0xffffffffu, rows | (col << 8) | (ch << 16));
      }
    }
    const unsigned chan6 = ((cls[0] >> 16) & 7u) | (((cls[1] >> 16) & 7u) << 3);
    const bool dc_tile = DCSKIP && chan6 == 0u;
    if (dc_tile) {
      uint4 og[2][3];
      if (SSE) {
        load_orig(tile, 0, og[0]);
        load_orig(tile, 1, og[1]);
      }
      dc_blocks(L, wd[0][0][0], wd[0][0][1], tb, 0, 0xffu);
      dc_blocks(L, wd[1][0][0], wd[1][0][1], tb, 1, 0xffu);
      __syncwarp();
      if (SSE) {
        sse_row48(acc, og[0], tb + y2 * (WT * 3) + 48 * G2);
        sse_row48(acc, og[1], tb + y2 * (WT * 3) + 192 + 48 * G2);
      }
    } else {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint4 og[3];
      if (SSE) load_orig(tile, h, og);
      // number of leading coefficient rows / columns that may hold non-zero indices: 2, 4 or 8
      const unsigned rows8 = cls[h] & 0xffu, cols8 = (cls[h] >> 8) & 0xffu;
      constexpr bool PRUNE4 = false;       // the 4-input codelets: 18 % fewer operations, but a third copy of both passes
      const int NU = !DCSKIP ? 8 : (rows8 & 0xfcu) == 0u ? 2 : (PRUNE4 && (rows8 & 0xf0u) == 0u) ? 4 : 8;
      const int NI = !DCSKIP ? 8 : (cols8 & 0xfcu) == 0u ? 2 : (PRUNE4 && (cols8 & 0xf0u) == 0u) ? 4 : 8;
      // ---- pass 1: dequantise + inverse DCT over u, 2 blocks per lane and channel ------
      auto pass1 = [&](auto nu_c) {
        constexpr int U = decltype(nu_c)::value;
        T* fw = F + i1 * H64_P + 2 * G1;
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
          if (DCSKIP && !((chan6 >> (3 * h + c)) & 1u)) continue;      // no AC index in this channel
          T v[2][8];
#pragma unroll
          for (int uu = 0; uu < U; ++uu) {
            const uint32_t sv = __funnelshift_rc(wd[h][uu][0], wd[h][uu][1], sh0 + 8 * c);
            // (byte - 128) * q: int16 * int of src/2D-DCT.py:398-410 (cannot wrap for q <= 255);
            // the int -> double conversion is exact and runs off the FP64 pipe
            v[0][uu] = __int2double_rn(int(sv & 255u) * q - 128 * q);
            v[1][uu] = __int2double_rn(int(sv >> 24) * q - 128 * q);
          }
          if (U == 2) {
            dct8_inv_low2<T, EXACT>(v[0]);
            dct8_inv_low2<T, EXACT>(v[1]);
          } else if (U == 4) {
            dct8_inv_low4<T, EXACT>(v[0]);
            dct8_inv_low4<T, EXACT>(v[1]);
          } else {
            dct8_inv<T, EXACT>(v[0]);
            dct8_inv<T, EXACT>(v[1]);
          }
#pragma unroll
          for (int yy = 0; yy < 8; ++yy)
            *reinterpret_cast<double2*>(fw + (c * 8 + yy) * H64_PP) = make_double2(v[0][yy], v[1][yy]);
        }
      };
      if (i1 < NI) {               // columns beyond NI are zero throughout: pass 2 does not read them
        if (NU == 2) pass1(std::integral_constant<int, 2>());
        else if (NU == 4) pass1(std::integral_constant<int, 4>());
        else pass1(std::integral_constant<int, 8>());
      }
      __syncwarp();
      // ---- pass 2: inverse DCT over i, to_RGB, +128, clip, truncate ----------------------
      {
        const T* fr = F + y2 * H64_PP + 2 * G2;
        T v[3][2][8];
        const unsigned chan3 = (chan6 >> (3 * h)) & 7u;
        uint32_t yccA = 0, yccB = 0;            // DC index bytes (Y, Co, Cg) of the lane's two blocks
        if (DCSKIP && chan3 != 7u) {
          const uint32_t lo = __funnelshift_r(wd[h][0][0], wd[h][0][1], sh0), hi = wd[h][0][1] >> sh0;
          yccA = __shfl_sync(0xffffffffu, lo, 8 * G2);                          // lane (i1 = 0, pair G2) holds them
          yccB = __shfl_sync(0xffffffffu, __byte_perm(lo, hi, 0x0543), 8 * G2);
        }
        auto pass2 = [&](auto ni_c) {
          constexpr int I = decltype(ni_c)::value;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            if (!DCSKIP || ((chan3 >> c) & 1u)) {
#pragma unroll
              for (int i = 0; i < I; ++i) {
                const double2 t2 = *reinterpret_cast<const double2*>(fr + c * 8 * H64_PP + i * H64_P);
                v[c][0][i] = t2.x;
                v[c][1][i] = t2.y;
              }
              if (I == 2) {
                dct8_inv_low2<T, EXACT>(v[c][0]);
                dct8_inv_low2<T, EXACT>(v[c][1]);
              } else if (I == 4) {
                dct8_inv_low4<T, EXACT>(v[c][0]);
                dct8_inv_low4<T, EXACT>(v[c][1]);
              } else {
                dct8_inv<T, EXACT>(v[c][0]);
                dct8_inv<T, EXACT>(v[c][1]);
              }
            } else {
              const T ka = dc_chain(int((yccA >> (8 * c)) & 255u), q), kb = dc_chain(int((yccB >> (8 * c)) & 255u), q);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                v[c][0][i] = ka;
                v[c][1][i] = kb;
              }
            }
          }
        };
        if (NI == 2) pass2(std::integral_constant<int, 2>());
        else if (NI == 4) pass2(std::integral_constant<int, 4>());
        else pass2(std::integral_constant<int, 8>());
        int px[2][8][3];
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int x = 0; x < 8; ++x) {
            const T Y = v[0][b][x], Co = v[1][b][x], Cg = v[2][b][x];
            const T R = O::fma(O::sub(O::add(Y, Co), Cg), SCALE, T(128));
            const T Gc = O::fma(O::add(Y, Cg), SCALE, T(128));
            const T Bc = O::fma(O::sub(O::sub(Y, Co), Cg), SCALE, T(128));
            px[b][x][0] = clamp255(__double2int_rz(R));
            px[b][x][1] = clamp255(__double2int_rz(Gc));
            px[b][x][2] = clamp255(__double2int_rz(Bc));
          }
        const int* p = &px[0][0][0];
        uint32_t ww[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) ww[j] = pack4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
        uint4* orow = reinterpret_cast<uint4*>(tb + y2 * (WT * 3) + 192 * h + 48 * G2);
        orow[0] = make_uint4(ww[0], ww[1], ww[2], ww[3]);
        orow[1] = make_uint4(ww[4], ww[5], ww[6], ww[7]);
        orow[2] = make_uint4(ww[8], ww[9], ww[10], ww[11]);
        if (SSE) sse_row48(acc, og, reinterpret_cast<const unsigned char*>(orow));   // the lane's own bytes
      }
      __syncwarp();
    }
    }
    tma::fence_proxy_async();
    __syncwarp();

    if (lane == 0) {
      issue_store(s, tile);
      tma::wait_group_read<1>();
      if (k >= 1) {
        const int nt = tile + (NSTAGE - 1) * w.stride;
        if (nt < w.ntiles) issue_load((k - 1) % NSTAGE, nt);
      }
    }
    __syncwarp();
  }
  if (lane == 0) tma::wait_group<0>();
  if (SSE) sse_finish(acc, a.stats, lane);
}

// ---- host side -----------------------------------------------------------------

bool fast_geometry_ok(const Geom& g, const void* p0, const void* p1) {
  if (g.W % 16 != 0 || g.nx % 16 != 0 || g.left != 0) return false;
  if ((reinterpret_cast<uintptr_t>(p0) & 15) || (reinterpret_cast<uintptr_t>(p1) & 15)) return false;
  return tma::encode_tiled_fn() != nullptr;
}

// RGB frames as (W*3/8 uint64, H, n); box = one warp tile (48 uint64 x 8 rows)
bool make_rgb_map(CUtensorMap* m, const Geom& g, int n, const void* base) {
  const uint64_t dims[3] = {uint64_t(g.W) * 3 / 8, uint64_t(g.H), uint64_t(n)};
  const uint64_t str[2] = {uint64_t(g.W) * 3, uint64_t(g.H) * g.W * 3};
  const uint32_t box[3] = {WT * 3 / 8, 8, 1};
  return tma::make_map(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(base), dims, str, box);
}

// Index planes sub[j*ny + y, i*nx + x, c] as a 5-D tensor.  i_major=false: dims (x bytes,
// i, y, j, frame) -> smem box [j][i][48]; i_major=true: dims (x bytes, j, y, i, frame)
// -> smem box [i][j][48] (the order that makes the encoder's stores conflict-free).
bool make_idx_map(CUtensorMap* m, const Geom& g, int n, const void* base, bool i_major) {
  const uint64_t si = uint64_t(g.nx) * 3, sy = uint64_t(g.Wp) * 3, sj = uint64_t(g.ny) * g.Wp * 3,
                 sf = uint64_t(g.Hp) * g.Wp * 3;
  const uint64_t dims[5] = {uint64_t(g.nx) * 3, 8, uint64_t(g.ny), 8, uint64_t(n)};
  const uint64_t str_j[4] = {si, sy, sj, sf};
  const uint64_t str_i[4] = {sj, sy, si, sf};
  const uint32_t box[5] = {WT / 8 * 3, 8, 1, 8, 1};
  return tma::make_map(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, const_cast<void*>(base), dims, i_major ? str_i : str_j,
                       box);
}

}  // namespace

int fast::sm_count() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

template <int NWARPS, int CTAS>
static int launch_enc_t(bool exact, bool qpow2, const CUtensorMap& in_map, const CUtensorMap& out_map,
                        const FastArgs& fa, cudaStream_t s) {
  int grid = sm_count() * CTAS;
  const int need = (fa.ntiles + NWARPS - 1) / NWARPS;
  if (grid > need) grid = need;
  void (*kern)(const CUtensorMap, const CUtensorMap, const FastArgs);
  static const bool scalar_only = getenv("VCFB_ENC_SCALAR") != nullptr;    // development knob
  if (exact && !scalar_only) return launch_encode_packed(NWARPS * 10 + CTAS, qpow2, in_map, out_map, fa, s);
  if (exact) kern = qpow2 ? enc8_fast_kernel<true, true, NWARPS, CTAS> : enc8_fast_kernel<true, false, NWARPS, CTAS>;
  else kern = qpow2 ? enc8_fast_kernel<false, true, NWARPS, CTAS> : enc8_fast_kernel<false, false, NWARPS, CTAS>;
  const int smem_bytes = NWARPS * ENC_WARP_SMEM;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(enc8_fast)");
  note_kernel("enc8_fast");
  kern<<<grid, NWARPS * 32, smem_bytes, s>>>(in_map, out_map, fa);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "enc8_fast_kernel launch");
  return VCFB_OK;
}

// Returns VCFB_E_UNSUPP when the request is outside the fast path (the caller then
// uses the general kernel), VCFB_OK after a launch, or an error.
int launch_encode_fast(const EncArgs& a, int B, cudaStream_t s) {
  if (B != 8 || a.color != VCFB_COLOR_YCOCG) return VCFB_E_UNSUPP;
  if (a.flags & (VCFB_F_NO_SUBBANDS | VCFB_F_PERCEPTUAL | VCFB_F_FP64)) return VCFB_E_UNSUPP;
  const Geom& g = a.g;
  if (!fast_geometry_ok(g, a.rgb, a.idx)) return VCFB_E_UNSUPP;
  const bool exact = !(a.flags & VCFB_F_CONTRACT);
  static const bool scalar_enc = getenv("VCFB_ENC_SCALAR") != nullptr;
  // the packed exact encoder counts non-zero indices and sum |k| itself; the histogram, the
  // scalar / contracted variants take a separate streaming pass over the indices (kernels_stats.cu)
  const bool fused_stats = a.stats && exact && !scalar_enc && !(a.flags & VCFB_F_HIST);
  if (a.stats && !fused_stats) {
    EncArgs b = a;
    b.stats = nullptr;
    int rc = launch_encode_fast(b, B, s);
    if (rc) return rc;
    return launch_index_stats(a.idx, (long long)a.n_frames * g.Hp * g.Wp * 3, (a.flags & VCFB_F_HIST) != 0, a.stats, s);
  }
  CUtensorMap in_map, out_map;
  if (!make_rgb_map(&in_map, g, a.n_frames, a.rgb)) return VCFB_E_UNSUPP;
  if (!make_idx_map(&out_map, g, a.n_frames, a.idx, true)) return VCFB_E_UNSUPP;

  FastArgs fa;
  fa.tiles_x = g.Wp / WT;
  fa.ny = g.ny;
  fa.top = g.top;
  const long long nt = (long long)a.n_frames * g.ny * fa.tiles_x;
  if (nt > 0x7fffffffLL - (1 << 20)) return VCFB_E_UNSUPP;
  fa.ntiles = int(nt);
  fa.q = float(a.q);
  for (int u = 0; u < 8; ++u)
    for (int c = 0; c < 3; ++c) {
      const int cexp = (c == 1) ? -1 : -2;                  // 2*Co, 4*Y, 4*Cg
      double sc = M8F::sgn(u) * p2(M8F::exp(u) + cexp + min_exp8());
      if (a.q_pow2) sc *= a.inv_q;
      fa.qtab[u][c] = float(sc);                             // a power of two: exact
    }

  fa.stats = fused_stats ? a.stats : nullptr;
  if (fused_stats) {
    int rc = launch_add_count(a.stats, VCFB_STAT_NINDICES, (unsigned long long)a.n_frames * g.Hp * g.Wp * 3, s);
    if (rc) return rc;
  }
  // development knob VCFB_ENC_CFG: 5x2 reproduces the scheduler imbalance of 5 warps per CTA,
  // 4x3 the 2-stage ring with 12 warps per SM (both measured slower)
  switch (dev_cfg("VCFB_ENC_CFG")) {
    case 52: return launch_enc_t<5, 2>(exact, a.q_pow2, in_map, out_map, fa, s);
    case 43: if (exact) return launch_encode_packed(43, a.q_pow2, in_map, out_map, fa, s);   // fallthrough
    default: return launch_enc_t<4, 2>(exact, a.q_pow2, in_map, out_map, fa, s);
  }
}

template <typename T, bool EXACT, int NWARPS, int CTAS>
static int launch_dec_t(const CUtensorMap& in_map, const CUtensorMap& out_map, const FastDecArgs& fa, cudaStream_t s) {
  using L = DecL<T>;
  int grid = sm_count() * CTAS;
  const int need = (fa.ntiles + NWARPS - 1) / NWARPS;
  if (grid > need) grid = need;
  auto kern = dec8_fast_kernel<T, EXACT, NWARPS, CTAS>;
  const int smem_bytes = NWARPS * L::WARP_SMEM;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dec8_fast)");
  note_kernel("dec8_fast");
  kern<<<grid, NWARPS * 32, smem_bytes, s>>>(in_map, out_map, fa);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "dec8_fast_kernel launch");
  return VCFB_OK;
}

template <int NWARPS, int CTAS, bool DCSKIP = true, bool SSE = false>
static int launch_dec_f64h(const CUtensorMap& in_map, const CUtensorMap& out_map, const FastDecArgs& fa,
                           cudaStream_t s) {
  int grid = sm_count() * CTAS;
  const int need = (fa.ntiles + NWARPS - 1) / NWARPS;
  if (grid > need) grid = need;
  auto kern = dec8_f64h_kernel<true, NWARPS, CTAS, DCSKIP, SSE>;
  const int smem_bytes = NWARPS * H64_WARP_SMEM;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dec8_f64h)");
  note_kernel("dec8_fast");
  kern<<<grid, NWARPS * 32, smem_bytes, s>>>(in_map, out_map, fa);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "dec8_f64h_kernel launch");
  return VCFB_OK;
}


// ---- probe: which float64 decoder suits this batch ------------------------------------------
// All three decoders give identical bytes; they differ in speed by content:
//   * exact chain (dec8_f64h_kernel): the reference's operation sequence for every sample;
//   * the same with the DC-only shortcut per (half-tile, channel): pays ~4 % where every channel
//     has AC indices, skips the transform of a channel that has none (the chroma planes of most
//     natural content from q ~ 24 up) and runs at HBM speed on tiles without any (smooth content
//     at a coarse step);
//   * two-tier (kernels_dec2t.cu): 27 % faster when exact-integer samples are rare, i.e. when
//     the indices are dense; up to 1.8x slower when they are sparse (+-1 indices that cancel).
// 32 CTAs read 256 tiles spread over the batch and count sparse blocks and (tile, channel) pairs
// without AC indices; the last one writes the choice to a slot in device memory.  The three
// kernels are launched behind it and the two that were not chosen return at once.  Correctness
// never depends on the choice.
constexpr int PROBE_TILES = 256;
constexpr int PROBE_CTAS = 32;             // x 8 warps x 1 tile
constexpr int PROBE_MIN_TILES = 4096;      // smaller jobs: not worth three extra launches

struct ProbeSlot {
  int sparse, dc_chan, low_tiles, ticket, choice;     // all but `choice` are zero between launches
};

struct ProbeArgs {
  const uint8_t* idx;
  long long frame_bytes;       // Hp * Wp * 3
  int row_bytes;               // Wp * 3
  int ny, nx, tiles_x, per_frame, ntiles;
  ProbeSlot* slot;
};

__device__ __forceinline__ uint32_t nonzero_flags(uint32_t x) {      // 0x80 in every byte of x that is not 0x80
  const uint32_t y = x ^ 0x80808080u;
  return (y | ((y & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;
}

// Counts, per sampled tile, the non-zero AC indices of each of its 16 blocks (all channels).
// A block with 1..6 of them is "sparse": the kind whose samples land on exact integers.
__global__ void __launch_bounds__(256, 1) dec8_probe_kernel(const ProbeArgs a) {
  __shared__ int s_sparse, s_dc, s_low;     // sparse blocks; (tile, channel) pairs without AC indices;
  if (threadIdx.x == 0) {                   // tiles whose rows / columns 4..7 are zero throughout
    s_sparse = 0;
    s_dc = 0;
    s_low = 0;
  }
  constexpr unsigned CH[3][3] = {{0xFF0000FFu, 0x0000FF00u, 0x00FF0000u},     // bytes of channel c in word k,
                                 {0x00FF0000u, 0xFF0000FFu, 0x0000FF00u},     // k mod 3 = 0, 1, 2
                                 {0x0000FF00u, 0x00FF0000u, 0xFF0000FFu}};
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int step = a.ntiles / PROBE_TILES;
  int nsparse = 0, ndc = 0, nlow = 0;
  for (int sidx = blockIdx.x * 8 + warp; sidx < PROBE_TILES; sidx += PROBE_CTAS * 8) {
    const int t = sidx * step + (sidx * 7) % step;
    const int f = t / a.per_frame, rem = t - f * a.per_frame;
    const int by = rem / a.tiles_x, tx = rem - by * a.tiles_x;
    uint32_t acc[4] = {0u, 0u, 0u, 0u};        // 16 byte-wide counters: block b in byte b & 3 of acc[b >> 2]
    uint32_t chf[3] = {0u, 0u, 0u};            // non-zero flags per channel
    uint32_t hif = 0u;                         // non-zero flags in coefficient rows / columns 4..7
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int seg = lane + 32 * r, j = seg >> 3, i = seg & 7;          // subband (j, i): 16 blocks x 3 bytes
      const uint4* p = reinterpret_cast<const uint4*>(a.idx + f * a.frame_bytes + (long long)(j * a.ny + by) * a.row_bytes +
                                                      (i * a.nx + tx * 16) * 3);
      const uint4 v0 = __ldg(p), v1 = __ldg(p + 1), v2 = __ldg(p + 2);
      const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
      if (seg != 0) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {           // 3 words = 4 blocks
          const uint32_t f0 = nonzero_flags(w[3 * g]), f1 = nonzero_flags(w[3 * g + 1]), f2 = nonzero_flags(w[3 * g + 2]);
          const uint32_t c0 = __popc(f0 & 0x00808080u);
          const uint32_t c1 = __popc(f0 & 0x80000000u) + __popc(f1 & 0x00008080u);
          const uint32_t c2 = __popc(f1 & 0x80800000u) + __popc(f2 & 0x00000080u);
          const uint32_t c3 = __popc(f2 & 0x80808000u);
          acc[g] += c0 | (c1 << 8) | (c2 << 16) | (c3 << 24);          // <= 3 per segment, 189 per tile: no carry
#pragma unroll
          for (int c = 0; c < 3; ++c) chf[c] |= (f0 & CH[0][c]) | (f1 & CH[1][c]) | (f2 & CH[2][c]);
          if (j >= 4 || i >= 4) hif |= f0 | f1 | f2;
        }
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) acc[g] = __reduce_add_sync(0xffffffffu, acc[g]);
    const uint32_t mine = lane < 16 ? (acc[lane >> 2] >> (8 * (lane & 3))) & 0xffu : 0xffu;
    nsparse += __popc(__ballot_sync(0xffffffffu, mine >= 1u && mine <= 6u));
#pragma unroll
    for (int c = 0; c < 3; ++c) ndc += !__any_sync(0xffffffffu, chf[c] != 0u);
    nlow += !__any_sync(0xffffffffu, hif != 0u);
  }
  if (lane == 0) {
    atomicAdd(&s_sparse, nsparse);
    atomicAdd(&s_dc, ndc);
    atomicAdd(&s_low, nlow);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ProbeSlot* sl = a.slot;
    atomicAdd(&sl->sparse, s_sparse);
    atomicAdd(&sl->dc_chan, s_dc);
    atomicAdd(&sl->low_tiles, s_low);
    __threadfence();
    if (atomicAdd(&sl->ticket, 1) == PROBE_CTAS - 1) {      // last CTA: decide, and leave the slot clean
      __threadfence();
      const int sparse = atomicExch(&sl->sparse, 0), dc = atomicExch(&sl->dc_chan, 0), low = atomicExch(&sl->low_tiles, 0);
      int kind = DEC_EXACT;
      if (50 * sparse <= PROBE_TILES * 16) kind = DEC_TWO_TIER;         // <= 2 % sparse blocks
      else if (8 * dc > 3 * PROBE_TILES ||                              // > 1/8 of the (tile, channel) pairs without AC,
               4 * low > PROBE_TILES)                                   // or > 1/4 of the tiles low-frequency only
        kind = DEC_EXACT_DCSKIP;
      sl->choice = kind;
      sl->ticket = 0;
    }
  }
}

// Slots for the probe's answer: one small ring per device, allocated on first use and kept for
// the life of the process; concurrent launches (other streams, other threads) get distinct slots.
static ProbeSlot* probe_slot() {
  constexpr int NSLOT = 4096, MAXDEV = 64;
  static std::mutex mu;
  static ProbeSlot* ring[MAXDEV] = {};
  static unsigned next[MAXDEV] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAXDEV) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!ring[dev]) {
    if (cudaMalloc(&ring[dev], NSLOT * sizeof(ProbeSlot)) != cudaSuccess ||
        cudaMemset(ring[dev], 0, NSLOT * sizeof(ProbeSlot)) != cudaSuccess ||
        cudaStreamSynchronize(0) != cudaSuccess) {       // once per device: the zeros are there before any probe
      cudaGetLastError();
      ring[dev] = nullptr;
      return nullptr;
    }
  }
  return ring[dev] + (next[dev]++ % NSLOT);
}

static int launch_decode_f64_probed(const DecArgs& a, const CUtensorMap& in_map, const CUtensorMap& out_map, FastDecArgs fa,
                             cudaStream_t s) {
  ProbeSlot* slot = fa.ntiles >= PROBE_MIN_TILES ? probe_slot() : nullptr;
  if (!slot) return fa.stats ? launch_dec_f64h<8, 1, false, true>(in_map, out_map, fa, s) : launch_dec_f64h<8, 1, false>(in_map, out_map, fa, s);
  const Geom& g = a.g;
  ProbeArgs pa;
  pa.idx = a.idx;
  pa.frame_bytes = (long long)g.Hp * g.Wp * 3;
  pa.row_bytes = g.Wp * 3;
  pa.ny = g.ny;
  pa.nx = g.nx;
  pa.tiles_x = fa.tiles_x;
  pa.per_frame = g.ny * fa.tiles_x;
  pa.ntiles = fa.ntiles;
  pa.slot = slot;
  dec8_probe_kernel<<<PROBE_CTAS, 256, 0, s>>>(pa);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "dec8_probe_kernel launch");
  note_extra_launches(1);
  fa.choice = &slot->choice;
  fa.kind = DEC_EXACT;
  int rc = fa.stats ? launch_dec_f64h<8, 1, false, true>(in_map, out_map, fa, s) : launch_dec_f64h<8, 1, false>(in_map, out_map, fa, s);
  if (rc) return rc;
  fa.kind = DEC_EXACT_DCSKIP;
  rc = fa.stats ? launch_dec_f64h<8, 1, true, true>(in_map, out_map, fa, s) : launch_dec_f64h<8, 1, true>(in_map, out_map, fa, s);
  if (rc) return rc;
  fa.kind = DEC_TWO_TIER;
  return launch_decode_2t(0, in_map, out_map, fa, s);
}

int launch_decode_fast(const DecArgs& a, int B, cudaStream_t s) {
  if (B != 8 || a.color != VCFB_COLOR_YCOCG) return VCFB_E_UNSUPP;
  if (a.flags & (VCFB_F_NO_SUBBANDS | VCFB_F_PERCEPTUAL)) return VCFB_E_UNSUPP;
  if (a.y_out || !a.rgb) return VCFB_E_UNSUPP;
  if ((a.stats != nullptr) != (a.original != nullptr)) return VCFB_E_UNSUPP;
  if (a.q_int < 1 || a.q_int > 255) return VCFB_E_UNSUPP;
  const Geom& g = a.g;
  const bool fused_sse = a.stats && (a.flags & VCFB_F_FP64) && dev_cfg("VCFB_DEC_CFG") == 0 &&
                         !(reinterpret_cast<uintptr_t>(a.original) & 15);
  if (a.stats && !fused_sse) {   // distortion = a separate streaming pass over (original, decoded)
    if (reinterpret_cast<uintptr_t>(a.original) & 15) return VCFB_E_UNSUPP;
    DecArgs b = a;
    b.stats = nullptr;
    b.original = nullptr;
    int rc = launch_decode_fast(b, B, s);
    if (rc) return rc;
    return launch_sse(a.original, a.rgb, (long long)a.n_frames * g.H * g.W * 3, a.stats, s);
  }
  // (a TMA store whose box starts at a negative row faults on sm_100a, so frames with
  //  vertical padding take the general kernel)
  if (!fast_geometry_ok(g, a.rgb, a.idx) || g.top != 0) return VCFB_E_UNSUPP;
  CUtensorMap in_map, out_map;
  if (!make_idx_map(&in_map, g, a.n_frames, a.idx, false)) return VCFB_E_UNSUPP;
  if (!make_rgb_map(&out_map, g, a.n_frames, a.rgb)) return VCFB_E_UNSUPP;
  FastDecArgs fa;
  fa.tiles_x = g.Wp / WT;
  fa.ny = g.ny;
  fa.top = g.top;
  const long long nt = (long long)a.n_frames * g.ny * fa.tiles_x;
  if (nt > 0x7fffffffLL - (1 << 20)) return VCFB_E_UNSUPP;
  fa.ntiles = int(nt);
  fa.q = a.q_int;
  fa.choice = nullptr;
  fa.kind = 0;
  fa.original = nullptr;
  fa.stats = nullptr;
  fa.frame_bytes = (long long)g.H * g.W * 3;
  fa.row_bytes = g.W * 3;
  if (fused_sse) {     // the float64 decoders accumulate SSE / SUMDIFF themselves; the sample count is known here
    fa.original = a.original;
    fa.stats = a.stats;
    int rc = launch_add_count(a.stats, VCFB_STAT_NSAMPLES, (unsigned long long)a.n_frames * g.H * g.W * 3, s);
    if (rc) return rc;
  }
  if (a.flags & VCFB_F_FP64) {
    // default: probe + device-side choice between the three float64 decoders.  Development knob
    // VCFB_DEC_CFG forces one: 9x2 exact chain, 9x1 exact chain + DC-only shortcut, 8x1 (and the
    // other shapes of kernels_dec2t.cu) two-tier, 4x1 the full-tile exact kernel (4 warps per SM)
    const int cfg = dev_cfg("VCFB_DEC_CFG");
    switch (cfg) {
      case 0: return launch_decode_f64_probed(a, in_map, out_map, fa, s);
      case 41: return launch_dec_t<double, true, 4, 1>(in_map, out_map, fa, s);
      case 91: return launch_dec_f64h<8, 1, true>(in_map, out_map, fa, s);
      case 92: return launch_dec_f64h<8, 1, false>(in_map, out_map, fa, s);
      default: return launch_decode_2t(cfg, in_map, out_map, fa, s);
    }
  }
  // float32 decode = the fast mode (+-1 LSB): scaled AAN transform + exact DC-only blocks
  // (kernels_dec32.cu).  Development knob VCFB_DEC32_CFG: 9x1 / 9x2 = the pocketfft codelets in
  // float32 (individually rounded / contracted), other values = launch shapes of the fast kernel.
  const int cfg32 = dev_cfg("VCFB_DEC32_CFG");
  // default: the tensor-core tier (kernels_tc.cu); VCFB_TC=0 or any VCFB_DEC32_CFG selects the CUDA-core kernels
  static const bool use_tc = !(getenv("VCFB_TC") && getenv("VCFB_TC")[0] == '0');
  if (use_tc && cfg32 == 0) {
    const int rc = launch_decode_tc(a, s);
    if (rc != VCFB_E_UNSUPP) return rc;
  }
  if (cfg32 == 91) return launch_dec_t<float, true, 4, 2>(in_map, out_map, fa, s);
  if (cfg32 == 92) return launch_dec_t<float, false, 4, 2>(in_map, out_map, fa, s);
  return launch_decode_f32a(cfg32, in_map, out_map, fa, s);
}

}  // namespace vcfb
