// Tensor-core tier of the B=8 fast path: the 8x8 inverse DCT of 128 blocks as one
// tcgen05 matrix product per channel,
//     D[128 blocks x 64 samples] (fp32, TMEM) = A[128 blocks x 64 indices] (fp16, TMEM) * M^T[64 x 64],
// M = the 64 x 64 Kronecker product of the orthonormal 8-point inverse DCT with itself, split into
// two fp16 limbs (hi + lo of 4096 * M: 22 significant bits, fp32-level) that accumulate into the
// same fp32 accumulator.  The indices are small integers, exact in fp16; the products are exact and
// the accumulation is fp32 (measured: profiles/microbench/tc_bringup.cu), so the result has the
// accuracy of a float32 transform -- this is a decoder of the north star's FAST MODE (pixels within
// +-1 LSB of the reference, PSNR within 0.01 dB), not of the bit-exact float64 mode.
//
// What is left on the CUDA cores is byte shuffling: index bytes -> fp16 (converter warps) and
// colour mix, * q, + 128, truncate, clip, pack (epilogue warps).  Blocks without AC indices are
// evaluated exactly like everywhere else (dec8_dc.cuh).
//
// One persistent CTA per SM, 10 warps:
//   warps 0-3  epilogue   thread = block = TMEM lane: tcgen05.ld 8 samples x 3 channels per pixel row,
//                         colour / scale / pack, 24 bytes per row into the RGB tile, TMA store
//   warps 4-7  converter  thread = block: 192 index bytes -> 96 packed fp16 pairs -> tcgen05.st (A lives
//                         in tensor memory: shared memory carries only the raw tiles and the matrix)
//   warp 8     TMA producer of index tiles (3-stage ring)
//   warp 9     MMA issuer (one thread), owner of the TMEM allocation
// Pipelines: idx_full/idx_empty (TMA <-> converter), a_full/a_empty (converter <-> MMA, A is single
// buffered), d_full/d_empty x 2 (MMA <-> epilogue, D is double buffered).
#include <cuda_fp16.h>
#include <math.h>

#include <mutex>
#include <vector>

#include "dec8_dc.cuh"
#include "fast_common.cuh"
#include "tc.cuh"

namespace vcfb {
using namespace fast;
namespace {

constexpr int TB = 128;                    // blocks per tile = TMEM lanes = MMA M
constexpr int IDX_TILE = 64 * TB * 3;      // [kk = 8 j + i][128 blocks x 3 bytes]
constexpr int OUT_TILE = 8 * TB * 24;      // two halves of [8 rows][64 blocks x 24 bytes]
constexpr int NSI = 3, NSO = 2;
constexpr int NLIMB = 2;
constexpr int LIMB_BYTES = 64 * 64 * 2;
constexpr int SCALE_LOG2 = 12;             // matrix entries are stored times 4096
constexpr int B_SBO = 128, B_LBO = 1024;   // canonical K-major, no swizzle: 8 x 16-byte rows per core matrix

constexpr int OFF_IDX = 0;
constexpr int OFF_OUT = OFF_IDX + NSI * IDX_TILE;
constexpr int OFF_B = OFF_OUT + NSO * OUT_TILE;
constexpr int OFF_DC = OFF_B + NLIMB * LIMB_BYTES;
constexpr int OFF_BAR = OFF_DC + 2 * TB * 4;
constexpr int SMEM_BYTES = OFF_BAR + 128;

constexpr int D_COLS = 192;                // 3 channels x 64 samples per stage
constexpr int A_COL = 2 * D_COLS;          // 3 channels x 32 packed columns behind the two D stages
constexpr int TMEM_COLS = 512;
constexpr int NTHREADS = 320;

struct TcDecArgs {
  int ntiles, tiles_x, ny, nx;
  int q;
  const unsigned char* btab;               // NLIMB * LIMB_BYTES, already in the canonical layout
};

struct Bars {
  uint64_t idx_full[NSI], idx_empty[NSI], a_full, a_empty, d_full[2], d_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ unsigned pack_sat_u8(int a, int b, unsigned c) {
  unsigned d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

__global__ void __launch_bounds__(NTHREADS, 1)
dec8_tc_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
               const TcDecArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  Bars* bars = reinterpret_cast<Bars*>(smem + OFF_BAR);
  uint32_t* dcinfo = reinterpret_cast<uint32_t*>(smem + OFF_DC);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    tma::prefetch_map(&in_map);
    tma::prefetch_map(&out_map);
    for (int s = 0; s < NSI; ++s) {
      tma::mbar_init(&bars->idx_full[s], 1);
      tma::mbar_init(&bars->idx_empty[s], 4);
    }
    tma::mbar_init(&bars->a_full, 4);
    tma::mbar_init(&bars->a_empty, 1);
    for (int s = 0; s < 2; ++s) {
      tma::mbar_init(&bars->d_full[s], 1);
      tma::mbar_init(&bars->d_empty[s], 4);
    }
    tma::fence_mbar_init();
  }
  {  // the matrix limbs, as prepared by the host
    const uint4* src = reinterpret_cast<const uint4*>(a.btab);
    uint4* dst = reinterpret_cast<uint4*>(smem + OFF_B);
    for (int i = threadIdx.x; i < NLIMB * LIMB_BYTES / 16; i += NTHREADS) dst[i] = __ldg(src + i);
  }
  tma::fence_proxy_async();               // generic writes of the matrix -> async proxy (tensor core)
  if (warp == 9) tc::tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(&bars->tmem_base);
  const int per_frame = a.ny * a.tiles_x;

  if (warp == 8) {
    // ===== TMA producer =====
    if (lane == 0) {
      int k = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
        const int s = k % NSI;
        tc::mbar_wait(&bars->idx_empty[s], ((k / NSI) & 1) ^ 1);
        const int f = tile / per_frame, rem = tile - f * per_frame, by = rem / a.tiles_x, tx = rem - by * a.tiles_x;
        tma::mbar_expect_tx(&bars->idx_full[s], IDX_TILE);
        tma::load_5d(smem + OFF_IDX + s * IDX_TILE, &in_map, &bars->idx_full[s], tx * (TB * 3 / 4), 0, by, 0, f);
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t IDESC = tc::idesc_f16(TB, 64);
      const uint32_t b_base = tma::smem_u32(smem + OFF_B);
      int k = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
        const int st = k & 1;
        tc::mbar_wait(&bars->d_empty[st], ((k >> 1) & 1) ^ 1);
        tc::mbar_wait(&bars->a_full, k & 1);
        tc::fence_after();
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int l = 0; l < NLIMB; ++l)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc::mma_ts(tbase + st * D_COLS + c * 64, tbase + A_COL + c * 32 + ks * 8,
                         tc::smem_desc(b_base + l * LIMB_BYTES + ks * 2 * B_LBO, B_LBO, B_SBO), IDESC, (l | ks) != 0);
        tc::commit(&bars->a_empty);
        tc::commit(&bars->d_full[st]);
      }
    }
  } else if (warp >= 4) {
    // ===== converter: thread = block =====
    const int b = threadIdx.x - 128;
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const __half2 bias = __floats2half2_rn(1152.0f, 1152.0f);       // 1024 (the magic) + 128 (the index bias)
    int k = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
      const int s = k % NSI;
      tc::mbar_wait(&bars->idx_full[s], (k / NSI) & 1);
      const unsigned char* tp = smem + OFF_IDX + s * IDX_TILE + 3 * b;
      uint32_t dcw = 0, flags = 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint32_t w[32];
        uint32_t nz = 0;
#pragma unroll
        for (int p = 0; p < 32; ++p) {
          const uint32_t b0 = tp[(2 * p) * (TB * 3) + c], b1 = tp[(2 * p + 1) * (TB * 3) + c];
          const uint32_t x = b0 | (b1 << 16);
          if (p == 0) {
            dcw |= b0 << (8 * c);
            nz |= b1 ^ 0x80u;
          } else {
            nz |= x ^ 0x00800080u;
          }
          // 0x6400 | byte is the fp16 number 1024 + byte: one packed subtraction gives the index, exactly
          const uint32_t m = x | 0x64006400u;
          const __half2 h = __hsub2(*reinterpret_cast<const __half2*>(&m), bias);
          w[p] = *reinterpret_cast<const uint32_t*>(&h);
        }
        flags |= (nz != 0u ? 1u : 0u) << c;
        if (c == 0) tc::mbar_wait(&bars->a_empty, (k & 1) ^ 1);       // the previous tile's MMAs have read A
#pragma unroll
        for (int j = 0; j < 4; ++j) tc::st8(tbase + lane_off + A_COL + c * 32 + 8 * j, &w[8 * j]);
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->idx_empty[s]);
      dcinfo[(k & 1) * TB + b] = dcw | (flags << 24);
      tc::wait_st();
      tc::fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->a_full);
    }
  } else {
    // ===== epilogue: thread = block = TMEM lane =====
    const int b = threadIdx.x;
    const uint32_t lane_off = uint32_t(warp * 32) << 16;
    const float qs = float(a.q) * float(1.0 / (1 << SCALE_LOG2));
    int k = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
      const int st = k & 1;
      tc::mbar_wait(&bars->d_full[st], (k >> 1) & 1);
      tc::fence_after();
      const uint32_t info = dcinfo[st * TB + b];
      const bool dconly = (info >> 24) == 0u;
      uint32_t ca = 0, cb = 0, cc = 0;
      if (dconly) {                        // the reference's float64 chain, exactly (dec8_dc.cuh)
        const unsigned rgb = dc_rgb(info & 0xffffffu, a.q);
        ca = __byte_perm(rgb, 0, 0x0210);
        cb = __byte_perm(rgb, 0, 0x1021);
        cc = __byte_perm(rgb, 0, 0x2102);
      }
      unsigned char* ob = smem + OFF_OUT + st * OUT_TILE + (b >> 6) * (OUT_TILE / 2) + (b & 63) * 24;
      const uint32_t dcol = tbase + lane_off + st * D_COLS;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        uint32_t y[8], co[8], cg[8];
        tc::ld8(dcol + 0 * 64 + r * 8, y);
        tc::ld8(dcol + 1 * 64 + r * 8, co);
        tc::ld8(dcol + 2 * 64 + r * 8, cg);
        tc::wait_ld();
        if (r == 0) tc::bar_sync(1, 128);  // thread 0 has waited for the store that last read this stage
        if (r == 7) {                      // D is in registers: hand the stage back to the MMA warp
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&bars->d_empty[st]);
        }
        int p[24];
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          const float Y = __uint_as_float(y[x]), Co = __uint_as_float(co[x]), Cg = __uint_as_float(cg[x]);
          const float t = Y - Cg;
          p[3 * x + 0] = __float2int_rz(fmaf(t + Co, qs, 128.0f));
          p[3 * x + 1] = __float2int_rz(fmaf(Y + Cg, qs, 128.0f));
          p[3 * x + 2] = __float2int_rz(fmaf(t - Co, qs, 128.0f));
        }
        uint32_t ww[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) ww[j] = pack_sat_u8(p[4 * j + 1], p[4 * j], pack_sat_u8(p[4 * j + 3], p[4 * j + 2], 0u));
        if (dconly) {
          ww[0] = ca; ww[1] = cb; ww[2] = cc; ww[3] = ca; ww[4] = cb; ww[5] = cc;
        }
        uint2* o = reinterpret_cast<uint2*>(ob + r * (64 * 24));
        o[0] = make_uint2(ww[0], ww[1]);
        o[1] = make_uint2(ww[2], ww[3]);
        o[2] = make_uint2(ww[4], ww[5]);
      }
      tma::fence_proxy_async();
      tc::bar_sync(1, 128);
      if (threadIdx.x == 0) {
        const int f = tile / per_frame, rem = tile - f * per_frame, by = rem / a.tiles_x, tx = rem - by * a.tiles_x;
        const unsigned char* src = smem + OFF_OUT + st * OUT_TILE;
        tma::store_3d(&out_map, src, (tx * TB) * 3, by * 8, f);
        if (tx * TB + 64 < a.nx) tma::store_3d(&out_map, src + OUT_TILE / 2, (tx * TB + 64) * 3, by * 8, f);
        tma::commit_group();
        tma::wait_group_read<1>();
      }
    }
    if (threadIdx.x == 0) tma::wait_group<0>();
  }

  tc::fence_before();
  __syncthreads();
  if (warp == 9) tc::tmem_dealloc<TMEM_COLS>(tbase);
}

// ---- host ---------------------------------------------------------------------------------

// 4096 * (orthonormal 8-point inverse DCT (x) itself), rows n = 8 r + x (sample), columns kk = 8 j + i
// (coefficient: j vertical, i horizontal frequency), as hi + lo fp16 limbs in the canonical K-major layout
std::vector<unsigned char> build_idct_limbs() {
  std::vector<unsigned char> t(NLIMB * LIMB_BYTES, 0);
  const double PI = 3.14159265358979323846;
  for (int n = 0; n < 64; ++n)
    for (int kk = 0; kk < 64; ++kk) {
      const int r = n / 8, x = n % 8, j = kk / 8, i = kk % 8;
      const double aj = j ? 0.5 : sqrt(0.125), ai = i ? 0.5 : sqrt(0.125);
      double v = double(1 << SCALE_LOG2) * aj * ai * cos((2 * r + 1) * j * PI / 16) * cos((2 * x + 1) * i * PI / 16);
      for (int l = 0; l < NLIMB; ++l) {
        const __half h = __float2half_rn(float(v));
        v -= double(__half2float(h));
        const unsigned short bits = *reinterpret_cast<const unsigned short*>(&h);
        *reinterpret_cast<unsigned short*>(&t[l * LIMB_BYTES + tc::off_kmajor16(n, kk, B_SBO, B_LBO)]) = bits;
      }
    }
  return t;
}

const unsigned char* device_idct_limbs() {
  static std::mutex mu;
  static unsigned char* tab[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!tab[dev]) {
    const std::vector<unsigned char> h = build_idct_limbs();
    unsigned char* d = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&d), h.size()) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaFree(d);
      return nullptr;
    }
    tab[dev] = d;
  }
  return tab[dev];
}

}  // namespace

// B = 8, YCoCg, subband layout, float32 fast mode; same preconditions as the other fast decoders
int launch_decode_tc(const DecArgs& a, cudaStream_t s) {
  const Geom& g = a.g;
  if (g.W % 16 != 0 || g.nx % 16 != 0 || g.left != 0 || g.top != 0) return VCFB_E_UNSUPP;
  if ((reinterpret_cast<uintptr_t>(a.rgb) & 15) || (reinterpret_cast<uintptr_t>(a.idx) & 15)) return VCFB_E_UNSUPP;
  if (a.q_int < 1 || a.q_int > 255 || tma::encode_tiled_fn() == nullptr) return VCFB_E_UNSUPP;
  const unsigned char* btab = device_idct_limbs();
  if (!btab) return VCFB_E_UNSUPP;
  CUtensorMap in_map, out_map;
  {   // index planes sub[j*ny + y, i*nx + x, c] as (x words, i, y, j, frame) -> smem [j][i][384 bytes]
    const uint64_t si = uint64_t(g.nx) * 3, sy = uint64_t(g.Wp) * 3, sj = uint64_t(g.ny) * g.Wp * 3,
                   sf = uint64_t(g.Hp) * g.Wp * 3;
    const uint64_t dims[5] = {uint64_t(g.nx) * 3 / 4, 8, uint64_t(g.ny), 8, uint64_t(a.n_frames)};
    const uint64_t str[4] = {si, sy, sj, sf};
    const uint32_t box[5] = {TB * 3 / 4, 8, 1, 8, 1};
    if (!tma::make_map(&in_map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, const_cast<uint8_t*>(a.idx), dims, str, box))
      return VCFB_E_UNSUPP;
  }
  {   // RGB frames as (W*3/8 uint64, H, n); box = 64 blocks x 8 rows (two per tile)
    const uint64_t dims[3] = {uint64_t(g.W) * 3 / 8, uint64_t(g.H), uint64_t(a.n_frames)};
    const uint64_t str[2] = {uint64_t(g.W) * 3, uint64_t(g.H) * g.W * 3};
    const uint32_t box[3] = {64 * 24 / 8, 8, 1};
    if (!tma::make_map(&out_map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, a.rgb, dims, str, box)) return VCFB_E_UNSUPP;
  }
  TcDecArgs ta;
  ta.tiles_x = (g.nx + TB - 1) / TB;
  ta.ny = g.ny;
  ta.nx = g.nx;
  const long long nt = (long long)a.n_frames * g.ny * ta.tiles_x;
  if (nt > 0x7fffffffLL - (1 << 20)) return VCFB_E_UNSUPP;
  ta.ntiles = int(nt);
  ta.q = a.q_int;
  ta.btab = btab;
  cudaError_t e = cudaFuncSetAttribute(dec8_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dec8_tc)");
  int grid = sm_count();
  if (grid > ta.ntiles) grid = ta.ntiles;
  note_kernel("dec8_tc");
  dec8_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, s>>>(in_map, out_map, ta);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "dec8_tc_kernel launch");
  return VCFB_OK;
}

}  // namespace vcfb
