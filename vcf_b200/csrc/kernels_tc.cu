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
// What is left on the CUDA cores is byte shuffling: index bytes -> colour mix of the indices -> fp16
// (converter warps) and * q, + 128, truncate, clip, pack (epilogue warps).  Blocks without AC indices are
// evaluated exactly like everywhere else (dec8_dc.cuh).
//
// One persistent CTA per SM, 4 (NGE + NGC) + 3 warps (a warp reaches only the TMEM lanes 32 (warp % 4) ... + 31):
//   epilogue   NGE x 4 warps, thread = (block = TMEM lane, 8 / NGE pixel rows): tcgen05.ld 8 samples x 3
//              channels per pixel row, colour / scale / pack, 24 bytes per row into the RGB tile, TMA store
//   converter  NGC x 4 warps, thread = (block, 64 / NGC coefficients): index bytes -> packed fp16 pairs ->
//              tcgen05.st (A lives in tensor memory: shared memory carries only raw tiles and the matrix)
//   1 warp     TMA producer of index tiles (3-stage ring)
//   1 warp     MMA issuer (one thread), owner of the TMEM allocation
// Pipelines: idx_full/idx_empty (TMA <-> converter), a_full/a_empty (converter <-> MMA, A is single
// buffered), d_full/d_empty x 2 (MMA <-> epilogue, D is double buffered).
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>

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
constexpr int NGC_MAX = 4;                 // warp groups per role are template parameters (NGE epilogue, NGC converter)
constexpr int LIMB_BYTES = 64 * 64 * 2;
constexpr int SCALE_LOG2 = 12;             // matrix entries are stored times 4096
constexpr int B_SBO = 128, B_LBO = 1024;   // canonical K-major, no swizzle: 8 x 16-byte rows per core matrix

constexpr int OFF_IDX = 0;
constexpr int OFF_OUT = OFF_IDX + NSI * IDX_TILE;
constexpr int OFF_B = OFF_OUT + NSO * OUT_TILE;
constexpr int OFF_DC = OFF_B + NLIMB * LIMB_BYTES;
constexpr int NDC = 4;                     // dcinfo ring: the converter runs up to three tiles ahead of the epilogue
constexpr int OFF_BAR = OFF_DC + NDC * NGC_MAX * TB * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256;
static_assert(sizeof(uint64_t) * (2 * NSI + 6 + 8) + 4 <= 256, "barrier block");

constexpr int D_COLS = 192;                // 3 channels x 64 samples per stage
constexpr int A_COL = 2 * D_COLS;          // 3 channels x 32 packed columns behind the two D stages
constexpr int TMEM_COLS = 512;

struct TcDecArgs {
  int ntiles, tiles_x, ny, nx;
  int q;
  const unsigned char* btab;               // NLIMB * LIMB_BYTES, already in the canonical layout
  long long* prof;                         // development: per-role cycle counters of CTA 0 (VCFB_TC_PROF)
};

struct Bars {
  uint64_t idx_full[NSI], idx_empty[NSI], a_full[3], a_empty[3], d_full[2], d_empty[2], out_full[2], out_free[2];
  uint32_t tmem_base;
};

// Development instrumentation (-DVCFB_TC_PROFILE): cycles each role spends in its waits, CTA 0
#ifdef VCFB_TC_PROFILE
#define TCP_WAIT(slot, stmt) do { if (a.prof) { const long long t0_ = clock64(); stmt; if (lane == 0) prof_acc[slot] += clock64() - t0_; } else { stmt; } } while (0)
#define TCP_ON(...) __VA_ARGS__
#else
#define TCP_WAIT(slot, stmt) do { stmt; } while (0)
#define TCP_ON(...)
#endif

__device__ __forceinline__ unsigned pack_sat_u8(int a, int b, unsigned c) {
  unsigned d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <int NGE, int NGC>
__global__ void __launch_bounds__((NGE + NGC) * 128 + 96, 1)
dec8_tc_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
               const TcDecArgs a) {
  constexpr int EPI_THREADS = NGE * 128, CONV_THREADS = NGC * 128, NTHREADS = EPI_THREADS + CONV_THREADS + 96;
  constexpr int W_CONV = EPI_THREADS / 32, W_TMA = (EPI_THREADS + CONV_THREADS) / 32, W_MMA = W_TMA + 1, W_ST = W_TMA + 2;
  static_assert(NGC <= NGC_MAX && 32 % NGC == 0 && (32 / NGC) % 8 == 0 && 8 % NGE == 0, "group counts");
  extern __shared__ __align__(128) unsigned char smem[];
  Bars* bars = reinterpret_cast<Bars*>(smem + OFF_BAR);
  uint32_t* dcinfo = reinterpret_cast<uint32_t*>(smem + OFF_DC);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    tma::prefetch_map(&in_map);
    tma::prefetch_map(&out_map);
    for (int s = 0; s < NSI; ++s) {
      tma::mbar_init(&bars->idx_full[s], 1);
      tma::mbar_init(&bars->idx_empty[s], 4 * NGC);
    }
    for (int c = 0; c < 3; ++c) {
      tma::mbar_init(&bars->a_full[c], 4 * NGC);
      tma::mbar_init(&bars->a_empty[c], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tma::mbar_init(&bars->d_full[s], 1);
      tma::mbar_init(&bars->d_empty[s], 4 * NGE);
      tma::mbar_init(&bars->out_full[s], 4 * NGE);
      tma::mbar_init(&bars->out_free[s], 1);
    }
    tma::fence_mbar_init();
  }
  {  // the matrix limbs, as prepared by the host
    const uint4* src = reinterpret_cast<const uint4*>(a.btab);
    uint4* dst = reinterpret_cast<uint4*>(smem + OFF_B);
    for (int i = threadIdx.x; i < NLIMB * LIMB_BYTES / 16; i += NTHREADS) dst[i] = __ldg(src + i);
  }
  tma::fence_proxy_async();               // generic writes of the matrix -> async proxy (tensor core)
  if (warp == W_MMA) tc::tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(&bars->tmem_base);
  const int per_frame = a.ny * a.tiles_x;
  TCP_ON(long long prof_acc[4] = {0, 0, 0, 0}; const long long t_start = clock64();)

  if (warp == W_TMA) {
    // ===== TMA producer =====
    if (lane == 0) {
      int k = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
        const int s = k % NSI;
        TCP_WAIT(0, tc::mbar_wait_sleep(&bars->idx_empty[s], ((k / NSI) & 1) ^ 1, 200));
        const int f = tile / per_frame, rem = tile - f * per_frame, by = rem / a.tiles_x, tx = rem - by * a.tiles_x;
        tma::mbar_expect_tx(&bars->idx_full[s], IDX_TILE);
        tma::load_5d(smem + OFF_IDX + s * IDX_TILE, &in_map, &bars->idx_full[s], tx * (TB * 3 / 4), 0, by, 0, f);
      }
    }
  } else if (warp == W_ST) {
    // ===== TMA store of finished RGB tiles =====
    if (lane == 0) {
      int k = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
        const int st = k & 1;
        TCP_WAIT(0, tc::mbar_wait_sleep(&bars->out_full[st], (k >> 1) & 1, 100));
        const int f = tile / per_frame, rem = tile - f * per_frame, by = rem / a.tiles_x, tx = rem - by * a.tiles_x;
        const unsigned char* src = smem + OFF_OUT + st * OUT_TILE;
        tma::store_3d(&out_map, src, (tx * TB) * 3, by * 8, f);
        if (tx * TB + 64 < a.nx) tma::store_3d(&out_map, src + OUT_TILE / 2, (tx * TB + 64) * 3, by * 8, f);
        tma::commit_group();
        if (k >= 1) {                      // the previous tile's stage has been read: the epilogue may refill it
          TCP_WAIT(1, tma::wait_group_read<1>());
          tc::mbar_arrive(&bars->out_free[st ^ 1]);
        }
      }
      tma::wait_group<0>();
    }
  } else if (warp == W_MMA) {
    // ===== MMA issuer: the whole warp walks the loop (uniform control flow, operands in uniform
    // registers), one elected lane issues =====
    {
      constexpr uint32_t IDESC = tc::idesc_f16(TB, 64);
      const uint32_t b_base = tma::smem_u32(smem + OFF_B);
      const uint32_t tb_u = __shfl_sync(0xffffffffu, tbase, 0);
      int k = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
        const int st = k & 1;
        TCP_WAIT(0, tc::mbar_wait(&bars->d_empty[st], ((k >> 1) & 1) ^ 1));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          TCP_WAIT(1, tc::mbar_wait(&bars->a_full[c], k & 1));
          tc::fence_after();
          if (tc::elect_one()) {
#pragma unroll
            for (int l = 0; l < NLIMB; ++l)
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                tc::mma_ts(tb_u + st * D_COLS + c * 64, tb_u + A_COL + c * 32 + ks * 8,
                           tc::smem_desc(b_base + l * LIMB_BYTES + ks * 2 * B_LBO, B_LBO, B_SBO), IDESC, (l | ks) != 0);
            tc::commit(&bars->a_empty[c]);   // channel c of A may be overwritten once these have completed
            if (c == 2) tc::commit(&bars->d_full[st]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= W_CONV) {
    // ===== converter: thread = (block, 64 / NG coefficients) =====
    const int b = threadIdx.x & 127, grp = (warp - W_CONV) >> 2;
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const __half2 bias_rb = __floats2half2_rn(1408.0f, 1408.0f), bias_g = __floats2half2_rn(1280.0f, 1280.0f);
    constexpr int NP = 32 / NGC;                                    // packed pairs per channel and thread
    const int woff = (3 * b) & ~3, sh = ((3 * b) & 3) * 8;          // the block's 3 bytes start inside this word
    int k = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
      const int s = k % NSI;
      TCP_WAIT(0, tc::mbar_wait(&bars->idx_full[s], (k / NSI) & 1));
      TCP_ON(const long long tc0 = a.prof ? clock64() : 0;)
      const unsigned char* tp = smem + OFF_IDX + s * IDX_TILE + grp * (2 * NP) * (TB * 3) + woff;
      uint32_t w[3][NP];
      uint32_t nz = 0, dcw = 0;
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const uint32_t* ra = reinterpret_cast<const uint32_t*>(tp + (2 * p) * (TB * 3));
        const uint32_t* rb = reinterpret_cast<const uint32_t*>(tp + (2 * p + 1) * (TB * 3));
        const uint32_t xa = __funnelshift_r(ra[0], ra[1], sh);      // bytes: index of Y, Co, Cg, (junk)
        const uint32_t xb = __funnelshift_r(rb[0], rb[1], sh);
        if (p == 0 && grp == 0) dcw = xa & 0x00ffffffu;             // the DC index does not count as "AC present"
        else nz |= xa ^ 0x00808080u;
        nz |= xb ^ 0x00808080u;
        // to_RGB on the integer INDICES (the transform is linear and q uniform): one dp4a per plane on the
        // bytes (Y, Co, Cg, junk) whose accumulator starts at 0x6400 + offset, so the result is the bit
        // pattern of the fp16 number 1024 + offset' + (index of the plane); after packing two of them, one
        // packed subtraction leaves the plane's indices, exact integers of at most 9 bits
        constexpr int MIX[3] = {0x00FF0101, 0x00010001, 0x00FFFF01};   // R = Y + Co - Cg, G = Y + Cg, B = Y - Co - Cg
        constexpr int MIXB[3] = {0x6400 + 256, 0x6400, 0x6400 + 512};  // bytes are index + 128
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const uint32_t m = __byte_perm(uint32_t(dp4a_us(xa, MIX[c], MIXB[c])), uint32_t(dp4a_us(xb, MIX[c], MIXB[c])), 0x5410);
          const __half2 h = __hsub2(*reinterpret_cast<const __half2*>(&m), c == 1 ? bias_g : bias_rb);
          w[c][p] = *reinterpret_cast<const uint32_t*>(&h);
        }
      }
      __syncwarp();
      TCP_ON(if (a.prof && lane == 0) prof_acc[2] += clock64() - tc0;)
      if (lane == 0) tc::mbar_arrive(&bars->idx_empty[s]);
      const uint32_t flags = ((nz & 0xffu) ? 1u : 0u) | ((nz & 0xff00u) ? 2u : 0u) | ((nz & 0xff0000u) ? 4u : 0u);
      dcinfo[((k % NDC) * NGC + grp) * TB + b] = dcw | (flags << 24);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        TCP_WAIT(1, tc::mbar_wait(&bars->a_empty[c], (k & 1) ^ 1));      // the previous tile's MMAs have read channel c of A
#pragma unroll
        TCP_ON(const long long ts0 = a.prof ? clock64() : 0;)
        for (int j = 0; j < NP / 8; ++j) tc::st8(tbase + lane_off + A_COL + c * 32 + grp * NP + 8 * j, &w[c][8 * j]);
        tc::wait_st();
        TCP_ON(if (a.prof && lane == 0) prof_acc[3] += clock64() - ts0;)
        tc::fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars->a_full[c]);
      }
    }
  } else {
    // ===== epilogue: thread = (block = TMEM lane, half of the 8 pixel rows) =====
    const int b = threadIdx.x & 127, grp = warp >> 2;
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const float qs = float(a.q) * float(1.0 / (1 << SCALE_LOG2));
    constexpr int NR = 8 / NGE;
    int k = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
      const int st = k & 1;
      TCP_WAIT(0, tc::mbar_wait(&bars->d_full[st], (k >> 1) & 1));
      tc::fence_after();
      uint32_t info = dcinfo[((k % NDC) * NGC) * TB + b];
#pragma unroll
      for (int g2 = 1; g2 < NGC; ++g2) info |= dcinfo[((k % NDC) * NGC + g2) * TB + b] & 0xff000000u;
      const bool dconly = (info >> 24) == 0u;
      uint32_t ca = 0, cb = 0, cc = 0;
      if (dconly) {                        // the reference's float64 chain, exactly (dec8_dc.cuh)
        const unsigned rgb = dc_rgb(info & 0xffffffu, a.q);
        ca = __byte_perm(rgb, 0, 0x0210);
        cb = __byte_perm(rgb, 0, 0x1021);
        cc = __byte_perm(rgb, 0, 0x2102);
      }
      unsigned char* ob = smem + OFF_OUT + st * OUT_TILE + (b >> 6) * (OUT_TILE / 2) + (b & 63) * 24;
      const uint32_t dcol = tbase + lane_off + st * D_COLS + grp * NR * 8;
      uint32_t d[2][3][8];
      tc::ld8(dcol + 0 * 64, d[0][0]);
      tc::ld8(dcol + 1 * 64, d[0][1]);
      tc::ld8(dcol + 2 * 64, d[0][2]);
      TCP_WAIT(2, tc::wait_ld());
      TCP_WAIT(1, tc::mbar_wait(&bars->out_free[st], ((k >> 1) & 1) ^ 1));   // the store that last read this stage is done
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        const int cur = rr & 1;
        if (rr + 1 < NR) {                 // next row's samples travel while this row is computed
          tc::ld8(dcol + 0 * 64 + (rr + 1) * 8, d[cur ^ 1][0]);
          tc::ld8(dcol + 1 * 64 + (rr + 1) * 8, d[cur ^ 1][1]);
          tc::ld8(dcol + 2 * 64 + (rr + 1) * 8, d[cur ^ 1][2]);
        } else {                           // D is in registers: hand the stage back to the MMA warp
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&bars->d_empty[st]);
        }
        int p[24];
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          p[3 * x + 0] = __float2int_rz(fmaf(__uint_as_float(d[cur][0][x]), qs, 128.0f));
          p[3 * x + 1] = __float2int_rz(fmaf(__uint_as_float(d[cur][1][x]), qs, 128.0f));
          p[3 * x + 2] = __float2int_rz(fmaf(__uint_as_float(d[cur][2][x]), qs, 128.0f));
        }
        uint32_t ww[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) ww[j] = pack_sat_u8(p[4 * j + 1], p[4 * j], pack_sat_u8(p[4 * j + 3], p[4 * j + 2], 0u));
        if (dconly) {
          ww[0] = ca; ww[1] = cb; ww[2] = cc; ww[3] = ca; ww[4] = cb; ww[5] = cc;
        }
        uint2* o = reinterpret_cast<uint2*>(ob + (grp * NR + rr) * (64 * 24));
        o[0] = make_uint2(ww[0], ww[1]);
        o[1] = make_uint2(ww[2], ww[3]);
        o[2] = make_uint2(ww[4], ww[5]);
        if (rr + 1 < NR) TCP_WAIT(2, tc::wait_ld());
      }
      tma::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->out_full[st]);
    }
  }

#ifdef VCFB_TC_PROFILE
  if (a.prof && blockIdx.x == 0 && lane == 0 && (warp & 3) == 0) {
    // role id: 0 epilogue g0, 1 epilogue g1, 2 conv g0, 3 conv g1, 4 = TMA/MMA warps (W_TMA % 4 == 0 only)
    long long* o = a.prof + (warp >> 2) * 8;
    o[0] = clock64() - t_start;
    for (int i = 0; i < 4; ++i) o[1 + i] = prof_acc[i];
  }
  if (a.prof && blockIdx.x == 0 && (warp == W_MMA || warp == W_ST) && lane == 0) {
    long long* o = a.prof + (warp == W_MMA ? 9 : 8) * 8;
    o[0] = clock64() - t_start;
    for (int i = 0; i < 4; ++i) o[1 + i] = prof_acc[i];
  }
#endif
  tc::fence_before();
  __syncthreads();
  if (warp == W_MMA) tc::tmem_dealloc<TMEM_COLS>(tbase);
}

// =============================================================================================
// Encoder, tensor-core tier (the north star's FAST MODE: fewer than 1e-6 of the indices differ from the
// reference's float32 path, and only at rounding boundaries; the bit-exact encoder is kernels_packed.cu).
//   converter  thread = block: 8 x 24 RGB bytes -> (4Y, 2Co, 4Cg) by one dp4a each (exact integers, fp16)
//              -> tcgen05.st, one A item (32 columns) per channel, ring of 6 items (two tiles)
//   MMA        per item D[128 blocks x 80] = A[128 x 64 samples] * F^T: rows 0..63 of F = 4096 * (8-point
//              DCT (x) itself) in two fp16 limbs, rows 64..79 = the sums the four RATIONAL positions need
//              (SURVEY 7.3/7.4): column sums S0[x] = sum_r p[r][x] and S4[x] = sum_r (+-) p[r][x] with the
//              signs of pocketfft's output 4 -- entries 0 / +-1, so these 16 columns are exact integers
//   epilogue   thread = block: coefficient * 2^-k / q -> truncate -> + 128 -> byte into the index tile
//              [8u + v][3 block + channel]; the coefficients (0,0) (0,4) (4,0) (4,4) are recomputed from S0 /
//              S4 with pocketfft's own sequence of individually rounded float32 operations (the same
//              operations dct_codelets.cuh::dct8_fwd performs for its outputs 0 and 4), because those
//              coefficients are exact rationals that land ON quantisation boundaries and the index then
//              follows the reference's last-bit rounding; everywhere else a boundary is hit with
//              probability ~1e-7 per index
// D items: ring of 4 x 80 columns; A items: ring of 6 x 32 columns (512 TMEM columns in all).
namespace enc {

constexpr int RGB_TILE = 8 * TB * 24;      // two halves of [8 rows][64 blocks x 24 bytes] (two TMA boxes)
constexpr int IDXT = 64 * TB * 3;          // [kk = 8u + v][128 blocks x 3 bytes]
constexpr int NSIN = 3, NSOUT = 3;          // one output stage per epilogue group
constexpr int NF = 80;                     // rows of the hi limb: 64 coefficients + 8 + 8 column sums
constexpr int HI_BYTES = NF * 64 * 2, LO_BYTES = 64 * 64 * 2;
constexpr int HI_LBO = NF / 8 * 128, LO_LBO = 1024;
constexpr int ND = 4, NA = 6;
// An increase is served from what the CTA's own warpgroups have handed back (not from registers the launch left
// unused): at 80 registers per thread and 768 threads the converters (80 -> 64) and the service group (80 -> 24) free
// 4 096 + 7 168, the three epilogue groups (80 -> 104) take 9 216.  (72 / 40 freed 7 168 < 9 216: the epilogue spun
// in its allocation loop for ever.)
constexpr int REGS_EPI = 104, REGS_CONV = 64, REGS_SERVICE = 24;
constexpr int D_ITEM = NF, A_ITEM = 32;
constexpr int A_COL0 = ND * D_ITEM;        // 320

constexpr int OFF_IN = 0;
constexpr int OFF_OUTI = OFF_IN + NSIN * RGB_TILE;
constexpr int OFF_F = OFF_OUTI + NSOUT * IDXT;
constexpr int OFF_EBAR = OFF_F + HI_BYTES + LO_BYTES;
constexpr int ESMEM = OFF_EBAR + 512;

struct EBars {
  // d_full is per (epilogue group, channel): every barrier then has ONE waiter that sees each of its phases in turn.
  // (With a d_full per D slot, a third group starts in the middle of the item sequence and meets barriers that are two
  // phases away from the one it last saw -- a parity wait cannot tell those apart, and the pipeline dead-locked.)
  uint64_t in_full[NSIN], in_empty[NSIN], a_full[NA], a_empty[NA], d_full[3 * 3], d_empty[ND], out_full[NSOUT], out_free[NSOUT];
  uint32_t tmem_base;
};
static_assert(sizeof(EBars) <= 512, "barrier block");

struct TcEncArgs {
  int ntiles, tiles_x, ny, nx, top;
  float inv_q;                             // 1 / q, q a power of two
  const unsigned char* ftab;               // HI_BYTES + LO_BYTES, canonical layout
  long long* prof;
  unsigned long long* stats;               // STATS: VCFB_STAT_NONZERO / VCFB_STAT_SUMABS accumulated in the epilogue
};

__device__ __forceinline__ unsigned pack_sat_s8(int a, int b, unsigned c) {
  unsigned d;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// Registers follow the roles (setmaxnreg, one instruction per warpgroup): the three epilogue groups take what the
// converter groups and the service group (TMA producer, MMA issuer, store warp and a fourth, idle warp that only
// completes the warpgroup) hand back.
#ifdef VCFB_NO_SETMAXNREG
template <int N> __device__ __forceinline__ void reg_inc() {}
template <int N> __device__ __forceinline__ void reg_dec() {}
#else
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
#endif

template <int NGE, int NGC, bool STATS = false>
__global__ void __launch_bounds__((NGE + NGC) * 128 + 128, 1)
enc8_tc_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
               const TcEncArgs a) {
  constexpr int EPI_THREADS = NGE * 128, CONV_THREADS = NGC * 128, NTHREADS = EPI_THREADS + CONV_THREADS + 128;
  constexpr int W_CONV = EPI_THREADS / 32, W_TMA = (EPI_THREADS + CONV_THREADS) / 32, W_MMA = W_TMA + 1, W_ST = W_TMA + 2;
  static_assert(NGE >= 1 && NGE <= 3, "epilogue groups take the tiles in turn (tile k: group k % NGE, output stage k % NSOUT)");
  static_assert(NGC == 1 || NGC == 2, "converter groups split the pixel rows 0-3 / 4-7");
  extern __shared__ __align__(128) unsigned char smem[];
  EBars* bars = reinterpret_cast<EBars*>(smem + OFF_EBAR);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    tma::prefetch_map(&in_map);
    tma::prefetch_map(&out_map);
    for (int s = 0; s < NSIN; ++s) {
      tma::mbar_init(&bars->in_full[s], 1);
      tma::mbar_init(&bars->in_empty[s], 4 * NGC);
    }
    for (int s = 0; s < NA; ++s) {
      tma::mbar_init(&bars->a_full[s], 4 * NGC);
      tma::mbar_init(&bars->a_empty[s], 1);
    }
    for (int s = 0; s < 9; ++s) tma::mbar_init(&bars->d_full[s], 1);
    for (int s = 0; s < ND; ++s) tma::mbar_init(&bars->d_empty[s], 4);
    for (int s = 0; s < NSOUT; ++s) {
      tma::mbar_init(&bars->out_full[s], 4);
      tma::mbar_init(&bars->out_free[s], 1);
    }
    tma::fence_mbar_init();
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.ftab);
    uint4* dst = reinterpret_cast<uint4*>(smem + OFF_F);
    for (int i = threadIdx.x; i < (HI_BYTES + LO_BYTES) / 16; i += NTHREADS) dst[i] = __ldg(src + i);
  }
  tma::fence_proxy_async();
  if (warp == W_MMA) tc::tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tbase = *reinterpret_cast<volatile uint32_t*>(&bars->tmem_base);
  const int per_frame = a.ny * a.tiles_x;

  if (warp >= W_TMA) {
  reg_dec<REGS_SERVICE>();   // the service warpgroup: TMA producer, store warp, MMA issuer, one idle warp
  if (warp == W_TMA) {
    // ===== TMA producer of RGB tiles (zero padding = out-of-bounds fill; rows may start above the frame) =====
    if (lane == 0) {
      int k = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
        const int s = k % NSIN;
        tc::mbar_wait_sleep(&bars->in_empty[s], ((k / NSIN) & 1) ^ 1, 200);
        const int f = tile / per_frame, rem = tile - f * per_frame, by = rem / a.tiles_x, tx = rem - by * a.tiles_x;
        unsigned char* dst = smem + OFF_IN + s * RGB_TILE;
        tma::mbar_expect_tx(&bars->in_full[s], RGB_TILE);
        tma::load_3d(dst, &in_map, &bars->in_full[s], (tx * TB) * 3, by * 8 - a.top, f);
        tma::load_3d(dst + RGB_TILE / 2, &in_map, &bars->in_full[s], (tx * TB + 64) * 3, by * 8 - a.top, f);
      }
    }
  } else if (warp == W_ST) {
    // ===== TMA store of finished index tiles =====
    if (lane == 0) {
      int k = 0;
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
        const int st = k % NSOUT;
        tc::mbar_wait_sleep(&bars->out_full[st], (k / NSOUT) & 1, 100);
        const int f = tile / per_frame, rem = tile - f * per_frame, by = rem / a.tiles_x, tx = rem - by * a.tiles_x;
        tma::store_5d(&out_map, smem + OFF_OUTI + st * IDXT, tx * (TB * 3 / 4), 0, by, 0, f);
        tma::commit_group();
        // hand the stage back as soon as THIS store has read it (a few hundred cycles): the epilogue groups own one
        // output stage each, and waiting for the next tile's store would chain their store phases to each other
        tma::wait_group_read<0>();
        tc::mbar_arrive(&bars->out_free[st]);
      }
      tma::wait_group<0>();
    }
  } else if (warp == W_MMA) {
    // ===== MMA issuer =====
    {
      constexpr uint32_t IDESC_HI = tc::idesc_f16(TB, NF), IDESC_LO = tc::idesc_f16(TB, 64);
      const uint32_t f_base = tma::smem_u32(smem + OFF_F);
      const uint32_t tb_u = __shfl_sync(0xffffffffu, tbase, 0);
      TCP_ON(long long eprof[2] = {0, 0}; const long long te0 = clock64();)
      int it = 0, kt = 0;                    // item = (tile, channel); kt = tiles done
      for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++kt) {
#pragma unroll
        for (int c = 0; c < 3; ++c, ++it) {
          const int sa = it % NA, sd = it % ND;
          TCP_ON(const long long tw0 = clock64();)
          tc::mbar_wait(&bars->d_empty[sd], ((it / ND) & 1) ^ 1);
          TCP_ON(const long long tw1 = clock64();)
          tc::mbar_wait(&bars->a_full[sa], (it / NA) & 1);
          TCP_ON(const long long tw2 = clock64(); eprof[0] += tw1 - tw0; eprof[1] += tw2 - tw1;)
          tc::fence_after();
          if (tc::elect_one()) {
            const uint32_t d_t = tb_u + sd * D_ITEM, a_t = tb_u + A_COL0 + sa * A_ITEM;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc::mma_ts(d_t, a_t + ks * 8, tc::smem_desc(f_base + ks * 2 * HI_LBO, HI_LBO, 128), IDESC_HI, ks != 0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc::mma_ts(d_t, a_t + ks * 8, tc::smem_desc(f_base + HI_BYTES + ks * 2 * LO_LBO, LO_LBO, 128), IDESC_LO, 1u);
            tc::commit(&bars->a_empty[sa]);
            tc::commit(&bars->d_full[(kt % NGE) * 3 + c]);
          }
          __syncwarp();
        }
      }
      TCP_ON(if (a.prof && blockIdx.x == 0 && lane == 0) { a.prof[0] = clock64() - te0; a.prof[1] = eprof[0]; a.prof[2] = eprof[1]; })
    }
  }
  } else if (warp >= W_CONV) {
    reg_dec<REGS_CONV>();
    // ===== converter: thread = (block, 8 / NGC pixel rows) =====
    const int b = threadIdx.x & 127, grp = (warp - W_CONV) >> 2;
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    constexpr int NRW = 8 / NGC;             // pixel rows per thread -> 4 * NRW packed columns per channel
    // half2 biases: the dp4a accumulators start at 0x6400 (+ an offset that keeps the sum positive)
    const __half2 sub_y = __floats2half2_rn(1536.0f, 1536.0f);      // 1024 + 512      : R + 2G + B - 512 = 4Y
    const __half2 sub_co = __floats2half2_rn(1280.0f, 1280.0f);     // 1024 + 256      : R - B = 2Co
    const __half2 sub_cg = __floats2half2_rn(1536.0f, 1536.0f);     // 1024 + 512      : -R + 2G - B = 4Cg
    TCP_ON(long long cw[3] = {0, 0, 0}; const long long tcv0 = clock64();)
    int k = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
      const int s = k % NSIN;
      TCP_ON(const long long tv0 = clock64();)
      tc::mbar_wait(&bars->in_full[s], (k / NSIN) & 1);
      TCP_ON(const long long tv1 = clock64(); cw[0] += tv1 - tv0;)
      const unsigned char* tp = smem + OFF_IN + s * RGB_TILE + (b >> 6) * (RGB_TILE / 2) + (b & 63) * 24 + grp * NRW * (64 * 24);
      uint32_t w[3][4 * NRW];
#pragma unroll
      for (int r = 0; r < NRW; ++r) {
        const uint2* rp = reinterpret_cast<const uint2*>(tp + r * (64 * 24));
        const uint2 q0 = rp[0], q1 = rp[1], q2 = rp[2];
        const uint32_t wd[6] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y};
#pragma unroll
        for (int xp = 0; xp < 4; ++xp) {       // pixels 2 xp and 2 xp + 1: bytes 6 xp .. 6 xp + 5 of the 24
          const int o0 = 6 * xp, o1 = 6 * xp + 3;
          const uint32_t pa = __funnelshift_r(wd[o0 >> 2], wd[(o0 >> 2) + ((o0 & 3) ? 1 : 0)], (o0 & 3) * 8);
          const uint32_t pb = __funnelshift_r(wd[o1 >> 2], wd[(o1 >> 2) + ((o1 & 3) > 1 ? 1 : 0)], (o1 & 3) * 8);
          // (R, G, B, junk) . (1, 2, 1, 0) etc.; the result is the bit pattern of the fp16 number 1024 + sum
          const uint32_t my = __byte_perm(uint32_t(dp4a_us(pa, 0x00010201, 0x6400)), uint32_t(dp4a_us(pb, 0x00010201, 0x6400)), 0x5410);
          const uint32_t mo = __byte_perm(uint32_t(dp4a_us(pa, 0x00FF0001, 0x6400 + 256)), uint32_t(dp4a_us(pb, 0x00FF0001, 0x6400 + 256)), 0x5410);
          const uint32_t mg = __byte_perm(uint32_t(dp4a_us(pa, 0x00FF02FF, 0x6400 + 512)), uint32_t(dp4a_us(pb, 0x00FF02FF, 0x6400 + 512)), 0x5410);
          const __half2 hy = __hsub2(*reinterpret_cast<const __half2*>(&my), sub_y);
          const __half2 ho = __hsub2(*reinterpret_cast<const __half2*>(&mo), sub_co);
          const __half2 hg = __hsub2(*reinterpret_cast<const __half2*>(&mg), sub_cg);
          w[0][4 * r + xp] = *reinterpret_cast<const uint32_t*>(&hy);
          w[1][4 * r + xp] = *reinterpret_cast<const uint32_t*>(&ho);
          w[2][4 * r + xp] = *reinterpret_cast<const uint32_t*>(&hg);
        }
      }
      __syncwarp();
      TCP_ON(cw[1] += clock64() - tv1;)
      if (lane == 0) tc::mbar_arrive(&bars->in_empty[s]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int it = 3 * k + c, sa = it % NA;
        TCP_ON(const long long tv2 = clock64();)
        tc::mbar_wait(&bars->a_empty[sa], ((it / NA) & 1) ^ 1);
        TCP_ON(cw[2] += clock64() - tv2;)
#pragma unroll
        for (int j = 0; j < NRW / 2; ++j) tc::st8(tbase + lane_off + A_COL0 + sa * A_ITEM + grp * 4 * NRW + 8 * j, &w[c][8 * j]);
        tc::wait_st();
        tc::fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars->a_full[sa]);
      }
    }
    TCP_ON(if (a.prof && blockIdx.x == 0 && threadIdx.x == W_CONV * 32) { a.prof[8] = clock64() - tcv0; a.prof[9] = cw[0]; a.prof[10] = cw[1]; a.prof[11] = cw[2]; })
  } else {
    reg_inc<REGS_EPI>();
    // ===== epilogue: thread = block = TMEM lane; the NGE warp groups take the tiles in turn, so the latency of a
    // tile's chain (barrier -> tensor-memory load -> quantise -> shuffle -> store) is hidden behind the other group =====
    const int b = threadIdx.x & 127, grp = warp >> 2;
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    constexpr float SQ2 = 0x1.6a09e6p+0f, SQ2H = 0x1.6a09e6p-1f;     // pocketfft's float32 constants (dct8_fwd: t67, t66)
    // the index tile is written as aligned words: the 4 blocks of a lane quad own 12 consecutive bytes of a row
    const int qj = lane & 3;                             // lane j < 3 of the quad writes word j
    const uint32_t wsel = qj == 0 ? 0x4210u : (qj == 1 ? 0x5421u : 0x6542u);
    TCP_ON(long long pw[3] = {0, 0, 0}; const long long tp0 = clock64();)
    unsigned st_nz = 0, st_abs = 0;                     // STATS: non-zero indices and sum |index| of this thread's blocks
    int k = 0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++k) {
      if (k % NGE != grp) continue;
      const int st = k % NSOUT, ku = k / NGE;            // ku: how many tiles this group has done
      uint32_t hold[3][2][4][2];                         // [channel][half][row][4 index bytes each]
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int it = 3 * k + c, sd = it % ND;
        // coefficient = D / (4096 * colour scale): the converter fed 4Y, 2Co, 4Cg
        const float sc = a.inv_q * (c == 1 ? 0x1p-13f : 0x1p-14f);
        TCP_ON(const long long tq0 = clock64();)
        tc::mbar_wait(&bars->d_full[grp * 3 + c], ku & 1);
        TCP_ON(pw[0] += clock64() - tq0;)
        tc::fence_after();
        const uint32_t dcol = tbase + lane_off + sd * D_ITEM;
        uint32_t sv[16];
        uint32_t dv[32];
        tc::ld16(dcol + 64, sv);                          // the sums of the rational positions travel with the first half
        tc::ld32(dcol, dv);
        tc::wait_ld();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int uu = 4 * h;                           // first row of the half: 0 or 4, the rows with rational positions
          // ---- the rational coefficients (uu, 0) and (uu, 4), exactly as dct_codelets.cuh::dct8_fwd rounds them ----
          float r0, r4;
          {
            float p[8];
            const float cm = uu == 0 ? SQ2 : SQ2H;          // pass 1 (over the rows): output 0 * sqrt(2), output 4 * sqrt(1/2)
#pragma unroll
            for (int x = 0; x < 8; ++x) p[x] = __fmul_rn(__uint_as_float(sv[8 * h + x]), cm);
            // pass 2 (over x): dct8_fwd's own additions for its outputs 0 and 4
            const float t9 = __fadd_rn(p[1], p[2]), t13 = __fadd_rn(p[5], p[6]), t11 = __fadd_rn(p[3], p[4]), t14 = __fadd_rn(p[0], p[7]);
            const float t16 = __fadd_rn(t9, t13), t27 = __fadd_rn(t11, t14);
            const float t29 = __fadd_rn(t16, t27), t30 = __fsub_rn(t27, t16);
            // lazy powers of two: output 0 carries 2^-2, output 4 carries 2^-1, per pass; colour scale; 1 / q
            const float e_u = (uu == 0 ? 0.25f : 0.5f) * a.inv_q * (c == 1 ? 0.5f : 0.25f);
            r0 = __fmul_rn(t29, SQ2) * (e_u * 0.25f);
            r4 = __fmul_rn(t30, SQ2H) * (e_u * 0.5f);
          }
#pragma unroll
          for (int ur = 0; ur < 4; ++ur) {
            float v[8];
#pragma unroll
            for (int x = 0; x < 8; ++x) v[x] = __uint_as_float(dv[8 * ur + x]) * sc;
            if (ur == 0) {
              v[0] = r0;
              v[4] = r4;
            }
            // truncate toward zero, saturating pack to int8 (|index| <= 128 here), + 128 as a flip of the top bit (the
            // reference's astype(uint8) wrap and this coincide for indices in [-128, 127])
            hold[c][h][ur][0] = pack_sat_s8(__float2int_rz(v[1]), __float2int_rz(v[0]),
                                            pack_sat_s8(__float2int_rz(v[3]), __float2int_rz(v[2]), 0u)) ^ 0x80808080u;
            hold[c][h][ur][1] = pack_sat_s8(__float2int_rz(v[5]), __float2int_rz(v[4]),
                                            pack_sat_s8(__float2int_rz(v[7]), __float2int_rz(v[6]), 0u)) ^ 0x80808080u;
          }
          if (h == 0) {
            tc::ld32(dcol + 32, dv);
            tc::wait_ld();
            tc::fence_before();              // D is in registers: hand the item back to the MMA warp
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&bars->d_empty[sd]);
          }
        }
      }
      // ---- the three channels of the block are in registers: write rows [8u + v][3 block + channel] as words ----
      TCP_ON(const long long tq1 = clock64();)
      tc::mbar_wait(&bars->out_free[st], ((k / NSOUT) & 1) ^ 1);    // the store that last read this stage is done
      TCP_ON(const long long tq2 = clock64(); pw[1] += tq2 - tq1;)
      uint32_t* ow = reinterpret_cast<uint32_t*>(smem + OFF_OUTI + st * IDXT) + 3 * (b >> 2) + qj;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int ur = 0; ur < 4; ++ur)
#pragma unroll
          for (int x = 0; x < 8; ++x) {
            const int u = 4 * h + ur;
            // this block's (Y, Co, Cg) bytes of coefficient (u, x)
            const uint32_t yc = __byte_perm(hold[0][h][ur][x >> 2], hold[1][h][ur][x >> 2], 0x0040 + 0x0011 * (x & 3));
            const uint32_t mine = __byte_perm(yc, hold[2][h][ur][x >> 2], 0x0010 + 0x0400 + 0x0100 * (x & 3));
            const uint32_t next = __shfl_down_sync(0xffffffffu, mine, 1);
            if (qj < 3) ow[(8 * u + x) * (TB * 3 / 4)] = __byte_perm(mine, next, wsel);
          }
      tma::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->out_full[st]);
      TCP_ON(pw[2] += clock64() - tq2;)
      if (STATS) {
        // the index bytes of this block are still in registers (byte = index + 128): |index| by a byte-wise absolute
        // difference, the count of non-zero ones by a byte-wise compare, both summed with dp4a.  Blocks beyond the
        // right edge of the frame (partial last tile: computed from zero fill, never stored) do not count.
        const int tx = tile % a.tiles_x;
        if (tx * TB + b < a.nx) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int ur = 0; ur < 4; ++ur)
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                  const uint32_t v = hold[c][h][ur][w];
                  st_abs = __dp4a(__vabsdiffu4(v, 0x80808080u), 0x01010101u, st_abs);
                  st_nz = __dp4a(__vsetne4(v, 0x80808080u), 0x01010101u, st_nz);
                }
        }
      }
      (void)ku;
    }
    if (STATS) {
      const unsigned nzw = __reduce_add_sync(0xffffffffu, st_nz), absw = __reduce_add_sync(0xffffffffu, st_abs);
      if (lane == 0) {
        if (nzw) atomicAdd(a.stats + VCFB_STAT_NONZERO, (unsigned long long)nzw);
        if (absw) atomicAdd(a.stats + VCFB_STAT_SUMABS, (unsigned long long)absw);
      }
    }
    TCP_ON(if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) { a.prof[4] = clock64() - tp0; a.prof[5] = pw[0]; a.prof[6] = pw[1]; a.prof[7] = pw[2]; })
  }

  tc::fence_before();
  __syncthreads();
  if (warp == W_MMA) tc::tmem_dealloc<TMEM_COLS>(tbase);
}

}  // namespace enc

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

// forward tables: hi limb 80 rows (64 coefficients kk = 8u + v of 4096 * DCT (x) DCT over the samples n = 8r + x,
// then S0[x] = sum_r p[r][x] and S4[x] = (p0 + p7 + p3 + p4) - (p1 + p2 + p5 + p6) of column x), lo limb 64 rows
std::vector<unsigned char> build_fdct_limbs() {
  std::vector<unsigned char> t(enc::HI_BYTES + enc::LO_BYTES, 0);
  const double PI = 3.14159265358979323846;
  auto put = [&](int off, float v) {
    const __half h = __float2half_rn(v);
    *reinterpret_cast<unsigned short*>(&t[off]) = *reinterpret_cast<const unsigned short*>(&h);
  };
  for (int kk = 0; kk < 64; ++kk)
    for (int n = 0; n < 64; ++n) {
      const int u = kk / 8, v = kk % 8, r = n / 8, x = n % 8;
      const double au = u ? 0.5 : sqrt(0.125), av = v ? 0.5 : sqrt(0.125);
      const double val = double(1 << SCALE_LOG2) * au * av * cos((2 * r + 1) * u * PI / 16) * cos((2 * x + 1) * v * PI / 16);
      const __half h = __float2half_rn(float(val));
      put(tc::off_kmajor16(kk, n, 128, enc::HI_LBO), __half2float(h));
      put(enc::HI_BYTES + tc::off_kmajor16(kk, n, 128, enc::LO_LBO), float(val - double(__half2float(h))));
    }
  for (int x = 0; x < 8; ++x)
    for (int r = 0; r < 8; ++r) {
      const bool plus = (r == 0 || r == 7 || r == 3 || r == 4);
      put(tc::off_kmajor16(64 + x, 8 * r + x, 128, enc::HI_LBO), 1.0f);
      put(tc::off_kmajor16(72 + x, 8 * r + x, 128, enc::HI_LBO), plus ? 1.0f : -1.0f);
    }
  return t;
}

const unsigned char* device_table(int which) {
  static std::mutex mu;
  static unsigned char* tab[2][64] = {{nullptr}};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!tab[which][dev]) {
    const std::vector<unsigned char> h = which ? build_fdct_limbs() : build_idct_limbs();
    unsigned char* d = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&d), h.size()) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaFree(d);
      return nullptr;
    }
    tab[which][dev] = d;
  }
  return tab[which][dev];
}

}  // namespace

// B = 8, YCoCg, subband layout, float32 fast mode; same preconditions as the other fast decoders
int launch_decode_tc(const DecArgs& a, cudaStream_t s) {
  const Geom& g = a.g;
  if (g.W % 16 != 0 || g.nx % 16 != 0 || g.left != 0 || g.top != 0) return VCFB_E_UNSUPP;
  if ((reinterpret_cast<uintptr_t>(a.rgb) & 15) || (reinterpret_cast<uintptr_t>(a.idx) & 15)) return VCFB_E_UNSUPP;
  if (a.q_int < 1 || a.q_int > 255 || tma::encode_tiled_fn() == nullptr) return VCFB_E_UNSUPP;
  const unsigned char* btab = device_table(0);
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
  ta.prof = nullptr;
  static const bool want_prof = getenv("VCFB_TC_PROF") != nullptr;
  static long long* prof_buf = nullptr;
  if (want_prof) {
    if (!prof_buf) cudaMalloc(reinterpret_cast<void**>(&prof_buf), 96 * 8);
    cudaMemsetAsync(prof_buf, 0, 96 * 8, s);
    ta.prof = prof_buf;
  }
  int grid = sm_count();
  if (grid > ta.ntiles) grid = ta.ntiles;
  // development knob VCFB_TC_CFG = "<epilogue groups>x<converter groups>"
  const int cfg = dev_cfg("VCFB_TC_CFG");
  void (*kern)(const CUtensorMap, const CUtensorMap, const TcDecArgs);
  int nthreads;
  switch (cfg) {
    case 22: kern = dec8_tc_kernel<2, 2>; nthreads = 4 * 128 + 96; break;
    case 42: kern = dec8_tc_kernel<4, 2>; nthreads = 6 * 128 + 96; break;
    case 14: kern = dec8_tc_kernel<1, 4>; nthreads = 5 * 128 + 96; break;
    case 24: kern = dec8_tc_kernel<2, 4>; nthreads = 6 * 128 + 96; break;
    default: kern = dec8_tc_kernel<1, 2>; nthreads = 3 * 128 + 96; break;      // memory-bound already: the fewest warps
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dec8_tc)");
  note_kernel("dec8_tc");
  kern<<<grid, nthreads, SMEM_BYTES, s>>>(in_map, out_map, ta);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "dec8_tc_kernel launch");
  if (want_prof) {
    long long h[96];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, prof_buf, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[8] = {"epi0", "epi1", "conv0", "conv1", "conv2", "conv3", "tma", "-"};
    const char* names2[2] = {"store", "mma"};
    for (int r = 0; r < 10; ++r)
      if (h[r * 8]) fprintf(stderr, "[tc prof] %-5s total %lld  wait0 %lld  wait1 %lld  wait2 %lld  wait3 %lld\n", r < 8 ? names[r] : names2[r - 8], h[r * 8], h[r * 8 + 1], h[r * 8 + 2], h[r * 8 + 3], h[r * 8 + 4]);
  }
  return VCFB_OK;
}

// B = 8, YCoCg, subband layout, q a power of two >= 8 (|index| <= 128 so the int8 pack cannot saturate):
// the fast-mode encoder (VCFB_F_FAST).  Everything else -> VCFB_E_UNSUPP -> the bit-exact encoders.
int launch_encode_tc(const EncArgs& a, cudaStream_t s) {
  const Geom& g = a.g;
  if (a.color != VCFB_COLOR_YCOCG || (a.flags & (VCFB_F_NO_SUBBANDS | VCFB_F_PERCEPTUAL | VCFB_F_FP64))) return VCFB_E_UNSUPP;
  if (g.W % 16 != 0 || g.nx % 16 != 0 || g.left != 0) return VCFB_E_UNSUPP;
  if ((reinterpret_cast<uintptr_t>(a.rgb) & 15) || (reinterpret_cast<uintptr_t>(a.idx) & 15)) return VCFB_E_UNSUPP;
  if (!a.q_pow2 || a.q < 8.0 || a.q > 1024.0 || tma::encode_tiled_fn() == nullptr) return VCFB_E_UNSUPP;
  const unsigned char* ftab = device_table(1);
  if (!ftab) return VCFB_E_UNSUPP;
  CUtensorMap in_map, out_map;
  {   // RGB frames as (W*3/8 uint64, H, n); box = 64 blocks x 8 rows (two per tile)
    const uint64_t dims[3] = {uint64_t(g.W) * 3 / 8, uint64_t(g.H), uint64_t(a.n_frames)};
    const uint64_t str[2] = {uint64_t(g.W) * 3, uint64_t(g.H) * g.W * 3};
    const uint32_t box[3] = {64 * 24 / 8, 8, 1};
    if (!tma::make_map(&in_map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint8_t*>(a.rgb), dims, str, box)) return VCFB_E_UNSUPP;
  }
  {   // index planes as (x words, v, y, u, frame) -> smem [u][v][384 bytes]
    const uint64_t si = uint64_t(g.nx) * 3, sy = uint64_t(g.Wp) * 3, sj = uint64_t(g.ny) * g.Wp * 3,
                   sf = uint64_t(g.Hp) * g.Wp * 3;
    const uint64_t dims[5] = {uint64_t(g.nx) * 3 / 4, 8, uint64_t(g.ny), 8, uint64_t(a.n_frames)};
    const uint64_t str[4] = {si, sy, sj, sf};
    const uint32_t box[5] = {TB * 3 / 4, 8, 1, 8, 1};
    if (!tma::make_map(&out_map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, a.idx, dims, str, box)) return VCFB_E_UNSUPP;
  }
  enc::TcEncArgs ta;
  ta.tiles_x = (g.nx + TB - 1) / TB;
  ta.ny = g.ny;
  ta.nx = g.nx;
  ta.top = g.top;
  const long long nt = (long long)a.n_frames * g.ny * ta.tiles_x;
  if (nt > (0x7fffffffLL - (1 << 20)) / 3) return VCFB_E_UNSUPP;
  ta.ntiles = int(nt);
  ta.inv_q = float(a.inv_q);
  ta.ftab = ftab;
  ta.prof = nullptr;
  ta.stats = a.stats;                      // non-NULL only when the caller wants the sums without the histogram (api.cu)
#ifdef VCFB_TC_PROFILE
  static long long* eprof_buf = nullptr;
  if (!eprof_buf) cudaMalloc(reinterpret_cast<void**>(&eprof_buf), 128);
  ta.prof = eprof_buf;
#endif
  int grid = sm_count();
  if (grid > ta.ntiles) grid = ta.ntiles;
  const int cfg = dev_cfg("VCFB_TC_ENC_CFG");
  void (*kern)(const CUtensorMap, const CUtensorMap, const enc::TcEncArgs);
  int nthreads;
  if (cfg == 22 && !a.stats) {
    kern = enc::enc8_tc_kernel<2, 2>;
    nthreads = 4 * 128 + 128;
  } else {
    kern = a.stats ? enc::enc8_tc_kernel<3, 2, true> : enc::enc8_tc_kernel<3, 2>;
    nthreads = 5 * 128 + 128;
  }
  if (a.stats) {
    const int rc = launch_add_count(a.stats, VCFB_STAT_NINDICES, (unsigned long long)a.n_frames * g.Hp * g.Wp * 3, s);
    if (rc) return rc;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, enc::ESMEM);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(enc8_tc)");
  note_kernel("enc8_tc");
  kern<<<grid, nthreads, enc::ESMEM, s>>>(in_map, out_map, ta);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "enc8_tc_kernel launch");
#ifdef VCFB_TC_PROFILE
  {
    long long h[12];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, ta.prof, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tc enc prof] mma total %lld  wait d_empty %lld  a_full %lld | epi total %lld  wait d_full %lld  out_free %lld  store phase %lld | conv total %lld  wait in_full %lld  convert %lld  wait a_empty %lld  (%d tiles per CTA)\n",
            h[0], h[1], h[2], h[4], h[5], h[6], h[7], h[8], h[9], h[10], h[11], (ta.ntiles + grid - 1) / grid);
  }
#endif
  return VCFB_OK;
}

}  // namespace vcfb
