// B = 16 fast path, pieces shared by kernels_b16.cu (exact float32 encoder, exact float64 decoder; compiled
// with -fmad=false) and kernels_b16f.cu (float32 fast-mode decoder; fused multiply-adds allowed).
#pragma once

#include <stdlib.h>

#include "dec8_dc.cuh"
#include "fast_common.cuh"

namespace vcfb {
namespace b16 {
using namespace fast;

constexpr int T16_W = 256;                    // pixels per tile
constexpr int T16_BYTES = 16 * T16_W * 3;     // 12288: RGB tile == index box
constexpr int T16_ROWW = T16_W * 3 / 4;       // words per RGB row
constexpr int NST16 = 2;
constexpr int ENC16_THREADS = 128;

// RGB frames as (W*3/8 uint64, H, n); box = one tile (96 uint64 x 16 rows)
inline bool make_rgb_map16(CUtensorMap* m, const Geom& g, int n, const void* base) {
  const uint64_t dims[3] = {uint64_t(g.W) * 3 / 8, uint64_t(g.H), uint64_t(n)};
  const uint64_t str[2] = {uint64_t(g.W) * 3, uint64_t(g.H) * g.W * 3};
  const uint32_t box[3] = {T16_W * 3 / 8, 16, 1};
  return tma::make_map(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(base), dims, str, box);
}

// ============================================================================================
// decode: T = double, EXACT (the reference's float64 chain, operation for operation; kernels_b16.cu) or
// T = float, contracted (the fast mode: +-1 LSB; kernels_b16f.cu, compiled with fused multiply-adds)
//
// The tile is processed as two halves of 8 blocks so that the float64 intermediate is 48 KB:
//   pass 1  thread = (coefficient column i, block b): 16 index triples down u, dequantise
//           (int16 * int as in src/2D-DCT.py:398-410), dct16_inv per channel, F[c][y][b][i]
//   pass 2  thread = (pixel row y, block b): dct16_inv along the row for the three channels,
//           to_RGB (YCoCg or the float YCrCb extension), +128, clip, truncate, 48 bytes of RGB
// The RGB tile aliases the index box, so a thread reads the index bytes of BOTH halves (12 packed
// words for the second) before the first barrier, i.e. before anybody writes pixels.
// ============================================================================================
template <typename T> struct Dec16L {
  static constexpr int VEC = 16 / int(sizeof(T));
  static constexpr int PITCH = 8 * 16 + VEC;           // elements per (c, y) row: 8 blocks x 16 columns, + 16 bytes
  static constexpr int F_BYTES = 3 * 16 * PITCH * int(sizeof(T));   // 49920 for double
  static constexpr int SMEM = NST16 * T16_BYTES + F_BYTES + 64 + 8 * 3 * 8;   // + the DC values of 8 blocks
};
constexpr int PRUNE16 = 4;                           // rows / columns kept by the pruned codelet (dct16_inv_low4)

struct Dec16Args {
  int ntiles, tiles_x, ny, top;
  int q;
  const uint8_t* original;      // fused distortion statistics when set (with stats)
  unsigned long long* stats;
  long long frame_bytes;
  int row_bytes;
};

__device__ __forceinline__ int clamp255(int v) { return min(max(v, 0), 255); }
__device__ __forceinline__ int to_int_rz16(double x) { return __double2int_rz(x); }
__device__ __forceinline__ int to_int_rz16(float x) { return __float2int_rz(x); }

template <typename T, bool EXACT, bool YCRCB, bool SSE>
__global__ void __launch_bounds__(ENC16_THREADS, 3)
dec16_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map, const Dec16Args a) {
  using O = Ops<T, EXACT>;
  using L = Dec16L<T>;
  constexpr int D16_PITCH = L::PITCH, D16_F_BYTES = L::F_BYTES, VEC = L::VEC;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ring = smem;
  T* F = reinterpret_cast<T*>(smem + NST16 * T16_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NST16 * T16_BYTES + D16_F_BYTES);
  double* DCV = reinterpret_cast<double*>(smem + NST16 * T16_BYTES + D16_F_BYTES + 64);   // [block of the half][channel]
  const int tid = threadIdx.x;

  if (tid == 0) {
    tma::prefetch_map(&in_map);
    tma::prefetch_map(&out_map);
#pragma unroll
    for (int s = 0; s < NST16; ++s) tma::mbar_init(&full[s], 1);
    tma::fence_mbar_init();
  }
  __syncthreads();

  const int per_frame = a.ny * a.tiles_x;
  auto coords = [&](int t, int& f, int& by, int& tx) {
    f = t / per_frame;
    const int rem = t - f * per_frame;
    by = rem / a.tiles_x;
    tx = rem - by * a.tiles_x;
  };
  auto issue_load = [&](int s, int t) {
    int f, by, tx;
    coords(t, f, by, tx);
    tma::mbar_expect_tx(&full[s], T16_BYTES);
    // in_map dims: (x bytes, i, block row, j, frame); smem box is [j][i][48 B]
    tma::load_5d(ring + s * T16_BYTES, &in_map, &full[s], tx * 16 * 3, 0, by, 0, f);
  };
  auto issue_store = [&](int s, int t) {
    int f, by, tx;
    coords(t, f, by, tx);
    tma::store_3d(&out_map, ring + s * T16_BYTES, tx * (T16_W * 3 / 8), by * 16 - a.top, f);
    tma::commit_group();
  };
  const int stride = gridDim.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST16; ++s) {
      const int t = blockIdx.x + s * stride;
      if (t < a.ntiles) issue_load(s, t);
    }
  }

  const int i1 = tid & 15, b1 = tid >> 4;          // pass 1: coefficient column, block of the half
  const int y2 = tid & 15, b2 = tid >> 4;          // pass 2: pixel row, block of the half
  const int q = a.q;
  SseAcc acc;
  sse_reset(acc);

  int k = 0;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += stride, ++k) {
    const int s = k % NST16;
    unsigned char* tb = ring + s * T16_BYTES;
    tma::mbar_wait(&full[s], (k / NST16) & 1);

    // index bytes (Y, Co, Cg) of (u, i1, block) for both halves: 3 bytes per u, 4 u per 3 words
    uint32_t kw[2][12];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int boff = (8 * h + b1) * 3;            // byte offset of the block inside the 48-byte run
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(tb + i1 * 48) + (boff >> 2);
      const int sh = (boff & 3) * 8;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t t3[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int uu = 4 * g + e;
          t3[e] = __funnelshift_r(rw[uu * (16 * 12)], rw[uu * (16 * 12) + 1], sh) & 0x00ffffffu;
        }
        kw[h][3 * g + 0] = t3[0] | (t3[1] << 24);
        kw[h][3 * g + 1] = (t3[1] >> 8) | (t3[2] << 16);
        kw[h][3 * g + 2] = (t3[2] >> 16) | (t3[3] << 8);
      }
    }

#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // Extent of the non-zero indices of this half-tile: when the coefficient rows (columns) from
      // PRUNE16 on are zero in all 8 blocks, pass 1 (pass 2) runs the pruned codelet -- the same
      // operations minus those on exact zeros -- and the threads of all-zero columns skip pass 1.
      // A block whose only non-zero indices are the three DC ones takes neither pass: pocketfft's sequence
      // then degenerates to two multiplications per pass (dct16_inv_low1: the operations on exact zeros
      // dropped, generated like the other pruned codelets) and every sample of a channel has the same value.
      bool hi_row = false, hi_col = false, dc_blk;
      {
        uint32_t lowu = 0, highu = 0;              // index bytes are 0x80 for a zero index
#pragma unroll
        for (int j = 0; j < 12; ++j) {
          const uint32_t t = kw[h][j] ^ 0x80808080u;
          if (j < (PRUNE16 * 3) / 4) lowu |= t;                                        // rows 0 .. PRUNE16-1
          else highu |= t;
        }
        hi_row = highu != 0u;
        hi_col = i1 >= PRUNE16 && (lowu | highu) != 0u;
        // AC indices of this thread's column: everything but bytes 0..2 of word 0 (u = 0) when i1 == 0
        uint32_t ac = highu | (kw[h][1] ^ 0x80808080u) | (kw[h][2] ^ 0x80808080u);
        ac |= (kw[h][0] ^ 0x80808080u) & (i1 == 0 ? 0xFF000000u : 0xFFFFFFFFu);
        const uint32_t bal = __ballot_sync(0xffffffffu, ac != 0u);                     // 16 threads = one block
        dc_blk = ((bal >> (tid & 16)) & 0xFFFFu) == 0u;
      }
      const bool full_u = __syncthreads_or(hi_row) != 0;
      const bool full_i = __syncthreads_or(hi_col) != 0;
      uint4 og[3];
      if (SSE) {
        int f, by, tx;
        coords(tile, f, by, tx);
        const uint4* p = reinterpret_cast<const uint4*>(a.original + f * a.frame_bytes + (long long)(by * 16 + y2) * a.row_bytes +
                                                        tx * (T16_W * 3) + (8 * h + b2) * 48);
        og[0] = __ldg(p);
        og[1] = __ldg(p + 1);
        og[2] = __ldg(p + 2);
      }
      // ---- pass 1 -------------------------------------------------------------------------
      if (dc_blk) {
        if (i1 == 0) {
          // (always the reference's float64 operations, also in the float32 kernel: these samples are often exact
          //  integers and the truncation then follows the last bit of the float64 chain)
          using OD = Ops<double, true>;
          constexpr double S2 = 0x1.6a09e667f3bcdp+0, S2_8 = 0x1.6a09e667f3bcdp-3;     // dct16_inv_low1
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int k0 = int((kw[h][0] >> (8 * c)) & 255u);
            const double e = OD::mul(OD::mul(__int2double_rn(k0 * q - 128 * q), S2), S2_8);   // axis 0
            DCV[b1 * 3 + c] = OD::mul(OD::mul(e, S2), S2_8);                                   // axis 1
          }
        }
      } else if (full_i || i1 < PRUNE16) {
        T* fw = F + b1 * 16 + i1;
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
          T v[16];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            // bytes 3e + c of the 12-byte group: shift the 96-bit group right by c bytes
            const uint32_t s0 = __funnelshift_r(kw[h][3 * g], kw[h][3 * g + 1], 8 * c);
            const uint32_t s1 = __funnelshift_r(kw[h][3 * g + 1], kw[h][3 * g + 2], 8 * c);
            const uint32_t s2 = kw[h][3 * g + 2] >> (8 * c);
            const int k0 = int(s0 & 255u), k1 = int(s0 >> 24), k2 = int((s1 >> 16) & 255u), k3 = int((s2 >> 8) & 255u);
            v[4 * g + 0] = T(k0 * q - 128 * q);      // (byte - 128) * q: cannot wrap for q <= 255
            v[4 * g + 1] = T(k1 * q - 128 * q);
            v[4 * g + 2] = T(k2 * q - 128 * q);
            v[4 * g + 3] = T(k3 * q - 128 * q);
          }
          if (full_u) dct16_inv<T, EXACT>(v);
          else dct16_inv_low4<T, EXACT>(v);
#pragma unroll
          for (int yy = 0; yy < 16; ++yy) fw[(c * 16 + yy) * D16_PITCH] = v[yy];
        }
      }
      __syncthreads();
      if (h == 1 && tid == 0 && k >= 1) {
        // every thread has taken its index bytes of this tile; the other stage (index box of the
        // previous tile, now pixels being stored) can be refilled once its store has read it
        const int nt = tile + stride;
        if (nt < a.ntiles) {
          tma::wait_group_read<0>();
          issue_load((k + 1) % NST16, nt);
        }
      }
      // ---- pass 2 -------------------------------------------------------------------------
      {
        const T* fr = F + y2 * D16_PITCH + b2 * 16;
        int px[16][3];
        if (dc_blk) {
          using OD = Ops<double, true>;
          const double c0 = DCV[b2 * 3], c1 = DCV[b2 * 3 + 1], c2 = DCV[b2 * 3 + 2];
          double R, G, Bv;
          if (!YCRCB) {
            R = OD::sub(OD::add(c0, c1), c2);
            G = OD::add(c0, c2);
            Bv = OD::sub(OD::sub(c0, c1), c2);
          } else {
            R = OD::add(c0, OD::mul(c1, 1.403));
            G = OD::add(OD::add(c0, OD::mul(c1, -0.714)), OD::mul(c2, -0.344));
            Bv = OD::add(c0, OD::mul(c2, 1.773));
          }
          const int r8 = clamp255(__double2int_rz(OD::add(R, 128.0)));
          const int g8 = clamp255(__double2int_rz(OD::add(G, 128.0)));
          const int b8 = clamp255(__double2int_rz(OD::add(Bv, 128.0)));
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            px[x][0] = r8;
            px[x][1] = g8;
            px[x][2] = b8;
          }
        } else {
        T v[3][16];
        if (full_i) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int m = 0; m < 16 / VEC; ++m) {
              const uint4 w4 = *reinterpret_cast<const uint4*>(fr + c * 16 * D16_PITCH + VEC * m);
              const T* wv = reinterpret_cast<const T*>(&w4);
#pragma unroll
              for (int e = 0; e < VEC; ++e) v[c][VEC * m + e] = wv[e];
            }
            dct16_inv<T, EXACT>(v[c]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int m = 0; m < PRUNE16 / VEC; ++m) {
              const uint4 w4 = *reinterpret_cast<const uint4*>(fr + c * 16 * D16_PITCH + VEC * m);
              const T* wv = reinterpret_cast<const T*>(&w4);
#pragma unroll
              for (int e = 0; e < VEC; ++e) v[c][VEC * m + e] = wv[e];
            }
            dct16_inv_low4<T, EXACT>(v[c]);
          }
        }
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          const T c0 = v[0][x], c1 = v[1][x], c2 = v[2][x];
          T R, G, Bv;
          if (!YCRCB) {            // Y + Co - Cg ; Y + Cg ; Y - Co - Cg, left to right
            R = O::sub(O::add(c0, c1), c2);
            G = O::add(c0, c2);
            Bv = O::sub(O::sub(c0, c1), c2);
          } else {                 // oracle ycrcb_to_rgb_float
            R = O::add(c0, O::mul(c1, T(1.403)));
            G = O::add(O::add(c0, O::mul(c1, T(-0.714))), O::mul(c2, T(-0.344)));
            Bv = O::add(c0, O::mul(c2, T(1.773)));
          }
          px[x][0] = clamp255(to_int_rz16(O::add(R, T(128))));      // :454, :466
          px[x][1] = clamp255(to_int_rz16(O::add(G, T(128))));
          px[x][2] = clamp255(to_int_rz16(O::add(Bv, T(128))));
        }
        }
        const int* p = &px[0][0];
        uint32_t ww[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) ww[j] = pack4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
        uint4* orow = reinterpret_cast<uint4*>(tb + y2 * (T16_W * 3) + (8 * h + b2) * 48);
        orow[0] = make_uint4(ww[0], ww[1], ww[2], ww[3]);
        orow[1] = make_uint4(ww[4], ww[5], ww[6], ww[7]);
        orow[2] = make_uint4(ww[8], ww[9], ww[10], ww[11]);
        if (SSE) sse_row48(acc, og, reinterpret_cast<const unsigned char*>(orow));
      }
      if (h == 0) __syncthreads();          // F is rewritten by pass 1 of the second half
    }
    tma::fence_proxy_async();
    __syncthreads();
    if (tid == 0) issue_store(s, tile);
  }
  if (tid == 0) tma::wait_group<0>();
  if (SSE) sse_finish(acc, a.stats, tid & 31);
}

// Index planes for the decoder: dims (x bytes, i, y, j, frame) -> smem box [j][i][48]
inline bool make_idx_map16_dec(CUtensorMap* m, const Geom& g, int n, const void* base) {
  const uint64_t si = uint64_t(g.nx) * 3, sy = uint64_t(g.Wp) * 3, sj = uint64_t(g.ny) * g.Wp * 3,
                 sf = uint64_t(g.Hp) * g.Wp * 3;
  const uint64_t dims[5] = {uint64_t(g.nx) * 3, 16, uint64_t(g.ny), 16, uint64_t(n)};
  const uint64_t str[4] = {si, sy, sj, sf};
  const uint32_t box[5] = {48, 16, 1, 16, 1};
  return tma::make_map(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, const_cast<void*>(base), dims, str, box);
}

template <typename T, bool EXACT, bool YCRCB>
int launch_dec16(const CUtensorMap& in_map, const CUtensorMap& out_map, const Dec16Args& da, cudaStream_t s, const char* name) {
  constexpr int DEC16_SMEM = Dec16L<T>::SMEM;
  void (*kern)(const CUtensorMap, const CUtensorMap, const Dec16Args) =
      da.stats ? dec16_kernel<T, EXACT, YCRCB, true> : dec16_kernel<T, EXACT, YCRCB, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC16_SMEM);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dec16)");
  int grid = sm_count() * 3;
  if (grid > da.ntiles) grid = da.ntiles;
  note_kernel(name);
  kern<<<grid, ENC16_THREADS, DEC16_SMEM, s>>>(in_map, out_map, da);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "dec16_kernel launch");
  return VCFB_OK;
}


// Preconditions + launch; VCFB_E_UNSUPP when the request is outside this fast path
template <typename T, bool EXACT>
int launch_decode16(const DecArgs& a, cudaStream_t s, const char* name) {
  if (a.flags & (VCFB_F_NO_SUBBANDS | VCFB_F_PERCEPTUAL)) return VCFB_E_UNSUPP;
  if (getenv("VCFB_NO_FAST16")) return VCFB_E_UNSUPP;          // development knob
  if (a.y_out || !a.rgb) return VCFB_E_UNSUPP;
  if ((a.stats != nullptr) != (a.original != nullptr)) return VCFB_E_UNSUPP;
  if (a.q_int < 1 || a.q_int > 255) return VCFB_E_UNSUPP;
  const Geom& g = a.g;
  // (a TMA store whose box starts at a negative row faults on sm_100a: no vertical padding here)
  if (g.W % T16_W != 0 || g.left != 0 || g.top != 0 || g.nx % 16 != 0) return VCFB_E_UNSUPP;
  if ((reinterpret_cast<uintptr_t>(a.rgb) & 15) || (reinterpret_cast<uintptr_t>(a.idx) & 15) ||
      (reinterpret_cast<uintptr_t>(a.original) & 15))
    return VCFB_E_UNSUPP;
  if (!tma::encode_tiled_fn()) return VCFB_E_UNSUPP;
  CUtensorMap in_map, out_map;
  if (!make_idx_map16_dec(&in_map, g, a.n_frames, a.idx)) return VCFB_E_UNSUPP;
  if (!make_rgb_map16(&out_map, g, a.n_frames, a.rgb)) return VCFB_E_UNSUPP;
  Dec16Args da;
  da.tiles_x = g.Wp / T16_W;
  da.ny = g.ny;
  da.top = g.top;
  const long long nt = (long long)a.n_frames * g.ny * da.tiles_x;
  if (nt > 0x7fffffffLL - (1 << 20)) return VCFB_E_UNSUPP;
  da.ntiles = int(nt);
  da.q = a.q_int;
  da.original = a.original;
  da.stats = a.stats;
  da.frame_bytes = (long long)g.H * g.W * 3;
  da.row_bytes = g.W * 3;
  if (a.stats) {
    int rc = launch_add_count(a.stats, VCFB_STAT_NSAMPLES, (unsigned long long)a.n_frames * g.H * g.W * 3, s);
    if (rc) return rc;
  }
  return a.color == VCFB_COLOR_YCRCB ? launch_dec16<T, EXACT, true>(in_map, out_map, da, s, name)
                                     : launch_dec16<T, EXACT, false>(in_map, out_map, da, s, name);
}

}  // namespace b16
}  // namespace vcfb
