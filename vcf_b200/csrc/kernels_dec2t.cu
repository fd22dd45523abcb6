// Two-tier float64 decoder of the B=8 fast path: bit-exact with the reference's float64
// chain (src/2D-DCT.py:398-466) at roughly half of its FP64 work.
//
// Why two tiers.  The reference truncates  clip(to_RGB(idct2(q*k)) + 128)  to uint8.  The
// truncated byte depends on the last-ulp rounding of pocketfft's operation sequence ONLY when
// the real value y* of the sample lies within rounding distance (~1e-9 at worst) of an
// integer; everywhere else any float64 evaluation with error << 1 gives the same byte.
//
//   tier 1 (every sample): a scaled Arai-Agui-Nakajima inverse DCT with fused
//     multiply-adds, 22.5 FP64 operations per pixel instead of pocketfft's 53:
//       * to_RGB is linear and q is uniform, so the colour transform is applied to the
//         integer INDICES (3 dp4a on the packed bytes) and the three inverse DCTs run on
//         R, G, B coefficient planes;
//       * the AAN pre-scale rides on fused multiply-adds of the first butterfly stage,
//         q/8 is folded into the per-lane scale constants;
//       * the constant 128 + 1.5*2^22 enters at the DC input, which reaches every output
//         through additions only: each result is therefore a fixed-point number with 30
//         fraction bits in the low mantissa word.  Integer ops extract floor(y+128), one
//         saturating pack clips to [0,255], and the fraction bits give the tie test
//         |y - rint(y)| < 2^-20 for free -- no conversion instruction, no FP64 compare.
//       Error budget: |values| < 2^21 for q <= 255 (fast-path precondition), so one rounding
//       is <= 2^-31 and a chain of < 16 roundings stays below 1e-8 in either tier; the tie
//       threshold 2^-20 ~ 9.5e-7 leaves a factor of 50.
//   tier 2 (samples whose tier-1 value is within 2^-20 of an integer): the reference's own
//     sequence of individually rounded operations.
//       * blocks without any AC index: pocketfft's DAG degenerates to two multiplications by
//         its sqrt(2) constant (every other operand is an exact zero), evaluated per block;
//         a half-tile that is DC-only throughout skips tier 1 altogether;
//       * otherwise the half-tile of 8 blocks is recomputed with the exact codelets.
//
// Pays off when exact-integer samples are rare (dense indices); sparse indices put them into
// most half-tiles and tier 1 is then wasted work, so the probe of kernels_fast.cu decides per
// batch whether this kernel or the single-tier exact one runs.  exact_half stays out of line:
// inlined, the loop body no longer fits the instruction cache and the kernel runs 2-3x slower.
//
// Pipeline, tile shape and lane mapping are those of dec8_f64h_kernel (kernels_fast.cu); the
// intermediate is XOR-swizzled instead of padded (12 KB per warp).
#include "fast_common.cuh"
#include "dec8_dc.cuh"

namespace vcfb {
using namespace fast;
namespace {

// Intermediate of one half-tile (8 blocks x 64 x 3 doubles = 12288 bytes, no padding):
// F[c][y >> 1][i][b] in 16-byte units (one unit = the two blocks of a pair), with
// b = (((y & 1) << 2) | pair) ^ i.  Pass-1 stores (a quarter-warp = 8 columns i of one pair)
// and pass-2 loads (a quarter-warp = 4 pairs x 2 rows y) both hit 8 distinct units mod 8.
constexpr int F_BYTES = 3 * 4 * 64 * 16;
__host__ __device__ constexpr int warp_smem(int nst) { return (nst * TILE + F_BYTES + 8 * nst + 127) / 128 * 128; }
__device__ __forceinline__ int f_store_off(int i1, int G1) { return (i1 * 8 + (G1 ^ i1)) * 2; }   // even y; odd y: ^ 8
__device__ __forceinline__ int f_load_off(int y2, int G2, int i) {
  return (y2 >> 1) * 128 + i * 16 + (((((y2 & 1) << 2) | G2) ^ i) << 1);
}

// sqrt(2) * cos(k pi / 16), k = 1..7 (AAN pre-scale); exactly 1 for k = 0 and k = 4
__constant__ double c_aan[8] = {1.0,
                                0x1.63150b15e8536p+0,
                                0x1.4e7ae9144f0fcp+0,
                                0x1.2d062ef88e31ap+0,
                                1.0,
                                0x1.92469c0dcf32fp-1,
                                0x1.1517a7bdb3896p-1,
                                0x1.1a855dec071b7p-2};
constexpr double SQRT2 = 0x1.6a09e667f3bcdp+0;
constexpr double K1 = 0x1.d906bcf328d46p+0;     //  2 cos(pi/8)
constexpr double K3 = -0x1.87de2a6aea964p-1;    //  2 (cos(pi/8) - cos(3pi/8)) - 2 cos(pi/8)

constexpr double T1_BIAS = 6291456.0 + 128.0;   // 1.5 * 2^22 (30 fraction bits) + the +128 of :454
constexpr int T1_IBASE = 0x05600000;            // ((0x415 << 22) mod 2^32) + 2^21
constexpr unsigned T1_TIE = 4096u;              // 4 * 2^10: |fraction| < 2^-20 (fraction bits << 2)

// Scaled inverse DCT, length 8.  PRE: the inputs are raw (integer-valued) and get their
// pre-scale c[k] inside the first butterflies; otherwise they are used as they are.
template <bool PRE>
__device__ __forceinline__ void aan8_inv(double (&v)[8], const double (&c)[8], double bias) {
  double x0, x1, x2, x5, t10, t11, t13, d26, z13, z10, z11, z12;
  if (PRE) {
    x0 = fma(v[0], c[0], bias);
    t10 = fma(v[4], c[4], x0);
    t11 = fma(v[4], -c[4], x0);
    x2 = v[2] * c[2];
    t13 = fma(v[6], c[6], x2);
    d26 = fma(v[6], -c[6], x2);
    x5 = v[5] * c[5];
    z13 = fma(v[3], c[3], x5);
    z10 = fma(v[3], -c[3], x5);
    x1 = v[1] * c[1];
    z11 = fma(v[7], c[7], x1);
    z12 = fma(v[7], -c[7], x1);
  } else {
    t10 = v[0] + v[4];
    t11 = v[0] - v[4];
    t13 = v[2] + v[6];
    d26 = v[2] - v[6];
    z13 = v[5] + v[3];
    z10 = v[5] - v[3];
    z11 = v[1] + v[7];
    z12 = v[1] - v[7];
  }
  const double t12 = fma(d26, SQRT2, -t13);
  const double e0 = t10 + t13, e3 = t10 - t13, e1 = t11 + t12, e2 = t11 - t12;
  const double o7 = z11 + z13;
  const double o6 = fma(z10, K3, fma(z12, K1, -o7));
  const double o5 = fma(z11 - z13, SQRT2, -o6);
  const double o4 = fma(z12, K3, fma(z10, -K1, o5));
  v[0] = e0 + o7;
  v[7] = e0 - o7;
  v[1] = e1 + o6;
  v[6] = e1 - o6;
  v[2] = e2 + o5;
  v[5] = e2 - o5;
  v[4] = e3 + o4;
  v[3] = e3 - o4;
}

// two saturated bytes on top of the low half of c: (c << 16) | sat_u8(a) << 8 | sat_u8(b)
__device__ __forceinline__ unsigned pack_sat_u8(int a, int b, unsigned c) {
  unsigned d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// ---- tier 2b: the reference's chain for one half-tile (same code as dec8_f64h_kernel) ------
struct HalfWords {
  uint32_t w[8][2];      // [coefficient row u][word]: the lane's 6-byte run of one half-tile
};

__device__ __noinline__ void exact_half(const HalfWords hw, double* F, unsigned char* tb, int h, int i1, int G1,
                                        int G2, int y2, int sh0, int q) {
  using O = Ops<double, true>;
  constexpr double SCALE = p2(2 * M8I::exp(0));
  {
    double* fw0 = F + f_store_off(i1, G1);
    double* fw1 = F + (f_store_off(i1, G1) ^ 8);
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
      double v[2][8];
#pragma unroll
      for (int uu = 0; uu < 8; ++uu) {
        const uint32_t sv = __funnelshift_rc(hw.w[uu][0], hw.w[uu][1], sh0 + 8 * c);
        v[0][uu] = __int2double_rn(int(sv & 255u) * q - 128 * q);     // int16 * int of :398-410
        v[1][uu] = __int2double_rn(int(sv >> 24) * q - 128 * q);
      }
      dct8_inv<double, true>(v[0]);
      dct8_inv<double, true>(v[1]);
#pragma unroll
      for (int yy = 0; yy < 8; ++yy)
        *reinterpret_cast<double2*>((yy & 1 ? fw1 : fw0) + (c * 4 + (yy >> 1)) * 128) = make_double2(v[0][yy], v[1][yy]);
    }
  }
  __syncwarp();
  {
    const double* fr = F;
    double v[3][2][8];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const double2 t2 = *reinterpret_cast<const double2*>(fr + c * 512 + f_load_off(y2, G2, i));
        v[c][0][i] = t2.x;
        v[c][1][i] = t2.y;
      }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      dct8_inv<double, true>(v[c][0]);
      dct8_inv<double, true>(v[c][1]);
    }
    int px[2][8][3];
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        const double Y = v[0][b][x], Co = v[1][b][x], Cg = v[2][b][x];
        const double R = O::fma(O::sub(O::add(Y, Co), Cg), SCALE, 128.0);
        const double Gc = O::fma(O::add(Y, Cg), SCALE, 128.0);
        const double Bc = O::fma(O::sub(O::sub(Y, Co), Cg), SCALE, 128.0);
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
  }
}

template <int NWARPS, int CTAS, int NST, bool UNROLL_C, bool SSE>
__global__ void __launch_bounds__(NWARPS * 32, CTAS)
dec8_2t_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
               const FastDecArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  if (not_chosen(a)) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* ring = smem + warp * warp_smem(NST);
  double* F = reinterpret_cast<double*>(ring + NST * TILE);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + NST * TILE + F_BYTES);

  if (lane == 0) {
    tma::prefetch_map(&in_map);
    tma::prefetch_map(&out_map);
#pragma unroll
    for (int s = 0; s < NST; ++s) tma::mbar_init(&full[s], 1);
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
    for (int s = 0; s < NST; ++s) {
      const int t = w.tile + s * w.stride;
      if (t < w.ntiles) issue_load(s, t);
    }
  }

  Lane L;
  L.i1 = lane & 7;
  L.G1 = lane >> 3;
  L.G2 = lane & 3;
  L.y2 = lane >> 2;
  L.sh0 = ((6 * L.G1) & 3) * 8;
  L.q = a.q;
  const int woff = (6 * L.G1) >> 2;
  // tier-1 scale of coefficient (u, i1): aan[u] * aan[i1] * q / 8
  double cu[8];
  {
    const double ai = c_aan[L.i1] * double(a.q) * 0.125;
#pragma unroll
    for (int u = 0; u < 8; ++u) cu[u] = c_aan[u] * ai;
  }
  const double bias1 = L.i1 == 0 ? T1_BIAS : 0.0;
  // index bytes are k + 128:  R = Y + Co - Cg,  G = Y + Cg,  B = Y - Co - Cg  on the indices
  constexpr int MIX[3] = {0x00FF0101, 0x00010001, 0x00FFFF01};
  constexpr int MIXB[3] = {-128, -256, 128};
  // bytes of a lane's 6-byte run that belong to an AC index (lane i1 == 0, row u == 0 holds the DC)
  const uint32_t dcm = L.i1 == 0 ? 0u : 0xffffffffu;
  SseAcc acc;
  sse_reset(acc);
  // the 48 bytes of the original frame under this lane's bytes of half h of a tile
  auto load_orig = [&](int t, int h, uint4 (&o)[3]) {
    int f, by, tx;
    w.coords(t, f, by, tx);
    const uint4* p = reinterpret_cast<const uint4*>(a.original + f * a.frame_bytes + (long long)(by * 8 + L.y2) * a.row_bytes +
                                                    tx * (WT * 3) + 192 * h + 48 * L.G2);
    o[0] = __ldg(p);
    o[1] = __ldg(p + 1);
    o[2] = __ldg(p + 2);
  };

  int k = 0;
  for (int tile = w.tile; tile < w.ntiles; tile += w.stride, ++k) {
    const int s = k % NST;
    unsigned char* tb = ring + s * TILE;
    tma::mbar_wait(&full[s], (k / NST) & 1);

    HalfWords wd[2];
    uint32_t nz[2][2];
    {
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(tb) + L.i1 * 12 + woff;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int uu = 0; uu < 8; ++uu) {
          wd[h].w[uu][0] = rw[uu * 96 + 6 * h];
          wd[h].w[uu][1] = rw[uu * 96 + 6 * h + 1];
        }
        nz[h][0] = (wd[h].w[0][0] ^ 0x80808080u) & dcm;
        nz[h][1] = (wd[h].w[0][1] ^ 0x80808080u) & dcm;
#pragma unroll
        for (int uu = 1; uu < 8; ++uu) {
          nz[h][0] |= wd[h].w[uu][0] ^ 0x80808080u;
          nz[h][1] |= wd[h].w[uu][1] ^ 0x80808080u;
        }
      }
    }
    __syncwarp();

#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint4 og[3];
      if (SSE) load_orig(tile, h, og);
      const unsigned char* mine = tb + L.y2 * (WT * 3) + 192 * h + 48 * L.G2;
      const unsigned ac24 = half_ac24(L, nz[h][0], nz[h][1]);
      if (ac24 == 0u) {                              // DC-only half-tile: tier 2a for all 8 blocks
        dc_blocks(L, wd[h].w[0][0], wd[h].w[0][1], tb, h, 0xffu);
        __syncwarp();
        if (SSE) sse_row48(acc, og, mine);
        continue;
      }
      // ---- tier 1, pass 1: index colour mix, int -> double, scaled inverse DCT over u ------
      {
        double* fw0 = F + f_store_off(L.i1, L.G1);
        double* fw1 = F + (f_store_off(L.i1, L.G1) ^ 8);
#pragma unroll(UNROLL_C ? 3 : 1)
        for (int c = 0; c < 3; ++c) {
          double v[2][8];
#pragma unroll
          for (int uu = 0; uu < 8; ++uu) {
            const uint32_t lo = __funnelshift_r(wd[h].w[uu][0], wd[h].w[uu][1], L.sh0), hi = wd[h].w[uu][1] >> L.sh0;
            const uint32_t pb = __byte_perm(lo, hi, 0x0543);
            v[0][uu] = __int2double_rn(dp4a_us(lo, MIX[c], MIXB[c]));
            v[1][uu] = __int2double_rn(dp4a_us(pb, MIX[c], MIXB[c]));
          }
          aan8_inv<true>(v[0], cu, bias1);
          aan8_inv<true>(v[1], cu, bias1);
#pragma unroll
          for (int yy = 0; yy < 8; ++yy)
            *reinterpret_cast<double2*>((yy & 1 ? fw1 : fw0) + (c * 4 + (yy >> 1)) * 128) =
                make_double2(v[0][yy], v[1][yy]);
        }
      }
      __syncwarp();
      // ---- tier 1, pass 2: inverse DCT over i, fixed-point extraction, tie test ------------
      unsigned tie8;
      {
        const double* fr = F;
        int px[2][8][3];
        unsigned dmin[2] = {0xffffffffu, 0xffffffffu};
#pragma unroll(UNROLL_C ? 3 : 1)
        for (int c = 0; c < 3; ++c) {
          double v[2][8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const double2 t2 = *reinterpret_cast<const double2*>(fr + c * 512 + f_load_off(L.y2, L.G2, i));
            v[0][i] = t2.x;
            v[1][i] = t2.y;
          }
          aan8_inv<false>(v[0], cu, 0.0);
          aan8_inv<false>(v[1], cu, 0.0);
#pragma unroll
          for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int x = 0; x < 8; ++x) {
              const unsigned lo = unsigned(__double2loint(v[b][x])), hi = unsigned(__double2hiint(v[b][x]));
              px[b][x][c] = int(__funnelshift_l(lo, hi, 2)) - T1_IBASE;      // floor(y + 128)
              dmin[b] = min(dmin[b], (lo << 2) + T1_TIE);
            }
        }
        const int* p = &px[0][0][0];
        uint32_t ww[12];
#pragma unroll
        for (int j = 0; j < 12; ++j)
          ww[j] = pack_sat_u8(p[4 * j + 1], p[4 * j], pack_sat_u8(p[4 * j + 3], p[4 * j + 2], 0u) );
        uint4* orow = reinterpret_cast<uint4*>(tb + L.y2 * (WT * 3) + 192 * h + 48 * L.G2);
        orow[0] = make_uint4(ww[0], ww[1], ww[2], ww[3]);
        orow[1] = make_uint4(ww[4], ww[5], ww[6], ww[7]);
        orow[2] = make_uint4(ww[8], ww[9], ww[10], ww[11]);
        tie8 = __reduce_or_sync(0xffffffffu, ((dmin[0] < 2u * T1_TIE ? 1u : 0u) | (dmin[1] < 2u * T1_TIE ? 2u : 0u))
                                                 << (2 * L.G2));
      }
      __syncwarp();
      // ---- tier 2 ---------------------------------------------------------------------------
      if (tie8) {
        // block-level AC mask: bit 2*pair + block
        unsigned ac8 = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) ac8 |= ((ac24 >> (3 * j)) & 1u) << j;
        if (tie8 & ac8) {
          exact_half(wd[h], F, tb, h, L.i1, L.G1, L.G2, L.y2, L.sh0, L.q);
        } else {
          dc_blocks(L, wd[h].w[0][0], wd[h].w[0][1], tb, h, tie8);
        }
        __syncwarp();
      }
      if (SSE) sse_row48(acc, og, mine);
    }
    tma::fence_proxy_async();
    __syncwarp();

    if (lane == 0) {
      issue_store(s, tile);
      tma::wait_group_read<1>();
      if (k >= 1) {
        const int nt = tile + (NST - 1) * w.stride;
        if (nt < w.ntiles) issue_load((k - 1) % NST, nt);
      }
    }
    __syncwarp();
  }
  if (lane == 0) tma::wait_group<0>();
  if (SSE) sse_finish(acc, a.stats, lane);
}

template <int NWARPS, int CTAS, int NST, bool UNROLL_C, bool SSE = false>
int launch_t(const CUtensorMap& in_map, const CUtensorMap& out_map, const FastDecArgs& fa, cudaStream_t s) {
  int grid = sm_count() * CTAS;
  const int need = (fa.ntiles + NWARPS - 1) / NWARPS;
  if (grid > need) grid = need;
  auto kern = dec8_2t_kernel<NWARPS, CTAS, NST, UNROLL_C, SSE>;
  const int smem_bytes = NWARPS * warp_smem(NST);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(dec8_2t)");
  note_kernel("dec8_fast");
  kern<<<grid, NWARPS * 32, smem_bytes, s>>>(in_map, out_map, fa);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "dec8_2t_kernel launch");
  return VCFB_OK;
}

}  // namespace

int launch_decode_2t(int cfg, const CUtensorMap& in_map, const CUtensorMap& out_map, const fast::FastDecArgs& fa,
                     cudaStream_t s) {
  if (fa.stats) return launch_t<8, 1, 3, true, true>(in_map, out_map, fa, s);   // fused distortion statistics
  switch (cfg) {
    case 42: return launch_t<4, 2, 3, true>(in_map, out_map, fa, s);
    case 43: return launch_t<4, 3, 2, true>(in_map, out_map, fa, s);     // 12 warps per SM, 2-stage ring
    case 82: return launch_t<8, 1, 2, true>(in_map, out_map, fa, s);
    default: return launch_t<8, 1, 3, true>(in_map, out_map, fa, s);
  }
}

}  // namespace vcfb
