// Packed (f32x2) exact encoder of the B=8 fast path.
//
// This translation unit is compiled with -fmad=false: ptxas 12.9 fuses mul.rn.f32x2 +
// add.rn.f32x2 (and fma.rn.f32x2 with a factor of 1) into one FFMA2 even though the
// rounding modifier is explicit -- which would change roundings and break bit-exactness
// (measured: 45 of 24.9 M indices at q = 1).  With contraction off for the whole unit the
// packed operations are issued exactly as written.
#include "fast_common.cuh"

namespace vcfb {
using namespace fast;
namespace {

// ============================================================================
// encode, packed variant: same pipeline, but every floating-point instruction works on
// two independent transforms at once (FADD2 / FMUL2 / FFMA2, IEEE-rounded per lane, so
// still bit-exact).  Floating-point work takes half the issue slots; the FP32 pipe itself
// becomes the limiter instead of instruction issue.
//   pass 1: lane = (pair of adjacent blocks bp, pair of columns ih): columns 2ih, 2ih+1
//           of block A = 2bp and of block B = 2bp+1; packed value = (A, B)
//   F[c][u][bp][ih] holds float4 (A.i, B.i, A.i+1, B.i+1), i = 2ih: pass 2 reads packed pairs
//   pass 2: lane = (row u, group G of 4 blocks = 2 block pairs)
// ============================================================================
__device__ __forceinline__ float2 dotf2(unsigned pa, unsigned pb, int coef, int bias) {
  const float2 r = make_float2(__int_as_float(dp4a_us(pa, coef, MAGIC_I + bias)),
                               __int_as_float(dp4a_us(pb, coef, MAGIC_I + bias)));
  return Ops<float2, true>::add(r, make_float2(-MAGIC_F, -MAGIC_F));
}

template <bool QPOW2, int NWARPS, int CTAS, int NST, bool STATS>
__global__ void __launch_bounds__(NWARPS * 32, CTAS)
enc8p_fast_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                  const FastArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* ring = smem + warp * enc_warp_smem(NST);
  float* F = reinterpret_cast<float*>(ring + NST * TILE);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + NST * TILE + ENC_F_BYTES);

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
    tma::load_3d(ring + s * TILE, &in_map, &full[s], tx * (WT * 3 / 8), by * 8 - w.top, f);
  };
  auto issue_store = [&](int s, int t) {
    int f, by, tx;
    w.coords(t, f, by, tx);
    tma::store_5d(&out_map, ring + s * TILE, tx * (WT / 8) * 3, 0, by, 0, f);
    tma::commit_group();
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      const int t = w.tile + s * w.stride;
      if (t < w.ntiles) issue_load(s, t);
    }
  }

  // pass 1 constants: 6-byte run of block A starts at byte 48*bp + 6*ih of the row, B is 24 bytes on
  const int bp = lane >> 2, ih = lane & 3;
  const int widx = 12 * bp + ((6 * ih) >> 2);
  const int sh = ((6 * ih) & 3) * 8;
  // pass 2 constants
  const int u = lane & 7, G = lane >> 3;
  float qs[3][2];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    qs[c][0] = a.qtab[u][c];
    qs[c][1] = a.qtab[u][c] * 2.0f;
  }
  const float qf = a.q;
  unsigned st_nz = 0, st_abs = 0;      // STATS: non-zero indices and sum |k| of this lane (src/IPP_DCT.py:273-292)

  int k = 0;
  for (int tile = w.tile; tile < w.ntiles; tile += w.stride, ++k) {
    const int s = k % NST;
    unsigned char* tb = ring + s * TILE;
    tma::mbar_wait(&full[s], (k / NST) & 1);

    // ---- pass 1 -----------------------------------------------------------------------
    {
      float2 v[3][2][8];     // [channel][column of the pair][row]  = (block A, block B)
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(tb) + widx;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const uint32_t a0 = rw[r * ROWW + 0], a1 = rw[r * ROWW + 1];
        const uint32_t b0 = rw[r * ROWW + 6], b1 = rw[r * ROWW + 7];
        const uint32_t alo = __funnelshift_r(a0, a1, sh), ahi = a1 >> sh;   // bytes 0-3, 4-5 of run A
        const uint32_t blo = __funnelshift_r(b0, b1, sh), bhi = b1 >> sh;
        const uint32_t a2 = __byte_perm(alo, ahi, 0x0543);                  // second pixel of run A
        const uint32_t b2 = __byte_perm(blo, bhi, 0x0543);
        v[0][0][r] = dotf2(alo, blo, 0x00010201, -512);
        v[1][0][r] = dotf2(alo, blo, 0x00FF0001, 0);
        v[2][0][r] = dotf2(alo, blo, 0x00FF02FF, 0);
        v[0][1][r] = dotf2(a2, b2, 0x00010201, -512);
        v[1][1][r] = dotf2(a2, b2, 0x00FF0001, 0);
        v[2][1][r] = dotf2(a2, b2, 0x00FF02FF, 0);
      }
      float* fw = F + 4 * lane;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        dct8_fwd<float2, true>(v[c][0]);
        dct8_fwd<float2, true>(v[c][1]);
#pragma unroll
        for (int uu = 0; uu < 8; ++uu)
          *reinterpret_cast<float4*>(fw + (c * 8 + uu) * ENC_FP) =
              make_float4(v[c][0][uu].x, v[c][0][uu].y, v[c][1][uu].x, v[c][1][uu].y);
      }
    }
    __syncwarp();

    // ---- pass 2 -----------------------------------------------------------------------
    {
      float2 v[3][2][8];     // [channel][block pair of the group][i] = (block 2pp, block 2pp+1)
      const float* fr = F + u * ENC_FP + 32 * G;
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int pp = 0; pp < 2; ++pp)
#pragma unroll
          for (int jh = 0; jh < 4; ++jh) {
            // one 128-bit load delivering two packed pairs (a plain float4 load gets split
            // into two LDS.64, which conflict 2-way in this layout)
            unsigned long long p0, p1;
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];"
                         : "=l"(p0), "=l"(p1)
                         : "r"(tma::smem_u32(fr + c * 8 * ENC_FP + 16 * pp + 4 * jh)));
            v[c][pp][2 * jh] = f2_from(p0);
            v[c][pp][2 * jh + 1] = f2_from(p1);
          }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        dct8_fwd<float2, true>(v[c][0]);
        dct8_fwd<float2, true>(v[c][1]);
      }
      uint32_t* ow = reinterpret_cast<uint32_t*>(tb) + u * 12 + 3 * G;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int kk[4][3];
#pragma unroll
        for (int pp = 0; pp < 2; ++pp)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float sc = M8F::sgn(i) > 0 ? qs[c][M8F::exp(i) - min_exp8()] : -qs[c][M8F::exp(i) - min_exp8()];
            const float2 t2 = Ops<float2, true>::mul(v[c][pp][i], make_float2(sc, sc));   // exact: power of two
            float tx = t2.x, ty = t2.y;
            if (!QPOW2) {
              tx = __fdiv_rn(tx, qf);
              ty = __fdiv_rn(ty, qf);
            }
            kk[2 * pp][c] = __float2int_rz(tx);
            kk[2 * pp + 1][c] = __float2int_rz(ty);
          }
        uint32_t* o = ow + i * 96;
        const uint32_t w0 = pack4(kk[0][0], kk[0][1], kk[0][2], kk[1][0]) ^ 0x80808080u;
        const uint32_t w1 = pack4(kk[1][1], kk[1][2], kk[2][0], kk[2][1]) ^ 0x80808080u;
        const uint32_t w2 = pack4(kk[2][2], kk[3][0], kk[3][1], kk[3][2]) ^ 0x80808080u;
        o[0] = w0;
        o[1] = w1;
        o[2] = w2;
        if (STATS) {
          // |k| per byte = |byte - 128| (wrapped indices count as the byte they became, like the
          // streaming pass over the stored array does)
          const uint32_t d0 = __vabsdiffu4(w0, 0x80808080u), d1 = __vabsdiffu4(w1, 0x80808080u),
                         d2 = __vabsdiffu4(w2, 0x80808080u);
          st_abs = __dp4a(d0, 0x01010101u, __dp4a(d1, 0x01010101u, __dp4a(d2, 0x01010101u, st_abs)));
          st_nz += __popc((d0 | ((d0 & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u) +
                   __popc((d1 | ((d1 & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u) +
                   __popc((d2 | ((d2 & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u);
        }
      }
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
  if (STATS) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      st_nz += __shfl_xor_sync(0xffffffffu, st_nz, o);
      st_abs += __shfl_xor_sync(0xffffffffu, st_abs, o);
    }
    if (lane == 0) {
      atomicAdd(a.stats + VCFB_STAT_NONZERO, (unsigned long long)st_nz);
      atomicAdd(a.stats + VCFB_STAT_SUMABS, (unsigned long long)st_abs);
    }
  }
}


template <int NWARPS, int CTAS, int NST = NSTAGE>
int launch_t(bool qpow2, const CUtensorMap& in_map, const CUtensorMap& out_map, const FastArgs& fa, cudaStream_t s) {
  int grid = sm_count() * CTAS;
  const int need = (fa.ntiles + NWARPS - 1) / NWARPS;
  if (grid > need) grid = need;
  void (*kern)(const CUtensorMap, const CUtensorMap, const FastArgs) =
      fa.stats ? (qpow2 ? enc8p_fast_kernel<true, NWARPS, CTAS, NST, true> : enc8p_fast_kernel<false, NWARPS, CTAS, NST, true>)
               : (qpow2 ? enc8p_fast_kernel<true, NWARPS, CTAS, NST, false> : enc8p_fast_kernel<false, NWARPS, CTAS, NST, false>);
  const int smem_bytes = NWARPS * enc_warp_smem(NST);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(enc8p_fast)");
  note_kernel("enc8_fast");
  kern<<<grid, NWARPS * 32, smem_bytes, s>>>(in_map, out_map, fa);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "enc8p_fast_kernel launch");
  return VCFB_OK;
}

}  // namespace

int launch_encode_packed(int cfg, bool qpow2, const CUtensorMap& in_map, const CUtensorMap& out_map,
                         const FastArgs& fa, cudaStream_t s) {
  switch (cfg) {
    case 25: return launch_t<2, 5>(qpow2, in_map, out_map, fa, s);
    case 33: return launch_t<3, 3>(qpow2, in_map, out_map, fa, s);
    case 19: return launch_t<1, 9>(qpow2, in_map, out_map, fa, s);
    case 52: return launch_t<5, 2>(qpow2, in_map, out_map, fa, s);
    case 43: return launch_t<4, 3, 2>(qpow2, in_map, out_map, fa, s);      // 2-stage ring: 12 warps / SM
    case 61: return launch_t<6, 2, 2>(qpow2, in_map, out_map, fa, s);
    default: return launch_t<4, 2>(qpow2, in_map, out_map, fa, s);
  }
}

}  // namespace vcfb
