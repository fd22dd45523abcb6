// Fast path for B = 16 (BASELINE configs 3 and 5): the exact float32 encoder.
//
// Same ingredients as the B = 8 fast path (kernels_fast.cu, kernels_packed.cu) -- TMA in, TMA
// out, dp4a colour transform, pocketfft-exact codelets on packed f32x2, subband box written in
// shared memory -- but a 16 x 16 block does not fit a warp-private pipeline: the smallest tile
// whose index rows are whole 16-byte units is 16 rows x 256 pixels (16 blocks, 12 KB) and its
// float32 intermediate is 48 KB.  So a CTA of 4 warps shares one tile:
//   pass 1  thread = (block pair bp, column i): column i of blocks 2bp and 2bp+1 as one packed
//           value, 16 rows x 3 channels, three dct16 codelets down the columns, F[c][u][bp][i]
//   pass 2  thread = (coefficient row u, block pair bp): one channel at a time, dct16 along the
//           row, scale (every lazy power of two and 1/q in one per-thread constant), truncate,
//           +128 with wrap, 6 bytes per (u, i) into the box [i][j = u][16 blocks x 3 B]
// with two CTA barriers per tile and three CTAs per SM to overlap them.  Persistent grid, 2-stage
// ring; the index box aliases the consumed RGB tile.
// Colour: YCoCg (exact integers through dp4a, lazy 1/4, 1/2) or the float YCrCb extension of
// oracle/vcf_oracle.py::ycrcb_from_rgb_float, operation for operation.
// This unit is compiled with -fmad=false (see kernels_packed.cu for why).
#include "dec16.cuh"

namespace vcfb {
using namespace fast;
using namespace b16;
namespace {

constexpr int F16_PITCH = 8 * 16 * 2 + 4;     // floats per (c, u) row: 8 pairs x 16 columns x float2, + 16 bytes
constexpr int F16_BYTES = 3 * 16 * F16_PITCH * 4;
constexpr int ENC16_SMEM = NST16 * T16_BYTES + F16_BYTES + 64;

using M16F = dct16_fwd_meta;
__host__ __device__ constexpr int min_exp16() {
  int m = M16F::exp(0);
  for (int i = 1; i < 16; ++i) m = M16F::exp(i) < m ? M16F::exp(i) : m;
  return m;
}
__host__ __device__ constexpr int max_exp16() {
  int m = M16F::exp(0);
  for (int i = 1; i < 16; ++i) m = M16F::exp(i) > m ? M16F::exp(i) : m;
  return m;
}
constexpr int NEXP16 = max_exp16() - min_exp16() + 1;

struct Enc16Args {
  int ntiles, tiles_x, ny, top;
  float q;
  float qtab[16][3];            // [u][c]: sgn_u * 2^(exp_u + colour exp_c + min_i exp_i) (/ q when q is 2^k)
  unsigned long long* stats;
};

using P = Ops<float2, true>;
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }

__device__ __forceinline__ float2 dotf2(unsigned pa, unsigned pb, int coef, int bias) {
  const float2 r = make_float2(__int_as_float(dp4a_us(pa, coef, MAGIC_I + bias)),
                               __int_as_float(dp4a_us(pb, coef, MAGIC_I + bias)));
  return P::add(r, f2(-MAGIC_F));
}

template <bool QPOW2, bool YCRCB, bool STATS>
__global__ void __launch_bounds__(ENC16_THREADS, 3)
enc16_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map, const Enc16Args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ring = smem;
  float* F = reinterpret_cast<float*>(smem + NST16 * T16_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NST16 * T16_BYTES + F16_BYTES);
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
    tma::load_3d(ring + s * T16_BYTES, &in_map, &full[s], tx * (T16_W * 3 / 8), by * 16 - a.top, f);
  };
  auto issue_store = [&](int s, int t) {
    int f, by, tx;
    coords(t, f, by, tx);
    // out_map dims: (x bytes, j, block row, i, frame); smem box is [i][j][48 B]
    tma::store_5d(&out_map, ring + s * T16_BYTES, tx * 16 * 3, 0, by, 0, f);
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

  // pass 1: column i of blocks 2bp, 2bp+1
  const int i1 = tid & 15, bp1 = tid >> 4;
  const int xA = bp1 * 32 + i1;                        // pixel column of block A; block B is 16 further
  const int wA = (3 * xA) >> 2, shA = ((3 * xA) & 3) * 8;
  // pass 2: coefficient row u of blocks 2bp, 2bp+1
  const int u2 = tid & 15, bp2 = tid >> 4;
  float qs[3][NEXP16];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float s = a.qtab[u2][c];
#pragma unroll
    for (int e = 0; e < NEXP16; ++e) {
      qs[c][e] = s;
      s *= 2.0f;
    }
  }
  const float qf = a.q;
  unsigned st_nz = 0, st_abs = 0;

  int k = 0;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += stride, ++k) {
    const int s = k % NST16;
    unsigned char* tb = ring + s * T16_BYTES;
    tma::mbar_wait(&full[s], (k / NST16) & 1);

    // ---- pass 1: colour transform + DCT down the columns ------------------------------------
    {
      float2 v[3][16];
      const uint32_t* rw = reinterpret_cast<const uint32_t*>(tb) + wA;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const uint32_t pa = __funnelshift_r(rw[r * T16_ROWW], rw[r * T16_ROWW + 1], shA);          // R G B . of block A
        const uint32_t pb = __funnelshift_r(rw[r * T16_ROWW + 12], rw[r * T16_ROWW + 13], shA);    // 48 bytes further
        if (!YCRCB) {
          // 4*Y = R + 2G + B - 512 ; 2*Co = R - B ; 4*Cg = -R + 2G - B   (centred pixel)
          v[0][r] = dotf2(pa, pb, 0x00010201, -512);
          v[1][r] = dotf2(pa, pb, 0x00FF0001, 0);
          v[2][r] = dotf2(pa, pb, 0x00FF02FF, 0);
        } else {
          const float2 rr = dotf2(pa, pb, 0x00000001, -128), gg = dotf2(pa, pb, 0x00000100, -128),
                       bb = dotf2(pa, pb, 0x00010000, -128);
          // (r*0.299 + g*0.587) + b*0.114 with every product rounded on its own: written as
          // a - (-c)*x, the form ptxas does not contract into FFMA2 (see exact_ops.cuh)
          const float2 y = P::sub(P::sub(P::mul(rr, f2(0.299f)), P::mul(gg, f2(-0.587f))), P::mul(bb, f2(-0.114f)));
          v[0][r] = y;
          v[1][r] = P::mul(P::sub(rr, y), f2(0.713f));
          v[2][r] = P::mul(P::sub(bb, y), f2(0.564f));
        }
      }
      float* fw = F + (bp1 * 16 + i1) * 2;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        dct16_fwd<float2, true>(v[c]);
#pragma unroll
        for (int uu = 0; uu < 16; ++uu) *reinterpret_cast<float2*>(fw + (c * 16 + uu) * F16_PITCH) = v[c][uu];
      }
    }
    __syncthreads();
    if (tid == 0 && k >= 1) {
      // the other stage held the previous tile's index box: once its store has read it, fetch the
      // next tile into it -- overlaps with pass 2
      const int nt = tile + stride;
      if (nt < a.ntiles) {
        tma::wait_group_read<0>();
        issue_load((k + 1) % NST16, nt);
      }
    }

    // ---- pass 2: DCT along the rows, quantise, pack -------------------------------------------
    {
      uint32_t res[3][8];          // per channel: (kA, kB) bytes of i = 2m in the low half, i = 2m + 1 in the high half
      const float* fr = F + u2 * F16_PITCH + bp2 * 32;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float2 v[16];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          unsigned long long p0, p1;
          asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];"
                       : "=l"(p0), "=l"(p1)
                       : "r"(tma::smem_u32(fr + c * 16 * F16_PITCH + 4 * m)));
          v[2 * m] = f2_from(p0);
          v[2 * m + 1] = f2_from(p1);
        }
        dct16_fwd<float2, true>(v);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          int kq[4];                                   // kA(2m), kB(2m), kA(2m+1), kB(2m+1)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int i = 2 * m + h;
            const float sc = M16F::sgn(i) > 0 ? qs[c][M16F::exp(i) - min_exp16()] : -qs[c][M16F::exp(i) - min_exp16()];
            const float2 t2 = P::mul(v[i], f2(sc));            // exact: power of two (times 1/q when q is one)
            float tx = t2.x, ty = t2.y;
            if (!QPOW2) {
              tx = __fdiv_rn(tx, qf);                          // src/deadzone.py:98  x / Q_step
              ty = __fdiv_rn(ty, qf);
            }
            kq[2 * h] = __float2int_rz(tx);                    // truncation = dead zone
            kq[2 * h + 1] = __float2int_rz(ty);
          }
          res[c][m] = pack4(kq[0], kq[1], kq[2], kq[3]);
        }
      }
      // box [i][j = u][16 blocks x 3 B]: this thread's 6 bytes (A.Y A.Co A.Cg B.Y B.Co B.Cg) at 6 * bp
      unsigned char* ob = tb + u2 * 48 + 6 * bp2;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t y = res[0][i >> 1] >> (16 * (i & 1)), co = res[1][i >> 1] >> (16 * (i & 1)),
                       cg = res[2][i >> 1] >> (16 * (i & 1));
        // bytes: y = (A.Y, B.Y), co = (A.Co, B.Co), cg = (A.Cg, B.Cg)
        const uint32_t w0 = (__byte_perm(__byte_perm(y, co, 0x0040), __byte_perm(cg, y, 0x0050), 0x5410)) ^ 0x80808080u;  // A.Y A.Co A.Cg B.Y
        const uint32_t h2 = (__byte_perm(co, cg, 0x0051)) ^ 0x00008080u;                                                  // B.Co B.Cg
        unsigned short* o = reinterpret_cast<unsigned short*>(ob + i * (16 * 48));
        o[0] = static_cast<unsigned short>(w0);
        o[1] = static_cast<unsigned short>(w0 >> 16);
        o[2] = static_cast<unsigned short>(h2);
        if (STATS) {
          const uint32_t d0 = __vabsdiffu4(w0, 0x80808080u), d1 = __vabsdiffu4(h2 & 0xffffu, 0x00008080u) & 0xffffu;
          st_abs = __dp4a(d0, 0x01010101u, __dp4a(d1, 0x01010101u, st_abs));
          st_nz += __popc((d0 | ((d0 & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u) +
                   __popc((d1 | ((d1 & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x00008080u);
        }
      }
    }
    tma::fence_proxy_async();
    __syncthreads();

    if (tid == 0) issue_store(s, tile);
  }
  if (tid == 0) tma::wait_group<0>();
  if (STATS) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      st_nz += __shfl_xor_sync(0xffffffffu, st_nz, o);
      st_abs += __shfl_xor_sync(0xffffffffu, st_abs, o);
    }
    if ((tid & 31) == 0) {
      atomicAdd(a.stats + VCFB_STAT_NONZERO, (unsigned long long)st_nz);
      atomicAdd(a.stats + VCFB_STAT_SUMABS, (unsigned long long)st_abs);
    }
  }
}


// Index planes sub[j*ny + y, i*nx + x, c]: dims (x bytes, j, y, i, frame) -> smem box [i][j][48]
bool make_idx_map16(CUtensorMap* m, const Geom& g, int n, const void* base) {
  const uint64_t si = uint64_t(g.nx) * 3, sy = uint64_t(g.Wp) * 3, sj = uint64_t(g.ny) * g.Wp * 3,
                 sf = uint64_t(g.Hp) * g.Wp * 3;
  const uint64_t dims[5] = {uint64_t(g.nx) * 3, 16, uint64_t(g.ny), 16, uint64_t(n)};
  const uint64_t str[4] = {sj, sy, si, sf};
  const uint32_t box[5] = {48, 16, 1, 16, 1};
  return tma::make_map(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, const_cast<void*>(base), dims, str, box);
}

template <bool QPOW2, bool YCRCB>
int launch_enc16(const CUtensorMap& in_map, const CUtensorMap& out_map, const Enc16Args& ea, cudaStream_t s) {
  void (*kern)(const CUtensorMap, const CUtensorMap, const Enc16Args) =
      ea.stats ? enc16_kernel<QPOW2, YCRCB, true> : enc16_kernel<QPOW2, YCRCB, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ENC16_SMEM);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(enc16)");
  int grid = sm_count() * 3;
  if (grid > ea.ntiles) grid = ea.ntiles;
  note_kernel("enc16_fast");
  kern<<<grid, ENC16_THREADS, ENC16_SMEM, s>>>(in_map, out_map, ea);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "enc16_kernel launch");
  return VCFB_OK;
}


}  // namespace

// Returns VCFB_E_UNSUPP when the request is outside this fast path (the caller then uses the
// general kernel), VCFB_OK after a launch, or an error.
int launch_encode_fast16(const EncArgs& a, cudaStream_t s) {
  if (a.flags & (VCFB_F_NO_SUBBANDS | VCFB_F_PERCEPTUAL | VCFB_F_FP64 | VCFB_F_CONTRACT)) return VCFB_E_UNSUPP;
  if (getenv("VCFB_NO_FAST16")) return VCFB_E_UNSUPP;          // development knob
  const Geom& g = a.g;
  if (a.stats && (a.flags & VCFB_F_HIST)) {
    // the histogram takes the streaming pass over the stored indices (kernels_stats.cu)
    if (reinterpret_cast<uintptr_t>(a.idx) & 15) return VCFB_E_UNSUPP;
    EncArgs b = a;
    b.stats = nullptr;
    int rc = launch_encode_fast16(b, s);
    if (rc) return rc;
    return launch_index_stats(a.idx, (long long)a.n_frames * g.Hp * g.Wp * 3, true, a.stats, s);
  }
  if (g.W % T16_W != 0 || g.left != 0 || g.nx % 16 != 0) return VCFB_E_UNSUPP;
  if ((reinterpret_cast<uintptr_t>(a.rgb) & 15) || (reinterpret_cast<uintptr_t>(a.idx) & 15)) return VCFB_E_UNSUPP;
  if (!tma::encode_tiled_fn()) return VCFB_E_UNSUPP;
  CUtensorMap in_map, out_map;
  if (!make_rgb_map16(&in_map, g, a.n_frames, a.rgb)) return VCFB_E_UNSUPP;
  if (!make_idx_map16(&out_map, g, a.n_frames, a.idx)) return VCFB_E_UNSUPP;
  Enc16Args ea;
  ea.tiles_x = g.Wp / T16_W;
  ea.ny = g.ny;
  ea.top = g.top;
  const long long nt = (long long)a.n_frames * g.ny * ea.tiles_x;
  if (nt > 0x7fffffffLL - (1 << 20)) return VCFB_E_UNSUPP;
  ea.ntiles = int(nt);
  ea.q = float(a.q);
  const bool ycrcb = a.color == VCFB_COLOR_YCRCB;
  for (int u = 0; u < 16; ++u)
    for (int c = 0; c < 3; ++c) {
      const int cexp = ycrcb ? 0 : ((c == 1) ? -1 : -2);         // 2*Co, 4*Y, 4*Cg; YCrCb is not lazy
      double sc = M16F::sgn(u) * p2(M16F::exp(u) + cexp + min_exp16());
      if (a.q_pow2) sc *= a.inv_q;
      ea.qtab[u][c] = float(sc);                                  // a power of two: exact
    }
  ea.stats = a.stats;
  if (a.stats) {
    int rc = launch_add_count(a.stats, VCFB_STAT_NINDICES, (unsigned long long)a.n_frames * g.Hp * g.Wp * 3, s);
    if (rc) return rc;
  }
  if (a.q_pow2) return ycrcb ? launch_enc16<true, true>(in_map, out_map, ea, s) : launch_enc16<true, false>(in_map, out_map, ea, s);
  return ycrcb ? launch_enc16<false, true>(in_map, out_map, ea, s) : launch_enc16<false, false>(in_map, out_map, ea, s);
}

}  // namespace vcfb

namespace vcfb {

// float64 decode, B = 16: VCFB_E_UNSUPP when outside this fast path
int launch_decode_fast16(const DecArgs& a, cudaStream_t s) {
  if (!(a.flags & VCFB_F_FP64)) return launch_decode_fast16_f32(a, s);      // float32 = the fast mode (kernels_b16f.cu)
  return b16::launch_decode16<double, true>(a, s, "dec16_fast");
}

}  // namespace vcfb
