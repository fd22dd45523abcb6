// Fused rate/distortion sweep: one pass over a frame per block size evaluates
// EVERY quantisation step -- the forward transform runs once, its coefficients stay in registers,
// and for each q the tile is quantised, dequantised and taken back through the reference's float64
// decode chain in shared memory; only the statistics leave the SM (3 B/pixel of HBM traffic per
// block size instead of 12 B/pixel per (B, q) point).
//
// What it replaces: the loop of src/2D-DCT.py:533-579 (optimize_block_size: per block size
// from_RGB -> space_analyze -> quantize -> dequantize -> space_synthesize -> to_RGB -> clip -> RMSE)
// and the per-point process runs of src/RDE.py:68-118.  Two dequantiser semantics exist in the
// reference and both are offered:
//   * default: what a decoder sees after the files were written -- the index wrapped to uint8
//     (src/2D-DCT.py:361), read back as int16 (:398) and multiplied in int16 (:410);
//   * VCFB_F_NOWRAP: the in-process loop of optimize_block_size, which dequantises the quantiser's own
//     (int64) indices before they are ever narrowed (:560-566).
// The two differ only where |k| > 127 (DC of large blocks at small q).
//
// Arithmetic is that of the general kernels, operation for operation: float32 pocketfft codelets
// forward (the reference's own precision, :276), float64 codelets inverse, so every statistic equals
// the one the separate encode + decode kernels accumulate -- tests/test_gpu_rd.py.
//
// Tile = one block row x TW pixels, one work item per thread in every phase:
//   forward 1   (channel, pixel column)      colour, DCT down the column          -> F (float, smem)
//   forward 2   (channel, block, row u)      DCT along the row                    -> registers (B floats)
//   per q:  A   (channel, block, row u)      quantise, statistics, dequantise     -> G (double, smem)
//           B1  (channel, coefficient column) inverse DCT, axis 0, in place
//           B2  (channel, block, pixel row)  inverse DCT, axis 1, in place
//           P   (pixel)                      to_RGB, +128, truncate, clip, SSE against the tile's own input
// F aliases G.  Pitches of TW + 1 elements and row-major lanes make every phase bank-conflict free.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "dct_codelets.cuh"

namespace vcfb {
namespace {

constexpr int MAXQ = VCFB_RD_MAX_STEPS;

struct RdArgs {
  const uint8_t* rgb;
  Geom g;
  int n_frames;
  int nq;
  int color;
  unsigned flags;
  unsigned long long* stats;   // nq x VCFB_STAT_LEN
  double q[MAXQ];
  float inv_q[MAXQ];
  int q_pow2[MAXQ];
  int q_int[MAXQ];
};

__host__ __device__ constexpr double p2(int e) {
  double r = 1.0;
  for (int i = 0; i < (e < 0 ? -e : e); ++i) r = e < 0 ? r * 0.5 : r * 2.0;
  return r;
}

template <int B> struct RdLayout {
  static constexpr int TW = B >= 32 ? 64 : 128;
  static constexpr int NT = 3 * TW;
  static constexpr int RAWP = TW * 3;
  static constexpr int GP = TW + 1;
  static constexpr int RAW_BYTES = B * RAWP;
  static constexpr int G_BYTES = 3 * B * GP * 8;
  static constexpr int SMEM = RAW_BYTES + G_BYTES + 3 * 256 * 4 + MAXQ * 8 * 4 + 64;
};

__device__ __forceinline__ unsigned warp_sum(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }   // one REDUX

// Phase A of one step for one row of B coefficients, with everything that depends only on the step a
// compile-time constant (the uniform branches otherwise sit inside the unrolled element loop).
template <int B, bool POW2, bool QINT, bool NOWRAP>
__device__ __forceinline__ void quantise_row(const float (&coef)[B], float q, float inv_q, int q_int, double qd, int poff,
                                             double* dst, unsigned* hist_c, bool do_hist, unsigned& nz, unsigned& sabs) {
  using OF = Ops<float, true>;
#pragma unroll
  for (int i = 0; i < B; ++i) {
    const float tq = POW2 ? OF::mul(coef[i], inv_q) : OF::div(coef[i], q);   // src/deadzone.py:98
    const int k = __float2int_rz(tq);                       // truncation = dead zone
    const unsigned byte = unsigned(k + poff) & 255u;        // src/2D-DCT.py:348,:361 (wraps)
    const int k8 = int((signed char)(byte ^ unsigned(poff)));   // what a decoder reads back (:398,:402): byte - 128 (or int8 without offset)
    nz += (k8 != 0);
    sabs += unsigned(abs(k8));
    if (do_hist && k8 != 0) atomicAdd(&hist_c[byte], 1u);   // the bin of the zero index is counted through nz
    double y;
    if (NOWRAP) y = QINT ? double((long long)k * q_int) : double(k) * qd;          // :565-568 (int64 indices)
    else y = QINT ? double(int(short(k8 * q_int))) : double(k8) * qd;              // int16 * python int stays int16
    dst[i] = y;
  }
}

template <int B>
__global__ void __launch_bounds__(RdLayout<B>::NT, 3) rd_sweep_kernel(const RdArgs a) {
  using L = RdLayout<B>;
  using OF = Ops<float, true>;
  using OD = Ops<double, true>;
  using DF = Dct<B, false>;
  using MF = typename DF::meta;
  using DI = Dct<B, true>;
  using MI = typename DI::meta;
  constexpr int TW = L::TW, NT = L::NT, RAWP = L::RAWP, GP = L::GP, TBX = TW / B;

  extern __shared__ __align__(16) unsigned char smem[];
  uint8_t* raw = smem;
  double* G = reinterpret_cast<double*>(smem + L::RAW_BYTES);
  float* F = reinterpret_cast<float*>(G);                       // forward intermediate, dead before the q loop
  unsigned* shist = reinterpret_cast<unsigned*>(smem + L::RAW_BYTES + L::G_BYTES);
  unsigned* sacc = shist + 3 * 256;                             // [q][8]: nz, sabs, sse r g b, sdiff, -, -

  const int tid = threadIdx.x;
  const int tile = blockIdx.x, by = blockIdx.y, f = blockIdx.z;
  const Geom g = a.g;
  const int x0 = tile * TW;
  const int bx0 = tile * TBX;
  const int nbx = min(TBX, g.nx - bx0);
  const bool do_hist = (a.flags & VCFB_F_HIST) != 0;
  const bool nowrap = (a.flags & VCFB_F_NOWRAP) != 0;
  const int poff = (a.flags & VCFB_F_NO_OFFSET) ? 0 : 128;   // the loop of optimize_block_size runs without offset (vcfb200.h)

  for (int i = tid; i < 3 * 256 + MAXQ * 8; i += NT) shist[i] = 0;

  // ---- load: B rows x TW pixels, zero outside the frame (padding, src/2D-DCT.py:216-227) ----
  {
    constexpr int CPR = RAWP / 16;
    const long long rowbytes = (long long)g.W * 3;
    for (int t = tid; t < B * CPR; t += NT) {
      const int r = t / CPR, ch = t % CPR;
      const int gy = by * B + r - g.top;
      const long long o0 = ((long long)x0 - g.left) * 3 + ch * 16;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (gy >= 0 && gy < g.H && o0 + 16 > 0 && o0 < rowbytes) {
        const uint8_t* rowp = a.rgb + (size_t(f) * g.H + gy) * size_t(rowbytes);
        if (o0 >= 0 && o0 + 16 <= rowbytes && (reinterpret_cast<uintptr_t>(rowp + o0) & 15) == 0) {
          val = __ldg(reinterpret_cast<const uint4*>(rowp + o0));
        } else {
          unsigned w[4] = {0, 0, 0, 0};
#pragma unroll
          for (int b = 0; b < 16; ++b) {
            const long long o = o0 + b;
            if (o >= 0 && o < rowbytes) w[b >> 2] |= unsigned(rowp[o]) << (8 * (b & 3));
          }
          val = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      *reinterpret_cast<uint4*>(raw + r * RAWP + ch * 16) = val;
    }
  }
  __syncthreads();

  // work item of the row phases: channel c, block bx, row u (lanes = consecutive rows)
  const int c = tid / TW;
  const int t = tid - c * TW;
  const int u = t % B, bx = t / B;
  const bool live = bx < nbx;

  // ---- forward 1: colour + DCT down each pixel column (axis 0) ----
  {
    const int x = t;
    const float cs = a.color == VCFB_COLOR_YCOCG ? (c == 1 ? 0.5f : 0.25f) : 1.0f;
    float v[B];
#pragma unroll
    for (int r = 0; r < B; ++r) {
      const uint8_t* px = raw + r * RAWP + x * 3;
      const int R = px[0], Gc = px[1], Bc = px[2];
      if (a.color == VCFB_COLOR_YCOCG) {
        v[r] = float((c == 0) ? (R + 2 * Gc + Bc - 4 * poff) : (c == 1) ? (R - Bc) : (2 * Gc - R - Bc));
      } else {
        const float r_ = float(R - poff), g_ = float(Gc - poff), b_ = float(Bc - poff);
        const float y = OF::add(OF::add(OF::mul(r_, 0.299f), OF::mul(g_, 0.587f)), OF::mul(b_, 0.114f));
        v[r] = c == 0 ? y : c == 1 ? OF::mul(OF::sub(r_, y), 0.713f) : OF::mul(OF::sub(b_, y), 0.564f);
      }
    }
    DF::template run<float, true>(v);
#pragma unroll
    for (int k = 0; k < B; ++k) F[(c * B + k) * GP + x] = OF::mul(v[k], float(MF::sgn(k) * p2(MF::exp(k))) * cs);
  }
  __syncthreads();

  // ---- forward 2: DCT along each block row (axis 1); the coefficients stay in registers ----
  float coef[B];
  {
    const float* src = F + (c * B + u) * GP + bx * B;
#pragma unroll
    for (int i = 0; i < B; ++i) coef[i] = live ? src[i] : 0.0f;
    DF::template run<float, true>(coef);
#pragma unroll
    for (int i = 0; i < B; ++i) coef[i] = OF::mul(coef[i], float(MF::sgn(i) * p2(MF::exp(i))));   // the coefficient scipy returns
  }
  __syncthreads();   // F is dead: G may be written

  for (int qi = 0; qi < a.nq; ++qi) {
    // ---- A: quantise (src/deadzone.py:98), statistics, dequantise (:115) ----
    {
      const float q = float(a.q[qi]), inv_q = a.inv_q[qi];
      const int q_int = a.q_int[qi];
      const double qd = a.q[qi];
      const bool pow2 = a.q_pow2[qi] != 0;
      unsigned nz = 0, sabs = 0;
      double* dst = G + (c * B + u) * GP + bx * B;
      unsigned* hc = shist + c * 256;
      // (lanes of blocks beyond the frame's right edge hold zero coefficients: they store zeros into their own
      //  part of the tile and count nothing)
#define VCFB_RD_ROW(P2, QI, NW) quantise_row<B, P2, QI, NW>(coef, q, inv_q, q_int, qd, poff, dst, hc, do_hist, nz, sabs)
      if (nowrap) {
        if (pow2) { if (q_int) VCFB_RD_ROW(true, true, true); else VCFB_RD_ROW(true, false, true); }
        else      { if (q_int) VCFB_RD_ROW(false, true, true); else VCFB_RD_ROW(false, false, true); }
      } else {
        if (pow2) { if (q_int) VCFB_RD_ROW(true, true, false); else VCFB_RD_ROW(true, false, false); }
        else      { if (q_int) VCFB_RD_ROW(false, true, false); else VCFB_RD_ROW(false, false, false); }
      }
#undef VCFB_RD_ROW
      // nz <= 32 and sabs <= 128 * 32 per lane: one reduction carries both; c is warp-uniform (TW is a multiple of 32)
      const unsigned packed = warp_sum((nz << 20) | sabs);
      const unsigned nlive = unsigned(__popc(__ballot_sync(0xffffffffu, live)));
      if ((tid & 31) == 0) {
        nz = packed >> 20;
        sabs = packed & 0xFFFFFu;
        if (nz) atomicAdd(&sacc[qi * 8 + 0], nz);
        if (sabs) atomicAdd(&sacc[qi * 8 + 1], sabs);
        if (do_hist) {
          const unsigned zeros = nlive * B - nz;
          if (zeros) atomicAdd(&hc[poff], zeros);
        }
      }
    }
    __syncthreads();

    // ---- B1: inverse DCT along axis 0 (coefficient columns), in place ----
    {
      if (do_hist) {   // flush the histogram of this step (every bin has one owner)
        for (int i = tid; i < 3 * 256; i += NT) {
          const unsigned h = shist[i];
          if (h) {
            atomicAdd(a.stats + size_t(qi) * VCFB_STAT_LEN + VCFB_STAT_HIST + i, (unsigned long long)h);
            shist[i] = 0;
          }
        }
      }
      const int x = t;
      if (x < nbx * B) {
        double v[B];
        double* col = G + (c * B) * GP + x;
#pragma unroll
        for (int k = 0; k < B; ++k) v[k] = col[k * GP];
        DI::template run<double, true>(v);
#pragma unroll
        for (int r = 0; r < B; ++r) col[r * GP] = OD::mul(v[r], MI::sgn(r) * p2(MI::exp(r)));
      }
    }
    __syncthreads();

    // ---- B2: inverse DCT along axis 1 (pixel rows of each block), in place ----
    if (live) {
      double v[B];
      double* row = G + (c * B + u) * GP + bx * B;
#pragma unroll
      for (int i = 0; i < B; ++i) v[i] = row[i];
      DI::template run<double, true>(v);
#pragma unroll
      for (int i = 0; i < B; ++i) row[i] = OD::mul(v[i], MI::sgn(i) * p2(MI::exp(i)));
    }
    __syncthreads();

    // ---- P: to_RGB, +128, truncate, clip (:449-466); SSE against the input (src/RDE.py:41-49) ----
    // thread = pixel column t, rows c, c + 3, c + 6, ...
    {
      unsigned sse[3] = {0, 0, 0};
      int sdiff = 0;
      const int x = t;
      const int gx = x0 + x - g.left;
      if (x < nbx * B && gx >= 0 && gx < g.W) {
#pragma unroll
        for (int j = 0; j < (B + 2) / 3; ++j) {
          const int r = c + 3 * j;
          const int gy = by * B + r - g.top;
          if (r >= B || gy < 0 || gy >= g.H) continue;
          const double c0 = G[(0 * B + r) * GP + x], c1 = G[(1 * B + r) * GP + x], c2 = G[(2 * B + r) * GP + x];
          double R, Gv, Bv;
          if (a.color == VCFB_COLOR_YCOCG) {   // Y + Co - Cg ; Y + Cg ; Y - Co - Cg, left to right
            R = OD::sub(OD::add(c0, c1), c2);
            Gv = OD::add(c0, c2);
            Bv = OD::sub(OD::sub(c0, c1), c2);
          } else {                             // oracle ycrcb_to_rgb_float
            R = OD::add(c0, OD::mul(c1, 1.403));
            Gv = OD::add(OD::add(c0, OD::mul(c1, -0.714)), OD::mul(c2, -0.344));
            Bv = OD::add(c0, OD::mul(c2, 1.773));
          }
          const double yoff = double(poff);
          const int v[3] = {min(max(__double2int_rz(OD::add(R, yoff)), 0), 255),
                            min(max(__double2int_rz(OD::add(Gv, yoff)), 0), 255),
                            min(max(__double2int_rz(OD::add(Bv, yoff)), 0), 255)};
          const uint8_t* px = raw + r * RAWP + x * 3;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int d = int(px[k]) - v[k];
            sse[k] += unsigned(d * d);
            sdiff += d;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const unsigned sv = warp_sum(sse[k]);
        if ((tid & 31) == 0 && sv) atomicAdd(&sacc[qi * 8 + 2 + k], sv);
      }
      const unsigned sd = warp_sum(unsigned(sdiff));             // two's complement sum
      if ((tid & 31) == 0 && sd) atomicAdd(&sacc[qi * 8 + 5], sd);
    }
    __syncthreads();
  }

  // ---- per-CTA totals -> global (integer atomics: order-independent, bit-reproducible) ----
  int rows = 0;
  for (int r = 0; r < B; ++r) {
    const int gy = by * B + r - g.top;
    rows += (gy >= 0 && gy < g.H);
  }
  const int gx0 = max(0, x0 - g.left), gx1 = min(g.W, x0 - g.left + nbx * B);
  const unsigned long long nsamp = (unsigned long long)rows * (gx1 > gx0 ? gx1 - gx0 : 0) * 3;
  for (int i = tid; i < a.nq * 8; i += NT) {
    const int qi = i >> 3, k = i & 7;
    unsigned long long* st = a.stats + size_t(qi) * VCFB_STAT_LEN;
    const unsigned v = sacc[i];
    switch (k) {
      case 0: if (v) atomicAdd(st + VCFB_STAT_NONZERO, (unsigned long long)v); break;
      case 1: if (v) atomicAdd(st + VCFB_STAT_SUMABS, (unsigned long long)v); break;
      case 2: case 3: case 4: if (v) atomicAdd(st + VCFB_STAT_SSE_R + (k - 2), (unsigned long long)v); break;
      case 5: if (v) atomicAdd(st + VCFB_STAT_SUMDIFF, (unsigned long long)(long long)(int)v); break;
      case 6: atomicAdd(st + VCFB_STAT_NINDICES, (unsigned long long)(nbx * B * B * 3)); break;
      case 7: if (nsamp) atomicAdd(st + VCFB_STAT_NSAMPLES, nsamp); break;
    }
  }
}

template <int B>
int launch_rd(const RdArgs& a, cudaStream_t s) {
  using L = RdLayout<B>;
  auto kern = rd_sweep_kernel<B>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(rd_sweep)");
  dim3 grid((a.g.Wp + L::TW - 1) / L::TW, a.g.ny, a.n_frames);
  note_kernel("rd_sweep");
  kern<<<grid, L::NT, L::SMEM, s>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "rd_sweep_kernel launch");
  return VCFB_OK;
}

}  // namespace

int launch_rd_sweep(const uint8_t* rgb, const Geom& g, int n_frames, int B, const double* qs, int nq, int color,
                    unsigned flags, unsigned long long* stats, cudaStream_t s) {
  RdArgs a;
  memset(&a, 0, sizeof(a));
  a.rgb = rgb;
  a.g = g;
  a.n_frames = n_frames;
  a.nq = nq;
  a.color = color;
  a.flags = flags;
  a.stats = stats;
  for (int i = 0; i < nq; ++i) {
    const double q = qs[i];
    int e2;
    a.q[i] = q;
    a.q_pow2[i] = frexp(q, &e2) == 0.5;
    a.inv_q[i] = float(1.0 / q);
    a.q_int[i] = (q == floor(q) && q < 32768.0) ? int(q) : 0;
    // float32 path: the reference divides the float32 coefficient by the python number q; numpy's weak scalar
    // keeps the array's float32, so q is rounded to float32 first -- as the general kernel's T(a.q) does.
  }
  switch (B) {
    case 4: return launch_rd<4>(a, s);
    case 8: return launch_rd<8>(a, s);
    case 16: return launch_rd<16>(a, s);
    case 32: return launch_rd<32>(a, s);
  }
  return VCFB_E_UNSUPP;
}

}  // namespace vcfb
