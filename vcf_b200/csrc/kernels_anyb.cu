// Block sizes without a straight-line codelet: B = 2, 64, 128.
//
// The reference's -L search (src/2D-DCT.py:533-579) tries B = 2^i for i = 1..7; 4..32 run on the
// generated codelets (kernels_general.cu and the fast paths).  The three remaining sizes are rare
// (one image per -L call) and big (a 128 x 128 x 3 float64 block is 393 KB), so they run as a
// short sequence of simple kernels over a per-frame scratch plane in HBM instead of a fused tile
// kernel: every thread interprets pocketfft's traced operation list (dag_programs.inc, generated
// by codegen/gen_dag_programs.py from the same DAG as the codelets and checked bit for bit
// against scipy) on one column / one row of one block, with individually rounded operations.
// Bit-exact in the same sense as the general kernels; every flag they take is taken here.
//
//   encode:  anyb_enc_cols   (channel, padded column x, block row): colour, DCT down the column
//            anyb_enc_rows   (channel, row y, block): DCT along the row, -p weights, quantise, store
//   decode:  anyb_dec_cols   (channel, coefficient column, block row): dequantise, inverse DCT (axis 0)
//            anyb_dec_rows   (channel, row, block): inverse DCT (axis 1), in place
//            anyb_dec_pixels (pixel): to_RGB, +128, float output, clip, truncate, SSE
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "exact_ops.cuh"
#include "dag_programs.inc"

namespace vcfb {
namespace {

constexpr int NT = 128;

template <typename T> __device__ __forceinline__ const T* dag_consts(const DagProgram& p);
template <> __device__ __forceinline__ const float* dag_consts<float>(const DagProgram& p) { return p.c32; }
template <> __device__ __forceinline__ const double* dag_consts<double>(const DagProgram& p) { return p.c64; }

template <typename T> __device__ __forceinline__ int to_int_rz(T x);
template <> __device__ __forceinline__ int to_int_rz<float>(float x) { return __float2int_rz(x); }
template <> __device__ __forceinline__ int to_int_rz<double>(double x) { return __double2int_rz(x); }

__device__ __forceinline__ double pow2i(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }

// Runs the program on the values in s[0..n) (slots); afterwards output k is out_scale(k) * s[out_slot(k)].
template <typename T, bool EXACT>
__device__ __forceinline__ void dag_run(const DagProgram& p, T* s) {
  using O = Ops<T, EXACT>;
  const T* cv = dag_consts<T>(p);
  for (int n = 0; n < p.nops; ++n) {
    const uint4 op = p.ops[n];
    const T a = s[op.z];
    T r;
    switch (op.x & 0xFFu) {
      case 1: r = O::add(a, s[op.w]); break;
      case 2: r = O::sub(a, s[op.w]); break;
      case 3: r = O::mul(a, cv[op.w]); break;
      default: {
        const T b = s[op.w];
        r = O::fma(a, cv[(op.x >> 8) & 0xFFFFu], ((op.x >> 24) & 1u) ? O::neg(b) : b);   // product exact: one rounding
      }
    }
    s[op.y] = r;
  }
}
template <typename T>
__device__ __forceinline__ T dag_out(const DagProgram& p, const T* s, int k) {
  // sign * 2^e: an exact factor (lazy power-of-two scale of the traced DAG)
  return Ops<T, true>::mul(s[p.outs[3 * k]], T(double(p.outs[3 * k + 1]) * pow2i(p.outs[3 * k + 2])));
}

// forward colour transform of the centred pixel, channel c (kernels_general.cu::color_fwd, not lazy)
template <typename T, bool EXACT>
__device__ __forceinline__ T color_fwd(int color, int c, int R, int G, int Bc, int poff) {
  using O = Ops<T, EXACT>;
  if (color == VCFB_COLOR_YCOCG) {   // exact for 8-bit input in any evaluation order (src/2D-DCT.py:292-298)
    const int v = (c == 0) ? (R + 2 * G + Bc - 4 * poff) : (c == 1) ? (R - Bc) : (2 * G - R - Bc);
    return T(v) * T(c == 1 ? 0.5 : 0.25);
  }
  const T r = T(R - poff), g = T(G - poff), b = T(Bc - poff);
  const T y = O::add(O::add(O::mul(r, T(0.299)), O::mul(g, T(0.587))), O::mul(b, T(0.114)));
  if (c == 0) return y;
  if (c == 1) return O::mul(O::sub(r, y), T(0.713));
  return O::mul(O::sub(b, y), T(0.564));
}

// scratch plane of one frame: F[c][y][x], Hp x Wp per channel
template <typename T> __device__ __forceinline__ T* plane(void* scratch, const Geom& g, int c) {
  return reinterpret_cast<T*>(scratch) + size_t(c) * g.Hp * g.Wp;
}

template <typename T, bool EXACT>
__global__ void __launch_bounds__(NT) anyb_enc_cols(const EncArgs a, int B, int f, void* scratch) {
  const Geom g = a.g;
  const long long item = (long long)blockIdx.x * NT + threadIdx.x;
  const long long total = 3LL * g.ny * g.Wp;
  if (item >= total) return;
  const int x = int(item % g.Wp);
  const int by = int((item / g.Wp) % g.ny);
  const int c = int(item / ((long long)g.Wp * g.ny));
  const DagProgram p = dag_program(B, false);
  T s[DAGP_MAX_SLOTS];
  const int gx = x - g.left;
  const int poff = (a.flags & VCFB_F_NO_OFFSET) ? 0 : 128;
  for (int r = 0; r < B; ++r) {
    const int gy = by * B + r - g.top;
    T v;
    if (gy >= 0 && gy < g.H && gx >= 0 && gx < g.W) {
      const uint8_t* px = a.rgb + ((size_t(f) * g.H + gy) * g.W + gx) * 3;
      v = color_fwd<T, EXACT>(a.color, c, px[0], px[1], px[2], poff);
    } else {
      v = color_fwd<T, EXACT>(a.color, c, 0, 0, 0, poff);   // zero padding BEFORE the -128 (:216-227, :292)
    }
    s[r] = v;
  }
  dag_run<T, EXACT>(p, s);
  T* F = plane<T>(scratch, g, c);
  for (int u = 0; u < B; ++u) F[(size_t(by) * B + u) * g.Wp + x] = dag_out<T>(p, s, u);
}

// KEEP: the coefficients go back into the plane (fused rate/distortion sweep) instead of being quantised
template <typename T, bool EXACT, bool KEEP = false>
__global__ void __launch_bounds__(NT) anyb_enc_rows(const EncArgs a, int B, int f, void* scratch) {
  using O = Ops<T, EXACT>;
  const Geom g = a.g;
  const long long item = (long long)blockIdx.x * NT + threadIdx.x;
  const long long total = 3LL * g.Hp * g.nx;
  unsigned nz = 0, sabs = 0;
  const bool do_stats = a.stats != nullptr;
  const bool do_hist = do_stats && (a.flags & VCFB_F_HIST) != 0;
  if (item < total) {
    const int bx = int(item % g.nx);
    const int y = int((item / g.nx) % g.Hp);
    const int c = int(item / ((long long)g.nx * g.Hp));
    const int by = y / B, u = y % B;
    const bool nosub = (a.flags & VCFB_F_NO_SUBBANDS) != 0;
    const bool percep = (a.flags & VCFB_F_PERCEPTUAL) != 0;
    const DagProgram p = dag_program(B, false);
    T s[DAGP_MAX_SLOTS];
    T* src = plane<T>(scratch, g, c) + size_t(y) * g.Wp + size_t(bx) * B;
    for (int i = 0; i < B; ++i) s[i] = src[i];
    dag_run<T, EXACT>(p, s);
    if (KEEP) {
      for (int i = 0; i < B; ++i) src[i] = dag_out<T>(p, s, i);
      return;
    }
    const T q = T(a.q), inv_q = T(a.inv_q);
    const double* wt = percep ? a.weights + (c ? B * B : 0) + u * B : nullptr;
    for (int i = 0; i < B; ++i) {
      T coef = dag_out<T>(p, s, i);
      if (percep) coef = T(double(coef) * wt[i]);                 // src/2D-DCT.py:322-324
      const T tq = a.q_pow2 ? O::mul(coef, inv_q) : O::div(coef, q);   // src/deadzone.py:98
      const int k = to_int_rz<T>(tq);
      const int poff = (a.flags & VCFB_F_NO_OFFSET) ? 0 : 128;
      const unsigned byte = unsigned(k + poff) & 255u;            // src/2D-DCT.py:348,:361 (wraps)
      size_t row, col;
      if (nosub) {
        row = y;
        col = size_t(bx) * B + i;
      } else {
        row = size_t(u) * g.ny + by;
        col = size_t(i) * g.nx + bx;
      }
      a.idx[((size_t(f) * g.Hp + row) * g.Wp + col) * 3 + c] = uint8_t(byte);
      if (do_stats) {
        const int kk = poff ? int(byte) - 128 : int((signed char)byte);
        nz += (kk != 0);
        sabs += unsigned(kk < 0 ? -kk : kk);
        if (do_hist) atomicAdd(a.stats + VCFB_STAT_HIST + c * 256 + byte, 1ULL);
      }
    }
  }
  if (do_stats && !KEEP) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      nz += __shfl_xor_sync(0xffffffffu, nz, o);
      sabs += __shfl_xor_sync(0xffffffffu, sabs, o);
    }
    if ((threadIdx.x & 31) == 0) {
      if (nz) atomicAdd(a.stats + VCFB_STAT_NONZERO, (unsigned long long)nz);
      if (sabs) atomicAdd(a.stats + VCFB_STAT_SUMABS, (unsigned long long)sabs);
    }
    if (item == 0) atomicAdd(a.stats + VCFB_STAT_NINDICES, (unsigned long long)g.Hp * g.Wp * 3);
  }
}

// Fused sweep, one step: quantise the float32 coefficient plane (block layout), statistics of the indices,
// dequantise into the float64 plane.  Semantics of kernels_rd.cu (wrapped / VCFB_F_NOWRAP).
__global__ void __launch_bounds__(NT) anyb_rd_quant(const float* coef, double* deq, long long n_per_channel, double qd,
                                                    int q_pow2, int q_int, unsigned flags, unsigned long long* stats) {
  using OF = Ops<float, true>;
  const long long item = (long long)blockIdx.x * NT + threadIdx.x;
  unsigned nz = 0, sabs = 0;
  if (item < 3 * n_per_channel) {
    const int c = int(item / n_per_channel);
    const float v = coef[item];
    const float tq = q_pow2 ? OF::mul(v, float(1.0 / qd)) : OF::div(v, float(qd));
    const int k = __float2int_rz(tq);
    const int poff = (flags & VCFB_F_NO_OFFSET) ? 0 : 128;
    const unsigned byte = unsigned(k + poff) & 255u;
    const int k8 = poff ? int(byte) - 128 : int((signed char)byte);
    nz = (k8 != 0);
    sabs = unsigned(k8 < 0 ? -k8 : k8);
    if (flags & VCFB_F_HIST) atomicAdd(stats + VCFB_STAT_HIST + c * 256 + byte, 1ULL);
    double y;
    if (flags & VCFB_F_NOWRAP) y = q_int ? double((long long)k * q_int) : double(k) * qd;
    else y = q_int ? double(int(short(k8 * q_int))) : double(k8) * qd;
    deq[item] = y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nz += __shfl_xor_sync(0xffffffffu, nz, o);
    sabs += __shfl_xor_sync(0xffffffffu, sabs, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (nz) atomicAdd(stats + VCFB_STAT_NONZERO, (unsigned long long)nz);
    if (sabs) atomicAdd(stats + VCFB_STAT_SUMABS, (unsigned long long)sabs);
  }
  if (item == 0) atomicAdd(stats + VCFB_STAT_NINDICES, (unsigned long long)(3 * n_per_channel));
}

// FROM_PLANE: the dequantised values are already in the plane (fused sweep), block layout
template <typename T, bool EXACT, bool FROM_PLANE = false>
__global__ void __launch_bounds__(NT) anyb_dec_cols(const DecArgs a, int B, int f, void* scratch) {
  const Geom g = a.g;
  const long long item = (long long)blockIdx.x * NT + threadIdx.x;
  const long long total = 3LL * g.ny * g.Wp;
  if (item >= total) return;
  const int x = int(item % g.Wp);
  const int by = int((item / g.Wp) % g.ny);
  const int c = int(item / ((long long)g.Wp * g.ny));
  const int bx = x / B, i = x % B;
  const bool nosub = (a.flags & VCFB_F_NO_SUBBANDS) != 0;
  const bool percep = (a.flags & VCFB_F_PERCEPTUAL) != 0;
  const DagProgram p = dag_program(B, true);
  T s[DAGP_MAX_SLOTS];
  if (FROM_PLANE) {
    const T* src = plane<T>(scratch, g, c) + size_t(by) * B * g.Wp + x;
    for (int u = 0; u < B; ++u) s[u] = src[size_t(u) * g.Wp];
  } else
  for (int u = 0; u < B; ++u) {
    size_t row, col;
    if (nosub) {
      row = size_t(by) * B + u;
      col = x;
    } else {
      row = size_t(u) * g.ny + by;
      col = size_t(i) * g.nx + bx;
    }
    const int k = int(a.idx[((size_t(f) * g.Hp + row) * g.Wp + col) * 3 + c]) - 128;   // :398,:402
    T y;
    if (a.q_int) y = T(int(short(k * a.q_int)));        // int16 * python int stays int16 (wraps)
    else y = T(double(k) * a.q);
    if (percep) {                                       // :421-435, stored back into the int16 array
      const float fv = float(double(float(y)) / a.weights[(c ? B * B : 0) + u * B + i]);
      y = T(int(short(__float2int_rz(fv))));
    }
    s[u] = y;
  }
  dag_run<T, EXACT>(p, s);
  T* F = plane<T>(scratch, g, c);
  for (int r = 0; r < B; ++r) F[(size_t(by) * B + r) * g.Wp + x] = dag_out<T>(p, s, r);
}

template <typename T, bool EXACT>
__global__ void __launch_bounds__(NT) anyb_dec_rows(const DecArgs a, int B, void* scratch) {
  const Geom g = a.g;
  const long long item = (long long)blockIdx.x * NT + threadIdx.x;
  const long long total = 3LL * g.Hp * g.nx;
  if (item >= total) return;
  const int bx = int(item % g.nx);
  const int y = int((item / g.nx) % g.Hp);
  const int c = int(item / ((long long)g.nx * g.Hp));
  const DagProgram p = dag_program(B, true);
  T s[DAGP_MAX_SLOTS];
  T* src = plane<T>(scratch, g, c) + size_t(y) * g.Wp + size_t(bx) * B;
  for (int i = 0; i < B; ++i) s[i] = src[i];
  dag_run<T, EXACT>(p, s);
  for (int i = 0; i < B; ++i) src[i] = dag_out<T>(p, s, i);
}

template <typename T, bool EXACT>
__global__ void __launch_bounds__(NT) anyb_dec_pixels(const DecArgs a, int f, void* scratch) {
  using O = Ops<T, EXACT>;
  const Geom g = a.g;
  const long long item = (long long)blockIdx.x * NT + threadIdx.x;
  const long long total = (long long)g.H * g.W;
  const bool do_sse = a.stats != nullptr && a.original != nullptr;
  unsigned sse[3] = {0, 0, 0};
  int sdiff = 0;
  if (item < total) {
    const int gx = int(item % g.W), gy = int(item / g.W);
    const size_t o = (size_t(gy) + g.top) * g.Wp + gx + g.left;     // crop (:444)
    const T c0 = plane<T>(scratch, g, 0)[o], c1 = plane<T>(scratch, g, 1)[o], c2 = plane<T>(scratch, g, 2)[o];
    T R, G, Bv;
    if (sizeof(T) == 8 && (a.flags & VCFB_F_SYNTH_F32)) {           // upstream variant, see kernels_general.cu
      using OF = Ops<float, true>;
      const float f0 = float(c0), f1 = float(c1), f2 = float(c2);
      float Rf, Gf, Bf;
      if (a.color == VCFB_COLOR_YCOCG) {
        Rf = OF::sub(OF::add(f0, f1), f2);
        Gf = OF::add(f0, f2);
        Bf = OF::sub(OF::sub(f0, f1), f2);
      } else {
        Rf = OF::add(f0, OF::mul(f1, 1.403f));
        Gf = OF::add(OF::add(f0, OF::mul(f1, -0.714f)), OF::mul(f2, -0.344f));
        Bf = OF::add(f0, OF::mul(f2, 1.773f));
      }
      R = T(OF::add(Rf, 128.0f));
      G = T(OF::add(Gf, 128.0f));
      Bv = T(OF::add(Bf, 128.0f));
    } else {
      if (a.color == VCFB_COLOR_YCOCG) {   // Y + Co - Cg ; Y + Cg ; Y - Co - Cg, left to right (:449)
        R = O::sub(O::add(c0, c1), c2);
        G = O::add(c0, c2);
        Bv = O::sub(O::sub(c0, c1), c2);
      } else {                             // oracle ycrcb_to_rgb_float
        R = O::add(c0, O::mul(c1, T(1.403)));
        G = O::add(O::add(c0, O::mul(c1, T(-0.714))), O::mul(c2, T(-0.344)));
        Bv = O::add(c0, O::mul(c2, T(1.773)));
      }
      const T yoff = T((a.flags & VCFB_F_NO_OFFSET) ? 0 : 128);
      R = O::add(R, yoff);                 // :454 (:572 in the loop of optimize_block_size)
      G = O::add(G, yoff);
      Bv = O::add(Bv, yoff);
    }
    const size_t po = ((size_t(f) * g.H + gy) * g.W + gx) * 3;
    if (a.y_out) {
      T* yo = reinterpret_cast<T*>(a.y_out) + po;
      yo[0] = R; yo[1] = G; yo[2] = Bv;
    }
    const int v[3] = {min(max(to_int_rz<T>(R), 0), 255), min(max(to_int_rz<T>(G), 0), 255),
                      min(max(to_int_rz<T>(Bv), 0), 255)};           // :466 (truncation, clip)
    if (a.rgb) {
      a.rgb[po] = uint8_t(v[0]); a.rgb[po + 1] = uint8_t(v[1]); a.rgb[po + 2] = uint8_t(v[2]);
    }
    if (do_sse) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int d = int(a.original[po + c]) - v[c];
        sse[c] = unsigned(d * d);
        sdiff += d;
      }
    }
  }
  if (do_sse) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int c = 0; c < 3; ++c) sse[c] += __shfl_xor_sync(0xffffffffu, sse[c], o);
      sdiff += __shfl_xor_sync(0xffffffffu, sdiff, o);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (sse[c]) atomicAdd(a.stats + VCFB_STAT_SSE_R + c, (unsigned long long)sse[c]);
      if (sdiff) atomicAdd(a.stats + VCFB_STAT_SUMDIFF, (unsigned long long)(long long)sdiff);
    }
    if (item == 0) atomicAdd(a.stats + VCFB_STAT_NSAMPLES, (unsigned long long)g.H * g.W * 3);
  }
}

inline unsigned blocks_for(long long items) { return unsigned((items + NT - 1) / NT); }

template <typename T, bool EXACT>
int run_encode(const EncArgs& a, int B, cudaStream_t s) {
  const Geom& g = a.g;
  void* scratch = nullptr;
  cudaError_t e = cudaMallocAsync(&scratch, size_t(3) * g.Hp * g.Wp * sizeof(T), s);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(any-B scratch)");
  for (int f = 0; f < a.n_frames; ++f) {
    note_kernel("encode_anyb");
    anyb_enc_cols<T, EXACT><<<blocks_for(3LL * g.ny * g.Wp), NT, 0, s>>>(a, B, f, scratch);
    anyb_enc_rows<T, EXACT><<<blocks_for(3LL * g.Hp * g.nx), NT, 0, s>>>(a, B, f, scratch);
    note_extra_launches(1);
  }
  e = cudaGetLastError();
  cudaFreeAsync(scratch, s);
  if (e != cudaSuccess) return cuda_fail(e, "any-B encode launch");
  return VCFB_OK;
}

template <typename T, bool EXACT>
int run_decode(const DecArgs& a, int B, cudaStream_t s) {
  const Geom& g = a.g;
  void* scratch = nullptr;
  cudaError_t e = cudaMallocAsync(&scratch, size_t(3) * g.Hp * g.Wp * sizeof(T), s);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(any-B scratch)");
  for (int f = 0; f < a.n_frames; ++f) {
    note_kernel("decode_anyb");
    anyb_dec_cols<T, EXACT><<<blocks_for(3LL * g.ny * g.Wp), NT, 0, s>>>(a, B, f, scratch);
    anyb_dec_rows<T, EXACT><<<blocks_for(3LL * g.Hp * g.nx), NT, 0, s>>>(a, B, scratch);
    anyb_dec_pixels<T, EXACT><<<blocks_for((long long)g.H * g.W), NT, 0, s>>>(a, f, scratch);
    note_extra_launches(2);
  }
  e = cudaGetLastError();
  cudaFreeAsync(scratch, s);
  if (e != cudaSuccess) return cuda_fail(e, "any-B decode launch");
  return VCFB_OK;
}

}  // namespace

// Fused rate/distortion sweep for the interpreted sizes: the forward transform once, then per step
// quantise / dequantise / inverse transform / SSE.  stats: nq x VCFB_STAT_LEN.
int launch_rd_sweep_anyb(const uint8_t* rgb, const Geom& g, int n_frames, int B, const double* qs, int nq, int color,
                         unsigned flags, unsigned long long* stats, cudaStream_t s) {
  const size_t npl = size_t(g.Hp) * g.Wp;
  void *fplane = nullptr, *dplane = nullptr;
  cudaError_t e = cudaMallocAsync(&fplane, 3 * npl * sizeof(float), s);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync(any-B scratch)");
  e = cudaMallocAsync(&dplane, 3 * npl * sizeof(double), s);
  if (e != cudaSuccess) { cudaFreeAsync(fplane, s); return cuda_fail(e, "cudaMallocAsync(any-B scratch)"); }
  EncArgs ea;
  memset(&ea, 0, sizeof(ea));
  ea.rgb = rgb; ea.g = g; ea.n_frames = n_frames; ea.color = color; ea.flags = flags & ~VCFB_F_NOWRAP;
  DecArgs da;
  memset(&da, 0, sizeof(da));
  da.g = g; da.n_frames = n_frames; da.color = color; da.flags = VCFB_F_FP64 | (flags & VCFB_F_NO_OFFSET); da.original = rgb;
  for (int f = 0; f < n_frames; ++f) {
    note_kernel("rd_sweep_anyb");
    anyb_enc_cols<float, true><<<blocks_for(3LL * g.ny * g.Wp), NT, 0, s>>>(ea, B, f, fplane);
    anyb_enc_rows<float, true, true><<<blocks_for(3LL * g.Hp * g.nx), NT, 0, s>>>(ea, B, f, fplane);
    note_extra_launches(1);
    for (int qi = 0; qi < nq; ++qi) {
      const double q = qs[qi];
      int e2;
      const int pow2 = frexp(q, &e2) == 0.5;
      const int q_int = (q == floor(q) && q < 32768.0) ? int(q) : 0;
      da.stats = stats + size_t(qi) * VCFB_STAT_LEN;
      anyb_rd_quant<<<blocks_for(3LL * npl), NT, 0, s>>>(static_cast<const float*>(fplane), static_cast<double*>(dplane),
                                                        (long long)npl, q, pow2, q_int, flags, da.stats);
      anyb_dec_cols<double, true, true><<<blocks_for(3LL * g.ny * g.Wp), NT, 0, s>>>(da, B, f, dplane);
      anyb_dec_rows<double, true><<<blocks_for(3LL * g.Hp * g.nx), NT, 0, s>>>(da, B, dplane);
      anyb_dec_pixels<double, true><<<blocks_for((long long)g.H * g.W), NT, 0, s>>>(da, f, dplane);
      note_extra_launches(4);
    }
  }
  e = cudaGetLastError();
  cudaFreeAsync(fplane, s);
  cudaFreeAsync(dplane, s);
  if (e != cudaSuccess) return cuda_fail(e, "any-B rd sweep launch");
  return VCFB_OK;
}

bool anyb_supported(int B) { return B == 2 || B == 64 || B == 128; }

int launch_encode_anyb(const EncArgs& a, int B, cudaStream_t s) {
  if (a.flags & VCFB_F_FP64) return run_encode<double, true>(a, B, s);
  if (a.flags & VCFB_F_CONTRACT) return run_encode<float, false>(a, B, s);
  return run_encode<float, true>(a, B, s);
}

int launch_decode_anyb(const DecArgs& a, int B, cudaStream_t s) {
  if (a.flags & VCFB_F_FP64) return run_decode<double, true>(a, B, s);
  if (a.flags & VCFB_F_CONTRACT) return run_decode<float, false>(a, B, s);
  return run_decode<float, true>(a, B, s);
}

}  // namespace vcfb
