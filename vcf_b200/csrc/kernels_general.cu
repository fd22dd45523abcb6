// General fused kernels: every block size B in {4,8,16,32}, float32 and float64,
// every flag, any frame shape and any pointer alignment.
//
// One CTA transforms one tile = one block-row (B rows) x TW pixels of one frame:
//
//   encode:  global RGB bytes -> smem (zero padded, src/2D-DCT.py:216-227)
//            pass 1  per pixel column: -128 + colour transform, length-B DCT
//                    down the column (axis 0)                      -> smem F
//            pass 2  per (block, row u): length-B DCT along the row (axis 1),
//                    perceptual scale, deadzone quantise, +128, wrap to uint8
//                                                                 -> smem runs
//            store   runs of contiguous bytes of the subband layout
//                    (sub[j*ny+y, i*nx+x, c]) -> global
//   decode:  the exact inverse, finishing with to_RGB, +128, clip, truncate.
//
// The DCT passes are the generated pocketfft-exact codelets, so with EXACT=true
// the arithmetic is the reference's operation for operation.  The lazy power-
// of-two output scales of the codelets are applied with exact multiplies.
//
// Global memory is touched only in 16-byte vectors where source and staging
// alignment agree (the staging copy of each run is shifted so that they do),
// bytes otherwise; nothing is read or written outside the arrays.
#include "common.cuh"
#include "dct_codelets.cuh"

namespace vcfb {
namespace {

// threads per CTA: B = 32 tiles fill the whole shared memory of an SM (one CTA per SM), so the CTA itself has to
// bring the warps: 384 threads = one round of the 384 (float64) or two of the 768 (float32) work items of a pass
#ifndef NT32
#define NT32 384
#endif
template <int B> __host__ __device__ constexpr int nthreads() { return B == 32 ? NT32 : 256; }

__host__ __device__ constexpr int tile_w(int B, int szT) { return (B >= 16 && szT == 8) ? 128 : 256; }

__host__ __device__ constexpr double p2(int e) {
  double r = 1.0;
  for (int i = 0; i < (e < 0 ? -e : e); ++i) r = e < 0 ? r * 0.5 : r * 2.0;
  return r;
}

__host__ __device__ constexpr int round16(int x) { return (x + 15) / 16 * 16; }
// smallest odd multiple of 16 bytes >= x: rows of such a pitch start in different banks
__host__ __device__ constexpr int round16_odd(int x) { return ((round16(x) / 16) | 1) * 16; }

template <typename T, int B> struct Layout {
  static constexpr int TW = tile_w(B, sizeof(T));
  static constexpr int TBX = TW / B;
  static constexpr int RAWP = TW * 3;                    // encode input row pitch (bytes)
  static constexpr int OUTP = TW * 3 + 16;               // decode output row pitch (bytes)
  static constexpr int FP = TW + 16 / int(sizeof(T));    // pitch of the intermediate (elements)
  static constexpr int RP_SUB = round16_odd(TBX * 3 + 15);   // run pitch, subband layout
  static constexpr int RP_NOSUB = round16(TW * 3 + 15);  // run pitch, -x layout
  static constexpr int STAGE = (B * B * RP_SUB > B * RP_NOSUB) ? B * B * RP_SUB : B * RP_NOSUB;
  static constexpr int F_BYTES = 3 * B * FP * int(sizeof(T));
  // 16-byte chunks per block row of the intermediate, and the XOR key that spreads the
  // chunks of 8 consecutive blocks over the 8 bank groups (pass 2 of the encoder reads one
  // block row per lane, lanes = consecutive blocks)
  static constexpr int VEC = 16 / int(sizeof(T));
  static constexpr int CPB = B / VEC;
  __host__ __device__ static constexpr int key(int bx) {
    return CPB <= 1 ? 0 : (CPB <= 8 ? ((bx * CPB / 8) & (CPB - 1)) : (bx & 7));
  }
  // physical column of logical column x (x = bx * B + i)
  __host__ __device__ static constexpr int swz(int x) {
    return (x / B) * B + ((((x % B) / VEC) ^ key(x / B)) * VEC) + (x % VEC);
  }
  static constexpr int SHIFT_BYTES = B * B * 4;
  static constexpr int ENC_SMEM = B * RAWP + F_BYTES + STAGE + SHIFT_BYTES + 4 * 256 * 3 + 64;
  static constexpr int DEC_SMEM = B * OUTP + F_BYTES + STAGE + SHIFT_BYTES + B * 4 + 64;
};

template <typename T> __device__ __forceinline__ int to_int_rz(T x);
template <> __device__ __forceinline__ int to_int_rz<float>(float x) { return __float2int_rz(x); }
template <> __device__ __forceinline__ int to_int_rz<double>(double x) { return __double2int_rz(x); }

// ---- colour ---------------------------------------------------------------

// Forward colour transform of one pixel, channel c.  YCoCg returns the exact
// integer 4*Y, 2*Co, 4*Cg of the centred pixel (lazy exponent -2, -1, -2):
// color_transforms.YCoCg.from_RGB on (RGB - 128) is exact in any evaluation
// order for 8-bit input, src/2D-DCT.py:292-298.  YCrCb is the float extension
// of oracle/vcf_oracle.py::ycrcb_from_rgb_float, operation for operation.
template <typename T, bool EXACT>
__device__ __forceinline__ T color_fwd(int color, int c, int R, int G, int Bc, int poff = 128) {   // poff: src/2D-DCT.py:292
  using O = Ops<T, EXACT>;
  if (color == VCFB_COLOR_YCOCG) {
    int v = (c == 0) ? (R + 2 * G + Bc - 4 * poff) : (c == 1) ? (R - Bc) : (2 * G - R - Bc);
    return T(v);
  }
  const T r = T(R - poff), g = T(G - poff), b = T(Bc - poff);
  const T y = O::add(O::add(O::mul(r, T(0.299)), O::mul(g, T(0.587))), O::mul(b, T(0.114)));
  if (c == 0) return y;
  if (c == 1) return O::mul(O::sub(r, y), T(0.713));
  return O::mul(O::sub(b, y), T(0.564));
}

// ---- small helpers ----------------------------------------------------------

__device__ __forceinline__ unsigned warp_sum(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ============================================================================
// encode
// ============================================================================

// MODE 0: subband layout, no perceptual weights, no statistics -- every flag a compile-time
// constant; MODE 1: the same with statistics; MODE 2: everything decided at run time.
template <typename T, int B, bool EXACT, int MODE>
__global__ void __launch_bounds__(nthreads<B>()) encode_kernel(const EncArgs a) {
  constexpr int NT = nthreads<B>();
  using L = Layout<T, B>;
  using O = Ops<T, EXACT>;
  using D = Dct<B, false>;
  using M = typename D::meta;
  constexpr int TW = L::TW, TBX = L::TBX, RAWP = L::RAWP, FP = L::FP;

  extern __shared__ __align__(16) unsigned char smem[];
  uint8_t* raw = smem;
  T* F = reinterpret_cast<T*>(smem + B * RAWP);
  uint8_t* stage = smem + B * RAWP + L::F_BYTES;
  int* rshift = reinterpret_cast<int*>(stage + L::STAGE);
  unsigned* shist = reinterpret_cast<unsigned*>(stage + L::STAGE + L::SHIFT_BYTES);
  unsigned* sacc = shist + 3 * 256;   // [0]=nonzero [1]=sumabs

  const int tid = threadIdx.x;
  const int tile = blockIdx.x, by = blockIdx.y, f = blockIdx.z;
  const Geom g = a.g;
  const int x0 = tile * TW;
  const int bx0 = tile * TBX;
  const int nbx = min(TBX, g.nx - bx0);
  const bool nosub = MODE == 2 && (a.flags & VCFB_F_NO_SUBBANDS) != 0;
  const bool percep = MODE == 2 && (a.flags & VCFB_F_PERCEPTUAL) != 0;
  const bool do_stats = MODE != 0 && a.stats != nullptr;
  const bool do_hist = do_stats && (a.flags & VCFB_F_HIST) != 0;
  const int poff = (MODE == 2 && (a.flags & VCFB_F_NO_OFFSET)) ? 0 : 128;   // the loop of optimize_block_size runs without offset
  const int nruns = nosub ? B : B * B;
  const int runlen = nosub ? nbx * B * 3 : nbx * 3;
  const int rpitch = nosub ? L::RP_NOSUB : L::RP_SUB;

  auto run_gptr = [&](int run) -> uint8_t* {
    size_t row, col;
    if (nosub) {
      row = size_t(by) * B + run;
      col = size_t(bx0) * B;
    } else {
      const int j = run / B, i = run % B;
      row = size_t(j) * g.ny + by;
      col = size_t(i) * g.nx + bx0;
    }
    return a.idx + ((size_t(f) * g.Hp + row) * g.Wp + col) * 3;
  };

  for (int r = tid; r < nruns; r += NT) rshift[r] = int(reinterpret_cast<uintptr_t>(run_gptr(r)) & 15);
  if (do_stats) {
    for (int i = tid; i < 3 * 256 + 2; i += NT) shist[i] = 0;
  }

  // ---- load: B rows x TW pixels, zero outside the frame (padding) -----------
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

  // ---- pass 1: colour + DCT down each pixel column (axis 0) -----------------
  constexpr bool JOINT = 3 * B * int(sizeof(T)) <= 4 * 96;     // all three channels in registers
  if (JOINT) {
    for (int x = tid; x < TW; x += NT) {
      T v[3][B];
#pragma unroll
      for (int r = 0; r < B; ++r) {
        const uint8_t* px = raw + r * RAWP + x * 3;
        const int R = px[0], G = px[1], Bc = px[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c][r] = color_fwd<T, EXACT>(a.color, c, R, G, Bc, poff);
      }
      const int xs = L::swz(x);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const T cs = T(a.color == VCFB_COLOR_YCOCG ? (c == 1 ? 0.5 : 0.25) : 1.0);
        D::template run<T, EXACT>(v[c]);
#pragma unroll
        for (int u = 0; u < B; ++u)
          F[(c * B + u) * FP + xs] = O::mul(v[c][u], T(M::sgn(u) * p2(M::exp(u))) * cs);   // exact factor
      }
    }
  } else {
    // work items = (channel, pixel column): all NT threads stay busy when TW < NT
    for (int wi = tid; wi < 3 * TW; wi += NT) {
      const int c = wi / TW, x = wi - c * TW;
      const T cs = T(a.color == VCFB_COLOR_YCOCG ? (c == 1 ? 0.5 : 0.25) : 1.0);
      T v[B];
#pragma unroll
      for (int r = 0; r < B; ++r) {
        const uint8_t* px = raw + r * RAWP + x * 3;
        v[r] = color_fwd<T, EXACT>(a.color, c, px[0], px[1], px[2], poff);
      }
      D::template run<T, EXACT>(v);
      const int xs = L::swz(x);
#pragma unroll
      for (int u = 0; u < B; ++u)
        F[(c * B + u) * FP + xs] = O::mul(v[u], T(M::sgn(u) * p2(M::exp(u))) * cs);       // exact factor
    }
  }
  __syncthreads();

  // ---- pass 2: DCT along each block row (axis 1), quantise ------------------
  unsigned nz = 0, sabs = 0;
  {
    const T q = T(a.q);
    const T inv_q = T(a.inv_q);
    const bool fastq = a.q_pow2 && !percep;
    {
      for (int wi = tid; wi < 3 * TW; wi += NT) {     // work items = (channel, block, row u)
        const int c = wi / TW, t = wi - c * TW;
        const int bx = t % TBX, u = t / TBX;      // lanes = consecutive blocks: conflict-free staging stores
        if (bx >= nbx) continue;
        T v[B];
        const T* src = F + (c * B + u) * FP + bx * B;
        constexpr int VEC = L::VEC;
        const int key = L::key(bx);
#pragma unroll
        for (int i = 0; i < B; i += VEC) {
          const uint4 w = *reinterpret_cast<const uint4*>(src + (((i / VEC) ^ key) * VEC));
          const T* wv = reinterpret_cast<const T*>(&w);
#pragma unroll
          for (int k = 0; k < VEC; ++k) v[i + k] = wv[k];
        }
        D::template run<T, EXACT>(v);
        const double* wt = percep ? a.weights + (c ? B * B : 0) + u * B : nullptr;
#pragma unroll
        for (int i = 0; i < B; ++i) {
          const T sc = T(M::sgn(i) * p2(M::exp(i)));   // exact factor
          T tq;
          if (fastq) {
            tq = O::mul(v[i], sc * inv_q);               // x / 2^k == x * 2^-k exactly
          } else {
            T coef = O::mul(v[i], sc);                   // the coefficient scipy returns
            if (percep) coef = T(double(coef) * wt[i]);  // src/2D-DCT.py:322-324
            tq = a.q_pow2 ? O::mul(coef, inv_q) : O::div(coef, q);  // src/deadzone.py:98
          }
          const int k = to_int_rz<T>(tq);                // truncation = dead zone
          const unsigned byte = unsigned(k + poff) & 255u;  // src/2D-DCT.py:348,:361 (wraps)
          int run, off;
          if (nosub) {
            run = u;
            off = (bx * B + i) * 3 + c;
          } else {
            run = u * B + i;
            off = bx * 3 + c;
          }
          stage[run * rpitch + rshift[run] + off] = uint8_t(byte);
          if (do_stats) {
            const int kk = poff ? int(byte) - 128 : int((signed char)byte);
            nz += (kk != 0);
            sabs += unsigned(kk < 0 ? -kk : kk);
            if (do_hist) atomicAdd(&shist[c * 256 + byte], 1u);
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- store runs -----------------------------------------------------------
  {
    const int chunks = rpitch / 16;
    for (int t = tid; t < nruns * chunks; t += NT) {
      const int run = t / chunks, ch = t % chunks;
      const int sh = rshift[run];
      const int lo = ch * 16, hi = lo + 16;
      const int vlo = max(lo, sh), vhi = min(hi, sh + runlen);
      if (vlo >= vhi) continue;
      uint8_t* gp = run_gptr(run) - sh;  // 16-byte aligned
      const uint8_t* sp = stage + run * rpitch;
      if (vlo == lo && vhi == hi) {
        *reinterpret_cast<uint4*>(gp + lo) = *reinterpret_cast<const uint4*>(sp + lo);
      } else {
        for (int b = vlo; b < vhi; ++b) gp[b] = sp[b];
      }
    }
  }

  if (do_stats) {
    nz = warp_sum(nz);
    sabs = warp_sum(sabs);
    if ((tid & 31) == 0) {
      atomicAdd(&sacc[0], nz);
      atomicAdd(&sacc[1], sabs);
    }
    __syncthreads();
    if (do_hist)
      for (int i = tid; i < 3 * 256; i += NT)
        if (shist[i]) atomicAdd(a.stats + VCFB_STAT_HIST + i, (unsigned long long)shist[i]);
    if (tid == 0) {
      atomicAdd(a.stats + VCFB_STAT_NONZERO, (unsigned long long)sacc[0]);
      atomicAdd(a.stats + VCFB_STAT_SUMABS, (unsigned long long)sacc[1]);
      atomicAdd(a.stats + VCFB_STAT_NINDICES, (unsigned long long)(nbx * B * B * 3));
    }
  }
}

// ============================================================================
// decode
// ============================================================================

// PLAIN: subband layout and no perceptual weights, known at compile time.
template <typename T, int B, bool EXACT, bool PLAIN>
__global__ void __launch_bounds__(nthreads<B>()) decode_kernel(const DecArgs a) {
  constexpr int NT = nthreads<B>();
  using L = Layout<T, B>;
  using O = Ops<T, EXACT>;
  using D = Dct<B, true>;
  using M = typename D::meta;
  constexpr int TW = L::TW, TBX = L::TBX, OUTP = L::OUTP, FP = L::FP;

  extern __shared__ __align__(16) unsigned char smem[];
  uint8_t* outb = smem;
  T* F = reinterpret_cast<T*>(smem + B * OUTP);
  uint8_t* stage = smem + B * OUTP + L::F_BYTES;
  int* rshift = reinterpret_cast<int*>(stage + L::STAGE);
  int* oshift = rshift + B * B;
  unsigned long long* ssse = reinterpret_cast<unsigned long long*>(
      smem + ((B * OUTP + L::F_BYTES + L::STAGE + L::SHIFT_BYTES + B * 4 + 7) / 8) * 8);

  const int tid = threadIdx.x;
  const int tile = blockIdx.x, by = blockIdx.y, f = blockIdx.z;
  const Geom g = a.g;
  const int x0 = tile * TW;
  const int bx0 = tile * TBX;
  const int nbx = min(TBX, g.nx - bx0);
  const bool nosub = !PLAIN && (a.flags & VCFB_F_NO_SUBBANDS) != 0;
  const bool percep = !PLAIN && (a.flags & VCFB_F_PERCEPTUAL) != 0;
  const int nruns = nosub ? B : B * B;
  const int runlen = nosub ? nbx * B * 3 : nbx * 3;
  const int rpitch = nosub ? L::RP_NOSUB : L::RP_SUB;

  // un-padded columns covered by this tile: [gx0, gx1)
  const int gx0 = max(0, x0 - g.left);
  const int gx1 = min(g.W, x0 - g.left + nbx * B);
  const int first = (gx0 - (x0 - g.left)) * 3;   // tile byte offset of the first kept byte
  const int outlen = max(0, gx1 - gx0) * 3;
  const uint8_t* align_base = a.rgb ? a.rgb : a.original;

  auto run_gptr = [&](int run) -> const uint8_t* {
    size_t row, col;
    if (nosub) {
      row = size_t(by) * B + run;
      col = size_t(bx0) * B;
    } else {
      const int j = run / B, i = run % B;
      row = size_t(j) * g.ny + by;
      col = size_t(i) * g.nx + bx0;
    }
    return a.idx + ((size_t(f) * g.Hp + row) * g.Wp + col) * 3;
  };
  auto out_goff = [&](int r) -> size_t {   // byte offset of the first kept byte of row r
    const int gy = by * B + r - g.top;
    return ((size_t(f) * g.H + gy) * g.W + gx0) * 3;
  };

  for (int r = tid; r < nruns; r += NT) rshift[r] = int(reinterpret_cast<uintptr_t>(run_gptr(r)) & 15);
  if (tid < B) {
    const int gy = by * B + tid - g.top;
    int sh = 0;
    if (gy >= 0 && gy < g.H && outlen > 0 && align_base)
      sh = int((reinterpret_cast<uintptr_t>(align_base + out_goff(tid)) - size_t(first)) & 15);
    oshift[tid] = sh;
  }
  if (tid < 4) ssse[tid] = 0;
  __syncthreads();

  // ---- gather runs ------------------------------------------------------------
  {
    const int chunks = rpitch / 16;
    for (int t = tid; t < nruns * chunks; t += NT) {
      const int run = t / chunks, ch = t % chunks;
      const int sh = rshift[run];
      const int lo = ch * 16, hi = lo + 16;
      const int vlo = max(lo, sh), vhi = min(hi, sh + runlen);
      if (vlo >= vhi) continue;
      const uint8_t* gp = run_gptr(run) - sh;
      uint8_t* sp = stage + run * rpitch;
      if (vlo == lo && vhi == hi) {
        *reinterpret_cast<uint4*>(sp + lo) = __ldg(reinterpret_cast<const uint4*>(gp + lo));
      } else {
        for (int b = vlo; b < vhi; ++b) sp[b] = gp[b];
      }
    }
  }
  __syncthreads();

  // ---- pass 1: dequantise, inverse DCT along axis 0 ---------------------------
  {
    for (int wi = tid; wi < 3 * TW; wi += NT) {       // work items = (channel, coefficient column)
      const int c = wi / TW, x = wi - c * TW;
      const int bx = x / B, i = x % B;
      if (bx >= nbx) continue;
      T v[B];
#pragma unroll
      for (int u = 0; u < B; ++u) {
        int run, off;
        if (nosub) {
          run = u;
          off = x * 3 + c;
        } else {
          run = u * B + i;
          off = bx * 3 + c;
        }
        const int k = int(stage[run * rpitch + rshift[run] + off]) - 128;  // :398,:402
        T y;
        if (a.q_int) {
          y = T(int(short(k * a.q_int)));        // int16 * python int stays int16 (wraps)
        } else {
          y = T(double(k) * a.q);
        }
        if (percep) {                            // :421-435, stored back into the int16 array
          const float fv = float(double(float(y)) / a.weights[(c ? B * B : 0) + u * B + i]);
          y = T(int(short(__float2int_rz(fv))));
        }
        v[u] = y;
      }
      D::template run<T, EXACT>(v);
#pragma unroll
      for (int r = 0; r < B; ++r)
        F[(c * B + r) * FP + x] = O::mul(v[r], T(M::sgn(r) * p2(M::exp(r))));
    }
  }
  __syncthreads();

  // ---- pass 2: inverse DCT along axis 1 (in place in smem) --------------------
  {
    for (int wi = tid; wi < 3 * TW; wi += NT) {       // work items = (channel, block, pixel row)
      const int c = wi / TW, t = wi - c * TW;
      const int r = t % B, bx = t / B;
      if (bx >= nbx) continue;
      T v[B];
      T* src = F + (c * B + r) * FP + bx * B;
      constexpr int VEC = 16 / int(sizeof(T));
#pragma unroll
      for (int i = 0; i < B; i += VEC) {
        const uint4 w = *reinterpret_cast<const uint4*>(src + i);
        const T* wv = reinterpret_cast<const T*>(&w);
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[i + k] = wv[k];
      }
      D::template run<T, EXACT>(v);
#pragma unroll
      for (int i = 0; i < B; i += VEC) {
        uint4 w;
        T* wv = reinterpret_cast<T*>(&w);
#pragma unroll
        for (int k = 0; k < VEC; ++k) wv[k] = O::mul(v[i + k], T(M::sgn(i + k) * p2(M::exp(i + k))));
        *reinterpret_cast<uint4*>(src + i) = w;
      }
    }
  }
  __syncthreads();

  // ---- pass 3: to_RGB, +128, clip, truncate (:449-466) ------------------------
  for (int t = tid; t < B * TW; t += NT) {
    const int r = t / TW, x = t % TW;
    const int gy = by * B + r - g.top;
    const int gx = x0 + x - g.left;
    if (x >= nbx * B || gy < 0 || gy >= g.H || gx < 0 || gx >= g.W) continue;
    const T c0 = F[(0 * B + r) * FP + x];
    const T c1 = F[(1 * B + r) * FP + x];
    const T c2 = F[(2 * B + r) * FP + x];
    T R, G, Bv;
    if (sizeof(T) == 8 && (a.flags & VCFB_F_SYNTH_F32)) {
      // upstream variant: synthesize_image stores its float64 result in a float32 array, so the
      // colour transform, the +128 and the truncation run on float32 (tests/test_oracle_variants.py)
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
    if (a.color == VCFB_COLOR_YCOCG) {   // Y + Co - Cg ; Y + Cg ; Y - Co - Cg, left to right
      R = O::sub(O::add(c0, c1), c2);
      G = O::add(c0, c2);
      Bv = O::sub(O::sub(c0, c1), c2);
    } else {                             // oracle ycrcb_to_rgb_float
      R = O::add(c0, O::mul(c1, T(1.403)));
      G = O::add(O::add(c0, O::mul(c1, T(-0.714))), O::mul(c2, T(-0.344)));
      Bv = O::add(c0, O::mul(c2, T(1.773)));
    }
    R = O::add(R, T(128));
    G = O::add(G, T(128));
    Bv = O::add(Bv, T(128));
    }
    if (a.y_out) {
      T* yo = reinterpret_cast<T*>(a.y_out) + ((size_t(f) * g.H + gy) * g.W + gx) * 3;
      yo[0] = R; yo[1] = G; yo[2] = Bv;
    }
    uint8_t* o = outb + r * OUTP + oshift[r] + x * 3;
    // np.clip(y, 0, 255).astype(uint8): truncation (saturating conversion) then the clip on integers
    o[0] = uint8_t(min(max(to_int_rz<T>(R), 0), 255));
    o[1] = uint8_t(min(max(to_int_rz<T>(G), 0), 255));
    o[2] = uint8_t(min(max(to_int_rz<T>(Bv), 0), 255));
  }
  __syncthreads();

  // ---- store rows, SSE against the original (src/RDE.py:41-49) -----------------
  {
    unsigned sse[3] = {0, 0, 0};
    int sdiff = 0;
    const bool do_sse = a.stats != nullptr && a.original != nullptr;
    constexpr int CH = OUTP / 16;
    if (outlen > 0) {
      for (int t = tid; t < B * CH; t += NT) {
        const int r = t / CH, ch = t % CH;
        const int gy = by * B + r - g.top;
        if (gy < 0 || gy >= g.H) continue;
        const int sh = oshift[r] + first;     // smem offset (in the row) of the first kept byte
        const int lo = ch * 16, hi = lo + 16;
        const int vlo = max(lo, sh), vhi = min(hi, sh + outlen);
        if (vlo >= vhi) continue;
        const uint8_t* sp = outb + r * OUTP;
        const size_t goff = out_goff(r);
        if (a.rgb) {
          uint8_t* gp = a.rgb + goff - sh;
          if (vlo == lo && vhi == hi) {
            *reinterpret_cast<uint4*>(gp + lo) = *reinterpret_cast<const uint4*>(sp + lo);
          } else {
            for (int b = vlo; b < vhi; ++b) gp[b] = sp[b];
          }
        }
        if (do_sse) {
          const uint8_t* op = a.original + goff - sh;
          if (vlo == lo && vhi == hi) {
            // whole 16-byte chunk: byte-wise |difference| in one instruction per word, squares summed
            // per byte phase with a masked dp4a; phase p of the chunk is channel (p + lo - sh) mod 3
            constexpr unsigned M[3][3] = {{0xFF0000FFu, 0x0000FF00u, 0x00FF0000u},
                                          {0x00FF0000u, 0xFF0000FFu, 0x0000FF00u},
                                          {0x0000FF00u, 0x00FF0000u, 0xFF0000FFu}};
            const uint4 xo = __ldg(reinterpret_cast<const uint4*>(op + lo));
            const uint4 yo = *reinterpret_cast<const uint4*>(sp + lo);
            const unsigned xa[4] = {xo.x, xo.y, xo.z, xo.w}, ya[4] = {yo.x, yo.y, yo.z, yo.w};
            unsigned ph[3] = {0, 0, 0};
            int sx = 0, sy = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const unsigned d = __vabsdiffu4(xa[k], ya[k]);
#pragma unroll
              for (int p = 0; p < 3; ++p) {
                const unsigned dm = d & M[(4 * k) % 3][p];
                ph[p] = __dp4a(dm, dm, ph[p]);
              }
              sx = __dp4a(xa[k], 0x01010101u, unsigned(sx));
              sy = __dp4a(ya[k], 0x01010101u, unsigned(sy));
            }
            const int rot = ((lo - sh) % 3 + 3) % 3;
            sse[rot] += ph[0];
            sse[rot == 2 ? 0 : rot + 1] += ph[1];
            sse[rot == 0 ? 2 : rot - 1] += ph[2];
            sdiff += sx - sy;
          } else {
            for (int b = vlo; b < vhi; ++b) {
              const int d = int(op[b]) - int(sp[b]);
              sse[(b - sh) % 3] += unsigned(d * d);
              sdiff += d;
            }
          }
        }
      }
    }
    if (do_sse) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const unsigned s = warp_sum(sse[c]);
        if ((tid & 31) == 0 && s) atomicAdd(&ssse[c], (unsigned long long)s);
      }
      {
        const unsigned sd = warp_sum(unsigned(sdiff));           // two's complement sum
        if ((tid & 31) == 0 && sd) atomicAdd(&ssse[3], (unsigned long long)(long long)(int)sd);
      }
      __syncthreads();
      if (tid < 3 && ssse[tid]) atomicAdd(a.stats + VCFB_STAT_SSE_R + tid, ssse[tid]);
      if (tid == 4 && ssse[3]) atomicAdd(a.stats + VCFB_STAT_SUMDIFF, ssse[3]);
      if (tid == 3) {
        int rows = 0;
        for (int r = 0; r < B; ++r) {
          const int gy = by * B + r - g.top;
          rows += (gy >= 0 && gy < g.H);
        }
        atomicAdd(a.stats + VCFB_STAT_NSAMPLES, (unsigned long long)rows * (outlen > 0 ? outlen : 0));
      }
    }
  }
}

// ---- launchers ---------------------------------------------------------------

template <typename T, int B, bool EXACT, int MODE>
int launch_enc_mode(const EncArgs& a, cudaStream_t s) {
  using L = Layout<T, B>;
  auto kern = encode_kernel<T, B, EXACT, MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::ENC_SMEM);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(encode)");
  dim3 grid((a.g.Wp + L::TW - 1) / L::TW, a.g.ny, a.n_frames);
  note_kernel("encode_general");
  kern<<<grid, nthreads<B>(), L::ENC_SMEM, s>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "encode_kernel launch");
  return VCFB_OK;
}

template <typename T, int B, bool EXACT>
int launch_enc(const EncArgs& a, cudaStream_t s) {
  const bool plain = !(a.flags & (VCFB_F_NO_SUBBANDS | VCFB_F_PERCEPTUAL | VCFB_F_NO_OFFSET));
  if (EXACT && plain) return a.stats ? launch_enc_mode<T, B, EXACT, 1>(a, s) : launch_enc_mode<T, B, EXACT, 0>(a, s);
  return launch_enc_mode<T, B, EXACT, 2>(a, s);
}

template <typename T, int B, bool EXACT, bool PLAIN>
int launch_dec_mode(const DecArgs& a, cudaStream_t s) {
  using L = Layout<T, B>;
  auto kern = decode_kernel<T, B, EXACT, PLAIN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DEC_SMEM);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(decode)");
  dim3 grid((a.g.Wp + L::TW - 1) / L::TW, a.g.ny, a.n_frames);
  note_kernel("decode_general");
  kern<<<grid, nthreads<B>(), L::DEC_SMEM, s>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "decode_kernel launch");
  return VCFB_OK;
}

template <typename T, int B, bool EXACT>
int launch_dec(const DecArgs& a, cudaStream_t s) {
  const bool plain = !(a.flags & (VCFB_F_NO_SUBBANDS | VCFB_F_PERCEPTUAL));
  if (EXACT && plain) return launch_dec_mode<T, B, EXACT, true>(a, s);
  return launch_dec_mode<T, B, EXACT, false>(a, s);
}

template <int B>
int enc_by_mode(const EncArgs& a, cudaStream_t s) {
  if (a.flags & VCFB_F_FP64) return launch_enc<double, B, true>(a, s);
  if (a.flags & VCFB_F_CONTRACT) return launch_enc<float, B, false>(a, s);
  return launch_enc<float, B, true>(a, s);
}

template <int B>
int dec_by_mode(const DecArgs& a, cudaStream_t s) {
  if (a.flags & VCFB_F_FP64) return launch_dec<double, B, true>(a, s);
  if (a.flags & VCFB_F_CONTRACT) return launch_dec<float, B, false>(a, s);
  return launch_dec<float, B, true>(a, s);
}

}  // namespace

int launch_encode_general(const EncArgs& a, int B, cudaStream_t s) {
  switch (B) {
    case 4: return enc_by_mode<4>(a, s);
    case 8: return enc_by_mode<8>(a, s);
    case 16: return enc_by_mode<16>(a, s);
    case 32: return enc_by_mode<32>(a, s);
  }
  set_error("unsupported block size (supported: 4, 8, 16, 32)");
  return VCFB_E_ARG;
}

int launch_decode_general(const DecArgs& a, int B, cudaStream_t s) {
  switch (B) {
    case 4: return dec_by_mode<4>(a, s);
    case 8: return dec_by_mode<8>(a, s);
    case 16: return dec_by_mode<16>(a, s);
    case 32: return dec_by_mode<32>(a, s);
  }
  set_error("unsupported block size (supported: 4, 8, 16, 32)");
  return VCFB_E_ARG;
}

}  // namespace vcfb
