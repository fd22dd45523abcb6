// Stand-alone colour codecs: the reference's src/YCoCg.py:33-85 and src/YCrCb.py:33-69 run
// as codecs of their own (colour transform + deadzone quantiser, no spatial transform).
// Pure integer arithmetic, 3 bytes in and 6 bytes out per pixel (or the reverse): HBM-bound
// elementwise kernels, 16 pixels per thread with 128-bit loads and stores.
//
//   YCoCg encode  img.astype(int16) -> from_RGB stored into an int16 array (fractions
//                 truncated toward zero on store) -> (x / q) truncated -> astype(uint16)
//   YCoCg decode  astype(int16) -> q*k (int16 wrap) -> to_RGB in int16 -> clip -> uint8
//   YCrCb encode  OpenCV 8-bit fixed point RGB2YCrCb -> int16 -> (x / q) truncated -> uint16
//   YCrCb decode  q*k in uint16 -> int16 -> uint8 (wraps) -> OpenCV 8-bit YCrCb2RGB
//                 (the uint8 cast before to_RGB is src/YCrCb.py:59)
#include "common.cuh"

namespace vcfb {
namespace {

__device__ __forceinline__ int rsh14(int v) { return (v + (1 << 13)) >> 14; }
__device__ __forceinline__ int sat8(int v) { return min(max(v, 0), 255); }

// truncating division as numpy performs it: float64 quotient, then astype(int)
__device__ __forceinline__ int quant(int x, int q_int, double q) {
  return q_int ? x / q_int : __double2int_rz(double(x) / q);
}

__device__ __forceinline__ void enc_px(int color, int R, int G, int B, int q_int, double q, unsigned short* o) {
  int c0, c1, c2;
  if (color == VCFB_COLOR_YCOCG) {
    c0 = (R + 2 * G + B) / 4;          // all C divisions truncate toward zero, like the int16 store
    c1 = (R - B) / 2;
    c2 = (2 * G - R - B) / 4;
  } else {
    c0 = rsh14(4899 * R + 9617 * G + 1868 * B);
    c1 = sat8(rsh14((R - c0) * 11682 + (128 << 14)));
    c2 = sat8(rsh14((B - c0) * 9241 + (128 << 14)));
    c0 = sat8(c0);
  }
  o[0] = (unsigned short)quant(c0, q_int, q);
  o[1] = (unsigned short)quant(c1, q_int, q);
  o[2] = (unsigned short)quant(c2, q_int, q);
}

__device__ __forceinline__ void dec_px(int color, const unsigned short* k, int q_int, double q, unsigned char* o) {
  if (color == VCFB_COLOR_YCOCG) {
    short y[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const short kk = short(k[c]);
      y[c] = q_int ? short(kk * q_int) : short(__double2int_rz(double(kk) * q));
    }
    const short r = short(short(y[0] + y[1]) - y[2]);     // int16 arithmetic wraps like numpy's
    const short g = short(y[0] + y[2]);
    const short b = short(short(y[0] - y[1]) - y[2]);
    o[0] = (unsigned char)sat8(r);
    o[1] = (unsigned char)sat8(g);
    o[2] = (unsigned char)sat8(b);
  } else {
    int v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const unsigned prod = q_int ? unsigned(k[c]) * unsigned(q_int) : unsigned(__double2int_rz(double(k[c]) * q));
      v[c] = int(prod & 255u);                              // uint16 -> int16 -> uint8
    }
    const int Y = v[0], Cr = v[1] - 128, Cb = v[2] - 128;
    o[0] = (unsigned char)sat8(Y + rsh14(Cr * 22987));
    o[1] = (unsigned char)sat8(Y + rsh14(Cb * -5636 + Cr * -11698));
    o[2] = (unsigned char)sat8(Y + rsh14(Cb * 29049));
  }
}

constexpr int PPT = 16;   // pixels per thread

__global__ void __launch_bounds__(256) color_encode_kernel(const uint8_t* __restrict__ rgb,
                                                           uint16_t* __restrict__ out, long long npx, int color,
                                                           int q_int, double q, int aligned) {
  const long long ngroups = (npx + PPT - 1) / PPT;
  for (long long gidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; gidx < ngroups;
       gidx += (long long)gridDim.x * blockDim.x) {
    const long long p0 = gidx * PPT;
    if (aligned && p0 + PPT <= npx) {
      uint4 in[3];
      const uint4* src = reinterpret_cast<const uint4*>(rgb + p0 * 3);
#pragma unroll
      for (int i = 0; i < 3; ++i) in[i] = __ldg(src + i);
      const unsigned char* b = reinterpret_cast<const unsigned char*>(in);
      __align__(16) unsigned short o[PPT * 3];
#pragma unroll
      for (int p = 0; p < PPT; ++p) enc_px(color, b[3 * p], b[3 * p + 1], b[3 * p + 2], q_int, q, o + 3 * p);
      uint4* dst = reinterpret_cast<uint4*>(out + p0 * 3);
#pragma unroll
      for (int i = 0; i < 6; ++i) dst[i] = reinterpret_cast<const uint4*>(o)[i];
    } else {
      for (long long p = p0; p < min(p0 + (long long)PPT, npx); ++p)
        enc_px(color, rgb[3 * p], rgb[3 * p + 1], rgb[3 * p + 2], q_int, q, out + 3 * p);
    }
  }
}

__global__ void __launch_bounds__(256) color_decode_kernel(const uint16_t* __restrict__ k,
                                                           uint8_t* __restrict__ rgb, long long npx, int color,
                                                           int q_int, double q, int aligned) {
  const long long ngroups = (npx + PPT - 1) / PPT;
  for (long long gidx = blockIdx.x * (long long)blockDim.x + threadIdx.x; gidx < ngroups;
       gidx += (long long)gridDim.x * blockDim.x) {
    const long long p0 = gidx * PPT;
    if (aligned && p0 + PPT <= npx) {
      uint4 in[6];
      const uint4* src = reinterpret_cast<const uint4*>(k + p0 * 3);
#pragma unroll
      for (int i = 0; i < 6; ++i) in[i] = __ldg(src + i);
      const unsigned short* kk = reinterpret_cast<const unsigned short*>(in);
      __align__(16) unsigned char o[PPT * 3];
#pragma unroll
      for (int p = 0; p < PPT; ++p) dec_px(color, kk + 3 * p, q_int, q, o + 3 * p);
      uint4* dst = reinterpret_cast<uint4*>(rgb + p0 * 3);
#pragma unroll
      for (int i = 0; i < 3; ++i) dst[i] = reinterpret_cast<const uint4*>(o)[i];
    } else {
      for (long long p = p0; p < min(p0 + (long long)PPT, npx); ++p) dec_px(color, k + 3 * p, q_int, q, rgb + 3 * p);
    }
  }
}

int grid_for(long long npx) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long groups = (npx + PPT - 1) / PPT;
  long long blocks = (groups + 255) / 256;
  const long long cap = (long long)sms * 8;
  if (blocks > cap) blocks = cap;
  return int(blocks < 1 ? 1 : blocks);
}

}  // namespace

int launch_color_encode(const uint8_t* rgb, long long npx, double q, int color, uint16_t* out, cudaStream_t s) {
  const int q_int = (q == floor(q) && q < 32768.0) ? int(q) : 0;
  const int aligned = ((reinterpret_cast<uintptr_t>(rgb) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  note_kernel("color_encode");
  color_encode_kernel<<<grid_for(npx), 256, 0, s>>>(rgb, out, npx, color, q_int, q, aligned);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? VCFB_OK : cuda_fail(e, "color_encode_kernel launch");
}

int launch_color_decode(const uint16_t* k, long long npx, double q, int color, uint8_t* rgb, cudaStream_t s) {
  const int q_int = (q == floor(q) && q < 32768.0) ? int(q) : 0;
  const int aligned = ((reinterpret_cast<uintptr_t>(rgb) | reinterpret_cast<uintptr_t>(k)) & 15) == 0;
  note_kernel("color_decode");
  color_decode_kernel<<<grid_for(npx), 256, 0, s>>>(k, rgb, npx, color, q_int, q, aligned);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? VCFB_OK : cuda_fail(e, "color_decode_kernel launch");
}

}  // namespace vcfb
