// Shared declarations of libvcfb200's translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "vcfb200.h"

namespace vcfb {

// Geometry of one frame: src/2D-DCT.py:208-219 (padding) and :306-307 (subband size).
struct Geom {
  int H, W;       // un-padded frame
  int Hp, Wp;     // padded to a multiple of B
  int top, left;  // offsets of the centred frame inside the padded one
  int ny, nx;     // blocks per column / row = subband size
};

struct EncArgs {
  const uint8_t* rgb;
  uint8_t* idx;
  Geom g;
  int n_frames;
  double q, inv_q;
  int q_pow2;
  int color;
  unsigned flags;
  const double* weights;
  unsigned long long* stats;
};

struct DecArgs {
  const uint8_t* idx;
  uint8_t* rgb;
  void* y_out;
  const uint8_t* original;
  Geom g;
  int n_frames;
  double q;
  int q_int;       // q as an integer when it is integral (numpy keeps int16 * int), else 0
  int color;
  unsigned flags;
  const double* weights;
  unsigned long long* stats;
};

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
void note_kernel(const char* name);   // remembered per thread for vcfb_last_kernel(); counts one launch
void note_extra_launches(int n);      // launches that do not go through note_kernel()

// general kernels (all B, float32 / float64): kernels_general.cu
int launch_encode_general(const EncArgs& a, int B, cudaStream_t s);
int launch_decode_general(const DecArgs& a, int B, cudaStream_t s);

// block sizes 2, 64, 128 (the rest of the reference's -L search set): interpreted pocketfft programs (kernels_anyb.cu)
bool anyb_supported(int B);
int launch_encode_anyb(const EncArgs& a, int B, cudaStream_t s);
int launch_decode_anyb(const DecArgs& a, int B, cudaStream_t s);

// fused rate/distortion sweep (kernels_rd.cu: B in {4, 8, 16, 32}; kernels_anyb.cu: 2, 64, 128)
int launch_rd_sweep(const uint8_t* rgb, const Geom& g, int n_frames, int B, const double* qs, int nq, int color,
                    unsigned flags, unsigned long long* stats, cudaStream_t s);
int launch_rd_sweep_anyb(const uint8_t* rgb, const Geom& g, int n_frames, int B, const double* qs, int nq, int color,
                         unsigned flags, unsigned long long* stats, cudaStream_t s);

// fast path for B = 32 (kernels_tile.cu): VCFB_E_UNSUPP means "not covered, use the general kernel"
int launch_encode_tile(const EncArgs& a, int B, cudaStream_t s);
int launch_decode_tile(const DecArgs& a, int B, cudaStream_t s);

// fast path (kernels_fast.cu): VCFB_E_UNSUPP means "not covered, use the general kernel"
int launch_encode_fast(const EncArgs& a, int B, cudaStream_t s);
int launch_decode_fast(const DecArgs& a, int B, cudaStream_t s);
// tensor-core tier of the B = 8 fast path, float32 fast mode (kernels_tc.cu)
int launch_decode_tc(const DecArgs& a, cudaStream_t s);
int launch_encode_tc(const EncArgs& a, cudaStream_t s);     // fast-mode encoder (VCFB_F_FAST)
// B = 16 fast path (kernels_b16.cu)
int launch_encode_fast16(const EncArgs& a, cudaStream_t s);
int launch_decode_fast16(const DecArgs& a, cudaStream_t s);
int launch_decode_fast16_f32(const DecArgs& a, cudaStream_t s);   // float32 fast mode (kernels_b16f.cu)

// streaming statistics behind the fast path (kernels_stats.cu); VCFB_E_UNSUPP if unaligned
int launch_index_stats(const uint8_t* idx, long long n_bytes, bool hist, unsigned long long* stats, cudaStream_t s);
int launch_sse(const uint8_t* a, const uint8_t* b, long long n_bytes, unsigned long long* stats, cudaStream_t s);
int launch_add_count(unsigned long long* stats, int slot, unsigned long long n, cudaStream_t s);

// stand-alone colour codecs (kernels_color.cu)
int launch_color_encode(const uint8_t* rgb, long long npx, double q, int color, uint16_t* out, cudaStream_t s);
int launch_color_decode(const uint16_t* k, long long npx, double q, int color, uint8_t* rgb, cudaStream_t s);


// motion estimation of the hybrid codec (kernels_motion.cu)
int launch_gray(const uint8_t* rgb, long long npx, uint8_t* gray, cudaStream_t s);
int launch_block_match(const uint8_t* ref, const uint8_t* cur, int n_frames, int H, int W, int bs, int sr, short* mv,
                       cudaStream_t s);
int launch_block_match_tss(const uint8_t* ref, const uint8_t* cur, int n_frames, int H, int W, int bs, int sr, short* mv,
                           cudaStream_t s);

}  // namespace vcfb
