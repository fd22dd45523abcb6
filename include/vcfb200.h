/* libvcfb200 -- C ABI of the B200-native colour + block-DCT + deadzone path of
 * the Visual Coding Framework (Sistemas-Multimedia/VCF).
 *
 * This is the drop-in boundary.  The reference has no FFI of its own (it is pure
 * Python); the entry points below are what a ctypes binding inside the
 * reference's spatial-transform stage binds in place of the numpy/scipy
 * arithmetic.  Each one names the reference code it replaces (paths relative
 * to the reference repository).  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C, no C++/torch types; every function returns 0 on success or a
 *     negative VCFB_E_* code and never throws; vcfb_last_error() gives the text
 *     of the last failure on the calling thread;
 *   - images are interleaved 8-bit RGB, shape (n_frames, H, W, 3), C order;
 *     index arrays are uint8, shape (n_frames, Hp, Wp, 3) with Hp, Wp = H, W
 *     rounded up to a multiple of B (vcfb_padded_dims);
 *   - the *_dev entry points take DEVICE pointers and a CUDA stream (0 = legacy
 *     default stream; pass torch's current stream handle from Python), are
 *     asynchronous, re-entrant per stream and keep no global mutable state;
 *   - the host entry points take HOST pointers, stage through pinned memory
 *     owned by an explicit context and return when the result is in the output
 *     buffer;
 *   - the caller owns every buffer.
 *   - There is no CPU fallback: without a CUDA device every compute entry point
 *     fails with VCFB_E_CUDA.
 */
#ifndef VCFB200_H
#define VCFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VCFB_VERSION 150 /* 0.5.0: + vcfb_deflate_rows_dev (row-above / previous-sample match candidates, cost model); 0.4.0: + vcfb_rd_sweep_dev, VCFB_F_NOWRAP, block sizes 2 / 64 / 128; 0.3.0: + VCFB_F_SYNTH_F32; 0.2.0: + vcfb_deflate_dev, vcfb_deflate_bound, vcfb_deflate_workspace, vcfb_crc32_dev, vcfb_adler32_dev (0.1.1: vcfb_launch_count, motion estimation) */

/* error codes */
#define VCFB_OK 0
#define VCFB_E_ARG (-1)     /* bad argument (NULL pointer, unsupported B, q <= 0, ...) */
#define VCFB_E_CUDA (-2)    /* CUDA runtime error (text in vcfb_last_error) */
#define VCFB_E_UNSUPP (-3)  /* valid request this build does not implement */

/* colour transform applied in front of the DCT */
#define VCFB_COLOR_YCOCG 0 /* color_transforms.YCoCg.from_RGB/to_RGB, the only one
                              src/2D-DCT.py reaches (:22-23, :298, :449) */
#define VCFB_COLOR_YCRCB 1 /* float BT.601 extension defined by oracle/vcf_oracle.py
                              (BASELINE.json config 5); not a reference behaviour */

/* flags */
#define VCFB_F_NO_SUBBANDS 1u /* -x / --disable_subbands (src/2D-DCT.py:40, :333, :413) */
#define VCFB_F_PERCEPTUAL 2u  /* -p / --perceptual_quantization (:38, :313-327, :421-435);
                                 needs `weights` */
#define VCFB_F_FP64 4u        /* encode: validation mode, float64 arithmetic (indices
                                 bit-exact with the oracle evaluated in float64);
                                 decode: the reference's own float64 chain
                                 (:398-466), pixels bit-exact */
#define VCFB_F_CONTRACT 8u    /* float32 only: let the compiler fuse multiply-adds in
                                 the DCT.  Without it the float32 encoder is bit-exact
                                 with the reference's float32 path (:276); with it
                                 < 1e-6 of the indices may differ. */

#define VCFB_F_SYNTH_F32 32u  /* decode, with VCFB_F_FP64: the upstream variant in which
                                 DCT2D.block_DCT.synthesize_image stores its float64 result in a
                                 float32 array (the un-vendored package cannot be read):
                                 to_RGB, +128 and the truncation then run on float32.  Differs from
                                 the default chain by <= 1 LSB per pixel, up to 0.03 dB on content
                                 dominated by DC-only blocks (tests/test_oracle_variants.py). */

#define VCFB_F_FAST 64u       /* encode: the north star's fast mode -- the tensor-core encoder
                                 (csrc/kernels_tc.cu: 64 x 64 Kronecker DCT on tcgen05, the four
                                 rational coefficients by pocketfft's own float32 sequence): fewer
                                 than 1e-6 of the indices differ from the reference's float32 path,
                                 only at rounding boundaries.  Applies to B = 8, YCoCg, subbands,
                                 q = 2^k >= 8; any other request is served by the bit-exact encoder.
                                 Decode: ignored (float32 decode IS the fast mode). */

#define VCFB_F_NOWRAP 128u    /* vcfb_rd_sweep_dev only: dequantise the quantiser's own indices, before they are
                                 narrowed to uint8 -- the in-process loop of optimize_block_size
                                 (src/2D-DCT.py:560-566).  Without it: the indices a decoder reads back from the
                                 code-stream (wrapped to uint8 :361, int16 arithmetic :398-410). */

#define VCFB_F_NO_OFFSET 256u  /* vcfb_encode_dev and vcfb_rd_sweep_dev: neither the -128 on the pixels nor the +128 on the
                                 indices.  This is what the reference's optimize_block_size actually runs with: it is
                                 called from __init__ (src/2D-DCT.py:99-103) BEFORE self.offset = 128 is assigned (:107-110),
                                 while self.offset still is the array [0, 0, 0] left by the colour stage (src/YCoCg.py:28-29) --
                                 found by running the unmodified reference (tests/golden/ref_flow_L_*.npz). */

#define VCFB_F_HIST 16u       /* statistics: also accumulate the 3 x 256 histogram of the indices
                                 (one shared-memory atomic per sample; off = only the sums) */

/* statistics vector (int64), accumulated with integer atomics; the caller zeroes it.
 * Integer sums make the multi-GPU all-reduce order-independent. */
#define VCFB_STAT_SSE_R 0      /* sum (original - decoded)^2, channel R  (src/RDE.py:41-49) */
#define VCFB_STAT_SSE_G 1
#define VCFB_STAT_SSE_B 2
#define VCFB_STAT_NSAMPLES 3   /* samples compared (H*W*3 per frame) */
#define VCFB_STAT_NONZERO 4    /* indices != 0 after removing the 128 bias */
#define VCFB_STAT_SUMABS 5     /* sum |index| */
#define VCFB_STAT_NINDICES 6   /* indices written (Hp*Wp*3 per frame) */
#define VCFB_STAT_SUMDIFF 7    /* signed sum (original - decoded): lets a caller form
                                  sum((original - c) - decoded)^2 for any constant c, which is
                                  what src/2D-DCT.py:572-574 compares (image still shifted by 128) */
#define VCFB_STAT_HIST 8       /* 3 x 256 histogram of the uint8 indices, channel-major */
#define VCFB_STAT_LEN (8 + 3 * 256)

int vcfb_version(void);
const char* vcfb_last_error(void);

/* Name of the kernel family the calling thread launched last ("enc8_fast",
 * "encode_general", "decode_general", ...): lets tests and benchmarks assert which
 * code path served a request. */
const char* vcfb_last_kernel(void);

/* Kernels launched so far by the calling thread through this library (monotonic).  The probed
 * float64 decode of the B=8 fast path counts 4: the probe and three decoders, two of which
 * return at once. */
long long vcfb_launch_count(void);

/* Number of CUDA devices visible (0 if none / no driver). */
int vcfb_device_count(void);

/* src/2D-DCT.py:208-219 (pad_and_center_to_multiple_of_block_size): padded dims
 * and the top/left offsets of the centred image. */
int vcfb_padded_dims(int H, int W, int B, int* Hp, int* Wp, int* top, int* left);

/* Encode: replaces src/2D-DCT.py:276-361 between encode_read_fn and compress --
 *   astype(float32) :276, pad :282, -=128 :292, from_RGB :298, space_analyze :303,
 *   perceptual scale :313-327, get_subbands :336, quantize_decom :343
 *   (src/deadzone.py:95-105), += 128 :348, astype(uint8) :361 (wraps, no clip).
 * rgb      (n_frames,H,W,3) uint8, device
 * q        quantisation step (-q / QSS, src/deadzone.py:30), > 0
 * weights  2*B*B doubles, device: Y_QSSs/121 then C_QSSs/99 (src/2D-DCT.py:322-324),
 *          row-major [j][i]; NULL unless VCFB_F_PERCEPTUAL
 * idx_out  (n_frames,Hp,Wp,3) uint8, device
 * stats    VCFB_STAT_LEN int64 on the device or NULL; adds NONZERO, SUMABS, NINDICES
 *          and, with VCFB_F_HIST, HIST */
int vcfb_encode_dev(const uint8_t* rgb, int n_frames, int H, int W, int B, double q,
                    int color, unsigned flags, const double* weights,
                    uint8_t* idx_out, int64_t* stats, void* cuda_stream);

/* Decode: replaces src/2D-DCT.py:398-466 between decompress and decode_write_fn --
 *   astype(int16) :398, -=128 :402, dequantize_decom :410 (src/deadzone.py:107-120),
 *   get_blocks :416, perceptual :421-435, space_synthesize :440, remove_padding :444,
 *   to_RGB :449, += 128 :454, [filter hook :461], clip + astype(uint8) :466.
 * idx        (n_frames,Hp,Wp,3) uint8, device
 * rgb_out    (n_frames,H,W,3) uint8, device, or NULL
 * y_out      (n_frames,H,W,3) un-clipped image handed to CT.CoDec.filter (:461):
 *            float64 with VCFB_F_FP64, float32 otherwise; device, or NULL
 * original   (n_frames,H,W,3) uint8, device, or NULL; with `stats` adds the SSE
 *            between it and the clipped uint8 result (src/RDE.py:12-55) */
int vcfb_decode_dev(const uint8_t* idx, int n_frames, int H, int W, int B, double q,
                    int color, unsigned flags, const double* weights,
                    uint8_t* rgb_out, void* y_out, const uint8_t* original,
                    int64_t* stats, void* cuda_stream);

/* Fused rate/distortion sweep: every quantisation step of one block size in ONE pass over
 * the frames -- forward transform once, then per step quantise / dequantise / the float64 decode chain / SSE
 * against the input, all on chip; nothing but statistics is written (3 B/pixel of HBM traffic per block size).
 * Replaces the body of the loop of src/2D-DCT.py:533-579 (optimize_block_size) and the encode + decode + RDE
 * runs per point of src/RDE.py:68-118; the numbers are those vcfb_encode_dev (float32, bit-exact) followed by
 * vcfb_decode_dev (VCFB_F_FP64, bit-exact) with `original` accumulate, step for step.
 * q_steps   n_steps quantisation steps (host array), n_steps <= VCFB_RD_MAX_STEPS
 * flags     VCFB_F_HIST, VCFB_F_NOWRAP, VCFB_F_NO_OFFSET; no perceptual weights (the reference's loop applies none)
 * stats     n_steps x VCFB_STAT_LEN int64 on the device, zeroed by the caller: row i is the statistics vector of
 *           q_steps[i] (SSE, NSAMPLES, SUMDIFF of the decoded image; NONZERO, SUMABS, NINDICES, HIST of the indices) */
#define VCFB_RD_MAX_STEPS 16
int vcfb_rd_sweep_dev(const uint8_t* rgb, int n_frames, int H, int W, int B, const double* q_steps, int n_steps,
                      int color, unsigned flags, int64_t* stats, void* cuda_stream);

/* Stand-alone colour codecs: the reference's colour stages run as codecs of their own
 * (`python YCoCg.py encode`, `python YCrCb.py encode`), colour transform + deadzone
 * quantiser without a spatial transform.
 *   encode replaces src/YCoCg.py:36-52 (color = VCFB_COLOR_YCOCG: astype(int16), from_RGB
 *          stored into int16, x / q truncated, astype(uint16)) or src/YCrCb.py:36-47
 *          (VCFB_COLOR_YCRCB: OpenCV 8-bit fixed-point RGB2YCrCb, then the same quantiser);
 *   decode replaces src/YCoCg.py:61-78 / src/YCrCb.py:56-66 (note the uint8 cast before
 *          to_RGB at :59).
 * rgb: n_pixels x 3 uint8; k: n_pixels x 3 uint16; device pointers. */
int vcfb_color_encode_dev(const uint8_t* rgb, long long n_pixels, double q, int color,
                          uint16_t* k_out, void* cuda_stream);
int vcfb_color_decode_dev(const uint16_t* k, long long n_pixels, double q, int color,
                          uint8_t* rgb_out, void* cuda_stream);

/* Motion estimation of the hybrid codec (src/IPP_DCT.py).
 *   vcfb_gray_dev         replaces cv2.cvtColor(frame, cv2.COLOR_RGB2GRAY) of
 *                         src/IPP_DCT.py:350-352 (8-bit fixed point, bit-identical with OpenCV).
 *   vcfb_block_match_dev  replaces IPP.block_matching's full search (src/IPP_DCT.py:217-244 and
 *                         :344-373): per bs x bs block the displacement in [-sr, sr]^2 minimising
 *                         the sum of absolute differences; candidates leaving the frame are
 *                         skipped, ties go to the first candidate of the dy-major scan.
 * ref, cur   (n_frames,H,W) uint8 gray frames, device; pair f is matched independently
 * mv_out     (n_frames, H/bs, W/bs, 2) int16, device: (dx, dy) as the reference stores them
 * bs a multiple of 4 in [4, 64], sr in [0, 31], H >= bs, W >= bs.
 *   vcfb_block_match_tss_dev  the `--fast` variant: `_three_step_search`, src/IPP_DCT.py:159-205, with
 *                         its moving centre and its walk at step 1 (the vector may leave [-sr, sr]);
 *                         bs in [1, 64], sr in [0, 32767]. */
int vcfb_gray_dev(const uint8_t* rgb, long long n_pixels, uint8_t* gray_out, void* cuda_stream);
int vcfb_block_match_dev(const uint8_t* ref, const uint8_t* cur, int n_frames, int H, int W, int bs, int sr,
                         int16_t* mv_out, void* cuda_stream);
int vcfb_block_match_tss_dev(const uint8_t* ref, const uint8_t* cur, int n_frames, int H, int W, int bs, int sr,
                             int16_t* mv_out, void* cuda_stream);

/* Entropy front-end: raw deflate (RFC 1951) of a byte array -- the zlib call
 * underneath the reference's entropy stage for the uint8 index planes (np.savez_compressed in
 * src/z_lib.py:19-23; tifffile's zlib codec in src/TIFF.py:23-31).  Run-length parse (distance-1
 * matches, zlib's Z_RLE strategy, each taken only where a sampled cost model says it is cheaper than
 * its bytes as literals) + one dynamic Huffman block per segment of 132 KB, stored blocks where
 * those are smaller.  The stream is complete (last block has BFINAL = 1) and any
 * inflate implementation reads it: zlib.decompress(stream, -15); prefix 78 9C and append the
 * big-endian Adler-32 for a zlib stream; wrap in a zip member with its CRC-32 for .npz.
 * It is NOT byte-identical with zlib's output -- the property kept is that the reference's decoder
 * (np.load / tifffile.imread / zlib.decompress) returns the same array.
 * src        n_bytes bytes, device, 8-byte aligned
 * dst        device, dst_capacity >= vcfb_deflate_bound(n_bytes)
 * out_bytes  one uint64 on the device: length of the stream in dst
 * workspace  device, 16-byte aligned, >= vcfb_deflate_workspace(n_bytes) bytes (about 3 * n_bytes:
 *            the segments' streams before they are packed, and 16 bits per byte for the tokens)
 * Asynchronous on cuda_stream; three kernels (segments, scan, gather). */
size_t vcfb_deflate_bound(size_t n_bytes);
size_t vcfb_deflate_workspace(size_t n_bytes);
int vcfb_deflate_dev(const uint8_t* src, size_t n_bytes, uint8_t* dst, size_t dst_capacity,
                     uint64_t* out_bytes, void* workspace, size_t workspace_bytes, void* cuda_stream);

/* The same for an array of rows: src is rows of row_bytes bytes whose samples lie sample_bytes
 * apart (the H x W x 3 uint8 index image the reference's entropy stage receives --
 * src/z_lib.py:19-23, src/TIFF.py:23-31 -- has row_bytes = 3 * W, sample_bytes = 3).  Besides runs
 * the parse then tries the previous sample of the same channel and the three samples above the
 * current one (distances sample_bytes, row_bytes, row_bytes -+ sample_bytes: where zlib's hash
 * chains find most of their matches in such planes) and takes a match when the cost model puts it
 * below the run-length parse of the bytes it covers.  Candidate distances beyond deflate's 32 KB
 * window are dropped; row_bytes = 0 is vcfb_deflate_dev.  Same buffers, bound and workspace. */
int vcfb_deflate_rows_dev(const uint8_t* src, size_t n_bytes, size_t row_bytes, int sample_bytes, uint8_t* dst,
                          size_t dst_capacity, uint64_t* out_bytes, void* workspace, size_t workspace_bytes,
                          void* cuda_stream);

/* CRC-32 (zip / zlib / PNG polynomial, the value zlib.crc32 returns) of n_bytes bytes on the device:
 * the checksum a zip member carries next to its deflate stream (np.savez_compressed,
 * src/z_lib.py:19-23).  src 8-byte aligned; out_crc one uint32 on the device.  Asynchronous on
 * cuda_stream (a 4-byte memset and one kernel).  CRCs of consecutive parts combine on the host:
 * vcf_b200.entropy.crc32_combine. */
int vcfb_crc32_dev(const uint8_t* src, size_t n_bytes, uint32_t* out_crc, void* cuda_stream);

/* Adler-32 (RFC 1950; the value zlib.adler32 returns) of n_bytes bytes on the device: the checksum
 * that closes a zlib stream -- the strips tifffile writes with compression='zlib', src/TIFF.py:23-31.
 * src 8-byte aligned; out_adler one uint32 on the device; workspace16 16 bytes on the device, 8-byte
 * aligned.  Asynchronous on cuda_stream (a memset and two kernels). */
int vcfb_adler32_dev(const uint8_t* src, size_t n_bytes, uint32_t* out_adler, void* workspace16, void* cuda_stream);

/* Host-buffer convenience layer (what a numpy caller binds).  A context owns one
 * CUDA stream plus pinned and device staging buffers that grow on demand. */
typedef struct vcfb_ctx vcfb_ctx;
int vcfb_ctx_create(int device, vcfb_ctx** out);
void vcfb_ctx_destroy(vcfb_ctx* ctx);

/* Pinned (page-locked) host memory.  Host buffers obtained here -- or pinned by
 * any other means -- are transferred in place; pageable buffers are staged through
 * the context's own pinned memory at the cost of one extra host copy. */
int vcfb_host_alloc(size_t bytes, void** out);
void vcfb_host_free(void* p);

/* Same contracts as the *_dev calls with HOST pointers (weights, stats too);
 * stats (may be NULL) receives this call's statistics (overwritten, not added).
 * A batch is cut into chunks of whole frames that are copied in, transformed and
 * copied out on three streams so both copy engines and the SMs overlap. */
int vcfb_encode_host(vcfb_ctx* ctx, const uint8_t* rgb, int n_frames, int H, int W, int B,
                     double q, int color, unsigned flags, const double* weights,
                     uint8_t* idx_out, int64_t* stats);
int vcfb_decode_host(vcfb_ctx* ctx, const uint8_t* idx, int n_frames, int H, int W, int B,
                     double q, int color, unsigned flags, const double* weights,
                     uint8_t* rgb_out, void* y_out, const uint8_t* original, int64_t* stats);

int vcfb_color_encode_host(vcfb_ctx* ctx, const uint8_t* rgb, long long n_pixels, double q, int color,
                           uint16_t* k_out);
int vcfb_color_decode_host(vcfb_ctx* ctx, const uint16_t* k, long long n_pixels, double q, int color,
                           uint8_t* rgb_out);

#ifdef __cplusplus
}
#endif
#endif /* VCFB200_H */
