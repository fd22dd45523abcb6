// Row F3 of SURVEY.md section 8: the motion-estimation hot spot of the reference's hybrid
// codec (src/IPP_DCT.py), the second caller of encode_fn / decode_fn.
//
//   gray_kernel          cv2.cvtColor(frame, COLOR_RGB2GRAY) on 8-bit data (src/IPP_DCT.py:350-352):
//                        OpenCV's fixed point  (R*9798 + G*19235 + B*3735 + 2^14) >> 15
//   block_match_kernel   full search of src/IPP_DCT.py:217-244 (`_process_block_row`, use_fast =
//                        False): for every bs x bs block of the current frame the displacement
//                        (dx, dy) in [-sr, sr]^2 with the smallest sum of absolute differences
//                        against the reference frame; candidates that leave the frame are skipped
//                        (:227-233), the scan is dy-major and a later candidate wins only when
//                        strictly better (:240) -- so ties go to the first candidate in scan order.
//
// Integer arithmetic throughout: results are bit-identical with the reference's numpy loop.
// One CTA per block: the search window and the block sit in shared memory, each thread owns
// candidates, and the minimum is taken over the key (SAD << 12 | scan index).
#include "common.cuh"

namespace vcfb {
namespace {

__global__ void gray_kernel(const uint8_t* __restrict__ rgb, uint8_t* __restrict__ gray, long long npx) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npx; p += stride) {
    const unsigned r = rgb[3 * p], g = rgb[3 * p + 1], b = rgb[3 * p + 2];
    gray[p] = uint8_t((r * 9798u + g * 19235u + b * 3735u + (1u << 14)) >> 15);
  }
}

constexpr int ME_THREADS = 256;

__global__ void __launch_bounds__(ME_THREADS)
block_match_kernel(const uint8_t* __restrict__ ref, const uint8_t* __restrict__ cur, int H, int W, int bs, int sr,
                   short* __restrict__ mv) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int ww = bs + 2 * sr;                 // window side
  const int wp = (ww + 3 + 4) & ~3;           // pitch: whole words, one spare word for the funnel shift
  unsigned char* win = sm;                    // ww x wp
  unsigned char* blk = sm + ww * wp;          // bs x bs (bs is a multiple of 4)
  __shared__ unsigned best_s[ME_THREADS / 32];

  const int bx = blockIdx.x, by = blockIdx.y, f = blockIdx.z;
  const int i0 = by * bs, j0 = bx * bs;
  const uint8_t* rf = ref + (size_t)f * H * W;
  const uint8_t* cf = cur + (size_t)f * H * W;
  for (int t = threadIdx.x; t < ww * wp; t += ME_THREADS) {
    const int y = t / wp, x = t - y * wp;
    const int gy = i0 - sr + y, gx = j0 - sr + x;
    win[t] = (x < ww && gy >= 0 && gy < H && gx >= 0 && gx < W) ? rf[(size_t)gy * W + gx] : 0;
  }
  for (int t = threadIdx.x; t < bs * bs; t += ME_THREADS) {
    const int y = t / bs, x = t - y * bs;
    blk[t] = cf[(size_t)(i0 + y) * W + j0 + x];
  }
  __syncthreads();

  const int nc1 = 2 * sr + 1, ncand = nc1 * nc1;
  const int bw = bs >> 2;                     // words per block row
  unsigned best = 0xffffffffu;
  for (int c = threadIdx.x; c < ncand; c += ME_THREADS) {
    const int dy = c / nc1 - sr, dx = c - (c / nc1) * nc1 - sr;
    const int ry = i0 + dy, rx = j0 + dx;
    if (ry < 0 || ry + bs > H || rx < 0 || rx + bs > W) continue;     // src/IPP_DCT.py:227-233
    const int off = dx + sr, sh = (off & 3) * 8;
    const uint32_t* wrow = reinterpret_cast<const uint32_t*>(win + (dy + sr) * wp) + (off >> 2);
    const uint32_t* brow = reinterpret_cast<const uint32_t*>(blk);
    unsigned sad = 0;
    for (int y = 0; y < bs; ++y) {
      uint32_t lo = wrow[0];
      for (int k = 0; k < bw; ++k) {
        const uint32_t hi = wrow[k + 1];
        sad = __vsadu4(brow[k], __funnelshift_r(lo, hi, sh)) + sad;     // 4 pixels per instruction
        lo = hi;
      }
      wrow += wp >> 2;
      brow += bw;
    }
    best = min(best, (sad << 12) | unsigned(c));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) best_s[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 1; k < ME_THREADS / 32; ++k) best = min(best, best_s[k]);
    const int c = int(best & 0xfffu);
    // (0,0) is always inside the frame, so `best` is never the empty key
    short* o = mv + (((size_t)f * gridDim.y + by) * gridDim.x + bx) * 2;
    o[0] = short(c - (c / nc1) * nc1 - sr);    // dx
    o[1] = short(c / nc1 - sr);                // dy
  }
}

// Three-step search of src/IPP_DCT.py:159-205 (`_three_step_search`), one warp per block, every
// quirk kept: the centre moves as soon as a neighbour improves the SAD, so the remaining
// neighbours of the same round are taken around the NEW centre; after a round with an
// improvement the step is max(1, step // 2), so the search keeps walking with step 1 until a
// round brings nothing -- the vector is not confined to the search range.  The reference frame
// is read from global memory (the walk has no fixed window).
__device__ __forceinline__ unsigned warp_sad(const uint8_t* __restrict__ cb, const uint8_t* __restrict__ rf, int W, int bs,
                                             int cy, int cx, int lane) {
  unsigned sad = 0;
  for (int p = lane; p < bs * bs; p += 32) {
    const int y = p / bs, x = p - y * bs;
    sad += __sad(int(cb[y * bs + x]), int(rf[(size_t)(cy + y) * W + cx + x]), 0u);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sad += __shfl_xor_sync(0xffffffffu, sad, o);
  return sad;
}

__global__ void __launch_bounds__(128)
tss_kernel(const uint8_t* __restrict__ ref, const uint8_t* __restrict__ cur, int H, int W, int bs, int sr, int nbx,
           int nby, int n_frames, short* __restrict__ mv) {
  extern __shared__ unsigned char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* cb = sm + warp * bs * bs;            // the current block of this warp
  const long long nblk = (long long)n_frames * nby * nbx;
  for (long long blk = (long long)blockIdx.x * 4 + warp; blk < nblk; blk += (long long)gridDim.x * 4) {
    const int f = int(blk / ((long long)nby * nbx));
    const int rem = int(blk - (long long)f * nby * nbx);
    const int by = rem / nbx, bx = rem - by * nbx;
    const int i = by * bs, j = bx * bs;
    const uint8_t* rf = ref + (size_t)f * H * W;
    const uint8_t* cf = cur + (size_t)f * H * W;
    __syncwarp();
    for (int p = lane; p < bs * bs; p += 32) cb[p] = cf[(size_t)(i + p / bs) * W + j + p % bs];
    __syncwarp();
    int step = sr / 2, cx = j, cy = i, bdx = 0, bdy = 0;
    unsigned min_sad = warp_sad(cb, rf, W, bs, cy, cx, lane);       // the block position itself is always inside
    while (step >= 1) {
      bool improved = false;
      for (int a = -1; a <= 1; ++a)
        for (int b = -1; b <= 1; ++b) {
          if (a == 0 && b == 0) continue;
          const int ry = cy + a * step, rx = cx + b * step;          // around the centre as it is NOW
          if (ry < 0 || ry + bs > H || rx < 0 || rx + bs > W) continue;
          const unsigned sad = warp_sad(cb, rf, W, bs, ry, rx, lane);
          if (sad < min_sad) {
            min_sad = sad;
            bdx = rx - j;
            bdy = ry - i;
            cx = rx;
            cy = ry;
            improved = true;
          }
        }
      step = improved ? max(1, step / 2) : step / 2;
    }
    if (lane == 0) {
      short* o = mv + blk * 2;
      o[0] = short(bdx);
      o[1] = short(bdy);
    }
  }
}

}  // namespace

int launch_gray(const uint8_t* rgb, long long npx, uint8_t* gray, cudaStream_t s) {
  long long g = (npx + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  note_kernel("gray");
  gray_kernel<<<int(g), 256, 0, s>>>(rgb, gray, npx);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "gray_kernel launch");
  return VCFB_OK;
}

int launch_block_match(const uint8_t* ref, const uint8_t* cur, int n_frames, int H, int W, int bs, int sr, short* mv,
                       cudaStream_t s) {
  const int ww = bs + 2 * sr;
  const int smem = ww * ((ww + 3 + 4) & ~3) + bs * bs;
  dim3 grid(W / bs, H / bs, n_frames);
  note_kernel("block_match");
  block_match_kernel<<<grid, ME_THREADS, smem, s>>>(ref, cur, H, W, bs, sr, mv);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "block_match_kernel launch");
  return VCFB_OK;
}

int launch_block_match_tss(const uint8_t* ref, const uint8_t* cur, int n_frames, int H, int W, int bs, int sr, short* mv,
                           cudaStream_t s) {
  const int nbx = W / bs, nby = H / bs;
  const long long nblk = (long long)n_frames * nby * nbx;
  long long grid = (nblk + 3) / 4;
  if (grid > 148 * 16) grid = 148 * 16;
  note_kernel("block_match_tss");
  tss_kernel<<<int(grid), 128, 4 * bs * bs, s>>>(ref, cur, H, W, bs, sr, nbx, nby, n_frames, mv);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "tss_kernel launch");
  return VCFB_OK;
}

}  // namespace vcfb
