// C ABI of libvcfb200 (include/vcfb200.h): argument checking, geometry, dispatch,
// and the host-buffer convenience layer (pinned staging + one stream per context).
#include <math.h>
#include <string.h>

#include <new>

#include "common.cuh"

namespace vcfb {

static thread_local std::string g_err;

void set_error(const std::string& msg) { g_err = msg; }

int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return VCFB_E_CUDA;
}

static bool make_geom(int H, int W, int B, Geom* g) {
  if (H <= 0 || W <= 0 || (B != 4 && B != 8 && B != 16 && B != 32)) return false;
  g->H = H;
  g->W = W;
  g->Hp = (H + B - 1) / B * B;   // src/2D-DCT.py:208-209
  g->Wp = (W + B - 1) / B * B;
  g->top = (g->Hp - H) / 2;      // :216-219 (remainder goes bottom / right)
  g->left = (g->Wp - W) / 2;
  g->ny = g->Hp / B;
  g->nx = g->Wp / B;
  return true;
}

static int check_common(const void* in, int n_frames, int H, int W, int B, double q, int color,
                        unsigned flags, const double* weights, Geom* g) {
  if (!in) { set_error("input pointer is NULL"); return VCFB_E_ARG; }
  if (n_frames <= 0 || n_frames > 65535) { set_error("n_frames must be in [1, 65535]"); return VCFB_E_ARG; }
  if (!make_geom(H, W, B, g)) { set_error("bad H/W or unsupported block size (supported B: 4, 8, 16, 32)"); return VCFB_E_ARG; }
  if (g->ny > 65535) { set_error("frame too tall for this block size"); return VCFB_E_ARG; }
  if (!(q > 0.0) || !isfinite(q)) { set_error("quantisation step q must be finite and > 0"); return VCFB_E_ARG; }
  if (color != VCFB_COLOR_YCOCG && color != VCFB_COLOR_YCRCB) { set_error("unknown colour transform"); return VCFB_E_ARG; }
  if ((flags & VCFB_F_PERCEPTUAL) && !weights) { set_error("VCFB_F_PERCEPTUAL needs weights"); return VCFB_E_ARG; }
  if ((flags & VCFB_F_FP64) && (flags & VCFB_F_CONTRACT)) { set_error("VCFB_F_CONTRACT is float32 only"); return VCFB_E_ARG; }
  if (flags & ~(VCFB_F_NO_SUBBANDS | VCFB_F_PERCEPTUAL | VCFB_F_FP64 | VCFB_F_CONTRACT)) { set_error("unknown flag bits"); return VCFB_E_ARG; }
  return VCFB_OK;
}

}  // namespace vcfb

using namespace vcfb;

extern "C" {

int vcfb_version(void) { return VCFB_VERSION; }

const char* vcfb_last_error(void) { return g_err.c_str(); }

int vcfb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int vcfb_padded_dims(int H, int W, int B, int* Hp, int* Wp, int* top, int* left) {
  Geom g;
  if (H <= 0 || W <= 0 || B <= 0) { set_error("bad H/W/B"); return VCFB_E_ARG; }
  g.Hp = (H + B - 1) / B * B;
  g.Wp = (W + B - 1) / B * B;
  if (Hp) *Hp = g.Hp;
  if (Wp) *Wp = g.Wp;
  if (top) *top = (g.Hp - H) / 2;
  if (left) *left = (g.Wp - W) / 2;
  return VCFB_OK;
}

int vcfb_encode_dev(const uint8_t* rgb, int n_frames, int H, int W, int B, double q, int color,
                    unsigned flags, const double* weights, uint8_t* idx_out, int64_t* stats,
                    void* cuda_stream) {
  EncArgs a;
  memset(&a, 0, sizeof(a));
  int rc = check_common(rgb, n_frames, H, W, B, q, color, flags, weights, &a.g);
  if (rc) return rc;
  if (!idx_out) { set_error("idx_out is NULL"); return VCFB_E_ARG; }
  a.rgb = rgb;
  a.idx = idx_out;
  a.n_frames = n_frames;
  a.q = q;
  a.inv_q = 1.0 / q;
  int e2;
  a.q_pow2 = (frexp(q, &e2) == 0.5);   // x / q == x * (1/q) exactly iff q is a power of two
  a.color = color;
  a.flags = flags;
  a.weights = weights;
  a.stats = reinterpret_cast<unsigned long long*>(stats);
  return launch_encode_general(a, B, static_cast<cudaStream_t>(cuda_stream));
}

int vcfb_decode_dev(const uint8_t* idx, int n_frames, int H, int W, int B, double q, int color,
                    unsigned flags, const double* weights, uint8_t* rgb_out, void* y_out,
                    const uint8_t* original, int64_t* stats, void* cuda_stream) {
  DecArgs a;
  memset(&a, 0, sizeof(a));
  int rc = check_common(idx, n_frames, H, W, B, q, color, flags, weights, &a.g);
  if (rc) return rc;
  if (!rgb_out && !y_out && !(original && stats)) { set_error("decode has no output"); return VCFB_E_ARG; }
  a.idx = idx;
  a.rgb = rgb_out;
  a.y_out = y_out;
  a.original = original;
  a.n_frames = n_frames;
  a.q = q;
  a.q_int = (q == floor(q) && q < 32768.0) ? int(q) : 0;
  a.color = color;
  a.flags = flags;
  a.weights = weights;
  a.stats = reinterpret_cast<unsigned long long*>(stats);
  return launch_decode_general(a, B, static_cast<cudaStream_t>(cuda_stream));
}

// ---- host-buffer layer ----------------------------------------------------------

struct vcfb_ctx {
  int device;
  cudaStream_t stream;
  void* pin;   size_t pin_cap;    // pinned host staging
  void* dev;   size_t dev_cap;    // device staging
};

static int ctx_reserve(vcfb_ctx* c, size_t pin_bytes, size_t dev_bytes) {
  cudaError_t e;
  if (pin_bytes > c->pin_cap) {
    if (c->pin) cudaFreeHost(c->pin);
    c->pin = nullptr; c->pin_cap = 0;
    e = cudaMallocHost(&c->pin, pin_bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMallocHost");
    c->pin_cap = pin_bytes;
  }
  if (dev_bytes > c->dev_cap) {
    if (c->dev) cudaFree(c->dev);
    c->dev = nullptr; c->dev_cap = 0;
    e = cudaMalloc(&c->dev, dev_bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    c->dev_cap = dev_bytes;
  }
  return VCFB_OK;
}

int vcfb_ctx_create(int device, vcfb_ctx** out) {
  if (!out) { set_error("out is NULL"); return VCFB_E_ARG; }
  *out = nullptr;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  vcfb_ctx* c = new (std::nothrow) vcfb_ctx();
  if (!c) { set_error("out of memory"); return VCFB_E_ARG; }
  memset(c, 0, sizeof(*c));
  c->device = device;
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate"); }
  *out = c;
  return VCFB_OK;
}

void vcfb_ctx_destroy(vcfb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->pin) cudaFreeHost(c->pin);
  if (c->dev) cudaFree(c->dev);
  delete c;
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

int vcfb_encode_host(vcfb_ctx* c, const uint8_t* rgb, int n_frames, int H, int W, int B, double q,
                     int color, unsigned flags, const double* weights, uint8_t* idx_out,
                     int64_t* stats) {
  if (!c) { set_error("ctx is NULL"); return VCFB_E_ARG; }
  Geom g;
  int rc = check_common(rgb, n_frames, H, W, B, q, color, flags, weights, &g);
  if (rc) return rc;
  if (!idx_out) { set_error("idx_out is NULL"); return VCFB_E_ARG; }
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  const size_t in_b = size_t(n_frames) * H * W * 3, out_b = size_t(n_frames) * g.Hp * g.Wp * 3;
  const size_t w_b = (flags & VCFB_F_PERCEPTUAL) ? size_t(2) * B * B * sizeof(double) : 0;
  const size_t st_b = stats ? VCFB_STAT_LEN * sizeof(int64_t) : 0;
  const size_t o_in = 0, o_out = align256(in_b), o_w = o_out + align256(out_b), o_st = o_w + align256(w_b);
  const size_t total = o_st + align256(st_b);
  rc = ctx_reserve(c, total, total);
  if (rc) return rc;
  char* hp = static_cast<char*>(c->pin);
  char* dp = static_cast<char*>(c->dev);
  memcpy(hp + o_in, rgb, in_b);
  if (w_b) memcpy(hp + o_w, weights, w_b);
  e = cudaMemcpyAsync(dp + o_in, hp + o_in, in_b, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess && w_b) e = cudaMemcpyAsync(dp + o_w, hp + o_w, w_b, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess && st_b) e = cudaMemsetAsync(dp + o_st, 0, st_b, c->stream);
  if (e != cudaSuccess) return cuda_fail(e, "host->device copy");
  rc = vcfb_encode_dev(reinterpret_cast<uint8_t*>(dp + o_in), n_frames, H, W, B, q, color, flags,
                       w_b ? reinterpret_cast<double*>(dp + o_w) : nullptr,
                       reinterpret_cast<uint8_t*>(dp + o_out),
                       st_b ? reinterpret_cast<int64_t*>(dp + o_st) : nullptr, c->stream);
  if (rc) return rc;
  e = cudaMemcpyAsync(hp + o_out, dp + o_out, out_b, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess && st_b) e = cudaMemcpyAsync(hp + o_st, dp + o_st, st_b, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) return cuda_fail(e, "encode (device->host / synchronize)");
  memcpy(idx_out, hp + o_out, out_b);
  if (st_b) memcpy(stats, hp + o_st, st_b);
  return VCFB_OK;
}

int vcfb_decode_host(vcfb_ctx* c, const uint8_t* idx, int n_frames, int H, int W, int B, double q,
                     int color, unsigned flags, const double* weights, uint8_t* rgb_out, void* y_out,
                     const uint8_t* original, int64_t* stats) {
  if (!c) { set_error("ctx is NULL"); return VCFB_E_ARG; }
  Geom g;
  int rc = check_common(idx, n_frames, H, W, B, q, color, flags, weights, &g);
  if (rc) return rc;
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  const size_t px_b = size_t(n_frames) * H * W * 3, idx_b = size_t(n_frames) * g.Hp * g.Wp * 3;
  const size_t y_b = y_out ? px_b * ((flags & VCFB_F_FP64) ? 8 : 4) : 0;
  const size_t or_b = original ? px_b : 0;
  const size_t w_b = (flags & VCFB_F_PERCEPTUAL) ? size_t(2) * B * B * sizeof(double) : 0;
  const size_t st_b = stats ? VCFB_STAT_LEN * sizeof(int64_t) : 0;
  const size_t o_idx = 0, o_rgb = align256(idx_b), o_y = o_rgb + align256(px_b), o_or = o_y + align256(y_b),
               o_w = o_or + align256(or_b), o_st = o_w + align256(w_b);
  const size_t total = o_st + align256(st_b);
  rc = ctx_reserve(c, total, total);
  if (rc) return rc;
  char* hp = static_cast<char*>(c->pin);
  char* dp = static_cast<char*>(c->dev);
  memcpy(hp + o_idx, idx, idx_b);
  if (or_b) memcpy(hp + o_or, original, or_b);
  if (w_b) memcpy(hp + o_w, weights, w_b);
  e = cudaMemcpyAsync(dp + o_idx, hp + o_idx, idx_b, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess && or_b) e = cudaMemcpyAsync(dp + o_or, hp + o_or, or_b, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess && w_b) e = cudaMemcpyAsync(dp + o_w, hp + o_w, w_b, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess && st_b) e = cudaMemsetAsync(dp + o_st, 0, st_b, c->stream);
  if (e != cudaSuccess) return cuda_fail(e, "host->device copy");
  rc = vcfb_decode_dev(reinterpret_cast<uint8_t*>(dp + o_idx), n_frames, H, W, B, q, color, flags,
                       w_b ? reinterpret_cast<double*>(dp + o_w) : nullptr,
                       rgb_out ? reinterpret_cast<uint8_t*>(dp + o_rgb) : nullptr,
                       y_b ? static_cast<void*>(dp + o_y) : nullptr,
                       or_b ? reinterpret_cast<uint8_t*>(dp + o_or) : nullptr,
                       st_b ? reinterpret_cast<int64_t*>(dp + o_st) : nullptr, c->stream);
  if (rc) return rc;
  if (rgb_out) e = cudaMemcpyAsync(hp + o_rgb, dp + o_rgb, px_b, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess && y_b) e = cudaMemcpyAsync(hp + o_y, dp + o_y, y_b, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess && st_b) e = cudaMemcpyAsync(hp + o_st, dp + o_st, st_b, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) return cuda_fail(e, "decode (device->host / synchronize)");
  if (rgb_out) memcpy(rgb_out, hp + o_rgb, px_b);
  if (y_b) memcpy(y_out, hp + o_y, y_b);
  if (st_b) memcpy(stats, hp + o_st, st_b);
  return VCFB_OK;
}

}  // extern "C"
