// C ABI of libvcfb200 (include/vcfb200.h): argument checking, geometry, dispatch,
// and the host-buffer convenience layer (pinned staging + one stream per context).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "common.cuh"

namespace vcfb {

static thread_local std::string g_err;

static thread_local const char* g_kernel = "";

void set_error(const std::string& msg) { g_err = msg; }
static thread_local long long g_launches = 0;
void note_kernel(const char* name) {
  g_kernel = name;
  ++g_launches;
}
void note_extra_launches(int n) { g_launches += n; }

int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return VCFB_E_CUDA;
}

static bool make_geom(int H, int W, int B, Geom* g) {
  if (H <= 0 || W <= 0 || (B != 4 && B != 8 && B != 16 && B != 32 && !anyb_supported(B))) return false;
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
  if (!make_geom(H, W, B, g)) { set_error("bad H/W or unsupported block size (supported B: 2, 4, 8, 16, 32, 64, 128)"); return VCFB_E_ARG; }
  if (g->ny > 65535) { set_error("frame too tall for this block size"); return VCFB_E_ARG; }
  if (!(q > 0.0) || !isfinite(q)) { set_error("quantisation step q must be finite and > 0"); return VCFB_E_ARG; }
  if (color != VCFB_COLOR_YCOCG && color != VCFB_COLOR_YCRCB) { set_error("unknown colour transform"); return VCFB_E_ARG; }
  if ((flags & VCFB_F_PERCEPTUAL) && !weights) { set_error("VCFB_F_PERCEPTUAL needs weights"); return VCFB_E_ARG; }
  if ((flags & VCFB_F_FP64) && (flags & VCFB_F_CONTRACT)) { set_error("VCFB_F_CONTRACT is float32 only"); return VCFB_E_ARG; }
  if ((flags & VCFB_F_SYNTH_F32) && !(flags & VCFB_F_FP64)) { set_error("VCFB_F_SYNTH_F32 is a variant of the float64 decoder: set VCFB_F_FP64 too"); return VCFB_E_ARG; }
  if (flags & ~(VCFB_F_NO_SUBBANDS | VCFB_F_PERCEPTUAL | VCFB_F_FP64 | VCFB_F_CONTRACT | VCFB_F_HIST | VCFB_F_SYNTH_F32 | VCFB_F_FAST | VCFB_F_NO_OFFSET)) { set_error("unknown flag bits"); return VCFB_E_ARG; }
  return VCFB_OK;
}

}  // namespace vcfb

using namespace vcfb;

extern "C" {

int vcfb_version(void) { return VCFB_VERSION; }

const char* vcfb_last_error(void) { return g_err.c_str(); }

const char* vcfb_last_kernel(void) { return g_kernel; }

long long vcfb_launch_count(void) { return g_launches; }

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
  if (flags & VCFB_F_SYNTH_F32) { set_error("VCFB_F_SYNTH_F32 is a decode flag"); return VCFB_E_ARG; }
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
  if (anyb_supported(B)) return launch_encode_anyb(a, B, static_cast<cudaStream_t>(cuda_stream));
  if (flags & VCFB_F_NO_OFFSET) return launch_encode_general(a, B, static_cast<cudaStream_t>(cuda_stream));
  if ((flags & VCFB_F_FAST) && B == 8 && !(flags & VCFB_F_FP64)) {
    // fast mode: tensor-core encoder; statistics, when asked for, by the streaming pass over the indices
    static const bool tc_off = getenv("VCFB_TC") && getenv("VCFB_TC")[0] == '0';
    if (!tc_off) {
      // statistics: the sums come out of the encoder's epilogue; the histogram, when asked for, takes the streaming
      // pass over the stored indices (kernels_stats.cu)
      const bool hist = a.stats && (flags & VCFB_F_HIST);
      EncArgs t = a;
      if (hist) t.stats = nullptr;
      rc = launch_encode_tc(t, static_cast<cudaStream_t>(cuda_stream));
      if (rc == VCFB_OK && hist)
        rc = launch_index_stats(a.idx, (long long)n_frames * a.g.Hp * a.g.Wp * 3, true, a.stats,
                                static_cast<cudaStream_t>(cuda_stream));
      if (rc != VCFB_E_UNSUPP) return rc;
    }
  }
  rc = B == 16 ? launch_encode_fast16(a, static_cast<cudaStream_t>(cuda_stream))
               : launch_encode_fast(a, B, static_cast<cudaStream_t>(cuda_stream));
  if (rc != VCFB_E_UNSUPP) return rc;
  rc = launch_encode_tile(a, B, static_cast<cudaStream_t>(cuda_stream));      // B = 32
  if (rc != VCFB_E_UNSUPP) return rc;
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
  if (flags & VCFB_F_NO_OFFSET) { set_error("VCFB_F_NO_OFFSET is an encode / rd_sweep flag (the reference's decoder always runs with offset 128)"); return VCFB_E_ARG; }
  if (q == floor(q) && q >= 32768.0) {
    // numpy refuses `python int * int16 array` when the int does not fit int16 (src/2D-DCT.py:410)
    set_error("integral quantisation step does not fit int16: the reference's dequantiser raises OverflowError");
    return VCFB_E_ARG;
  }
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
  if (anyb_supported(B)) return launch_decode_anyb(a, B, static_cast<cudaStream_t>(cuda_stream));
  if (!(flags & VCFB_F_SYNTH_F32)) {      // the upstream-variant decoder exists in the general kernel only
    rc = B == 16 ? launch_decode_fast16(a, static_cast<cudaStream_t>(cuda_stream))
                 : launch_decode_fast(a, B, static_cast<cudaStream_t>(cuda_stream));
    if (rc != VCFB_E_UNSUPP) return rc;
    rc = launch_decode_tile(a, B, static_cast<cudaStream_t>(cuda_stream));    // B = 32
    if (rc != VCFB_E_UNSUPP) return rc;
  }
  return launch_decode_general(a, B, static_cast<cudaStream_t>(cuda_stream));
}

int vcfb_rd_sweep_dev(const uint8_t* rgb, int n_frames, int H, int W, int B, const double* q_steps, int n_steps,
                      int color, unsigned flags, int64_t* stats, void* cuda_stream) {
  Geom g;
  if (!q_steps || n_steps < 1 || n_steps > VCFB_RD_MAX_STEPS) { set_error("n_steps must be in [1, VCFB_RD_MAX_STEPS]"); return VCFB_E_ARG; }
  if (!stats) { set_error("stats is NULL"); return VCFB_E_ARG; }
  if (flags & ~(VCFB_F_HIST | VCFB_F_NOWRAP | VCFB_F_NO_OFFSET)) { set_error("vcfb_rd_sweep_dev takes VCFB_F_HIST, VCFB_F_NOWRAP and VCFB_F_NO_OFFSET only"); return VCFB_E_ARG; }
  for (int i = 0; i < n_steps; ++i) {
    int rc = check_common(rgb, n_frames, H, W, B, q_steps[i], color, 0, nullptr, &g);
    if (rc) return rc;
    if (!(flags & VCFB_F_NOWRAP) && q_steps[i] == floor(q_steps[i]) && q_steps[i] >= 32768.0) {
      set_error("integral quantisation step does not fit int16: the reference's dequantiser raises OverflowError");
      return VCFB_E_ARG;
    }
  }
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  unsigned long long* st = reinterpret_cast<unsigned long long*>(stats);
  if (anyb_supported(B)) return launch_rd_sweep_anyb(rgb, g, n_frames, B, q_steps, n_steps, color, flags, st, s);
  return launch_rd_sweep(rgb, g, n_frames, B, q_steps, n_steps, color, flags, st, s);
}

int vcfb_color_encode_dev(const uint8_t* rgb, long long n_pixels, double q, int color, uint16_t* k_out,
                          void* cuda_stream) {
  if (!rgb || !k_out) { set_error("NULL pointer"); return VCFB_E_ARG; }
  if (n_pixels <= 0) { set_error("n_pixels must be > 0"); return VCFB_E_ARG; }
  if (!(q > 0.0) || !isfinite(q)) { set_error("quantisation step q must be finite and > 0"); return VCFB_E_ARG; }
  if (color != VCFB_COLOR_YCOCG && color != VCFB_COLOR_YCRCB) { set_error("unknown colour transform"); return VCFB_E_ARG; }
  return launch_color_encode(rgb, n_pixels, q, color, k_out, static_cast<cudaStream_t>(cuda_stream));
}

int vcfb_color_decode_dev(const uint16_t* k, long long n_pixels, double q, int color, uint8_t* rgb_out,
                          void* cuda_stream) {
  if (!k || !rgb_out) { set_error("NULL pointer"); return VCFB_E_ARG; }
  if (n_pixels <= 0) { set_error("n_pixels must be > 0"); return VCFB_E_ARG; }
  if (!(q > 0.0) || !isfinite(q)) { set_error("quantisation step q must be finite and > 0"); return VCFB_E_ARG; }
  if (color != VCFB_COLOR_YCOCG && color != VCFB_COLOR_YCRCB) { set_error("unknown colour transform"); return VCFB_E_ARG; }
  return launch_color_decode(k, n_pixels, q, color, rgb_out, static_cast<cudaStream_t>(cuda_stream));
}

int vcfb_gray_dev(const uint8_t* rgb, long long n_pixels, uint8_t* gray_out, void* cuda_stream) {
  if (!rgb || !gray_out) { set_error("NULL pointer"); return VCFB_E_ARG; }
  if (n_pixels <= 0) { set_error("n_pixels must be > 0"); return VCFB_E_ARG; }
  return launch_gray(rgb, n_pixels, gray_out, static_cast<cudaStream_t>(cuda_stream));
}

int vcfb_block_match_dev(const uint8_t* ref, const uint8_t* cur, int n_frames, int H, int W, int bs, int sr,
                         int16_t* mv_out, void* cuda_stream) {
  if (!ref || !cur || !mv_out) { set_error("NULL pointer"); return VCFB_E_ARG; }
  if (n_frames <= 0 || n_frames > 65535) { set_error("n_frames must be in [1, 65535]"); return VCFB_E_ARG; }
  if (bs < 4 || bs > 64 || (bs & 3)) { set_error("motion block size must be a multiple of 4 in [4, 64]"); return VCFB_E_ARG; }
  if (sr < 0 || sr > 31) { set_error("search range must be in [0, 31]"); return VCFB_E_ARG; }
  if (H < bs || W < bs || H / bs > 65535) { set_error("frame smaller than one motion block (or too tall)"); return VCFB_E_ARG; }
  return launch_block_match(ref, cur, n_frames, H, W, bs, sr, mv_out, static_cast<cudaStream_t>(cuda_stream));
}

int vcfb_block_match_tss_dev(const uint8_t* ref, const uint8_t* cur, int n_frames, int H, int W, int bs, int sr,
                             int16_t* mv_out, void* cuda_stream) {
  if (!ref || !cur || !mv_out) { set_error("NULL pointer"); return VCFB_E_ARG; }
  if (n_frames <= 0 || n_frames > 65535) { set_error("n_frames must be in [1, 65535]"); return VCFB_E_ARG; }
  if (bs < 1 || bs > 64) { set_error("motion block size must be in [1, 64]"); return VCFB_E_ARG; }
  if (sr < 0 || sr > 32767) { set_error("search range must be in [0, 32767]"); return VCFB_E_ARG; }
  if (H < bs || W < bs || H > 32767 || W > 32767) { set_error("frame smaller than one motion block (or larger than 32767)"); return VCFB_E_ARG; }
  return launch_block_match_tss(ref, cur, n_frames, H, W, bs, sr, mv_out, static_cast<cudaStream_t>(cuda_stream));
}

// ---- host-buffer layer ----------------------------------------------------------
//
// A batch is cut into chunks of whole frames; chunk i runs on slot i % NSLOT
// (one stream + device staging per slot), so the host->device copy of the next
// chunk, the kernel and the device->host copy of the previous chunk overlap on the
// two copy engines.  Pinned user buffers (cudaHostAlloc / cudaHostRegister /
// vcfb_host_alloc / torch pin_memory) are used in place; pageable ones are staged
// through the slot's own pinned buffer.

namespace {
constexpr int NSLOT = 3;
static size_t chunk_target() {
  // bytes of input per chunk; VCFB_CHUNK_MB is a development knob
  static size_t v = []() { const char* e = getenv("VCFB_CHUNK_MB"); return size_t(e ? atoi(e) : 32) << 20; }();
  return v;
}
#define CHUNK_TARGET chunk_target()

struct Slot {
  cudaStream_t s;
  char* dev; size_t dev_cap;
  char* pin; size_t pin_cap;
  // pageable output waiting for the stream: copy pin+off -> dst after sync
  struct Pending { void* dst; size_t off, bytes; } pend[2];
  int npend;
};
}  // namespace

struct vcfb_ctx {
  int device;
  Slot slot[NSLOT];
  char* aux;        // device: weights + stats
  size_t aux_cap;
  cudaEvent_t ready;
};

// Pageable user buffers: by default handed to cudaMemcpyAsync as they are (the driver stages
// them through its own pinned pool); VCFB_STAGE=1 stages explicitly through the slot buffer.
static bool stage_pageable() {
  static const bool v = getenv("VCFB_STAGE") != nullptr;
  return v;
}

static bool is_pinned(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

static int slot_reserve(Slot* sl, size_t dev_bytes, size_t pin_bytes) {
  cudaError_t e;
  if (dev_bytes > sl->dev_cap) {
    if (sl->dev) cudaFree(sl->dev);
    sl->dev = nullptr; sl->dev_cap = 0;
    e = cudaMalloc(reinterpret_cast<void**>(&sl->dev), dev_bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    sl->dev_cap = dev_bytes;
  }
  if (pin_bytes > sl->pin_cap) {
    if (sl->pin) cudaFreeHost(sl->pin);
    sl->pin = nullptr; sl->pin_cap = 0;
    e = cudaMallocHost(reinterpret_cast<void**>(&sl->pin), pin_bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMallocHost");
    sl->pin_cap = pin_bytes;
  }
  return VCFB_OK;
}

static int slot_drain(Slot* sl) {
  cudaError_t e = cudaStreamSynchronize(sl->s);
  if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
  for (int i = 0; i < sl->npend; ++i) memcpy(sl->pend[i].dst, sl->pin + sl->pend[i].off, sl->pend[i].bytes);
  sl->npend = 0;
  return VCFB_OK;
}


// Scope guard of the host entry points: remembers the calling thread's current device, makes the
// context's device current, and on the way out (a) on any path that did not finish cleanly waits
// for the slot streams and forgets the pending pinned->pageable copies -- their destinations are
// the caller's buffers of THIS call and must never be written by a later one -- and (b) restores
// the caller's device, so a torch process sitting on cuda:N is not silently moved.
struct HostScope {
  vcfb_ctx* c;
  int prev = -1;
  bool ok = false;
  cudaError_t err = cudaSuccess;
  explicit HostScope(vcfb_ctx* ctx) : c(ctx) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    err = cudaSetDevice(c->device);
  }
  ~HostScope() {
    if (!ok && err == cudaSuccess) {
      for (int i = 0; i < NSLOT; ++i) {
        if (c->slot[i].s) cudaStreamSynchronize(c->slot[i].s);
        c->slot[i].npend = 0;
      }
    }
    if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
  }
};

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

int vcfb_ctx_create(int device, vcfb_ctx** out) {
  if (!out) { set_error("out is NULL"); return VCFB_E_ARG; }
  *out = nullptr;
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  vcfb_ctx* c = new (std::nothrow) vcfb_ctx();
  if (!c) { set_error("out of memory"); return VCFB_E_ARG; }
  memset(c, 0, sizeof(*c));
  c->device = device;
  for (int i = 0; i < NSLOT; ++i) {
    e = cudaStreamCreateWithFlags(&c->slot[i].s, cudaStreamNonBlocking);
    if (e != cudaSuccess) { vcfb_ctx_destroy(c); return cuda_fail(e, "cudaStreamCreate"); }
  }
  e = cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
  if (e != cudaSuccess) { vcfb_ctx_destroy(c); return cuda_fail(e, "cudaEventCreate"); }
  *out = c;
  return VCFB_OK;
}

void vcfb_ctx_destroy(vcfb_ctx* c) {
  if (!c) return;
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
  cudaSetDevice(c->device);
  for (int i = 0; i < NSLOT; ++i) {
    if (c->slot[i].s) { cudaStreamSynchronize(c->slot[i].s); cudaStreamDestroy(c->slot[i].s); }
    if (c->slot[i].dev) cudaFree(c->slot[i].dev);
    if (c->slot[i].pin) cudaFreeHost(c->slot[i].pin);
  }
  if (c->aux) cudaFree(c->aux);
  if (c->ready) cudaEventDestroy(c->ready);
  if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
  delete c;
}

int vcfb_host_alloc(size_t bytes, void** out) {
  if (!out) { set_error("out is NULL"); return VCFB_E_ARG; }
  cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMallocHost");
  return VCFB_OK;
}

void vcfb_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// Upload weights, zero the statistics vector; every slot stream waits for both.
static int aux_prepare(vcfb_ctx* c, const double* weights, size_t w_b, bool want_stats,
                       double** d_w, int64_t** d_st) {
  const size_t st_off = align256(w_b);
  const size_t need = st_off + VCFB_STAT_LEN * sizeof(int64_t);
  cudaError_t e;
  if (need > c->aux_cap) {
    for (int i = 0; i < NSLOT; ++i) cudaStreamSynchronize(c->slot[i].s);
    if (c->aux) cudaFree(c->aux);
    c->aux = nullptr; c->aux_cap = 0;
    e = cudaMalloc(reinterpret_cast<void**>(&c->aux), need);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    c->aux_cap = need;
  }
  *d_w = w_b ? reinterpret_cast<double*>(c->aux) : nullptr;
  *d_st = want_stats ? reinterpret_cast<int64_t*>(c->aux + st_off) : nullptr;
  cudaStream_t s0 = c->slot[0].s;
  e = cudaSuccess;
  if (w_b) e = cudaMemcpyAsync(c->aux, weights, w_b, cudaMemcpyHostToDevice, s0);
  if (e == cudaSuccess && want_stats) e = cudaMemsetAsync(c->aux + st_off, 0, VCFB_STAT_LEN * sizeof(int64_t), s0);
  if (e == cudaSuccess) e = cudaEventRecord(c->ready, s0);
  for (int i = 1; i < NSLOT && e == cudaSuccess; ++i) e = cudaStreamWaitEvent(c->slot[i].s, c->ready, 0);
  if (e != cudaSuccess) return cuda_fail(e, "weights / statistics setup");
  return VCFB_OK;
}

static int finish(vcfb_ctx* c, int64_t* d_st, int64_t* stats) {
  for (int i = 0; i < NSLOT; ++i) {
    int rc = slot_drain(&c->slot[i]);
    if (rc) return rc;
  }
  if (d_st && stats) {
    cudaError_t e = cudaMemcpy(stats, d_st, VCFB_STAT_LEN * sizeof(int64_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail(e, "statistics device->host");
  }
  return VCFB_OK;
}

int vcfb_encode_host(vcfb_ctx* c, const uint8_t* rgb, int n_frames, int H, int W, int B, double q,
                     int color, unsigned flags, const double* weights, uint8_t* idx_out,
                     int64_t* stats) {
  if (!c) { set_error("ctx is NULL"); return VCFB_E_ARG; }
  Geom g;
  int rc = check_common(rgb, n_frames, H, W, B, q, color, flags, weights, &g);
  if (rc) return rc;
  if (!idx_out) { set_error("idx_out is NULL"); return VCFB_E_ARG; }
  HostScope scope(c);
  cudaError_t e = scope.err;
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  const size_t in_f = size_t(H) * W * 3, out_f = size_t(g.Hp) * g.Wp * 3;
  const size_t w_b = (flags & VCFB_F_PERCEPTUAL) ? size_t(2) * B * B * sizeof(double) : 0;
  double* d_w; int64_t* d_st;
  rc = aux_prepare(c, weights, w_b, stats != nullptr, &d_w, &d_st);
  if (rc) return rc;
  const bool pin_in = is_pinned(rgb) || !stage_pageable(), pin_out = is_pinned(idx_out) || !stage_pageable();
  int cf = int(CHUNK_TARGET / in_f);
  cf = cf < 1 ? 1 : cf;
  if (cf > (n_frames + NSLOT - 1) / NSLOT) cf = (n_frames + NSLOT - 1) / NSLOT;
  const size_t o_in = 0, o_out = align256(in_f * cf), slot_b = o_out + align256(out_f * cf);
  for (int f0 = 0, ci = 0; f0 < n_frames; f0 += cf, ++ci) {
    Slot* sl = &c->slot[ci % NSLOT];
    const int nf = (n_frames - f0 < cf) ? n_frames - f0 : cf;
    rc = slot_drain(sl);
    if (rc) return rc;
    rc = slot_reserve(sl, slot_b, (pin_in && pin_out) ? 0 : slot_b);
    if (rc) return rc;
    const uint8_t* src = rgb + size_t(f0) * in_f;
    if (!pin_in) { memcpy(sl->pin + o_in, src, in_f * nf); src = reinterpret_cast<uint8_t*>(sl->pin + o_in); }
    e = cudaMemcpyAsync(sl->dev + o_in, src, in_f * nf, cudaMemcpyHostToDevice, sl->s);
    if (e != cudaSuccess) return cuda_fail(e, "host->device copy");
    rc = vcfb_encode_dev(reinterpret_cast<uint8_t*>(sl->dev + o_in), nf, H, W, B, q, color, flags, d_w,
                         reinterpret_cast<uint8_t*>(sl->dev + o_out), d_st, sl->s);
    if (rc) return rc;
    uint8_t* dst = idx_out + size_t(f0) * out_f;
    if (pin_out) {
      e = cudaMemcpyAsync(dst, sl->dev + o_out, out_f * nf, cudaMemcpyDeviceToHost, sl->s);
    } else {
      e = cudaMemcpyAsync(sl->pin + o_out, sl->dev + o_out, out_f * nf, cudaMemcpyDeviceToHost, sl->s);
      sl->pend[sl->npend++] = {dst, o_out, out_f * nf};
    }
    if (e != cudaSuccess) return cuda_fail(e, "device->host copy");
  }
  rc = finish(c, d_st, stats);
  scope.ok = (rc == VCFB_OK);
  return rc;
}

int vcfb_decode_host(vcfb_ctx* c, const uint8_t* idx, int n_frames, int H, int W, int B, double q,
                     int color, unsigned flags, const double* weights, uint8_t* rgb_out, void* y_out,
                     const uint8_t* original, int64_t* stats) {
  if (!c) { set_error("ctx is NULL"); return VCFB_E_ARG; }
  Geom g;
  int rc = check_common(idx, n_frames, H, W, B, q, color, flags, weights, &g);
  if (rc) return rc;
  if (!rgb_out && !y_out && !(original && stats)) { set_error("decode has no output"); return VCFB_E_ARG; }
  if (flags & VCFB_F_NO_OFFSET) { set_error("VCFB_F_NO_OFFSET is an encode / rd_sweep flag (the reference's decoder always runs with offset 128)"); return VCFB_E_ARG; }
  HostScope scope(c);
  cudaError_t e = scope.err;
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  const size_t px_f = size_t(H) * W * 3, idx_f = size_t(g.Hp) * g.Wp * 3;
  const size_t y_f = y_out ? px_f * ((flags & VCFB_F_FP64) ? 8 : 4) : 0;
  const size_t or_f = original ? px_f : 0;
  const size_t w_b = (flags & VCFB_F_PERCEPTUAL) ? size_t(2) * B * B * sizeof(double) : 0;
  double* d_w; int64_t* d_st;
  rc = aux_prepare(c, weights, w_b, stats != nullptr, &d_w, &d_st);
  if (rc) return rc;
  const bool all_pinned = !stage_pageable() ||
                          (is_pinned(idx) && (!rgb_out || is_pinned(rgb_out)) && (!y_out || is_pinned(y_out)) &&
                           (!original || is_pinned(original)));
  int cf = int(CHUNK_TARGET / idx_f);
  cf = cf < 1 ? 1 : cf;
  if (cf > (n_frames + NSLOT - 1) / NSLOT) cf = (n_frames + NSLOT - 1) / NSLOT;
  const size_t o_idx = 0, o_rgb = align256(idx_f * cf), o_y = o_rgb + align256(px_f * cf),
               o_or = o_y + align256(y_f * cf), slot_b = o_or + align256(or_f * cf);
  for (int f0 = 0, ci = 0; f0 < n_frames; f0 += cf, ++ci) {
    Slot* sl = &c->slot[ci % NSLOT];
    const int nf = (n_frames - f0 < cf) ? n_frames - f0 : cf;
    rc = slot_drain(sl);
    if (rc) return rc;
    rc = slot_reserve(sl, slot_b, all_pinned ? 0 : slot_b);
    if (rc) return rc;
    const uint8_t* src = idx + size_t(f0) * idx_f;
    const uint8_t* osrc = original ? original + size_t(f0) * px_f : nullptr;
    if (!all_pinned) {
      memcpy(sl->pin + o_idx, src, idx_f * nf);
      src = reinterpret_cast<uint8_t*>(sl->pin + o_idx);
      if (osrc) { memcpy(sl->pin + o_or, osrc, px_f * nf); osrc = reinterpret_cast<uint8_t*>(sl->pin + o_or); }
    }
    e = cudaMemcpyAsync(sl->dev + o_idx, src, idx_f * nf, cudaMemcpyHostToDevice, sl->s);
    if (e == cudaSuccess && osrc) e = cudaMemcpyAsync(sl->dev + o_or, osrc, px_f * nf, cudaMemcpyHostToDevice, sl->s);
    if (e != cudaSuccess) return cuda_fail(e, "host->device copy");
    rc = vcfb_decode_dev(reinterpret_cast<uint8_t*>(sl->dev + o_idx), nf, H, W, B, q, color, flags, d_w,
                         rgb_out ? reinterpret_cast<uint8_t*>(sl->dev + o_rgb) : nullptr,
                         y_out ? static_cast<void*>(sl->dev + o_y) : nullptr,
                         original ? reinterpret_cast<uint8_t*>(sl->dev + o_or) : nullptr, d_st, sl->s);
    if (rc) return rc;
    if (rgb_out) {
      uint8_t* dst = rgb_out + size_t(f0) * px_f;
      if (all_pinned) {
        e = cudaMemcpyAsync(dst, sl->dev + o_rgb, px_f * nf, cudaMemcpyDeviceToHost, sl->s);
      } else {
        e = cudaMemcpyAsync(sl->pin + o_rgb, sl->dev + o_rgb, px_f * nf, cudaMemcpyDeviceToHost, sl->s);
        sl->pend[sl->npend++] = {dst, o_rgb, px_f * nf};
      }
      if (e != cudaSuccess) return cuda_fail(e, "device->host copy");
    }
    if (y_out) {
      char* dst = static_cast<char*>(y_out) + size_t(f0) * y_f;
      if (all_pinned) {
        e = cudaMemcpyAsync(dst, sl->dev + o_y, y_f * nf, cudaMemcpyDeviceToHost, sl->s);
      } else {
        e = cudaMemcpyAsync(sl->pin + o_y, sl->dev + o_y, y_f * nf, cudaMemcpyDeviceToHost, sl->s);
        sl->pend[sl->npend++] = {dst, o_y, y_f * nf};
      }
      if (e != cudaSuccess) return cuda_fail(e, "device->host copy");
    }
  }
  rc = finish(c, d_st, stats);
  scope.ok = (rc == VCFB_OK);
  return rc;
}

int vcfb_color_encode_host(vcfb_ctx* c, const uint8_t* rgb, long long n_pixels, double q, int color,
                           uint16_t* k_out) {
  if (!c) { set_error("ctx is NULL"); return VCFB_E_ARG; }
  if (!rgb || !k_out || n_pixels <= 0) { set_error("bad argument"); return VCFB_E_ARG; }
  HostScope scope(c);
  cudaError_t e = scope.err;
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  Slot* sl = &c->slot[0];
  int rc = slot_drain(sl);
  if (rc) return rc;
  const size_t in_b = size_t(n_pixels) * 3, out_b = size_t(n_pixels) * 6, o_out = align256(in_b);
  rc = slot_reserve(sl, o_out + out_b, 0);
  if (rc) return rc;
  e = cudaMemcpyAsync(sl->dev, rgb, in_b, cudaMemcpyHostToDevice, sl->s);
  if (e != cudaSuccess) return cuda_fail(e, "host->device copy");
  rc = vcfb_color_encode_dev(reinterpret_cast<uint8_t*>(sl->dev), n_pixels, q, color,
                             reinterpret_cast<uint16_t*>(sl->dev + o_out), sl->s);
  if (rc) return rc;
  e = cudaMemcpyAsync(k_out, sl->dev + o_out, out_b, cudaMemcpyDeviceToHost, sl->s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(sl->s);
  if (e != cudaSuccess) return cuda_fail(e, "colour encode (device->host / synchronize)");
  scope.ok = true;
  return VCFB_OK;
}

int vcfb_color_decode_host(vcfb_ctx* c, const uint16_t* k, long long n_pixels, double q, int color,
                           uint8_t* rgb_out) {
  if (!c) { set_error("ctx is NULL"); return VCFB_E_ARG; }
  if (!k || !rgb_out || n_pixels <= 0) { set_error("bad argument"); return VCFB_E_ARG; }
  HostScope scope(c);
  cudaError_t e = scope.err;
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  Slot* sl = &c->slot[0];
  int rc = slot_drain(sl);
  if (rc) return rc;
  const size_t in_b = size_t(n_pixels) * 6, out_b = size_t(n_pixels) * 3, o_out = align256(in_b);
  rc = slot_reserve(sl, o_out + out_b, 0);
  if (rc) return rc;
  e = cudaMemcpyAsync(sl->dev, k, in_b, cudaMemcpyHostToDevice, sl->s);
  if (e != cudaSuccess) return cuda_fail(e, "host->device copy");
  rc = vcfb_color_decode_dev(reinterpret_cast<uint16_t*>(sl->dev), n_pixels, q, color,
                             reinterpret_cast<uint8_t*>(sl->dev + o_out), sl->s);
  if (rc) return rc;
  e = cudaMemcpyAsync(rgb_out, sl->dev + o_out, out_b, cudaMemcpyDeviceToHost, sl->s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(sl->s);
  if (e != cudaSuccess) return cuda_fail(e, "colour decode (device->host / synchronize)");
  scope.ok = true;
  return VCFB_OK;
}

}  // extern "C"
