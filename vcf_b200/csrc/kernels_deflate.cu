// Entropy front-end: raw deflate (RFC 1951) of a byte array on the GPU.
//
// Replaces the zlib call underneath the reference's entropy stage for the uint8 index planes of
// the transform path (np.savez_compressed, src/z_lib.py:19-23; the zlib codec of tifffile,
// src/TIFF.py:23-31).  The output is an ordinary deflate stream: zlib.decompress / np.load /
// tifffile read it.  It is not byte-identical with zlib's output (no two deflate encoders are);
// the drop-in property is "the reference's decoder returns the same bytes".
//
//   deflate_segments_kernel  one CTA per segment of NT pieces.  Phase 0: the CTA samples the segment
//       (every fourth word) for the cost model of the parse.  Phase 1: every thread parses its
//       piece (greedy over the candidate distances, deflate_core.cuh), keeps the tokens (16 bits
//       each, token j of all threads side by side) and counts them into shared histograms.  Phase 2:
//       Huffman codes -- literal/length in CTA-parallel steps (rank sort, serial two-queue merge,
//       lengths, canonical codes), distances by one thread beside it, block header by thread 0.
//       Phase 3: every thread sizes its tokens under the codes; exclusive scan -> bit offsets; the
//       segment falls back to stored blocks when those are smaller.  Phase 4: the CTA zeroes exactly
//       the bytes the block takes, every thread writes its tokens at its bit offset (first and last
//       word with atomicOr, the rest with plain stores).
//   deflate_scan_kernel      exclusive scan of the segment sizes, total size, closing block.
//   deflate_gather_kernel    copies every segment's bytes to its place in the output stream.
#include "common.cuh"
#include "deflate_core.cuh"
#include "crc32_core.cuh"

#include <stdlib.h>

namespace vcfb {
namespace fast { int sm_count(); }      // kernels_fast.cu
namespace {

constexpr int NT = 512;

struct Plan {
  int piece;              // bytes per thread
  long long seg_bytes;    // bytes per segment = NT * piece
  long long nseg;
  long long stride;       // bytes of scratch per segment
  long long sizes_off, offs_off, regions_off, tokens_off, total;   // workspace layout
  long long bound;        // largest possible stream
};

Plan make_plan(unsigned long long n) {
  Plan p;
  // a multiple of the longest match (258): a piece that lies inside a long run becomes matches of
  // length 258 (symbol 285, no extra bits), as in a sequential parse.  258 itself: the time of a
  // segment is the serial walk of a thread over its piece, and longer pieces were slower at every
  // input size on the B200 (profiles/r1h_deflate_piece_sweep.txt) while saving < 1 % of the stream.
  p.piece = dfl::MAX_MATCH;
  if (const char* env = getenv("VCFB_DEFLATE_PIECE")) {      // measurement knob: multiples of 258 only
    const int k = atoi(env) / dfl::MAX_MATCH;
    if (k >= 1 && k <= 16) p.piece = k * dfl::MAX_MATCH;
  }
  p.seg_bytes = (long long)NT * p.piece;
  p.nseg = (long long)((n + p.seg_bytes - 1) / p.seg_bytes);
  p.stride = (dfl::stored_size(p.seg_bytes) + 15) / 16 * 16 + 16;
  p.sizes_off = 0;
  p.offs_off = (p.nseg * 4 + 15) / 16 * 16;
  p.regions_off = p.offs_off + (p.nseg * 8 + 15) / 16 * 16;
  p.tokens_off = p.regions_off + p.nseg * p.stride;            // 16 bits per input byte at most
  p.total = p.tokens_off + p.nseg * p.seg_bytes * 2 + 16;
  p.bound = (long long)n + 5 * p.nseg * ((p.seg_bytes + dfl::STORED_MAX - 1) / dfl::STORED_MAX) + 2;
  return p;
}

struct SegShared {
  uint32_t hist[288];
  uint32_t dhist[32];
  uint32_t scnt[256];     // sampled literals
  uint32_t stot[4];       // their number, the number of sampled run starts, of words that are not a run, of those found above
  dfl::MatchParams mp;
  dfl::CostModel cm;
  dfl::Codes codes;
  dfl::Header hdr;
  dfl::BuildScratch scratch;
  dfl::BuildScratch dscratch;
  uint32_t off[NT];
  uint32_t wsum[NT / 32];
  long long total_bytes;
  int stored;
};

struct SampleAdd {
  uint32_t* cnt;
  uint32_t* tot;
  __device__ void lit(int b) { atomicAdd(cnt + b, 1u); atomicAdd(tot, 1u); }
  __device__ void run() { atomicAdd(tot + 1, 1u); }
  __device__ void dense() { atomicAdd(tot + 2, 1u); }
  __device__ void hit() { atomicAdd(tot + 3, 1u); }
};

// the kept tokens of a thread, four loads in flight
template <class V>
__device__ __forceinline__ void walk_tokens(const uint16_t* __restrict__ tok, int ntok, V& v) {
  int j = 0;
  for (; j + 4 <= ntok; j += 4) {
    const uint16_t t0 = tok[(long long)j * NT], t1 = tok[(long long)(j + 1) * NT], t2 = tok[(long long)(j + 2) * NT],
                   t3 = tok[(long long)(j + 3) * NT];
    dfl::visit_token(t0, v);
    dfl::visit_token(t1, v);
    dfl::visit_token(t2, v);
    dfl::visit_token(t3, v);
  }
  for (; j < ntok; ++j) dfl::visit_token(tok[(long long)j * NT], v);
}

struct TokenStore {        // token j of this thread; the tokens of a segment are laid out j-major
  uint16_t* base;
  __device__ void operator()(int j, uint16_t t) const { base[(long long)j * NT] = t; }
};

// MINB = 3 (40 registers, some spills): the most segments in flight, for batches; MINB = 2 (64 registers, no
// spills): faster per segment, for inputs that do not fill the GPU anyway (one 4K frame is 189 segments).
template <int MINB>
__global__ void __launch_bounds__(NT, MINB)
deflate_segments_kernel(const uint8_t* __restrict__ src, long long n, int piece, dfl::MatchParams mp,
                        uint8_t* __restrict__ regions, long long stride, uint16_t* __restrict__ tokens,
                        uint32_t* __restrict__ seg_size) {
  __shared__ SegShared sh;
  const int tid = threadIdx.x;
  const long long seg = blockIdx.x;
  const long long seg_bytes = (long long)NT * piece;
  const long long s0 = seg * seg_bytes;
  const long long nseg = min(n - s0, seg_bytes);
  const long long s = min(n, s0 + (long long)tid * piece);
  const long long e = min(n, s + piece);
  uint8_t* out = regions + seg * stride;

  uint16_t* tok = tokens + seg * seg_bytes + tid;

  if (tid == 64) sh.mp = mp;
  __syncthreads();
  for (int i = tid; i < 288; i += NT) sh.hist[i] = 0;
  if (tid < 256) sh.scnt[tid] = 0;
  if (tid < 32) sh.dhist[tid] = 0;
  if (tid <= dfl::MAX_LIT_BITS) sh.scratch.cnt[tid] = 0;
  if (tid == 32) { sh.scratch.m = 0; sh.scratch.hi = 0; sh.stot[0] = 0; sh.stot[1] = 0; sh.stot[2] = 0; sh.stot[3] = 0; }
  if (tid == 0) sh.hist[dfl::EOB] = 1;

  // phase 0: cost model of the parse from a sample of the segment
  {
    SampleAdd add{sh.scnt, sh.stot};
    dfl::sample_segment(src, n, s0, nseg, tid, NT, sh.mp, add);
  }
  __syncthreads();
  if (tid < 256) sh.cm.lit8[tid] = dfl::model_lit8(sh.scnt, sh.stot[0], sh.stot[1], tid);
  if (tid == 256) { sh.cm.len8 = dfl::model_len8(sh.stot[0], sh.stot[1]); sh.cm.far_on = dfl::model_far_on(sh.stot[2], sh.stot[3]); }
  __syncthreads();

  // phase 1: parse; tokens and their frequencies
  int ntok;
  {
    dfl::TokenVisitor<TokenStore> tv;
    tv.st.base = tok;
    tv.n = 0;
    tv.cv.init(sh.hist, sh.dhist, &sh.mp);
    dfl::parse_piece(src, n, s, e, sh.mp, &sh.cm, tv);
    tv.cv.flush();
    ntok = tv.n;
  }
  __syncthreads();
  // phase 2: code construction
  //  (the steps of deflate_core.cuh: only the two-queue merge and the Kraft fix are serial; the distance code
  //  -- a few symbols -- is built by warps 1 and 2 beside the first three steps)
  dfl::dpar_prepare(sh.dhist, sh.dscratch, tid - 32);
  dfl::par_rank(sh.hist, dfl::NLIT, sh.scratch, sh.codes.code, sh.codes.len, tid, NT);
  __syncthreads();
  if (tid >= 64 && tid < 96) dfl::dpar_rank(sh.dhist, sh.dscratch, tid - 64, 32);
  dfl::par_tree(sh.scratch, tid);
  __syncthreads();
  dfl::dpar_build(sh.dhist, sh.dscratch, sh.codes, tid - 32);
  dfl::par_count(sh.scratch, dfl::MAX_LIT_BITS, tid, NT);
  __syncthreads();
  dfl::par_limit(sh.scratch, dfl::MAX_LIT_BITS, tid);
  __syncthreads();
  dfl::par_lengths(sh.scratch, dfl::MAX_LIT_BITS, sh.codes.len, tid, NT);
  __syncthreads();
  dfl::par_codes(sh.scratch, dfl::NLIT, sh.codes.len, sh.codes.code, tid, NT);
  __syncthreads();
  // the block header in CTA-parallel steps (run-length tokens of the two length arrays, ranks of the
  // code-length symbols); the construction of the code-length code needs one thread and nothing below needs it
  // before the sizes are summed: the other warps size their pieces meanwhile
  dfl::hpar_prepare(sh.codes, sh.hdr, tid, NT);
  __syncthreads();
  dfl::hpar_fill(sh.codes, sh.hdr, tid, NT);
  __syncthreads();
  dfl::hpar_count(sh.hdr, tid, NT);
  __syncthreads();
  dfl::hpar_blocks(sh.hdr, tid, NT);
  __syncthreads();
  dfl::hpar_tokens(sh.hdr, tid, NT);
  __syncthreads();
  dfl::hpar_clprepare(sh.hdr, tid);
  __syncthreads();
  dfl::hpar_clrank(sh.scratch, sh.hdr, tid, NT);
  __syncthreads();
  dfl::hpar_clcode(sh.scratch, sh.hdr, tid);

  // phase 3: sizes and bit offsets
  {
    dfl::SizeVisitor sv;
    sv.c = &sh.codes;
    sv.P = &sh.mp;
    sv.bits = 0;
    walk_tokens(tok, ntok, sv);
    sh.off[tid] = sv.bits;
  }
  __syncthreads();
  dfl::hpar_size(sh.hdr, tid, NT);
  __syncthreads();
  {
    // exclusive scan of the pieces' bit counts (a segment stays below 2^32 bits): warp shuffles,
    // then the 16 warp totals
    const int lane = tid & 31, wid = tid >> 5;
    const uint32_t mine = sh.off[tid];
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) sh.wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const uint32_t v = lane < NT / 32 ? sh.wsum[lane] : 0u;
      uint32_t wi = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += t;
      }
      if (lane < NT / 32) sh.wsum[lane] = wi - v;
      if (lane == 31) {
        dfl::hpar_finish(sh.hdr, 0);
        const long long dyn = dfl::dynamic_size(sh.hdr, sh.codes, (long long)wi);
        const long long st = dfl::stored_size(nseg);
        sh.stored = dyn >= st;
        sh.total_bytes = sh.stored ? st : dyn;
        seg_size[seg] = uint32_t(sh.total_bytes);
      }
    }
    __syncthreads();
    sh.off[tid] = uint32_t(sh.hdr.bits) + sh.wsum[wid] + (incl - mine);
  }
  __syncthreads();

  // phase 4: output
  if (sh.stored) {
    dfl::stored_copy(src + s0, nseg, s - s0, e - s0, out);
    return;
  }
  {
    const int n16 = int((sh.total_bytes + 15) / 16);
    uint4* o4 = reinterpret_cast<uint4*>(out);
    for (int i = tid; i < n16; i += NT) o4[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  {
    dfl::BitWriter bw;
    bw.init(reinterpret_cast<uint32_t*>(out), tid == 0 ? 0 : (long long)sh.off[tid]);
    if (tid == 0) dfl::header_emit(sh.hdr, bw);
    dfl::EmitVisitor ev;
    ev.c = &sh.codes;
    ev.P = &sh.mp;
    ev.bw = &bw;
    walk_tokens(tok, ntok, ev);
    if (tid == NT - 1) dfl::segment_close(sh.codes, bw);
    bw.finish();
  }
}

// one CTA: exclusive scan of the segment sizes; the stream is closed with an empty fixed-Huffman
// block with BFINAL = 1 (bits 1, 01, 0000000 -> bytes 03 00)
__global__ void __launch_bounds__(NT)
deflate_scan_kernel(const uint32_t* __restrict__ seg_size, long long nseg, unsigned long long* __restrict__ seg_off,
                    uint8_t* __restrict__ dst, unsigned long long* __restrict__ out_bytes) {
  __shared__ unsigned long long part[NT];
  const int tid = threadIdx.x;
  const long long per = (nseg + NT - 1) / NT;
  const long long a = min(nseg, tid * per), b = min(nseg, a + per);
  unsigned long long sum = 0;
  for (long long i = a; i < b; ++i) sum += seg_size[i];
  part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    unsigned long long acc = 0;
    for (int t = 0; t < NT; ++t) {
      const unsigned long long v = part[t];
      part[t] = acc;
      acc += v;
    }
    dst[acc] = 0x03;
    dst[acc + 1] = 0x00;
    *out_bytes = acc + 2;
  }
  __syncthreads();
  unsigned long long acc = part[tid];
  for (long long i = a; i < b; ++i) {
    seg_off[i] = acc;
    acc += seg_size[i];
  }
}

__global__ void __launch_bounds__(NT)
deflate_gather_kernel(const uint8_t* __restrict__ regions, long long stride, const uint32_t* __restrict__ seg_size,
                      const unsigned long long* __restrict__ seg_off, uint8_t* __restrict__ dst) {
  const long long seg = blockIdx.x;
  const uint8_t* in = regions + seg * stride;
  uint8_t* out = dst + seg_off[seg];
  const unsigned nb = seg_size[seg];
  // bytes up to the first 16-byte boundary of the destination, then 16-byte stores fed by byte-aligned
  // 4-byte reads would need a shifter; the compressed stream is small, so plain words where both sides
  // are 4-byte aligned and bytes otherwise
  if (((reinterpret_cast<uintptr_t>(out) & 3) == 0)) {
    const unsigned nw = nb / 4;
    const uint32_t* in4 = reinterpret_cast<const uint32_t*>(in);
    uint32_t* out4 = reinterpret_cast<uint32_t*>(out);
    for (unsigned i = threadIdx.x; i < nw; i += NT) out4[i] = in4[i];
    for (unsigned i = nw * 4 + threadIdx.x; i < nb; i += NT) out[i] = in[i];
  } else {
    for (unsigned i = threadIdx.x; i < nb; i += NT) out[i] = in[i];
  }
}

// CRC-32 of a byte array (crc32_core.cuh): one 512-byte chunk per thread and grid-stride step, the
// chunk's finished CRC shifted to the end of the input, XOR over the grid.  *out is zero on entry.
constexpr int CRC_NT = 256;
constexpr int CRC_CHUNK = 512;

__global__ void __launch_bounds__(CRC_NT)
crc32_kernel(const uint8_t* __restrict__ src, long long n, crc::Powers P, uint32_t* __restrict__ out) {
  __shared__ uint32_t tab[256];
  __shared__ crc::Powers Ps;
  tab[threadIdx.x] = crc::table_entry(threadIdx.x);
  if (threadIdx.x < 32) Ps.x2n[threadIdx.x] = P.x2n[threadIdx.x];
  __syncthreads();
  const long long nchunks = (n + CRC_CHUNK - 1) / CRC_CHUNK;
  uint32_t acc = 0;
  for (long long c = (long long)blockIdx.x * CRC_NT + threadIdx.x; c < nchunks; c += (long long)gridDim.x * CRC_NT) {
    const long long s = c * CRC_CHUNK;
    const long long e = min(n, s + CRC_CHUNK);
    acc ^= crc::chunk_term(src, n, s, e, tab, Ps);
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, d);
  if ((threadIdx.x & 31) == 0 && acc) atomicXor(out, acc);
}

// Adler-32 (RFC 1950, the checksum that closes a zlib stream -- tifffile's zlib codec,
// src/TIFF.py:23-31): A = 1 + sum b_i, B = n + sum (n - i) b_i, both mod 65521.  Eight bytes per
// thread and grid-stride step; sums[0], sums[1] are zero on entry.
constexpr unsigned ADLER_MOD = 65521u;

__global__ void __launch_bounds__(256)
adler32_sums_kernel(const uint8_t* __restrict__ src, long long n, unsigned long long* __restrict__ sums) {
  const long long nw = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long s1 = 0, s2 = 0;
  unsigned wgt = unsigned((unsigned long long)(n - 8 * i) % ADLER_MOD);     // (n - position of the word) mod 65521
  const unsigned dec = unsigned((unsigned long long)(8 * stride) % ADLER_MOD);
  for (; i < nw; i += stride) {
    uint64_t w = *reinterpret_cast<const uint64_t*>(src + 8 * i);
    unsigned sb = 0, sk = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const unsigned b = unsigned(w) & 0xffu;
      sb += b;
      sk += unsigned(k) * b;
      w >>= 8;
    }
    s1 += sb;
    s2 += (unsigned long long)wgt * sb + (unsigned long long)ADLER_MOD * 8u - sk;   // sk <= 7140 < 8 * 65521
    wgt = wgt >= dec ? wgt - dec : wgt + ADLER_MOD - dec;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long p = nw << 3; p < n; ++p) { s1 += src[p]; s2 += (unsigned long long)(n - p) * src[p] % ADLER_MOD; }
  s1 %= ADLER_MOD;
  s2 %= ADLER_MOD;
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(sums, s1);
    atomicAdd(sums + 1, s2);
  }
}

__global__ void adler32_finish_kernel(const unsigned long long* __restrict__ sums, long long n, uint32_t* __restrict__ out) {
  const unsigned a = unsigned((1ull + sums[0]) % ADLER_MOD);
  const unsigned b = unsigned(((unsigned long long)n % ADLER_MOD + sums[1]) % ADLER_MOD);
  *out = (b << 16) | a;
}

}  // namespace
}  // namespace vcfb

using namespace vcfb;

extern "C" {

size_t vcfb_deflate_bound(size_t n_bytes) { return size_t(make_plan(n_bytes).bound); }

size_t vcfb_deflate_workspace(size_t n_bytes) { return size_t(make_plan(n_bytes).total); }

int vcfb_deflate_dev(const uint8_t* src, size_t n_bytes, uint8_t* dst, size_t dst_capacity,
                     uint64_t* out_bytes, void* workspace, size_t workspace_bytes, void* cuda_stream) {
  return vcfb_deflate_rows_dev(src, n_bytes, 0, 1, dst, dst_capacity, out_bytes, workspace, workspace_bytes, cuda_stream);
}

int vcfb_deflate_rows_dev(const uint8_t* src, size_t n_bytes, size_t row_bytes, int sample_bytes, uint8_t* dst,
                          size_t dst_capacity, uint64_t* out_bytes, void* workspace, size_t workspace_bytes,
                          void* cuda_stream) {
  if (!dst || !out_bytes) { set_error("output pointer is NULL"); return VCFB_E_ARG; }
  if (n_bytes && !src) { set_error("input pointer is NULL"); return VCFB_E_ARG; }
  if (n_bytes >= (1ull << 40)) { set_error("input too large"); return VCFB_E_ARG; }
  if (reinterpret_cast<uintptr_t>(src) & 7) { set_error("vcfb_deflate_dev: src must be 8-byte aligned"); return VCFB_E_ARG; }
  const Plan p = make_plan(n_bytes);
  if (dst_capacity < size_t(p.bound)) { set_error("vcfb_deflate_dev: dst_capacity is below vcfb_deflate_bound()"); return VCFB_E_ARG; }
  if (workspace_bytes < size_t(p.total) || (!workspace && p.total)) { set_error("vcfb_deflate_dev: workspace is below vcfb_deflate_workspace()"); return VCFB_E_ARG; }
  if (reinterpret_cast<uintptr_t>(workspace) & 15) { set_error("vcfb_deflate_dev: workspace must be 16-byte aligned"); return VCFB_E_ARG; }
  if (p.nseg > 0x7fffffffLL) { set_error("input too large"); return VCFB_E_ARG; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  uint32_t* seg_size = reinterpret_cast<uint32_t*>(ws + p.sizes_off);
  unsigned long long* seg_off = reinterpret_cast<unsigned long long*>(ws + p.offs_off);
  uint8_t* regions = ws + p.regions_off;
  uint16_t* tokens = reinterpret_cast<uint16_t*>(ws + p.tokens_off);
  if (sample_bytes < 1 || sample_bytes > 16) { set_error("vcfb_deflate_rows_dev: sample_bytes must be in [1, 16]"); return VCFB_E_ARG; }
  // rows too long for deflate's 32 KB window: runs and the previous sample only
  const dfl::MatchParams mp = dfl::make_match_params((long long)row_bytes, sample_bytes);
  cudaError_t e;
  if (p.nseg) {
    note_kernel("deflate_segments");
    if (p.nseg <= 2ll * fast::sm_count())
      deflate_segments_kernel<2><<<unsigned(p.nseg), NT, 0, s>>>(src, (long long)n_bytes, p.piece, mp, regions, p.stride,
                                                                 tokens, seg_size);
    else
      deflate_segments_kernel<3><<<unsigned(p.nseg), NT, 0, s>>>(src, (long long)n_bytes, p.piece, mp, regions, p.stride,
                                                                 tokens, seg_size);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "deflate_segments_kernel launch");
  }
  note_kernel("deflate_scan");
  deflate_scan_kernel<<<1, NT, 0, s>>>(seg_size, p.nseg, seg_off, dst, reinterpret_cast<unsigned long long*>(out_bytes));
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "deflate_scan_kernel launch");
  if (p.nseg) {
    note_kernel("deflate_gather");
    deflate_gather_kernel<<<unsigned(p.nseg), NT, 0, s>>>(regions, p.stride, seg_size, seg_off, dst);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "deflate_gather_kernel launch");
  }
  return VCFB_OK;
}

int vcfb_crc32_dev(const uint8_t* src, size_t n_bytes, uint32_t* out_crc, void* cuda_stream) {
  if (!out_crc) { set_error("output pointer is NULL"); return VCFB_E_ARG; }
  if (n_bytes && !src) { set_error("input pointer is NULL"); return VCFB_E_ARG; }
  if (n_bytes >= (1ull << 40)) { set_error("input too large"); return VCFB_E_ARG; }
  if (reinterpret_cast<uintptr_t>(src) & 7) { set_error("vcfb_crc32_dev: src must be 8-byte aligned"); return VCFB_E_ARG; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
  cudaError_t e = cudaMemsetAsync(out_crc, 0, sizeof(uint32_t), s);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
  if (!n_bytes) return VCFB_OK;                       // crc32 of nothing is 0
  crc::Powers P;
  crc::make_powers(P);
  const long long nchunks = (long long)((n_bytes + CRC_CHUNK - 1) / CRC_CHUNK);
  const long long want = (nchunks + CRC_NT - 1) / CRC_NT;
  const unsigned grid = unsigned(want < 148 * 8 ? want : 148 * 8);
  note_kernel("crc32");
  crc32_kernel<<<grid, CRC_NT, 0, s>>>(src, (long long)n_bytes, P, out_crc);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "crc32_kernel launch");
  return VCFB_OK;
}

int vcfb_adler32_dev(const uint8_t* src, size_t n_bytes, uint32_t* out_adler, void* workspace16, void* cuda_stream) {
  if (!out_adler || !workspace16) { set_error("output or workspace pointer is NULL"); return VCFB_E_ARG; }
  if (n_bytes && !src) { set_error("input pointer is NULL"); return VCFB_E_ARG; }
  if (n_bytes >= (1ull << 40)) { set_error("input too large"); return VCFB_E_ARG; }
  if ((reinterpret_cast<uintptr_t>(src) & 7) || (reinterpret_cast<uintptr_t>(workspace16) & 7)) {
    set_error("vcfb_adler32_dev: src and workspace must be 8-byte aligned"); return VCFB_E_ARG;
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
  unsigned long long* sums = static_cast<unsigned long long*>(workspace16);
  cudaError_t e = cudaMemsetAsync(sums, 0, 16, s);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync");
  const long long nw = (long long)(n_bytes >> 3);
  const long long want = (nw + 256 * 8 - 1) / (256 * 8);           // about 8 words per thread
  const unsigned grid = unsigned(want < 1 ? 1 : want < 148 * 8 ? want : 148 * 8);
  note_kernel("adler32");
  adler32_sums_kernel<<<grid, 256, 0, s>>>(src, (long long)n_bytes, sums);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "adler32_sums_kernel launch");
  adler32_finish_kernel<<<1, 1, 0, s>>>(sums, (long long)n_bytes, out_adler);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "adler32_finish_kernel launch");
  return VCFB_OK;
}

}  // extern "C"
