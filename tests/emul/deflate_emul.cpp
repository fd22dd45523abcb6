// Host emulation of deflate_segments_kernel / deflate_scan_kernel / deflate_gather_kernel
// (vcf_b200/csrc/kernels_deflate.cu): the same __host__ __device__ code from
// vcf_b200/csrc/deflate_core.cuh, compiled with g++, the CTA's phases run one "thread" after the
// other.  TEST INFRASTRUCTURE ONLY: built and loaded by tests/test_deflate_core.py so that the
// parsing / Huffman / header / bit-packing logic is checked against zlib's decoder on machines
// without a GPU.  Nothing in vcf_b200/ uses it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../vcf_b200/csrc/deflate_core.cuh"

using namespace vcfb;

namespace {
struct Store {                 // token j of a thread, laid out like the kernel's scratch (j-major)
  uint16_t* base;
  int NT;
  void operator()(int j, uint16_t t) const { base[(long long)j * NT] = t; }
};

struct SampleAdd {
  uint32_t* cnt;
  uint32_t* tot;               // [0] literals, [1] run starts, [2] words that are not a run, [3] of those found above
  void lit(int b) { ++cnt[b]; ++tot[0]; }
  void run() { ++tot[1]; }
  void dense() { ++tot[2]; }
  void hit() { ++tot[3]; }
};
}  // namespace

// row, pixel: make_match_params(); dists: overrides the candidate distances of the parse (nd of them, the
// first one 1) when nd > 0.  model = 0: plain run-length parse (no cost model; runs only).
extern "C" long long dfl_emul(const uint8_t* src, long long n, int piece, int NT, long long row, int pixel, const int* dists,
                              int nd, int model, uint8_t* dst, long long cap, long long* n_stored_segments) {
  dfl::MatchParams P = dfl::make_match_params(row, pixel);
  if (nd > 0) {
    P.nd = 1;
    P.rows3 = 0;
    for (int i = 1; i < nd; ++i) dfl::match_params_add(P, dists[i]);
  }
  if (!model) P.nd = 1;
  std::vector<uint16_t> tokens((size_t)NT * piece);
  std::vector<int> ntok(NT);
  const long long seg_bytes = (long long)NT * piece;
  const long long nsegs = (n + seg_bytes - 1) / seg_bytes;
  const long long stride = (dfl::stored_size(seg_bytes) + 15) / 16 * 16 + 16;
  std::vector<uint8_t> region(stride);
  long long pos = 0, stored_count = 0;
  for (long long seg = 0; seg < nsegs; ++seg) {
    uint32_t hist[288], dhist[32];
    dfl::Codes codes;
    dfl::Header hdr;
    dfl::BuildScratch scratch;
    std::vector<uint32_t> off(NT);
    const long long s0 = seg * seg_bytes;
    const long long nseg = std::min(n - s0, seg_bytes);
    auto S = [&](int tid) { return std::min(n, s0 + (long long)tid * piece); };
    auto E = [&](int tid) { return std::min(n, S(tid) + piece); };
    for (int i = 0; i < 288; ++i) hist[i] = 0;
    for (int i = 0; i < 32; ++i) dhist[i] = 0;
    hist[dfl::EOB] = 1;
    dfl::CostModel cm;
    if (model) {
      uint32_t cnt[256] = {0}, tot[4] = {0, 0, 0, 0};
      SampleAdd add{cnt, tot};
      for (int tid = 0; tid < NT; ++tid) dfl::sample_segment(src, n, s0, nseg, tid, NT, P, add);
      cm.far_on = dfl::model_far_on(tot[2], tot[3]);
      for (int b = 0; b < 256; ++b) cm.lit8[b] = dfl::model_lit8(cnt, tot[0], tot[1], b);
      cm.len8 = dfl::model_len8(tot[0], tot[1]);
    }
    for (int tid = 0; tid < NT; ++tid) {
      dfl::TokenVisitor<Store> tv;
      tv.st.base = tokens.data() + tid;
      tv.st.NT = NT;
      tv.n = 0;
      tv.cv.init(hist, dhist, &P);
      dfl::parse_piece(src, n, S(tid), E(tid), P, model ? &cm : nullptr, tv);
      tv.cv.flush();
      ntok[tid] = tv.n;
      if (tv.n > piece) return -6;
    }
    {
      // the kernel's steps ...
      dfl::BuildScratch sd;
      uint32_t dh2[32];
      memcpy(dh2, dhist, sizeof dh2);
      for (int tid = 0; tid < NT; ++tid) dfl::dpar_prepare(dhist, sd, tid);
      for (int tid = NT - 1; tid >= 0; --tid) dfl::dpar_rank(dhist, sd, tid, NT);
      for (int tid = 0; tid < NT; ++tid) dfl::dpar_build(dhist, sd, codes, tid);
      // ... give the code of the one-thread function
      dfl::Codes cd;
      dfl::BuildScratch sd2;
      dfl::distance_code(dh2, sd2, cd);
      if (memcmp(cd.dlen, codes.dlen, dfl::NDIST) || memcmp(cd.dcode, codes.dcode, dfl::NDIST * sizeof(uint16_t))) return -8;
    }
    // the kernel's CTA-parallel construction, one step after the other ...
    scratch.m = 0;
    scratch.hi = 0;
    for (int i = 0; i <= dfl::MAX_LIT_BITS; ++i) scratch.cnt[i] = 0;
    for (int tid = 0; tid < NT; ++tid) dfl::par_rank(hist, dfl::NLIT, scratch, codes.code, codes.len, tid, NT);
    for (int tid = 0; tid < NT; ++tid) dfl::par_tree(scratch, tid);
    for (int tid = 0; tid < NT; ++tid) dfl::par_count(scratch, dfl::MAX_LIT_BITS, tid, NT);
    for (int tid = 0; tid < NT; ++tid) dfl::par_limit(scratch, dfl::MAX_LIT_BITS, tid);
    for (int tid = 0; tid < NT; ++tid) dfl::par_lengths(scratch, dfl::MAX_LIT_BITS, codes.len, tid, NT);
    for (int tid = 0; tid < NT; ++tid) dfl::par_codes(scratch, dfl::NLIT, codes.len, codes.code, tid, NT);
    for (int tid = 0; tid < NT; ++tid) dfl::hpar_prepare(codes, hdr, tid, NT);
    for (int tid = NT - 1; tid >= 0; --tid) dfl::hpar_fill(codes, hdr, tid, NT);
    for (int tid = 0; tid < NT; ++tid) dfl::hpar_count(hdr, tid, NT);
    for (int tid = NT - 1; tid >= 0; --tid) dfl::hpar_blocks(hdr, tid, NT);
    for (int tid = NT - 1; tid >= 0; --tid) dfl::hpar_tokens(hdr, tid, NT);
    for (int tid = 0; tid < NT; ++tid) dfl::hpar_clprepare(hdr, tid);
    for (int tid = NT - 1; tid >= 0; --tid) dfl::hpar_clrank(scratch, hdr, tid, NT);
    for (int tid = 0; tid < NT; ++tid) dfl::hpar_clcode(scratch, hdr, tid);
    for (int tid = NT - 1; tid >= 0; --tid) dfl::hpar_size(hdr, tid, NT);
    for (int tid = 0; tid < NT; ++tid) dfl::hpar_finish(hdr, tid);
    {
      // ... gives the code of the serial build_code()
      dfl::Codes c2 = codes;
      dfl::Header h2;
      dfl::BuildScratch s2;
      int hi = 0;
      for (int i = 0; i < dfl::NLIT; ++i)
        if (hist[i]) { s2.sorted[dfl::rank_of(hist, dfl::NLIT, i)] = uint16_t(i); hi = i; }
      dfl::segment_build(hist, s2, c2, h2);
      if (memcmp(c2.len, codes.len, sizeof c2.len) || memcmp(c2.code, codes.code, sizeof c2.code)) return -4;
      if (h2.bits != hdr.bits || h2.ntok != hdr.ntok || int(scratch.hi) != hi) return -5;
      if (h2.hlit != hdr.hlit || h2.hdist != hdr.hdist || h2.hclen != hdr.hclen || memcmp(h2.tok_sym, hdr.tok_sym, size_t(h2.ntok)) ||
          memcmp(h2.tok_ext, hdr.tok_ext, size_t(h2.ntok)) || memcmp(h2.cllen, hdr.cllen, sizeof h2.cllen) ||
          memcmp(h2.clcode, hdr.clcode, sizeof h2.clcode))
        return -7;
    }
    for (int tid = 0; tid < NT; ++tid) {
      dfl::SizeVisitor sv;
      sv.c = &codes;
      sv.P = &P;
      sv.bits = 0;
      for (int j = 0; j < ntok[tid]; ++j) dfl::visit_token(tokens[(size_t)j * NT + tid], sv);
      off[tid] = sv.bits;
    }
    long long acc = hdr.bits;
    for (int t = 0; t < NT; ++t) {
      const uint32_t b = off[t];
      off[t] = uint32_t(acc);
      acc += b;
    }
    const long long dyn = dfl::dynamic_size(hdr, codes, acc - hdr.bits);
    const long long st = dfl::stored_size(nseg);
    const bool stored = dyn >= st;
    const long long total = stored ? st : dyn;
    std::fill(region.begin(), region.end(), 0xAA);     // the kernel's scratch is not zero either
    if (stored) {
      ++stored_count;
      for (int tid = 0; tid < NT; ++tid) dfl::stored_copy(src + s0, nseg, S(tid) - s0, E(tid) - s0, region.data());
    } else {
      memset(region.data(), 0, size_t((total + 15) / 16 * 16));
      // threads in reverse order: the result must not depend on who writes a shared word first
      for (int tid = NT - 1; tid >= 0; --tid) {
        dfl::BitWriter bw;
        bw.init(reinterpret_cast<uint32_t*>(region.data()), tid == 0 ? 0 : (long long)off[tid]);
        if (tid == 0) dfl::header_emit(hdr, bw);
        dfl::EmitVisitor ev;
        ev.c = &codes;
        ev.P = &P;
        ev.bw = &bw;
        for (int j = 0; j < ntok[tid]; ++j) dfl::visit_token(tokens[(size_t)j * NT + tid], ev);
        if (tid == 0 && bw.bitpos() != (NT > 1 ? (long long)off[1] : bw.bitpos()) && S(1) < E(1)) return -2;
        if (tid == NT - 1) {
          dfl::segment_close(codes, bw);
          if (bw.bitpos() != total * 8) return -3;
        }
        bw.finish();
      }
    }
    if (pos + total + 2 > cap) return -1;
    memcpy(dst + pos, region.data(), size_t(total));
    pos += total;
  }
  if (pos + 2 > cap) return -1;
  dst[pos++] = 0x03;
  dst[pos++] = 0x00;
  if (n_stored_segments) *n_stored_segments = stored_count;
  return pos;
}

// CRC-32 of src[0, n) the way crc32_kernel (vcf_b200/csrc/kernels_deflate.cu) forms it: chunks of
// `chunk` bytes, one per "thread", XOR of their shifted CRCs.
#include "../../vcf_b200/csrc/crc32_core.cuh"

extern "C" uint32_t crc_emul(const uint8_t* src, long long n, int chunk) {
  uint32_t tab[256];
  for (uint32_t i = 0; i < 256; ++i) tab[i] = crc::table_entry(i);
  crc::Powers P;
  crc::make_powers(P);
  uint32_t total = 0;
  for (long long s = 0; s < n; s += chunk) total ^= crc::chunk_term(src, n, s, std::min(n, s + chunk), tab, P);
  return total;
}
