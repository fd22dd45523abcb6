// B = 16 fast path, float32 decoder = the fast mode of the north star (pixels within +-1 LSB of the reference's
// float64 chain): the kernel of dec16.cuh instantiated for float with contraction allowed (this unit is compiled
// WITH fused multiply-adds, unlike kernels_b16.cu).  Blocks whose only indices are the DC ones are evaluated by
// two multiplications per pass, as in the exact decoder.
#include "dec16.cuh"

namespace vcfb {

int launch_decode_fast16_f32(const DecArgs& a, cudaStream_t s) {
  if (a.flags & (VCFB_F_FP64 | VCFB_F_SYNTH_F32)) return VCFB_E_UNSUPP;
  return b16::launch_decode16<float, false>(a, s, "dec16_f32");
}

}  // namespace vcfb
