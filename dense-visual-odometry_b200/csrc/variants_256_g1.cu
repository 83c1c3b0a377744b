// One slice of the alignment-kernel instantiations (see variants.cuh).
#include "variants.cuh"

namespace dvo {
align_fn pick_align_256_g1(int w, int oob) { return pick_variants<256, 1, 1>(w, oob, 0); }
}  // namespace dvo
