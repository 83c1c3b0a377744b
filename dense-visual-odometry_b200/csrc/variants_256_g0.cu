// One slice of the alignment-kernel instantiations (see variants.cuh).
#include "variants.cuh"

namespace dvo {
align_fn pick_align_256_g0(int w, int oob, int depth) { return pick_variants<256, 1, 0>(w, oob, depth); }
}  // namespace dvo
