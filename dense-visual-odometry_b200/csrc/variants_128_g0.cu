// One slice of the alignment-kernel instantiations (see variants.cuh).
#include "variants.cuh"

namespace dvo {
align_fn pick_align_128_g0(int w, int oob, int depth) { return pick_variants<128, 2, 0>(w, oob, depth); }
}  // namespace dvo
