// One slice of the alignment-kernel instantiations (see variants.cuh).
#include "variants.cuh"

namespace dvo {
align_fn pick_align_128_g0(int w, int oob, int depth) { return pick_variants<DVO_T128, DVO_MINB_128, 0>(w, oob, depth); }
}  // namespace dvo
