// One slice of the alignment-kernel instantiations (see variants.cuh).
#include "variants.cuh"

namespace dvo {
align_fn pick_align_128_g1(int w, int oob) { return pick_variants<DVO_T128, DVO_MINB_128, 1>(w, oob, 0); }
}  // namespace dvo
