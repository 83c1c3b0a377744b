// One slice of the alignment-kernel instantiations (see variants.cuh): the tail kernels of approximate_image2_gradient.
#include "variants.cuh"

namespace dvo {
align_fn pick_tail_g1(int w, int oob) { return pick_tail_variants<1>(w, oob, 0); }
}  // namespace dvo
