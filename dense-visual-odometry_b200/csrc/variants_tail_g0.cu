// One slice of the alignment-kernel instantiations (see variants.cuh): the tail kernels of the default gradient mode.
#include "variants.cuh"

namespace dvo {
align_fn pick_tail_g0(int w, int oob, int depth) { return pick_tail_variants<0>(w, oob, depth); }
}  // namespace dvo
