// One slice of the alignment-kernel instantiations (see variants.cuh).
#include "variants.cuh"

namespace dvo {
align_fn pick_cluster_g0(int w, int oob, int depth) { return pick_cluster_variants<0>(w, oob, depth); }
}  // namespace dvo
