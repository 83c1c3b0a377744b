// One slice of the alignment-kernel instantiations (see variants.cuh).
#include "variants.cuh"

namespace dvo {
align_fn pick_cluster_g1(int w, int oob) { return pick_cluster_variants<1>(w, oob, 0); }
}  // namespace dvo
