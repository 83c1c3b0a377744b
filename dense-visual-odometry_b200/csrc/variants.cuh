// Kernel dispatch tables, split over several translation units so that the ~60 template instantiations of the
// alignment kernel compile in parallel (__graft_entry__.build): each variants_*.cu instantiates one slice and
// exports one pick function; dvo_b200.cu only calls them.
#pragma once
#include "align_kernel.cuh"

namespace dvo {

typedef void (*align_fn)(const AlignParams);

// CTAs per SM the 128-thread kernels are compiled for (__launch_bounds__ minimum): 2 = 8 warps per SM at 255 registers
#ifndef DVO_MINB_128
#define DVO_MINB_128 2
#endif
#ifndef DVO_T128   // developer knob: threads per CTA of the default launch shape
#define DVO_T128 128
#endif

// weights x oob_mode for one launch shape (T threads, B CTAs per SM) and one gradient mode G; depth != 0 selects the
// photometric + depth residual variants (G = 0, unweighted or fixed-threshold Huber only).  nullptr = not built.
template <int T, int B, int G>
static align_fn pick_variants(int w, int oob, int depth) {
#ifdef DVO_FAST_BUILD   // developer build: the headline variant only
    if (G == 0 && !depth && w == DVO_W_NONE && oob == DVO_OOB_INCLUSIVE)
        return (align_fn)align_kernel<DVO_W_NONE, DVO_OOB_INCLUSIVE, G, T, B>;
    return nullptr;
#else
    if (depth) {
        if constexpr (G == 0) {
#define DVO_PICKD(WM, OM) \
    if (w == WM && oob == OM) return (align_fn)align_kernel<WM, OM, 0, T, B, 1>;
            DVO_PICKD(DVO_W_NONE, DVO_OOB_INCLUSIVE)
            DVO_PICKD(DVO_W_NONE, DVO_OOB_STRICT)
            DVO_PICKD(DVO_W_HUBER, DVO_OOB_INCLUSIVE)
            DVO_PICKD(DVO_W_HUBER, DVO_OOB_STRICT)
#undef DVO_PICKD
        }
        return nullptr;
    }
#define DVO_PICK(WM, OM) \
    if (w == WM && oob == OM) return (align_fn)align_kernel<WM, OM, G, T, B>;
    DVO_PICK(DVO_W_NONE, DVO_OOB_INCLUSIVE)
    DVO_PICK(DVO_W_NONE, DVO_OOB_STRICT)
    DVO_PICK(DVO_W_TDIST_REF, DVO_OOB_INCLUSIVE)
    DVO_PICK(DVO_W_TDIST_REF, DVO_OOB_STRICT)
    DVO_PICK(DVO_W_HUBER, DVO_OOB_INCLUSIVE)
    DVO_PICK(DVO_W_HUBER, DVO_OOB_STRICT)
    DVO_PICK(DVO_W_HUBER_MAD, DVO_OOB_INCLUSIVE)
    DVO_PICK(DVO_W_HUBER_MAD, DVO_OOB_STRICT)
#undef DVO_PICK
    return nullptr;
#endif
}

// Cluster-mode kernels (one thread-block cluster per pair) for one gradient mode; not built for Huber/MAD.
// depth != 0: the photometric + depth residual variants (G = 0, unweighted or fixed-threshold Huber).
template <int G>
static align_fn pick_cluster_variants(int w, int oob, int depth) {
#ifdef DVO_FAST_BUILD
    if (G == 0 && !depth && w == DVO_W_NONE && oob == DVO_OOB_INCLUSIVE)
        return (align_fn)align_cluster_kernel<DVO_W_NONE, DVO_OOB_INCLUSIVE, G>;
    return nullptr;
#else
    if (depth) {
        if constexpr (G == 0) {
#define DVO_PICKCD(WM, OM) \
    if (w == WM && oob == OM) return (align_fn)align_cluster_kernel<WM, OM, 0, 1>;
            DVO_PICKCD(DVO_W_NONE, DVO_OOB_INCLUSIVE)
            DVO_PICKCD(DVO_W_NONE, DVO_OOB_STRICT)
            DVO_PICKCD(DVO_W_HUBER, DVO_OOB_INCLUSIVE)
            DVO_PICKCD(DVO_W_HUBER, DVO_OOB_STRICT)
#undef DVO_PICKCD
        }
        return nullptr;
    }
#define DVO_PICKC(WM, OM) \
    if (w == WM && oob == OM) return (align_fn)align_cluster_kernel<WM, OM, G>;
    DVO_PICKC(DVO_W_NONE, DVO_OOB_INCLUSIVE)
    DVO_PICKC(DVO_W_NONE, DVO_OOB_STRICT)
    DVO_PICKC(DVO_W_HUBER, DVO_OOB_INCLUSIVE)
    DVO_PICKC(DVO_W_HUBER, DVO_OOB_STRICT)
    DVO_PICKC(DVO_W_TDIST_REF, DVO_OOB_INCLUSIVE)
    DVO_PICKC(DVO_W_TDIST_REF, DVO_OOB_STRICT)
#undef DVO_PICKC
    return nullptr;
#endif
}

// Tail kernels (align_kernel with CL = 1: one thread-block cluster finishes one pair the persistent kernel left
// unfinished), for one gradient mode; the same combinations as the cluster-mode kernels.
template <int G>
static align_fn pick_tail_variants(int w, int oob, int depth) {
#ifdef DVO_FAST_BUILD
    if (G == 0 && !depth && w == DVO_W_NONE && oob == DVO_OOB_INCLUSIVE)
        return (align_fn)align_kernel<DVO_W_NONE, DVO_OOB_INCLUSIVE, G, 128, 2, 0, 1>;
    return nullptr;
#else
    if (depth) {
        if constexpr (G == 0) {
#define DVO_PICKTD(WM, OM) \
    if (w == WM && oob == OM) return (align_fn)align_kernel<WM, OM, 0, 128, 2, 1, 1>;
            DVO_PICKTD(DVO_W_NONE, DVO_OOB_INCLUSIVE)
            DVO_PICKTD(DVO_W_NONE, DVO_OOB_STRICT)
            DVO_PICKTD(DVO_W_HUBER, DVO_OOB_INCLUSIVE)
            DVO_PICKTD(DVO_W_HUBER, DVO_OOB_STRICT)
#undef DVO_PICKTD
        }
        return nullptr;
    }
#define DVO_PICKT(WM, OM) \
    if (w == WM && oob == OM) return (align_fn)align_kernel<WM, OM, G, 128, 2, 0, 1>;
    DVO_PICKT(DVO_W_NONE, DVO_OOB_INCLUSIVE)
    DVO_PICKT(DVO_W_NONE, DVO_OOB_STRICT)
    DVO_PICKT(DVO_W_HUBER, DVO_OOB_INCLUSIVE)
    DVO_PICKT(DVO_W_HUBER, DVO_OOB_STRICT)
    DVO_PICKT(DVO_W_TDIST_REF, DVO_OOB_INCLUSIVE)
    DVO_PICKT(DVO_W_TDIST_REF, DVO_OOB_STRICT)
#undef DVO_PICKT
    return nullptr;
#endif
}

// one definition per variants_*.cu
align_fn pick_align_128_g0(int w, int oob, int depth);
align_fn pick_align_128_g1(int w, int oob);
align_fn pick_align_256_g0(int w, int oob, int depth);
align_fn pick_align_256_g1(int w, int oob);
align_fn pick_cluster_g0(int w, int oob, int depth);
align_fn pick_cluster_g1(int w, int oob);
align_fn pick_tail_g0(int w, int oob, int depth);
align_fn pick_tail_g1(int w, int oob);

}  // namespace dvo
