// libdvo_b200.so — C ABI (include/dvo_b200.h) over the sm_100a kernels.
// Build: every csrc/*.cu with nvcc -c -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3, linked with
// nvcc -shared (see __graft_entry__.build, which compiles the translation units in parallel).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "../../include/dvo_b200.h"
#include "align_kernel.cuh"
#include "pyramid_kernels.cuh"
#include "variants.cuh"

using namespace dvo;

struct dvo_handle {
    int device = 0, H = 0, W = 0, levels = 0, max_frames = 0, max_pairs = 0;
    dvo_config cfg{};
    bool intrinsics_set = false;
    float fx = 0, fy = 0, cx = 0, cy = 0;
    double depth_scale = 0;
    int clamp_thr = 65536;
    int lw[DVO_MAX_LEVELS]{}, lh[DVO_MAX_LEVELS]{}, lpitch[DVO_MAX_LEVELS]{};
    size_t lplane[DVO_MAX_LEVELS]{};
    size_t n_rec[DVO_MAX_LEVELS]{}, n_raw[DVO_MAX_LEVELS]{};   // elements allocated per level (frames + slack rows)
    unsigned long long* dbg_violations = nullptr;            // DVO_BOUNDS_CHECK builds only
    uint8_t* gray[DVO_MAX_LEVELS]{};
    uint16_t* depth[DVO_MAX_LEVELS]{};
    uint2* rec[DVO_MAX_LEVELS]{};
    float* prec[DVO_MAX_LEVELS]{};   // previous-frame point lists, room for 2 words per pixel (align_kernel.cuh, pt_pack)
    int* pt_tiles[DVO_MAX_LEVELS]{};   // [frame] tiles of the frame's point list
    std::vector<uint8_t> slot_built;   // per frame slot: kHasPoints | kHasTaps, what the last build of the slot produced
    float k4[DVO_MAX_LEVELS][4]{}, kinv4[DVO_MAX_LEVELS][4]{};
    int* queue = nullptr;   // kQueueSlots x 4 work-queue counters; concurrent dvo_estimate calls (different streams) rotate through them
    int queue_next = 0;
    unsigned long long* ring = nullptr;   // work-queue cells, one per frame slot (a launch uses those of its previous frames)
    int* dlist = nullptr;                 // pairs left to the tail kernel, one entry per frame slot
    PairState* pstate = nullptr;          // saved Gauss-Newton state of a pair between two time slices, one per frame slot
    float* chunk_sums = nullptr;          // per-chunk sums of a pair's current pass (canonical summation order), per frame slot
    int sm_count = 0, threads = 256, blocks_per_sm = 2, grid_max = 0;
    uint8_t* stage_bgr = nullptr;
    uint16_t* stage_depth = nullptr;
    float* qt_init = nullptr;
    float* qt_last = nullptr;
    float* qt_out = nullptr;
    float* qt_one = nullptr;  // 7 + 12 floats for dvo_residuals_jacobian
    dvo_pair_stats* stats = nullptr;
    cudaEvent_t ev0[8]{}, ev1[8]{};   // event pairs of the last launches (ring): calls in flight on different streams
    int ev_last = -1;                 // never share a pair; dvo_last_estimate_ms reads the most recent one
    long long launches = 0;
    std::string err;
};

static const int kQueueSlots = 256;
static const int kMaxPrefetchRows = 32;  // upper bound of dvo_config.reserved[0]; the plane slack is derived from it
static const char* kNullHandle = "null handle";

#define DVO_CUDA(h, call)                                                                             \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return DVO_ERR_CUDA;                                                                      \
        }                                                                                             \
    } while (0)

static int fail(dvo_handle* h, int code, const char* msg) {
    if (h) h->err = msg;
    return code;
}

extern "C" void dvo_default_config(dvo_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->max_iterations = 100;
    cfg->max_increased_steps = 0;
    cfg->tolerance = 1e-6f;
    cfg->sigma_prior = -1.0f;
    cfg->weights = DVO_W_NONE;
    cfg->oob_mode = DVO_OOB_INCLUSIVE;
    cfg->tdist_dof = 5.0f;
    cfg->tdist_init_sigma = 5.0f;
    cfg->tdist_tolerance = 1e-3f;
    cfg->tdist_max_iterations = 50;
    cfg->huber_k = 1.345f * 5.0f;
    cfg->max_distance = 5.0f;
    cfg->depth_weight = 2500.0f;
}

extern "C" const char* dvo_last_error(const dvo_handle* h) { return h ? h->err.c_str() : kNullHandle; }

// ---- kernel dispatch ---------------------------------------------------------------------------
// The alignment-kernel instantiations live in variants_*.cu (compiled in parallel); see variants.cuh.
// Launch shapes: 128 threads x 2 CTAs per SM (default: one CTA's reduction / solve overlaps the other's streaming)
// or 256 threads x 1 CTA per SM (lower latency for a single pair).  Both run 8 warps per SM at 255 registers;
// shapes with more warps per SM spill inside the pipelined loop and measured slower (profiles/r1/SUMMARY.md).
static align_fn get_align(const dvo_handle* h) {
    const int w = h->cfg.weights, o = h->cfg.oob_mode, gm = h->cfg.approximate_image2_gradient ? 1 : 0;
    const int dz = h->cfg.use_depth_residual ? 1 : 0;
    if (gm && dz) return nullptr;
    if (h->threads == 256) return gm ? pick_align_256_g1(w, o) : pick_align_256_g0(w, o, dz);
    return gm ? pick_align_128_g1(w, o) : pick_align_128_g0(w, o, dz);
}

// Cluster-mode kernel (one thread-block cluster per pair); not built for the Huber/MAD weights.
static align_fn get_cluster(const dvo_handle* h) {
    const int w = h->cfg.weights, o = h->cfg.oob_mode;
    const int dz = h->cfg.use_depth_residual ? 1 : 0;
    if (h->cfg.approximate_image2_gradient) return dz ? nullptr : pick_cluster_g1(w, o);
    return pick_cluster_g0(w, o, dz);
}

// Tail kernel (one thread-block cluster per pair the persistent kernel left unfinished); 128-thread shape only.
static align_fn get_tail(const dvo_handle* h) {
    const int w = h->cfg.weights, o = h->cfg.oob_mode;
    const int dz = h->cfg.use_depth_residual ? 1 : 0;
    if (h->threads != 128 || w == DVO_W_HUBER_MAD) return nullptr;
    if (h->cfg.approximate_image2_gradient) return dz ? nullptr : pick_tail_g1(w, o);
    return pick_tail_g0(w, o, dz);
}

typedef void (*dump_fn)(const AlignParams, int, int, int, const float*, float, float*, float*, uint8_t*, uint8_t*,
                        double*);
static dump_fn get_dump(const dvo_handle* h) {
    const bool hub = h->cfg.weights == DVO_W_HUBER;  // dump is unweighted unless Huber
    const bool strict = h->cfg.oob_mode == DVO_OOB_STRICT;
    const bool ap = h->cfg.approximate_image2_gradient != 0;
#define DVO_DUMP(WM, OM, GM) (dump_fn) dump_kernel<WM, OM, GM>
    if (ap) {
        if (hub) return strict ? DVO_DUMP(DVO_W_HUBER, DVO_OOB_STRICT, 1) : DVO_DUMP(DVO_W_HUBER, DVO_OOB_INCLUSIVE, 1);
        return strict ? DVO_DUMP(DVO_W_NONE, DVO_OOB_STRICT, 1) : DVO_DUMP(DVO_W_NONE, DVO_OOB_INCLUSIVE, 1);
    }
    if (hub) return strict ? DVO_DUMP(DVO_W_HUBER, DVO_OOB_STRICT, 0) : DVO_DUMP(DVO_W_HUBER, DVO_OOB_INCLUSIVE, 0);
    return strict ? DVO_DUMP(DVO_W_NONE, DVO_OOB_STRICT, 0) : DVO_DUMP(DVO_W_NONE, DVO_OOB_INCLUSIVE, 0);
#undef DVO_DUMP
}

__global__ void pose_matrix_kernel(const float* qt, float* T12) {
    PoseQT p;
    for (int i = 0; i < 4; ++i) p.q[i] = qt[i];
    for (int i = 0; i < 3; ++i) p.t[i] = qt[4 + i];
    pose_matrix(p, T12);
}

// ---- lifetime ----------------------------------------------------------------------------------
extern "C" int dvo_destroy(dvo_handle* h) {
    if (!h) return DVO_ERR_INVALID;
    cudaSetDevice(h->device);
    for (int l = 0; l < DVO_MAX_LEVELS; ++l) {
        cudaFree(h->gray[l]);
        cudaFree(h->depth[l]);
        cudaFree(h->rec[l]);
        cudaFree(h->prec[l]);
        cudaFree(h->pt_tiles[l]);
    }
    cudaFree(h->queue);
    cudaFree(h->ring);
    cudaFree(h->dlist);
    cudaFree(h->pstate);
    cudaFree(h->chunk_sums);
    cudaFree(h->dbg_violations);
    cudaFree(h->stage_bgr);
    cudaFree(h->stage_depth);
    cudaFree(h->qt_init);
    cudaFree(h->qt_last);
    cudaFree(h->qt_out);
    cudaFree(h->qt_one);
    cudaFree(h->stats);
    for (int i = 0; i < 8; ++i) {
        if (h->ev0[i]) cudaEventDestroy(h->ev0[i]);
        if (h->ev1[i]) cudaEventDestroy(h->ev1[i]);
    }
    delete h;
    return DVO_OK;
}

constexpr uint8_t kHasPoints = 1, kHasTaps = 2;

// The roles the frames of an estimate must have been built for (dvo_build_pyramids, with_gradients).
static int check_roles(dvo_handle* h, int prev_base, int cur_base, int n, bool prev_needs_taps) {
    const uint8_t prev_need = kHasPoints | (prev_needs_taps ? kHasTaps : 0);
    for (int i = 0; i < n; ++i) {
        if ((h->slot_built[prev_base + i] & prev_need) != prev_need)
            return fail(h, DVO_ERR_STATE, prev_needs_taps
                ? "a previous-frame slot was not built with with_gradients = 1 (approximate_image2_gradient needs both roles)"
                : "a previous-frame slot was not built as a previous frame (with_gradients 0 or 1)");
        if (!(h->slot_built[cur_base + i] & kHasTaps))
            return fail(h, DVO_ERR_STATE, "a current-frame slot was not built as a current frame (with_gradients 1 or 2)");
    }
    return DVO_OK;
}

static int create_impl(dvo_handle* h) {
    DVO_CUDA(h, cudaSetDevice(h->device));
    cudaDeviceProp prop;
    DVO_CUDA(h, cudaGetDeviceProperties(&prop, h->device));
    h->sm_count = prop.multiProcessorCount;
    if (h->cfg.reserved[0] > kMaxPrefetchRows) {   // the kernels read that many rows ahead; the plane slack is sized for it
        h->err = "reserved[0] (L1 prefetch distance in rows) must be <= 32";
        return DVO_ERR_INVALID;
    }
    if (h->W > kPtMaxDim + 1 || h->H > kPtMaxDim + 1) {   // a point word holds 11-bit column and row indices (pt_pack)
        h->err = "image too large: width and height must be <= 2048";
        return DVO_ERR_INVALID;
    }
    int w = h->W, hh = h->H;
    for (int l = 0; l < h->levels; ++l) {
        h->lw[l] = w;
        h->lh[l] = hh;
        h->lpitch[l] = (w + kTile - 1) / kTile * kTile;
        h->lplane[l] = (size_t)hh * h->lpitch[l];
        // tap indices travel as 2^23 + index in a float32 (align_kernel.cuh, prep_pair)
        if (h->lplane[l] + (size_t)h->lpitch[l] + 2 >= (1u << 23)) {
            h->err = "image too large: a level plane must hold fewer than 2^23 pixels";
            return DVO_ERR_INVALID;
        }
        const size_t n = h->lplane[l] * h->max_frames;
        // Reads past the end of the last frame slot (align_kernel.cuh, fused_pass), in rows of this level:
        //   tap records / current depth: tap coordinates are validated (inside the plane, or (0,0)), the unclamped
        //     (x0 + 1, y0 + 1) tap reads at most pitch + 1 elements past the plane, and the L1 touches run up to
        //     kMaxPrefetchRows + 1 rows ahead of a tap: kMaxPrefetchRows + 2 rows + 1 element.
        //   point lists (2 words per pixel of the plane; a list never has more tiles than the plane has 128-pixel
        //     groups): the software pipeline loads 2 tiles past a chunk's last tile and touches up to kMaxPrefetchRows
        //     tiles (+ a tile of lane spread) ahead of that: (kMaxPrefetchRows + 3) * 256 words, which the
        //     2 * (kMaxPrefetchRows + 4) * pitch words of slack cover because pitch >= 128.
        // One row more than needed is allocated for each.
        const size_t n_rec = n + (size_t)(kMaxPrefetchRows + 3) * (size_t)h->lpitch[l];
        const size_t n_raw = n + (size_t)(kMaxPrefetchRows + 4) * (size_t)h->lpitch[l];
        h->n_rec[l] = n_rec;
        h->n_raw[l] = n_raw;
        DVO_CUDA(h, cudaMalloc(&h->gray[l], n_raw));
        DVO_CUDA(h, cudaMalloc(&h->depth[l], n_raw * sizeof(uint16_t)));
        DVO_CUDA(h, cudaMalloc(&h->rec[l], n_rec * sizeof(uint2)));
        DVO_CUDA(h, cudaMemset(h->gray[l], 0, n_raw));
        DVO_CUDA(h, cudaMemset(h->depth[l], 0, n_raw * sizeof(uint16_t)));
        DVO_CUDA(h, cudaMemset(h->rec[l], 0, n_rec * sizeof(uint2)));
        DVO_CUDA(h, cudaMalloc(&h->prec[l], 2 * n_raw * sizeof(float)));
        prec_fill_kernel<<<h->sm_count * 8, 256>>>(h->prec[l], 2 * n_raw);   // "no depth" points everywhere
        DVO_CUDA(h, cudaMalloc(&h->pt_tiles[l], (size_t)h->max_frames * sizeof(int)));
        DVO_CUDA(h, cudaMemset(h->pt_tiles[l], 0, (size_t)h->max_frames * sizeof(int)));
        DVO_CUDA(h, cudaGetLastError());
        {   // strip = umulhi(t, floor(2^32/h)+1) must be exact for every tile index of the plane
            const unsigned magic = (unsigned)((1ull << 32) / (unsigned)hh) + 1u;
            const int strips = h->lpitch[l] / kTile;
            for (int s = 1; s <= strips; ++s) {
                const unsigned long long t1 = (unsigned long long)s * hh - 1, t2 = t1 + 1;
                if ((unsigned)((t1 * magic) >> 32) != (unsigned)(s - 1) || (s < strips && (unsigned)((t2 * magic) >> 32) != (unsigned)s)) {
                    h->err = "image too large for the 32-bit tile index arithmetic";
                    return DVO_ERR_INVALID;
                }
            }
        }
        w = (w + 1) / 2;   // image_pyramid.py:21 / :84-85 (ceil division)
        hh = (hh + 1) / 2;
    }
    h->threads = h->cfg.threads_per_block ? h->cfg.threads_per_block : DVO_T128;
    if (h->threads != DVO_T128 && h->threads != 256) {
        h->err = "threads_per_block must be 0, 128 or 256";
        return DVO_ERR_INVALID;
    }
    if (h->cfg.cluster_size == -1) {   // the largest cluster this device can co-schedule for the latency kernel
        h->cfg.cluster_size = 8;
        align_fn cfn16 = get_cluster(h);
        if (cfn16 && cudaFuncSetAttribute((const void*)cfn16, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(16);
            lc.blockDim = dim3(kClusterThreads);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 16;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            lc.attrs = at;
            lc.numAttrs = 1;
            int n16 = 0;
            if (cudaOccupancyMaxActiveClusters(&n16, (const void*)cfn16, &lc) == cudaSuccess && n16 >= 1) h->cfg.cluster_size = 16;
        }
        (void)cudaGetLastError();   // a failed query is not an error of the handle
    }
    if (h->cfg.cluster_size != 0 && h->cfg.cluster_size != 1 && h->cfg.cluster_size != 2 && h->cfg.cluster_size != 4 &&
        h->cfg.cluster_size != 8 && h->cfg.cluster_size != 16) {
        h->err = "cluster_size must be -1, 0, 1, 2, 4, 8 or 16";
        return DVO_ERR_INVALID;
    }
    {   // points_kernel keeps one count per (strip, row) of a level in shared memory
        const size_t smem = points_smem_bytes(h->lpitch[0] / kTile, h->lh[0]);
        if (smem > 200 * 1024) {
            h->err = "image too large for the point-list build (strips x rows of level 0 must fit shared memory)";
            return DVO_ERR_INVALID;
        }
        // the attribute belongs to the function on this device, not to the handle: handles of different image sizes
        // share it, so it is only ever raised
        static size_t granted[64] = {};
        const int di = h->device & 63;
        if (smem > granted[di]) {
            DVO_CUDA(h, cudaFuncSetAttribute((const void*)points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            granted[di] = smem;
        }
    }
    align_fn fn = get_align(h);
    if (!fn) {
        h->err = "unsupported weights / oob_mode / approximate_image2_gradient / use_depth_residual combination";
        return DVO_ERR_INVALID;
    }
    int occ = 0;
    DVO_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, h->threads, 0));
    if (occ < 1) occ = 1;
    h->blocks_per_sm = h->cfg.blocks_per_sm > 0 ? (h->cfg.blocks_per_sm < occ ? h->cfg.blocks_per_sm : occ) : occ;
    h->grid_max = h->sm_count * h->blocks_per_sm;
    DVO_CUDA(h, cudaMalloc(&h->queue, sizeof(int) * 8 * kQueueSlots));
    DVO_CUDA(h, cudaMalloc(&h->dlist, sizeof(int) * h->max_frames));
    DVO_CUDA(h, cudaMalloc(&h->ring, sizeof(unsigned long long) * h->max_frames));
    DVO_CUDA(h, cudaMalloc(&h->pstate, sizeof(PairState) * h->max_frames));
    DVO_CUDA(h, cudaMalloc(&h->chunk_sums, sizeof(float) * kMaxChunks * kChunkFloats * (size_t)h->max_frames));
#ifdef DVO_BOUNDS_CHECK
    DVO_CUDA(h, cudaMalloc(&h->dbg_violations, sizeof(unsigned long long)));
    DVO_CUDA(h, cudaMemset(h->dbg_violations, 0, sizeof(unsigned long long)));
#endif
    DVO_CUDA(h, cudaMalloc(&h->qt_init, sizeof(float) * 7 * h->max_pairs));
    DVO_CUDA(h, cudaMalloc(&h->qt_last, sizeof(float) * 7 * h->max_pairs));
    DVO_CUDA(h, cudaMalloc(&h->qt_out, sizeof(float) * 7 * h->max_pairs));
    DVO_CUDA(h, cudaMalloc(&h->qt_one, sizeof(float) * 32));
    DVO_CUDA(h, cudaMalloc(&h->stats, sizeof(dvo_pair_stats) * h->max_pairs));
    for (int i = 0; i < 8; ++i) {
        DVO_CUDA(h, cudaEventCreate(&h->ev0[i]));
        DVO_CUDA(h, cudaEventCreate(&h->ev1[i]));
    }
    return DVO_OK;
}

extern "C" int dvo_create(dvo_handle** out, int device, int height, int width, int levels, int max_frames,
                          int max_pairs, const dvo_config* cfg) {
    if (!out) return DVO_ERR_INVALID;
    *out = nullptr;
    if (height < 1 || width < 1 || levels < 1 || levels > DVO_MAX_LEVELS || max_frames < 1 || max_pairs < 1)
        return DVO_ERR_INVALID;
    dvo_handle* h = new (std::nothrow) dvo_handle();
    if (!h) return DVO_ERR_INVALID;
    h->device = device;
    h->H = height;
    h->W = width;
    h->levels = levels;
    h->max_frames = max_frames;
    h->slot_built.assign((size_t)max_frames, 0);
    h->max_pairs = max_pairs;
    if (cfg)
        h->cfg = *cfg;
    else
        dvo_default_config(&h->cfg);
    *out = h;  // returned even on failure so the caller can read dvo_last_error, then dvo_destroy
    return create_impl(h);
}

extern "C" int dvo_set_intrinsics(dvo_handle* h, float fx, float fy, float cx, float cy, double depth_scale) {
    if (!h) return DVO_ERR_INVALID;
    if (!(depth_scale >= 0.0)) return fail(h, DVO_ERR_INVALID, "depth_scale must be >= 0");
    h->fx = fx; h->fy = fy; h->cx = cx; h->cy = cy;
    h->depth_scale = depth_scale;
    for (int l = 0; l < h->levels; ++l) {
        // RGBDCameraModel.at (camera_model.py:62-79): K_l = S_l K in float32
        const float s = ldexpf(1.0f, -l);
        const float o = (l == 0) ? 0.0f : (ldexpf(1.0f, -l - 1) - 0.5f);
        volatile float sx = s * cx, sy = s * cy;
        const float fxl = s * fx, fyl = s * fy;
        const float cxl = (l == 0) ? cx : (float)(sx + o), cyl = (l == 0) ? cy : (float)(sy + o);
        h->k4[l][0] = fxl; h->k4[l][1] = fyl; h->k4[l][2] = cxl; h->k4[l][3] = cyl;
        // np.linalg.inv(K_l) in float32 (camera_model.py:216): 1/fx and -(cx/fx), as LAPACK's gesv yields
        volatile float ifx = 1.0f / fxl, ify = 1.0f / fyl;
        volatile float qx = cxl / fxl, qy = cyl / fyl;
        h->kinv4[l][0] = ifx; h->kinv4[l][1] = ify; h->kinv4[l][2] = -qx; h->kinv4[l][3] = -qy;
    }
    // base_dense_visual_odometry.py:59: (depth * scale) > max_distance in float64
    h->clamp_thr = 65536;
    for (int d = 0; d < 65536; ++d) {
        volatile double m = (double)d * depth_scale;
        if (m > (double)h->cfg.max_distance) {
            h->clamp_thr = d;
            break;
        }
    }
    h->intrinsics_set = true;
    return DVO_OK;
}

extern "C" int dvo_depth_clamp_threshold(const dvo_handle* h, int* threshold) {
    if (!h || !threshold) return DVO_ERR_INVALID;
    if (!h->intrinsics_set) return DVO_ERR_STATE;
    *threshold = h->clamp_thr;
    return DVO_OK;
}

extern "C" int dvo_level_shape(const dvo_handle* h, int level, int* height, int* width) {
    if (!h || level < 0 || level >= h->levels) return DVO_ERR_RANGE;
    if (height) *height = h->lh[level];
    if (width) *width = h->lw[level];
    return DVO_OK;
}

extern "C" int dvo_level_intrinsics(const dvo_handle* h, int level, float* k4) {
    if (!h || !k4 || level < 0 || level >= h->levels) return DVO_ERR_RANGE;
    if (!h->intrinsics_set) return DVO_ERR_STATE;
    for (int i = 0; i < 4; ++i) k4[i] = h->k4[level][i];
    return DVO_OK;
}

extern "C" long long dvo_launch_count(const dvo_handle* h) { return h ? h->launches : 0; }

// ---- pyramids ----------------------------------------------------------------------------------
// with_gradients (the `roles` of the frames): 0 = used as previous frames only (previous-frame records, no tap
// records), 1 = both roles, 2 = used as current frames only (tap records, no previous-frame records).
static int build_levels(dvo_handle* h, int frame_base, int n_frames, int with_gradients, cudaStream_t st) {
    for (int l = 1; l < h->levels; ++l) {
        const int groups = (h->lw[l] + 3) / 4 * h->lh[l];
        dim3 grid((groups + 127) / 128, n_frames);
        median3_down_pair_kernel<<<grid, 128, 0, st>>>(
            h->gray[l - 1] + (size_t)frame_base * h->lplane[l - 1], h->gray[l] + (size_t)frame_base * h->lplane[l],
            h->depth[l - 1] + (size_t)frame_base * h->lplane[l - 1], h->depth[l] + (size_t)frame_base * h->lplane[l],
            h->lw[l - 1], h->lh[l - 1], h->lpitch[l - 1], h->lplane[l - 1], h->lw[l], h->lh[l], h->lpitch[l], h->lplane[l]);
        h->launches += 1;
    }
    if (with_gradients != 2) {   // frames used as previous frames: their point lists
        for (int l = 0; l < h->levels; ++l) {
            const size_t smem = points_smem_bytes((h->lw[l] + 127) / 128, h->lh[l]);
            points_kernel<<<n_frames, 1024, smem, st>>>(h->gray[l] + (size_t)frame_base * h->lplane[l],
                                                     h->depth[l] + (size_t)frame_base * h->lplane[l],
                                                     h->prec[l] + 2 * (size_t)frame_base * h->lplane[l],
                                                     h->pt_tiles[l] + frame_base, h->depth_scale, h->lw[l], h->lh[l],
                                                     h->lpitch[l], h->lplane[l]);
            h->launches += 1;
        }
    }
    if (with_gradients) {
        for (int l = 0; l < h->levels; ++l) {
            dim3 grid(((h->lw[l] + 7) / 8 * h->lh[l] + 127) / 128, n_frames);
            sobel3_kernel<<<grid, 128, 0, st>>>(h->gray[l] + (size_t)frame_base * h->lplane[l],
                                                h->depth[l] + (size_t)frame_base * h->lplane[l],
                                                h->rec[l] + (size_t)frame_base * h->lplane[l], h->lw[l], h->lh[l],
                                                h->lpitch[l], h->lplane[l]);
            h->launches += 1;
        }
    }
    DVO_CUDA(h, cudaGetLastError());
    return DVO_OK;
}

static int build_impl(dvo_handle* h, int frame_base, const uint8_t* img, uint16_t* depth, int n_frames,
                      int with_gradients, bool has_bgr, bool clamp, cudaStream_t st) {
    if (!h) return DVO_ERR_INVALID;
    if (!img || !depth) return fail(h, DVO_ERR_INVALID, "null image pointer");
    if (n_frames < 1 || frame_base < 0 || frame_base + n_frames > h->max_frames)
        return fail(h, DVO_ERR_RANGE, "frame slots out of range");
    if (n_frames > 65535) return fail(h, DVO_ERR_RANGE, "at most 65535 frames per build call (the frame index is gridDim.y)");
    if (!h->intrinsics_set) return fail(h, DVO_ERR_STATE, "dvo_set_intrinsics must be called first");
    if (with_gradients < 0 || with_gradients > 2) return fail(h, DVO_ERR_INVALID, "with_gradients must be 0, 1 or 2");
    DVO_CUDA(h, cudaSetDevice(h->device));
    const int gpr = (h->W + 3) / 4;
    dim3 grid((gpr * h->H + 255) / 256, n_frames);
    const bool vec = (h->W % 4 == 0) && (((uintptr_t)img & 3) == 0) && (((uintptr_t)depth & 7) == 0);
    uint8_t* g0 = h->gray[0] + (size_t)frame_base * h->lplane[0];
    uint16_t* d0 = h->depth[0] + (size_t)frame_base * h->lplane[0];
#define DVO_LAUNCH_GC(V, B)                                                                                  \
    gray_clamp_kernel<V, B><<<grid, 256, 0, st>>>(img, depth, g0, d0, h->W, h->H, h->lpitch[0], h->lplane[0], \
                                                  h->clamp_thr, clamp ? 1 : 0)
    if (vec && has_bgr) DVO_LAUNCH_GC(true, true);
    else if (vec) DVO_LAUNCH_GC(true, false);
    else if (has_bgr) DVO_LAUNCH_GC(false, true);
    else DVO_LAUNCH_GC(false, false);
#undef DVO_LAUNCH_GC
    h->launches += 1;
    for (int i = 0; i < n_frames; ++i)
        h->slot_built[frame_base + i] = (uint8_t)((with_gradients != 2 ? kHasPoints : 0) | (with_gradients != 0 ? kHasTaps : 0));
    return build_levels(h, frame_base, n_frames, with_gradients, st);
}

extern "C" int dvo_build_pyramids(dvo_handle* h, int frame_base, const uint8_t* bgr_dev, uint16_t* depth_dev,
                                  int n_frames, int with_gradients, void* stream) {
    return build_impl(h, frame_base, bgr_dev, depth_dev, n_frames, with_gradients, true, true, (cudaStream_t)stream);
}

extern "C" int dvo_build_pyramids_gray(dvo_handle* h, int frame_base, const uint8_t* gray_dev,
                                       const uint16_t* depth_dev, int n_frames, int with_gradients, void* stream) {
    return build_impl(h, frame_base, gray_dev, const_cast<uint16_t*>(depth_dev), n_frames, with_gradients, false,
                      false, (cudaStream_t)stream);
}

static int stage_ready(dvo_handle* h, int frame_base, int n_frames) {
    if (n_frames < 1 || frame_base < 0 || frame_base + n_frames > h->max_frames)
        return fail(h, DVO_ERR_RANGE, "frame slots out of range");
    DVO_CUDA(h, cudaSetDevice(h->device));
    if (!h->stage_bgr) {
        const size_t px = (size_t)h->H * h->W;
        DVO_CUDA(h, cudaMalloc(&h->stage_bgr, px * 3 * h->max_frames));
        DVO_CUDA(h, cudaMalloc(&h->stage_depth, px * sizeof(uint16_t) * h->max_frames));
    }
    return DVO_OK;
}

extern "C" int dvo_upload_frames(dvo_handle* h, int frame_base, const uint8_t* bgr_host, const uint16_t* depth_host,
                                 int n_frames, void* stream) {
    if (!h) return DVO_ERR_INVALID;
    if (!bgr_host || !depth_host) return fail(h, DVO_ERR_INVALID, "null image pointer");
    const int rc = stage_ready(h, frame_base, n_frames);
    if (rc != DVO_OK) return rc;
    const size_t px = (size_t)h->H * h->W;
    cudaStream_t st = (cudaStream_t)stream;
    DVO_CUDA(h, cudaMemcpyAsync(h->stage_bgr + px * 3 * frame_base, bgr_host, px * 3 * n_frames, cudaMemcpyHostToDevice, st));
    DVO_CUDA(h, cudaMemcpyAsync(h->stage_depth + px * frame_base, depth_host, px * sizeof(uint16_t) * n_frames,
                                cudaMemcpyHostToDevice, st));
    return DVO_OK;
}

extern "C" int dvo_build_pyramids_staged(dvo_handle* h, int frame_base, int n_frames, int with_gradients, void* stream) {
    if (!h) return DVO_ERR_INVALID;
    const int rc = stage_ready(h, frame_base, n_frames);
    if (rc != DVO_OK) return rc;
    const size_t px = (size_t)h->H * h->W;
    return build_impl(h, frame_base, h->stage_bgr + px * 3 * frame_base, h->stage_depth + px * frame_base, n_frames,
                      with_gradients, true, true, (cudaStream_t)stream);
}

extern "C" int dvo_build_pyramids_host(dvo_handle* h, int frame_base, const uint8_t* bgr_host,
                                       const uint16_t* depth_host, int n_frames, int with_gradients, void* stream) {
    const int rc = dvo_upload_frames(h, frame_base, bgr_host, depth_host, n_frames, stream);
    if (rc != DVO_OK) return rc;
    return dvo_build_pyramids_staged(h, frame_base, n_frames, with_gradients, stream);
}

extern "C" int dvo_get_pyramid(dvo_handle* h, int slot, int level, uint8_t* gray_dev, uint16_t* depth_dev,
                               float* gx_dev, float* gy_dev, void* stream) {
    if (!h) return DVO_ERR_INVALID;
    if (slot < 0 || slot >= h->max_frames || level < 0 || level >= h->levels)
        return fail(h, DVO_ERR_RANGE, "slot / level out of range");
    DVO_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int w = h->lw[level], hh = h->lh[level], pitch = h->lpitch[level];
    dim3 grid((w + 255) / 256, hh);
    const size_t off = (size_t)slot * h->lplane[level];
    if (gray_dev) unpitch_kernel<uint8_t><<<grid, 256, 0, st>>>(h->gray[level] + off, gray_dev, w, hh, pitch);
    if (depth_dev) unpitch_kernel<uint16_t><<<grid, 256, 0, st>>>(h->depth[level] + off, depth_dev, w, hh, pitch);
    if (gx_dev || gy_dev) unpitch_grad_kernel<<<grid, 256, 0, st>>>(h->rec[level] + off, gx_dev, gy_dev, w, hh, pitch);
    h->launches += (gray_dev != nullptr) + (depth_dev != nullptr) + ((gx_dev || gy_dev) ? 1 : 0);
    DVO_CUDA(h, cudaGetLastError());
    return DVO_OK;
}

extern "C" int dvo_get_point_list(dvo_handle* h, int slot, int level, float* z_dev, int* col_dev, int* row_dev,
                                  uint8_t* intensity_dev, int* n_dev, void* stream) {
    if (!h) return DVO_ERR_INVALID;
    if (slot < 0 || slot >= h->max_frames || level < 0 || level >= h->levels)
        return fail(h, DVO_ERR_RANGE, "slot / level out of range");
    if (!(h->slot_built[slot] & kHasPoints))
        return fail(h, DVO_ERR_STATE, "the slot was not built as a previous frame (with_gradients 0 or 1)");
    DVO_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int cap = h->lw[level] * h->lh[level];
    if (n_dev) DVO_CUDA(h, cudaMemsetAsync(n_dev, 0, 2 * sizeof(int), st));
    point_list_dump_kernel<<<(cap + 127 + 255) / 256, 256, 0, st>>>(h->prec[level] + 2 * (size_t)slot * h->lplane[level],
                                                                    h->pt_tiles[level] + slot, z_dev, col_dev, row_dev,
                                                                    intensity_dev, n_dev, cap);
    h->launches += 1;
    DVO_CUDA(h, cudaGetLastError());
    return DVO_OK;
}

// ---- estimate ----------------------------------------------------------------------------------
static void fill_params(const dvo_handle* h, AlignParams& p) {
    memset(&p, 0, sizeof(p));
    for (int l = 0; l < h->levels; ++l) {
        LevelGeom& g = p.lv[l];
        g.gray = h->gray[l];
        g.depth = h->depth[l];
        g.rec = h->rec[l];
        g.prec = h->prec[l];
        g.pt_tiles = h->pt_tiles[l];
        g.plane = h->lplane[l];
        g.w = h->lw[l];
        g.h = h->lh[l];
        g.pitch = h->lpitch[l];
        g.strips = h->lpitch[l] / kTile;
        g.n_tiles = g.strips * h->lh[l];
        g.h_magic = (unsigned)((1ull << 32) / (unsigned)h->lh[l]) + 1u;
        {   // n_chunks = NW * k with about 75 tiles per chunk of a full point list (align_kernel.cuh, fused_pass): 32
            // chunks at 640x480, one per warp of the tail kernel's clusters.  The chunking fixes the summation order,
            // so it depends on the image size and the launch shape only, never on the batch.  Measured
            // (profiles/r2/kernel_experiments.jsonl, tags chunk*): 60 / 75 / 90 tiles: 18.9 / 18.0 / 18.0 ms on 512
            // pairs, 40.2 / 39.1 / 39.1 ms on 1184.
            const int nw = h->threads / 32;
            static const int target = [] {   // developer knob DVO_TUNE_CHUNK_ROWS (tiles per chunk aimed at)
                const char* e = getenv("DVO_TUNE_CHUNK_ROWS");
                const int v = e ? atoi(e) : 0;
                return (v >= 8 && v <= 2048) ? v : 75;
            }();
            const int dense_tiles = (h->lw[l] * h->lh[l] + kTile - 1) / kTile;
            int k = (dense_tiles + nw * (target / 2)) / (nw * target);
            if (k < 1) k = 1;
            while (k > 1 && nw * k > kMaxChunks) --k;   // the chunk table holds kMaxChunks sums per level
            g.n_chunks = nw * k;
        }
#ifdef DVO_BOUNDS_CHECK
        p.dbg_rec_lo[l] = reinterpret_cast<const char*>(h->rec[l]);
        p.dbg_rec_hi[l] = p.dbg_rec_lo[l] + h->n_rec[l] * sizeof(uint2);
        p.dbg_prec_lo[l] = reinterpret_cast<const char*>(h->prec[l]);
        p.dbg_prec_hi[l] = p.dbg_prec_lo[l] + 2 * h->n_raw[l] * sizeof(float);
        {   // negative control: DVO_DEBUG_SHRINK_ROWS pretends the allocations are that many rows shorter
            static const int shrink = [] { const char* e = getenv("DVO_DEBUG_SHRINK_ROWS"); return e ? atoi(e) : 0; }();
            p.dbg_rec_hi[l] -= (size_t)shrink * h->lpitch[l] * sizeof(uint2);
            p.dbg_prec_hi[l] -= (size_t)shrink * h->lpitch[l] * 2 * sizeof(float);
        }
#endif
        g.fx = h->k4[l][0]; g.fy = h->k4[l][1]; g.cx = h->k4[l][2]; g.cy = h->k4[l][3];
        g.ifx = h->kinv4[l][0]; g.ify = h->kinv4[l][1]; g.icx = h->kinv4[l][2]; g.icy = h->kinv4[l][3];
    }
    p.levels = h->levels;
    p.depth_scale = h->depth_scale;
    p.max_iterations = h->cfg.max_iterations;
    p.max_increased_steps = h->cfg.max_increased_steps;
    p.tolerance = h->cfg.tolerance;
    p.sigma_prior = h->cfg.sigma_prior;
    p.tdist_dof = h->cfg.tdist_dof;
    p.tdist_lambda0 = 1.0f / (h->cfg.tdist_init_sigma * h->cfg.tdist_init_sigma);
    p.tdist_tol = h->cfg.tdist_tolerance;
    p.tdist_max_iter = h->cfg.tdist_max_iterations;
    p.tdist_mean = h->cfg.tdist_mean ? 1 : 0;
    {   // developer knob DVO_TUNE_SPEC_TOL: acceptance bound of the speculated t-distribution weights
        static const float tol = [] {
            const char* e = getenv("DVO_TUNE_SPEC_TOL");
            const double v = e ? atof(e) : 0.0;
            return (v > 0.0 && v <= 1e-2) ? (float)v : (float)kTdSpecTol;
        }();
        p.tdist_spec_tol = tol;
    }
    p.huber_k = h->cfg.huber_k;
    const float s_hi = (float)h->depth_scale;
    p.scale_hi = s_hi;
    p.scale_lo = (float)(h->depth_scale - (double)s_hi);
    p.queue = h->queue;
    // tuning knob (dvo_config.reserved[0]): L1 prefetch distance in rows; 0 = default (2), < 0 = off
    p.prefetch_rows = h->cfg.reserved[0] > 0 ? h->cfg.reserved[0] : (h->cfg.reserved[0] < 0 ? 0 : 2);
    p.prefetch_raw_rows = p.prefetch_rows;
    {   // developer knobs DVO_TUNE_PF (tap rows ahead) / DVO_TUNE_RAW_PF (point-list tiles ahead), 1..32
        static const int pf = [] { const char* e = getenv("DVO_TUNE_PF"); return e ? atoi(e) : 0; }();
        static const int raw = [] { const char* e = getenv("DVO_TUNE_RAW_PF"); return e ? atoi(e) : 0; }();
        if (pf >= 1 && pf <= 32 && p.prefetch_rows > 0) p.prefetch_rows = pf;
        if (raw >= 1 && raw <= 32 && p.prefetch_rows > 0) p.prefetch_raw_rows = raw;
    }
    {   // residual-only passes run further ahead; developer knob DVO_TUNE_RES_PF (rows, 1..32)
        static const int res_rows = [] {
            const char* e = getenv("DVO_TUNE_RES_PF");
            const int v = e ? atoi(e) : 0;
            return (v >= 1 && v <= 32) ? v : 4;
        }();
        p.prefetch_res_rows = res_rows;
    }
    p.depth_weight = h->cfg.depth_weight;
#ifdef DVO_BOUNDS_CHECK
    p.dbg_violations = h->dbg_violations;
#endif
}

extern "C" int dvo_debug_bounds_violations(dvo_handle* h, unsigned long long* count) {
    if (!h || !count) return DVO_ERR_INVALID;
#ifdef DVO_BOUNDS_CHECK
    DVO_CUDA(h, cudaSetDevice(h->device));
    DVO_CUDA(h, cudaDeviceSynchronize());
    DVO_CUDA(h, cudaMemcpy(count, h->dbg_violations, sizeof(*count), cudaMemcpyDeviceToHost));
    return DVO_OK;
#else
    return fail(h, DVO_ERR_STATE, "the library was built without -DDVO_BOUNDS_CHECK");
#endif
}

extern "C" int dvo_estimate(dvo_handle* h, int prev_base, int cur_base, int n_pairs, const float* init_qt_dev,
                            const float* last_qt_dev, float* out_qt_dev, dvo_pair_stats* stats_dev, void* stream) {
    if (!h) return DVO_ERR_INVALID;
    if (!out_qt_dev) return fail(h, DVO_ERR_INVALID, "out_qt is null");
    if (!h->intrinsics_set) return fail(h, DVO_ERR_STATE, "dvo_set_intrinsics must be called first");
    if (n_pairs < 1 || prev_base < 0 || cur_base < 0 || prev_base + n_pairs > h->max_frames ||
        cur_base + n_pairs > h->max_frames)
        return fail(h, DVO_ERR_RANGE, "pair range exceeds the frame slots");
    if (n_pairs > h->max_pairs) return fail(h, DVO_ERR_RANGE, "n_pairs exceeds max_pairs");
    {
        const int rc = check_roles(h, prev_base, cur_base, n_pairs, h->cfg.approximate_image2_gradient != 0);
        if (rc != DVO_OK) return rc;
    }
    DVO_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    AlignParams p;
    fill_params(h, p);
    p.n_pairs = n_pairs;
    p.prev_base = prev_base;
    p.cur_base = cur_base;
    p.init_qt = init_qt_dev;
    p.last_qt = last_qt_dev;
    p.out_qt = out_qt_dev;
    p.stats = stats_dev;
    p.queue = h->queue + 8 * h->queue_next;
    p.dlist = h->dlist + prev_base;
    h->queue_next = (h->queue_next + 1) % kQueueSlots;
    p.ring = h->ring + prev_base;      // disjoint between launches in flight: they use disjoint frame slots
    p.pstate = h->pstate + prev_base;
    p.chunk_sums = h->chunk_sums + (size_t)prev_base * (kMaxChunks * kChunkFloats);
    const int grid = n_pairs < h->grid_max ? n_pairs : h->grid_max;
    {   // time slice of the persistent kernel, in level-0 iterations (developer knob DVO_TUNE_QUANTUM, 0 = run every
        // pair to completion); pointless when every pair has a CTA of its own
        static const int q0 = [] {
            const char* e = getenv("DVO_TUNE_QUANTUM");
            const int v = e ? atoi(e) : 4;
            return v < 0 ? 0 : v;
        }();
        p.quantum_tiles = (n_pairs > grid) ? q0 * p.lv[0].n_tiles : 0;
        // Launches in which every pair has a CTA of its own (n_pairs <= grid) run in time slices too: once a tenth of
        // the pairs are finished, the CTAs that find the queue empty raise the draining flag, the others park their pair
        // at the end of its slice and the tail kernel's clusters finish the rest on all SMs (instead of the launch ending
        // on its longest pair with most SMs idle).  Measured (profiles/r2/kernel_experiments.jsonl, tags rtc / slice_*):
        // 8 pairs 6.6 -> 4.1 ms, 32 pairs 8.7 -> 6.0, 256 pairs 15.1 -> 10.9, 296 pairs at 1920x1080 122 -> 68.
        // Developer knobs: DVO_TUNE_SLICE_MIN = smallest such launch (pairs; 0 = never), DVO_TUNE_DRAIN_PCT.
        static const int slice_min = [] { const char* e = getenv("DVO_TUNE_SLICE_MIN"); return e ? atoi(e) : 2; }();
        static const int drain_pct = [] { const char* e = getenv("DVO_TUNE_DRAIN_PCT"); return e ? atoi(e) : 10; }();
        p.drain_after = 0;
        if (n_pairs <= grid && slice_min > 0 && n_pairs >= slice_min) {
            p.quantum_tiles = q0 * p.lv[0].n_tiles;
            p.drain_after = (int)((long long)n_pairs * drain_pct / 100);
        }
    }
    // tail kernel: clusters of `tail_c` CTAs finish the pairs still running when the persistent kernel's queue runs
    // dry (developer knob DVO_TUNE_TAIL_CLUSTER: 0 = off, 2, 4 or 8)
    static const int tail_c = [] {
        const char* e = getenv("DVO_TUNE_TAIL_CLUSTER");
        const int v = e ? atoi(e) : 8;
        return (v == 2 || v == 4 || v == 8) ? v : 0;
    }();
    align_fn tfn = (p.quantum_tiles > 0 && tail_c > 0) ? get_tail(h) : nullptr;
    p.defer = tfn ? 1 : 0;
    if (!tfn && n_pairs <= grid) p.quantum_tiles = 0;   // no tail kernel for this variant: small launches run to completion
    align_fn fn = get_align(h);
    align_fn cfn = h->cfg.cluster_size > 1 ? get_cluster(h) : nullptr;
    const int ev = (h->ev_last + 1) & 7;
    DVO_CUDA(h, cudaEventRecord(h->ev0[ev], st));
    void* args[] = {&p};
    if (cfn) {
        const int C = h->cfg.cluster_size;
        if (C > 8) DVO_CUDA(h, cudaFuncSetAttribute((const void*)cfn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)n_pairs * (unsigned)C);
        lc.blockDim = dim3(kClusterThreads);
        lc.dynamicSmemBytes = 0;
        lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)C;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        lc.attrs = at;
        lc.numAttrs = 1;
        DVO_CUDA(h, cudaLaunchKernelExC(&lc, (const void*)cfn, args));
    } else {
        work_init_kernel<<<(n_pairs + 255) / 256, 256, 0, st>>>(p.queue, p.ring, p.pstate, n_pairs);
        h->launches += 1;
        DVO_CUDA(h, cudaLaunchKernel((const void*)fn, dim3(grid), dim3(h->threads), args, 0, st));
        if (tfn) {   // at most one deferred pair per CTA of the persistent grid
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3((unsigned)grid * (unsigned)tail_c);
            lc.blockDim = dim3(128);
            lc.dynamicSmemBytes = 0;
            lc.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)tail_c;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            lc.attrs = at;
            lc.numAttrs = 1;
            DVO_CUDA(h, cudaLaunchKernelExC(&lc, (const void*)tfn, args));
            h->launches += 1;
        }
    }
    DVO_CUDA(h, cudaEventRecord(h->ev1[ev], st));
    h->ev_last = ev;
    h->launches += 1;
    return DVO_OK;
}

extern "C" int dvo_estimate_host(dvo_handle* h, int prev_base, int cur_base, int n_pairs, const float* init_qt_host,
                                 const float* last_qt_host, float* out_qt_host, dvo_pair_stats* stats_host,
                                 void* stream) {
    if (!h) return DVO_ERR_INVALID;
    if (!out_qt_host) return fail(h, DVO_ERR_INVALID, "out_qt is null");
    if (n_pairs < 1 || n_pairs > h->max_pairs) return fail(h, DVO_ERR_RANGE, "n_pairs exceeds max_pairs");
    DVO_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t qb = sizeof(float) * 7 * n_pairs;
    if (init_qt_host) DVO_CUDA(h, cudaMemcpyAsync(h->qt_init, init_qt_host, qb, cudaMemcpyHostToDevice, st));
    if (last_qt_host) DVO_CUDA(h, cudaMemcpyAsync(h->qt_last, last_qt_host, qb, cudaMemcpyHostToDevice, st));
    const int rc = dvo_estimate(h, prev_base, cur_base, n_pairs, init_qt_host ? h->qt_init : nullptr,
                                last_qt_host ? h->qt_last : nullptr, h->qt_out, h->stats, st);
    if (rc != DVO_OK) return rc;
    DVO_CUDA(h, cudaMemcpyAsync(out_qt_host, h->qt_out, qb, cudaMemcpyDeviceToHost, st));
    if (stats_host)
        DVO_CUDA(h, cudaMemcpyAsync(stats_host, h->stats, sizeof(dvo_pair_stats) * n_pairs, cudaMemcpyDeviceToHost, st));
    return DVO_OK;
}

extern "C" int dvo_last_estimate_ms(dvo_handle* h, float* ms) {
    if (!h || !ms) return DVO_ERR_INVALID;
    if (h->ev_last < 0) return fail(h, DVO_ERR_STATE, "no estimate has been launched");
    DVO_CUDA(h, cudaSetDevice(h->device));
    DVO_CUDA(h, cudaEventSynchronize(h->ev1[h->ev_last]));
    DVO_CUDA(h, cudaEventElapsedTime(ms, h->ev0[h->ev_last], h->ev1[h->ev_last]));
    return DVO_OK;
}

extern "C" int dvo_residuals_jacobian(dvo_handle* h, int prev_slot, int cur_slot, int level, const float* qt_host,
                                      float* r_dev, float* J_dev, uint8_t* depth_mask_dev, uint8_t* warp_valid_dev,
                                      double* acc_dev, void* stream) {
    if (!h) return DVO_ERR_INVALID;
    if (!qt_host) return fail(h, DVO_ERR_INVALID, "qt is null");
    if (!h->intrinsics_set) return fail(h, DVO_ERR_STATE, "dvo_set_intrinsics must be called first");
    if (prev_slot < 0 || prev_slot >= h->max_frames || cur_slot < 0 || cur_slot >= h->max_frames || level < 0 ||
        level >= h->levels)
        return fail(h, DVO_ERR_RANGE, "slot / level out of range");
    if (!h->slot_built[prev_slot] || !(h->slot_built[cur_slot] & kHasTaps) ||
        (h->cfg.approximate_image2_gradient && !(h->slot_built[prev_slot] & kHasTaps)))
        return fail(h, DVO_ERR_STATE, "the slots were not built for these roles (dvo_build_pyramids, with_gradients)");
    DVO_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    AlignParams p;
    fill_params(h, p);
    DVO_CUDA(h, cudaMemcpyAsync(h->qt_one, qt_host, sizeof(float) * 7, cudaMemcpyHostToDevice, st));
    pose_matrix_kernel<<<1, 1, 0, st>>>(h->qt_one, h->qt_one + 16);
    if (acc_dev) DVO_CUDA(h, cudaMemsetAsync(acc_dev, 0, sizeof(double) * DVO_ACC_TERMS, st));
    dump_fn fn = get_dump(h);
    const int n_tiles = (int)(h->lplane[level] / kTile);
    const float* T12 = h->qt_one + 16;
    float lambda = 0.0f;
    void* args[] = {&p, &level, &prev_slot, &cur_slot, &T12, &lambda, &r_dev, &J_dev, &depth_mask_dev,
                    &warp_valid_dev, &acc_dev};
    DVO_CUDA(h, cudaLaunchKernel((const void*)fn, dim3((n_tiles + 7) / 8), dim3(256), args, 0, st));
    h->launches += 2;
    return DVO_OK;
}

extern "C" int dvo_depth_residuals_jacobian(dvo_handle* h, int prev_slot, int cur_slot, int level, const float* qt_host,
                                            float* rz_dev, float* Jz_dev, uint8_t* valid_dev, double* acc_dev,
                                            void* stream) {
    if (!h) return DVO_ERR_INVALID;
    if (!qt_host) return fail(h, DVO_ERR_INVALID, "qt is null");
    if (!h->intrinsics_set) return fail(h, DVO_ERR_STATE, "dvo_set_intrinsics must be called first");
    if (prev_slot < 0 || prev_slot >= h->max_frames || cur_slot < 0 || cur_slot >= h->max_frames || level < 0 ||
        level >= h->levels)
        return fail(h, DVO_ERR_RANGE, "slot / level out of range");
    if (!h->slot_built[prev_slot] || !(h->slot_built[cur_slot] & kHasTaps) ||
        (h->cfg.approximate_image2_gradient && !(h->slot_built[prev_slot] & kHasTaps)))
        return fail(h, DVO_ERR_STATE, "the slots were not built for these roles (dvo_build_pyramids, with_gradients)");
    DVO_CUDA(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    AlignParams p;
    fill_params(h, p);
    DVO_CUDA(h, cudaMemcpyAsync(h->qt_one, qt_host, sizeof(float) * 7, cudaMemcpyHostToDevice, st));
    pose_matrix_kernel<<<1, 1, 0, st>>>(h->qt_one, h->qt_one + 16);
    if (acc_dev) DVO_CUDA(h, cudaMemsetAsync(acc_dev, 0, sizeof(double) * DVO_ACC_TERMS, st));
    const int n_tiles = (int)(h->lplane[level] / kTile);
    const float* T12 = h->qt_one + 16;
    const dim3 grid((n_tiles + 7) / 8);
    if (h->cfg.oob_mode == DVO_OOB_STRICT)
        depth_dump_kernel<DVO_OOB_STRICT><<<grid, 256, 0, st>>>(p, level, prev_slot, cur_slot, T12, rz_dev, Jz_dev,
                                                                valid_dev, acc_dev);
    else
        depth_dump_kernel<DVO_OOB_INCLUSIVE><<<grid, 256, 0, st>>>(p, level, prev_slot, cur_slot, T12, rz_dev, Jz_dev,
                                                                   valid_dev, acc_dev);
    DVO_CUDA(h, cudaGetLastError());
    h->launches += 2;
    return DVO_OK;
}
