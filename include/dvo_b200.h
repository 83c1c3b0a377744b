/*
 * dvo_b200.h — C ABI of the B200-native photometric-alignment hot path.
 *
 * The reference (pfontana96/dense-visual-odometry) is pure Python and has no FFI; its plugin
 * boundary is the backend contract of `BaseRobustDVO` (four hooks) behind `robust_dvo_factory` /
 * `get_dvo`.  Each entry point below names the reference interface it replaces (paths relative to
 * /root/reference/src/dense_visual_odometry/).  INTEGRATION.md shows the ctypes binding a
 * maintainer would add on the reference side.
 *
 * Conventions
 *  - Plain C types only.  Every function returns 0 on success or a negative dvo_status; it never
 *    throws.  dvo_last_error() gives a message owned by the handle (or a static string if the
 *    handle is NULL).
 *  - Pointers named *_dev are device pointers on the handle's GPU, *_host are host pointers
 *    (pinned memory makes the copies asynchronous).  `stream` is a cudaStream_t passed as void*
 *    (NULL = the legacy default stream).  Calls only enqueue work unless stated otherwise.
 *  - A handle is bound to one device and is not thread-safe: one handle per GPU/stream.
 *  - Frames live in handle-owned "frame slots" 0..max_frames-1 (all pyramid levels, plus what the two roles
 *    of a frame need: metric depth + intensity planes as PREVIOUS frame, packed gradient + intensity records as
 *    CURRENT frame).  A pair p of an estimate call aligns slot prev_base+p (previous frame, provides
 *    intensity + depth) against slot cur_base+p (current frame, provides intensity + gradients).
 *    Independent pairs: prev_base = 0, cur_base = B.  A sequence: prev_base = 0, cur_base = 1.
 *  - Poses cross the boundary the way the reference's `Se3` stores them
 *    (utils/lie_algebra/special_euclidean_group.py:16-27): 7 floats [qw qx qy qz tx ty tz], the
 *    quaternion NOT renormalised, mapping previous-camera points into the current camera.
 */
#ifndef DVO_B200_H
#define DVO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVO_MAX_LEVELS 8
#define DVO_ACC_TERMS 29 /* 21 upper-triangular H entries, 6 J^T W r, sum w r^2, count */

typedef enum dvo_status {
    DVO_OK = 0,
    DVO_ERR_INVALID = -1, /* bad argument */
    DVO_ERR_CUDA = -2,    /* a CUDA runtime call failed; see dvo_last_error */
    DVO_ERR_RANGE = -3,   /* frame slot / pair / level out of range */
    DVO_ERR_STATE = -4    /* call sequence error, e.g. intrinsics not set */
} dvo_status;

typedef enum dvo_weights {
    DVO_W_NONE = 0,      /* reference default, use_weighter=False (base_robust_dvo.py:182-184) */
    DVO_W_TDIST_REF = 1, /* reference TDistributionWeighter as written (weighter/t_weighter.py) */
    DVO_W_HUBER = 2,     /* extension, not in the reference: Huber weights, fixed threshold huber_k (intensity units) */
    DVO_W_HUBER_MAD = 3  /* extension: Huber threshold huber_k * 1.4826 * median|r| re-estimated every iteration;
                          * huber_k is then the tuning constant (1.345) */
} dvo_weights;

typedef enum dvo_oob_mode {
    DVO_OOB_INCLUSIVE = 0, /* valid iff 0 <= x <= W-1 and 0 <= y <= H-1 (passes the reference tests) */
    DVO_OOB_STRICT = 1     /* valid iff floor(x)+1 < W and floor(y)+1 < H (the docstring's rule) */
} dvo_oob_mode;

/* Estimator options: the constructor arguments of BaseRobustDVO (base_robust_dvo.py:34-83),
 * BaseDenseVisualOdometry (base_dense_visual_odometry.py:25-45) and TDistributionWeighter
 * (weighter/t_weighter.py:11-19). */
typedef struct dvo_config {
    int32_t max_iterations;      /* default 100 */
    int32_t max_increased_steps; /* max_increased_steps_allowed, default 0 */
    float tolerance;             /* default 1e-6 */
    float sigma_prior;           /* `sigma`; <= 0 disables the motion prior (default) */
    int32_t weights;             /* dvo_weights */
    int32_t oob_mode;            /* dvo_oob_mode */
    float tdist_dof;             /* 5 */
    float tdist_init_sigma;      /* 5 */
    float tdist_tolerance;       /* 1e-3 */
    int32_t tdist_max_iterations; /* 50 */
    float huber_k;               /* threshold in intensity units (extension) */
    float max_distance;          /* depth clamp in metres, default 5 */
    int32_t threads_per_block;   /* 0 = library default */
    int32_t blocks_per_sm;       /* 0 = library default */
    int32_t approximate_image2_gradient; /* base_robust_dvo.py:34-83 kwarg (cpu_...py:60-77): image Jacobian from the
                                          * PREVIOUS frame's Sobel gradients at the unwarped pixel; frames used as
                                          * "previous" must then be built with_gradients != 0. default 0 */
    int32_t cluster_size;        /* 0/1 = one CTA per pair (throughput); 2, 4, 8 or 16 = one thread-block cluster per
                                  * pair (latency of single pairs and short batches); -1 = the largest of those the
                                  * device can co-schedule (16 needs the non-portable cluster size: 16 on a B200, else 8);
                                  * ignored with the Huber/MAD weights */
    int32_t tdist_mean;          /* extension, with DVO_W_TDIST_REF only: 1 = the textbook t-distribution scale
                                  * (MEAN of the weighted squared residuals) instead of the reference's sum (default 0) */
    int32_t use_depth_residual;  /* extension, not in the reference: 1 = add the depth (geometric) residual
                                  * r_Z = Z2(w(x)) - [T P]_z to the normal equations with weight depth_weight; available
                                  * with DVO_W_NONE / DVO_W_HUBER and approximate_image2_gradient = 0.
                                  * default 0 */
    float depth_weight;          /* lambda_Z: a depth residual of 1/sqrt(lambda_Z) metres weighs like one grey level.
                                  * default 2500 (2 cm) */
    int32_t reserved[1];         /* tuning: [0] L1 prefetch distance in rows (0 default, < 0 off, at most 32) */
} dvo_config;

/* Per-pair statistics written by dvo_estimate (index = pyramid level). 128 bytes. */
typedef struct dvo_pair_stats {
    int32_t iters[DVO_MAX_LEVELS];   /* Gauss-Newton iterations run at the level */
    int32_t n_valid[DVO_MAX_LEVELS]; /* residual count at the level's last iteration */
    float err[DVO_MAX_LEVELS];       /* error value at the level's last iteration */
    int32_t flags;                   /* bit0 non-finite error, bit1 singular H, bit2 max_iterations hit */
    int32_t reserved[7];
} dvo_pair_stats;

typedef struct dvo_handle dvo_handle;

/* Fills cfg with the reference defaults. */
void dvo_default_config(dvo_config* cfg);

/* Replaces the estimator constructor (RobustDVOGPU.__init__, gpu_robust_dense_visual_odometry.py:16-47:
 * fixed resolution, everything preallocated).  Allocates pyramids for max_frames frame slots and
 * state for max_pairs pairs per estimate call.  height and width at most 2048 (DVO_ERR_INVALID beyond). */
int dvo_create(dvo_handle** out, int device, int height, int width, int levels, int max_frames, int max_pairs,
               const dvo_config* cfg);
int dvo_destroy(dvo_handle* h);
const char* dvo_last_error(const dvo_handle* h);

/* Replaces RGBDCameraModel.__init__/at (camera_model.py:28-79): level-0 pinhole intrinsics and the
 * depth scale; per-level K is derived inside exactly as `at` does. */
int dvo_set_intrinsics(dvo_handle* h, float fx, float fy, float cx, float cy, double depth_scale);

/* Replaces BaseDenseVisualOdometry.step's preprocessing (base_dense_visual_odometry.py:58-59: BGR->gray,
 * far depth -> 0 IN PLACE) followed by _build_pyramids (cpu_...py:44-52 -> image_pyramid.py:19-54) and
 * _setup's Sobel planes (cpu_...py:58 -> jacobian.py:70-71) for n_frames frames stored in slots
 * frame_base..frame_base+n_frames-1.  bgr: [n,H,W,3] u8, depth: [n,H,W] u16.  with_gradients names the role(s)
 * the frames will play: 0 = previous frames only (point lists, no gradient planes), 1 = both roles (a sequence: every frame is
 * first "current", then "previous"), 2 = current frames only (gradient planes, no point lists).
 * dvo_set_intrinsics must have been called: the depth scale enters the previous-frame point lists
 * (camera_model.py:199-200, z = depth * depth_scale, is evaluated here, once per frame, not per iteration).
 * The handle remembers what every slot was built for: dvo_estimate and the dense hooks return DVO_ERR_STATE when a slot
 * is used in a role it was not built for. */
/* At most 65535 frames per call. */
int dvo_build_pyramids(dvo_handle* h, int frame_base, const uint8_t* bgr_dev, uint16_t* depth_dev, int n_frames,
                       int with_gradients, void* stream);
/* Same from a gray image, no clamp: the backend hook `_build_pyramids(gray_image, depth_image)`
 * (base_robust_dvo.py:119-125). gray: [n,H,W] u8. */
int dvo_build_pyramids_gray(dvo_handle* h, int frame_base, const uint8_t* gray_dev, const uint16_t* depth_dev,
                            int n_frames, int with_gradients, void* stream);
/* Host-buffer variant of dvo_build_pyramids: copies the frames to a handle-owned staging area
 * (asynchronous if the host memory is pinned) and builds.  The host depth is NOT written back; the
 * caller applies the reference's in-place clamp with dvo_depth_clamp_threshold if it needs it. */
int dvo_build_pyramids_host(dvo_handle* h, int frame_base, const uint8_t* bgr_host, const uint16_t* depth_host,
                            int n_frames, int with_gradients, void* stream);
/* The two halves of dvo_build_pyramids_host, for callers that keep the copy engine busy on a dedicated stream
 * while other streams build and estimate (order them with events): upload = the two host->device copies into
 * the staging area of slots frame_base.., staged = gray/clamp/pyramids/gradients from that staging area. */
int dvo_upload_frames(dvo_handle* h, int frame_base, const uint8_t* bgr_host, const uint16_t* depth_host, int n_frames,
                      void* stream);
int dvo_build_pyramids_staged(dvo_handle* h, int frame_base, int n_frames, int with_gradients, void* stream);
/* Smallest digital number d with (double)d * depth_scale > max_distance (65536 if none). */
int dvo_depth_clamp_threshold(const dvo_handle* h, int* threshold);

/* Replaces BaseRobustDVO._step (base_robust_dvo.py:137-236): the whole coarse-to-fine Gauss-Newton
 * estimate for n_pairs independent pairs in ONE kernel launch, no host synchronisation.
 *   init_qt_dev  [n,7] initial guesses or NULL (identity)
 *   last_qt_dev  [n,7] previous estimates for the sigma prior or NULL (identity)
 *   out_qt_dev   [n,7] estimates
 *   stats_dev    [n] dvo_pair_stats or NULL */
int dvo_estimate(dvo_handle* h, int prev_base, int cur_base, int n_pairs, const float* init_qt_dev,
                 const float* last_qt_dev, float* out_qt_dev, dvo_pair_stats* stats_dev, void* stream);
/* Calls on DIFFERENT streams may be in flight together (a caller pipelining host->device copies of the
 * next pairs behind the estimate of the previous ones), with disjoint frame slots / output ranges per call; the
 * only per-launch state is one of 256 work counters, so at most 256 launches may be pending at a time. */
/* Same, results copied to host memory (pinned => asynchronous); waits for nothing. */
int dvo_estimate_host(dvo_handle* h, int prev_base, int cur_base, int n_pairs, const float* init_qt_host,
                      const float* last_qt_host, float* out_qt_host, dvo_pair_stats* stats_host, void* stream);

/* Replaces the backend hook compute_residuals_and_jacobian (base_robust_dvo.py:91-117,
 * cpu_...py:134-200) in dense ("dump") form for one pair at one level and one pose:
 *   r_dev [H_l*W_l] f32, J_dev [H_l*W_l,6] f32, depth_mask_dev / warp_valid_dev [H_l*W_l] u8 (any may be NULL);
 *   acc_dev [DVO_ACC_TERMS] f64 = the fused reduction of the same pass (H upper triangle row-major,
 *   J^T W r, sum w r^2, count), produced by the same per-pixel code as dvo_estimate. */
int dvo_residuals_jacobian(dvo_handle* h, int prev_slot, int cur_slot, int level, const float* qt_host,
                           float* r_dev, float* J_dev, uint8_t* depth_mask_dev, uint8_t* warp_valid_dev,
                           double* acc_dev, void* stream);

/* Dense form of the depth-residual extension (dvo_config.use_depth_residual; no reference counterpart, defined in
 * align_kernel.cuh / oracle depth_residuals_and_jacobian) for one pair at one level and one pose:
 *   rz_dev [H_l*W_l] f32 metres (NaN where undefined), Jz_dev [H_l*W_l,6] f32, valid_dev [H_l*W_l] u8 (any may be
 *   NULL); acc_dev [DVO_ACC_TERMS] f64 = the term's contribution to the normal equations, depth_weight included,
 *   [28] = number of depth residuals.  Works whether or not the handle was created with use_depth_residual. */
int dvo_depth_residuals_jacobian(dvo_handle* h, int prev_slot, int cur_slot, int level, const float* qt_host,
                                 float* rz_dev, float* Jz_dev, uint8_t* valid_dev, double* acc_dev, void* stream);

/* Reads back one pyramid level of a frame slot (ImagePyramid.at, image_pyramid.py:60-65, and the Sobel
 * planes of jacobian.py:70-71).  Outputs are dense [H_l,W_l]; any may be NULL. */
int dvo_get_pyramid(dvo_handle* h, int slot, int level, uint8_t* gray_dev, uint16_t* depth_dev, float* gx_dev,
                    float* gy_dev, void* stream);
int dvo_level_shape(const dvo_handle* h, int level, int* height, int* width);
/* Reads back the point list of one level of a frame slot built with with_gradients 0 or 1: the pixels with depth, the
 * form in which the alignment kernel consumes a PREVIOUS frame (the masked point cloud of RGBDCameraModel.deproject,
 * camera_model.py:171-226, with z = fl32(float64(d) * depth_scale) of :199-200).  Order: 128-pixel-wide column strips
 * left to right, row-major inside a strip.  z_dev f32, col_dev / row_dev i32, intensity_dev u8: [H_l*W_l] each, the
 * first n entries are written (any may be NULL); n_dev i32[2] = {n = number of points, number of 128-point tiles}. */
int dvo_get_point_list(dvo_handle* h, int slot, int level, float* z_dev, int* col_dev, int* row_dev,
                       uint8_t* intensity_dev, int* n_dev, void* stream);
/* Per-level intrinsics fx, fy, cx, cy as the kernels use them (camera_model.py:62-79). */
int dvo_level_intrinsics(const dvo_handle* h, int level, float* k4);

/* Number of kernel launches this handle has issued (for bench.py's gpu_launches). */
long long dvo_launch_count(const dvo_handle* h);

/* Timing of the most recently LAUNCHED dvo_estimate kernel, measured with CUDA events recorded around that launch
 * on its stream (every launch has its own pair of events).  Synchronises on the end event.  With several
 * launches in flight on different streams the interval includes time the kernel shared the GPU with the others. */
int dvo_last_estimate_ms(dvo_handle* h, float* ms);

/* Debug builds only (nvcc -DDVO_BOUNDS_CHECK; tools/sanitize_cases.py): the alignment kernel tests every load and
 * prefetch address of its streaming pass against the extents of the handle's allocations; this returns the number of
 * misses since dvo_create (after a device synchronisation).  DVO_ERR_STATE in a normal build. */
int dvo_debug_bounds_violations(dvo_handle* h, unsigned long long* count);

#ifdef __cplusplus
}
#endif
#endif /* DVO_B200_H */
