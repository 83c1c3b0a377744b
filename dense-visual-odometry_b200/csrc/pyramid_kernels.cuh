// Image / depth pyramid construction.  All integer, bit-exact.
//
//   gray_clamp_kernel     BGR u8 -> gray u8 (OpenCV fixed point) + far-depth -> 0, both into level 0
//                         (reference: core/base_dense_visual_odometry.py:58-59)
//   median3_down_pair_kernel  3x3 median, replicated border, keep even rows/cols (gray and depth level in one launch)
//                         (reference: utils/image_pyramid.py:19-21, cv2.medianBlur(.,3)[::2, ::2])
//   points_kernel         the level's pixels with depth, compacted into the previous-frame POINT LIST the alignment
//                         kernel walks (pt_pack in align_kernel.cuh; camera_model.py:171-226: the masked point cloud,
//                         with z = fl32(float64(d) * scale) of :199-200 hoisted out of the Gauss-Newton loop)
//   sobel3_kernel         3x3 Sobel dx/dy, gain 8, replicated border -> packed 8-byte records {gx, gy, I, depth}
//                         (reference: utils/jacobian.py:70-71; layout: rec_pack in align_kernel.cuh); the
//                         intensity rides along so that one 8-byte load per bilinear tap feeds the alignment kernel
//
// Plane layout: every level plane is [frame][h][pitch] with pitch a multiple of 16 elements; padding
// columns stay as initialised at creation (kernels only write col < w): zero in the gray / depth / tap-record
// planes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dvo {

// ---- a1 -----------------------------------------------------------------------------------------
// One thread converts 4 consecutive pixels of one row.  VEC = row starts are 4-pixel aligned (W % 4 == 0):
// 12 B of BGR are read as three 32-bit words and depth as one 64-bit word.
template <bool VEC, bool HAS_BGR>
__global__ void __launch_bounds__(256) gray_clamp_kernel(const uint8_t* __restrict__ bgr_or_gray,
                                                         uint16_t* __restrict__ depth_io, uint8_t* __restrict__ gray0,
                                                         uint16_t* __restrict__ depth0, int w, int h, int pitch,
                                                         size_t plane, int clamp_thr, int do_clamp) {
    const int gpr = (w + 3) >> 2;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int frame = blockIdx.y;
    if (g >= gpr * h) return;
    const int row = g / gpr;
    const int col = (g - row * gpr) << 2;
    const size_t in_px = ((size_t)frame * h + row) * w + col;
    uint8_t gr[4];
    uint16_t d[4];
    bool changed = false;
    if (VEC) {
        if (HAS_BGR) {
            const uint32_t* p = reinterpret_cast<const uint32_t*>(bgr_or_gray + in_px * 3);
            const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
            const uint32_t b0 = w0 & 255u, g0 = (w0 >> 8) & 255u, r0 = (w0 >> 16) & 255u;
            const uint32_t b1 = w0 >> 24, g1 = w1 & 255u, r1 = (w1 >> 8) & 255u;
            const uint32_t b2 = (w1 >> 16) & 255u, g2 = w1 >> 24, r2 = w2 & 255u;
            const uint32_t b3 = (w2 >> 8) & 255u, g3 = (w2 >> 16) & 255u, r3 = w2 >> 24;
            gr[0] = (uint8_t)((b0 * 3735u + g0 * 19235u + r0 * 9798u + 16384u) >> 15);
            gr[1] = (uint8_t)((b1 * 3735u + g1 * 19235u + r1 * 9798u + 16384u) >> 15);
            gr[2] = (uint8_t)((b2 * 3735u + g2 * 19235u + r2 * 9798u + 16384u) >> 15);
            gr[3] = (uint8_t)((b3 * 3735u + g3 * 19235u + r3 * 9798u + 16384u) >> 15);
        } else {
            const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(bgr_or_gray + in_px));
            gr[0] = v.x; gr[1] = v.y; gr[2] = v.z; gr[3] = v.w;
        }
        const ushort4 dv = *reinterpret_cast<const ushort4*>(depth_io + in_px);
        d[0] = dv.x; d[1] = dv.y; d[2] = dv.z; d[3] = dv.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            gr[k] = 0;
            d[k] = 0;
            if (col + k < w) {
                if (HAS_BGR) {
                    const uint8_t* p = bgr_or_gray + (in_px + k) * 3;
                    gr[k] = (uint8_t)(((uint32_t)p[0] * 3735u + (uint32_t)p[1] * 19235u + (uint32_t)p[2] * 9798u +
                                       16384u) >> 15);
                } else {
                    gr[k] = bgr_or_gray[in_px + k];
                }
                d[k] = depth_io[in_px + k];
            }
        }
    }
    if (do_clamp) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if ((int)d[k] >= clamp_thr) {
                d[k] = 0;
                changed = true;
            }
    }
    const size_t out = (size_t)frame * plane + (size_t)row * pitch + col;
    if (VEC) {
        *reinterpret_cast<uchar4*>(gray0 + out) = make_uchar4(gr[0], gr[1], gr[2], gr[3]);
        *reinterpret_cast<ushort4*>(depth0 + out) = make_ushort4(d[0], d[1], d[2], d[3]);
        if (changed) *reinterpret_cast<ushort4*>(depth_io + in_px) = make_ushort4(d[0], d[1], d[2], d[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (col + k < w) {
                gray0[out + k] = gr[k];
                depth0[out + k] = d[k];
                if (changed) depth_io[in_px + k] = d[k];
            }
    }
}

// Initial state of a point-list buffer: "no depth" points everywhere (tiles of 128 z values, 128 point words).
__global__ void prec_fill_kernel(float* __restrict__ p, size_t n_floats) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_floats; i += stride)
        p[i] = __uint_as_float(((i >> 7) & 1) ? 0u : kPrecNoDepth);
}

// ---- a4 -----------------------------------------------------------------------------------------
// The point list of one level of every frame (layout and order: align_kernel.cuh, "previous-frame point lists").
// One CTA of 32 warps per frame; a SEGMENT is one row of one 128-pixel strip (warp = segment, lane = four columns),
// segments numbered strip-major.  Three phases, no barrier inside any loop:
//   1  every warp counts the pixels with depth of its segments               -> s_cnt[segment]   (shared memory)
//   2  block-wide exclusive scan of s_cnt: the place of every segment's first point in the list
//   3  every warp re-reads its segments (L2 hits) and writes their points
// The last tile is padded with "no depth" points; pt_tiles[frame] receives the number of tiles.
// Dynamic shared memory: (strips * h + 64) ints + 32 KB of staging.
inline size_t points_smem_bytes(int strips, int h) {
    return ((size_t)strips * h + 64) * sizeof(int) + 32 * 256 * sizeof(float);
}
__device__ __forceinline__ int points_count4(uint2 dv) {
    return ((dv.x & 0xffffu) != 0u) + ((dv.x >> 16) != 0u) + ((dv.y & 0xffffu) != 0u) + ((dv.y >> 16) != 0u);
}
__global__ void __launch_bounds__(1024) points_kernel(const uint8_t* __restrict__ gray, const uint16_t* __restrict__ depth,
                                                      float* __restrict__ list, int* __restrict__ pt_tiles,
                                                      double depth_scale, int w, int h, int pitch, size_t plane) {
    extern __shared__ int s_cnt[];   // [n_seg] counts, then exclusive offsets; [n_seg .. n_seg + 31] warp totals of the scan
    const int frame = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint8_t* g8 = gray + (size_t)frame * plane;
    const uint16_t* d16 = depth + (size_t)frame * plane;
    float* out = list + 2u * (size_t)frame * plane;
    const int strips = (w + 127) >> 7;
    const int n_seg = strips * h;
    int* s_warp = s_cnt + n_seg;
    // the padding columns of the depth plane are zero (no depth), so whole 128-column segments are read
    auto seg_elem = [&](int seg) {
        const int s = seg / h, row = seg - s * h;
        return (size_t)row * pitch + (size_t)(s * 128 + 4 * lane);
    };
    // ---- 1: counts (four segments of the warp in flight)
    for (int seg0 = wid; seg0 < n_seg; seg0 += 4 * 32) {
        uint2 dv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int seg = seg0 + 32 * k;
            dv[k] = seg < n_seg ? __ldg(reinterpret_cast<const uint2*>(d16 + seg_elem(seg))) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int seg = seg0 + 32 * k;
            const int c = __reduce_add_sync(0xffffffffu, points_count4(dv[k]));
            if (lane == 0 && seg < n_seg) s_cnt[seg] = c;
        }
    }
    __syncthreads();
    // ---- 2: exclusive scan of s_cnt (each thread a contiguous run, then warp and block level)
    const int per = (n_seg + 1023) >> 10;
    const int i0 = min(tid * per, n_seg), i1 = min(i0 + per, n_seg);
    int run = 0;
    for (int i = i0; i < i1; ++i) run += s_cnt[i];
    int incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int wincl = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, wincl, o);
        if (lane >= o) wincl += v;
    }
    const int before = __shfl_sync(0xffffffffu, wincl, (wid + 31) & 31);
    const int total = __shfl_sync(0xffffffffu, wincl, 31);
    int off = (wid ? before : 0) + incl - run;
    for (int i = i0; i < i1; ++i) {
        const int c = s_cnt[i];
        s_cnt[i] = off;
        off += c;
    }
    __syncthreads();
    // ---- 3: the points.  A segment's points are consecutive in the list; the warp first lines them up in its
    // staging buffer, then stores them tile word by tile word, 32 consecutive words per instruction.
    float* stage_z = reinterpret_cast<float*>(s_warp + 32) + wid * 256;
    unsigned* stage_w = reinterpret_cast<unsigned*>(stage_z) + 128;
    for (int seg0 = wid; seg0 < n_seg; seg0 += 2 * 32) {
        uint2 dv[2];
        unsigned gw[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int seg = seg0 + 32 * k;
            dv[k] = make_uint2(0u, 0u);
            gw[k] = 0u;
            if (seg < n_seg) {
                const size_t e = seg_elem(seg);
                dv[k] = __ldg(reinterpret_cast<const uint2*>(d16 + e));
                gw[k] = __ldg(reinterpret_cast<const unsigned*>(g8 + e));
            }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int seg = seg0 + 32 * k;
            if (seg >= n_seg) break;
            const int s = seg / h, row = seg - s * h;
            const int col = s * 128 + 4 * lane;
            const unsigned dd[4] = {dv[k].x & 0xffffu, dv[k].x >> 16, dv[k].y & 0xffffu, dv[k].y >> 16};
            const int cnt = points_count4(dv[k]);
            int lincl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, lincl, o);
                if (lane >= o) lincl += v;
            }
            const int n_pts = __shfl_sync(0xffffffffu, lincl, 31);
            int q = lincl - cnt;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (dd[j] != 0u) {
                    stage_z[q] = (float)((double)dd[j] * depth_scale);
                    stage_w[q] = pt_pack(col + j, row, (gw[k] >> (8 * j)) & 255u);
                    ++q;
                }
            __syncwarp();
            const int first = s_cnt[seg], last = first + n_pts;   // list positions [first, last)
            for (int tile = first >> 7; tile * 128 < last; ++tile) {
                float* t = out + (size_t)tile * 256u;
#pragma unroll
                for (int i0w = 0; i0w < 128; i0w += 32) {
                    const int i = i0w + lane;                                            // word of the tile's z half
                    const int j = ((i >> 6) << 6) | ((i & 1) << 5) | ((i >> 1) & 31);    // inverse of pt_index
                    const int pos = tile * 128 + j;
                    if (pos >= first && pos < last) {
                        t[i] = stage_z[pos - first];
                        t[i + 128] = __uint_as_float(stage_w[pos - first]);
                    }
                }
            }
            __syncwarp();
        }
    }
    const int n_tiles = (total + 127) >> 7;
    for (int pos = total + tid; pos < n_tiles * 128; pos += 1024) {
        float* t = out + (size_t)(pos >> 7) * 256u + pt_index(pos & 127);
        t[0] = __uint_as_float(kPrecNoDepth);
        t[128] = 0.0f;
    }
    if (tid == 0) pt_tiles[frame] = n_tiles;
}

// ---- a2 -----------------------------------------------------------------------------------------
// Medians run two at a time in the 16-bit halves of a register (VIMNMX.U16x2: one instruction is a packed
// min or max).  u16 depths are the halves themselves; a u8 intensity v travels as the key v * 257 (its byte
// duplicated by the same byte permute that extracts it), which orders like v and carries v in its low byte.
__device__ __forceinline__ void cswap2(unsigned& a, unsigned& b) {
    const unsigned lo = __vminu2(a, b);
    const unsigned hi = __vmaxu2(a, b);
    a = lo;
    b = hi;
}

// Median of nine with the classic 19 compare-exchange network, on two independent sets at once.
__device__ __forceinline__ unsigned median9x2(unsigned p0, unsigned p1, unsigned p2, unsigned p3, unsigned p4, unsigned p5,
                                              unsigned p6, unsigned p7, unsigned p8) {
    cswap2(p1, p2); cswap2(p4, p5); cswap2(p7, p8);
    cswap2(p0, p1); cswap2(p3, p4); cswap2(p6, p7);
    cswap2(p1, p2); cswap2(p4, p5); cswap2(p7, p8);
    cswap2(p0, p3); cswap2(p5, p8); cswap2(p4, p7);
    cswap2(p3, p6); cswap2(p1, p4); cswap2(p2, p5);
    cswap2(p4, p7); cswap2(p4, p2); cswap2(p6, p4);
    cswap2(p4, p2);
    return p4;
}

// FOUR consecutive outputs ox = 4k .. 4k+3 of a row need source columns v_0 .. v_8 = 8k-1 .. 8k+7 (border replicated):
// output j uses v_2j, v_2j+1, v_2j+2.  Outputs (0, 2) and (1, 3) are paired, so one source row contributes the five
// packed words A_c = (v_c, v_c+4), c = 0..4: outputs (0, 2) take A_0 A_1 A_2, outputs (1, 3) take A_2 A_3 A_4.
// Interior threads use one scalar + one 8-element vector load (rows start 16-byte aligned because every pitch is a
// multiple of 64) and five byte permutes; the first and the last thread of a row clamp.
template <typename T>
__device__ __forceinline__ void load_row5x2(const T* __restrict__ row, int k, int sw, unsigned* A) {
    const int c0 = 8 * k;
    if (k > 0 && c0 + 7 < sw) {
        const unsigned v0 = (unsigned)__ldg(row + c0 - 1);
        if (sizeof(T) == 1) {
            const uint2 w = __ldg(reinterpret_cast<const uint2*>(row + c0));   // bytes v1..v4, v5..v8
            A[0] = __byte_perm(v0, w.x, 0x7700);
            A[1] = __byte_perm(w.x, w.y, 0x4400);
            A[2] = __byte_perm(w.x, w.y, 0x5511);
            A[3] = __byte_perm(w.x, w.y, 0x6622);
            A[4] = __byte_perm(w.x, w.y, 0x7733);
        } else {
            const uint4 w = __ldg(reinterpret_cast<const uint4*>(row + c0));   // halves (v1,v2) (v3,v4) (v5,v6) (v7,v8)
            A[0] = __byte_perm(v0, w.y, 0x7610);
            A[1] = __byte_perm(w.x, w.z, 0x5410);
            A[2] = __byte_perm(w.x, w.z, 0x7632);
            A[3] = __byte_perm(w.y, w.w, 0x5410);
            A[4] = __byte_perm(w.y, w.w, 0x7632);
        }
    } else {
        unsigned v[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) v[j] = (unsigned)__ldg(row + min(max(c0 - 1 + j, 0), sw - 1));
#pragma unroll
        for (int c = 0; c < 5; ++c) A[c] = v[c] | (v[c + 4] << 16);
    }
}

// One thread per FOUR output pixels of a row; only the kept (even, even) medians are computed.
// Both planes of a frame in one launch: the gray (u8) and depth (u16) levels share their geometry, so one thread
// produces four gray and four depth outputs, with all six source-row loads issued before the first median; threads
// are numbered linearly over (output row, 4-pixel group) so that the narrow coarse levels still fill their blocks
// (one block per row left 37-84 % of the threads idle there).
__global__ void __launch_bounds__(128) median3_down_pair_kernel(const uint8_t* __restrict__ src8, uint8_t* __restrict__ dst8,
                                                                const uint16_t* __restrict__ src16,
                                                                uint16_t* __restrict__ dst16, int sw, int sh, int spitch,
                                                                size_t splane, int dw, int dh, int dpitch, size_t dplane) {
    const int kpr = (dw + 3) >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kpr * dh) return;
    const int oy = idx / kpr;
    const int k = idx - oy * kpr;
    const int frame = blockIdx.y;
    const int ox = 4 * k;
    const int cy = oy * 2;
    const size_t ra = (size_t)max(cy - 1, 0) * spitch, rb = (size_t)cy * spitch, rc = (size_t)min(cy + 1, sh - 1) * spitch;
    const uint8_t* s8 = src8 + (size_t)frame * splane;
    const uint16_t* s16 = src16 + (size_t)frame * splane;
    unsigned a0[5], a1[5], a2[5], b0[5], b1[5], b2[5];
    load_row5x2(s8 + ra, k, sw, a0);
    load_row5x2(s8 + rb, k, sw, a1);
    load_row5x2(s8 + rc, k, sw, a2);
    load_row5x2(s16 + ra, k, sw, b0);
    load_row5x2(s16 + rb, k, sw, b1);
    load_row5x2(s16 + rc, k, sw, b2);
    // medians of outputs (0, 2) and (1, 3)
    const unsigned m02 = median9x2(a0[0], a0[1], a0[2], a1[0], a1[1], a1[2], a2[0], a2[1], a2[2]);
    const unsigned m13 = median9x2(a0[2], a0[3], a0[4], a1[2], a1[3], a1[4], a2[2], a2[3], a2[4]);
    const unsigned n02 = median9x2(b0[0], b0[1], b0[2], b1[0], b1[1], b1[2], b2[0], b2[1], b2[2]);
    const unsigned n13 = median9x2(b0[2], b0[3], b0[4], b1[2], b1[3], b1[4], b2[2], b2[3], b2[4]);
    const size_t o = (size_t)frame * dplane + (size_t)oy * dpitch + ox;
    if (ox + 3 < dw) {
        *reinterpret_cast<uint32_t*>(dst8 + o) = __byte_perm(m02, m13, 0x6240);   // the low bytes of the four keys
        *reinterpret_cast<uint2*>(dst16 + o) = make_uint2(__byte_perm(n02, n13, 0x5410), __byte_perm(n02, n13, 0x7632));
    } else {
        const unsigned m[4] = {m02 & 255u, m13 & 255u, (m02 >> 16) & 255u, (m13 >> 16) & 255u};
        const unsigned n[4] = {n02 & 0xffffu, n13 & 0xffffu, n02 >> 16, n13 >> 16};
        for (int j = 0; j < 4 && ox + j < dw; ++j) {   // padding columns keep their initial state
            dst8[o + j] = (uint8_t)m[j];
            dst16[o + j] = (uint16_t)n[j];
        }
    }
}

// ---- a9 -----------------------------------------------------------------------------------------
// One thread per EIGHT pixels of a row: columns 8k-1 .. 8k+8 of three rows (one 64-bit load + two bytes each),
// eight packed records out as four 16-byte stores.
__device__ __forceinline__ void load_row10(const uint8_t* __restrict__ row, int x0, int w, int* v) {
    v[0] = (int)__ldg(row + max(x0 - 1, 0));
    if (x0 + 7 < w) {
        const uint2 q = __ldg(reinterpret_cast<const uint2*>(row + x0));
        v[1] = q.x & 255u; v[2] = (q.x >> 8) & 255u; v[3] = (q.x >> 16) & 255u; v[4] = q.x >> 24;
        v[5] = q.y & 255u; v[6] = (q.y >> 8) & 255u; v[7] = (q.y >> 16) & 255u; v[8] = q.y >> 24;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[1 + j] = (int)__ldg(row + min(x0 + j, w - 1));
    }
    v[9] = (int)__ldg(row + min(x0 + 8, w - 1));
}

__global__ void __launch_bounds__(128) sobel3_kernel(const uint8_t* __restrict__ gray, const uint16_t* __restrict__ depth,
                                                     uint2* __restrict__ rec, int w, int h, int pitch, size_t plane) {
    // threads numbered linearly over (row, 8-pixel group): narrow levels still fill their blocks
    const int gpr = (w + 7) >> 3;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= gpr * h) return;
    const int y = idx / gpr;
    const int x0 = 8 * (idx - y * gpr);
    const int frame = blockIdx.y;
    const uint8_t* s = gray + (size_t)frame * plane;
    int a[10], b[10], c[10];
    load_row10(s + (size_t)max(y - 1, 0) * pitch, x0, w, a);
    load_row10(s + (size_t)y * pitch, x0, w, b);
    load_row10(s + (size_t)min(y + 1, h - 1) * pitch, x0, w, c);
    // the pixel's depth rides in the record too (padding columns of the depth plane are zero, and so are the
    // values of a partial group beyond the image: the 16-byte load stays inside the row's pitch)
    const uint4 dw = __ldg(reinterpret_cast<const uint4*>(depth + (size_t)frame * plane + (size_t)y * pitch + x0));
    const unsigned dd[8] = {dw.x & 0xffffu, dw.x >> 16, dw.y & 0xffffu, dw.y >> 16,
                            dw.z & 0xffffu, dw.z >> 16, dw.w & 0xffffu, dw.w >> 16};
    uint2 out[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int gx = (a[j + 2] + 2 * b[j + 2] + c[j + 2]) - (a[j] + 2 * b[j] + c[j]);
        const int gy = (c[j] + 2 * c[j + 1] + c[j + 2]) - (a[j] + 2 * a[j + 1] + a[j + 2]);
        out[j] = rec_pack(gx, gy, b[j + 1], dd[j]);
    }
    uint2* d = rec + (size_t)frame * plane + (size_t)y * pitch + x0;
    if (x0 + 7 < w) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            reinterpret_cast<uint4*>(d)[j] = make_uint4(out[2 * j].x, out[2 * j].y, out[2 * j + 1].x, out[2 * j + 1].y);
    } else {
        for (int j = 0; j < 8 && x0 + j < w; ++j) d[j] = out[j];   // padding columns stay zero
    }
}

// Dense read-back of one level plane (drops the pitch padding); used by dvo_get_pyramid.
template <typename T>
__global__ void unpitch_kernel(const T* __restrict__ src, T* __restrict__ dst, int w, int h, int pitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x < w) dst[(size_t)y * w + x] = src[(size_t)y * pitch + x];
}
__global__ void unpitch_grad_kernel(const uint2* __restrict__ src, float* __restrict__ gx, float* __restrict__ gy,
                                    int w, int h, int pitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x < w) {
        int a, b, c;
        rec_unpack(src[(size_t)y * pitch + x], a, b, c);
        if (gx) gx[(size_t)y * w + x] = (float)a;
        if (gy) gy[(size_t)y * w + x] = (float)b;
    }
}

// Read-back of one point list in list order (dvo_get_point_list); the padding of the last tile is dropped.
__global__ void point_list_dump_kernel(const float* __restrict__ list, const int* __restrict__ pt_tiles,
                                       float* __restrict__ z_out, int* __restrict__ col_out, int* __restrict__ row_out,
                                       uint8_t* __restrict__ i_out, int* __restrict__ n_out, int capacity) {
    const int n_tiles = *pt_tiles;
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos == 0 && n_out) n_out[1] = n_tiles;
    if (pos >= n_tiles * 128) return;
    const float* t = list + (size_t)(pos >> 7) * 256u + pt_index(pos & 127);
    const float z = t[0];
    if (__float_as_uint(z) == kPrecNoDepth) return;   // padding (only at the end of the list)
    if (n_out) atomicAdd(n_out, 1);
    if (pos >= capacity) return;
    const unsigned w = __float_as_uint(t[128]);
    if (z_out) z_out[pos] = z;
    if (col_out) col_out[pos] = pt_col(w);
    if (row_out) row_out[pos] = pt_row(w);
    if (i_out) i_out[pos] = (uint8_t)((w >> 11) & 255u);
}

}  // namespace dvo
