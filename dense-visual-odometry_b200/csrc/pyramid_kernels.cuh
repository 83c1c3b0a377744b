// Image / depth pyramid construction.  All integer, bit-exact.
//
//   gray_clamp_kernel     BGR u8 -> gray u8 (OpenCV fixed point) + far-depth -> 0, both into level 0
//                         (reference: core/base_dense_visual_odometry.py:58-59)
//   median3_down_kernel   3x3 median, replicated border, keep even rows/cols
//                         (reference: utils/image_pyramid.py:19-21, cv2.medianBlur(.,3)[::2, ::2])
//   sobel3_kernel         3x3 Sobel dx/dy, gain 8, replicated border -> packed 8-byte records {gx, gy, I}
//                         (reference: utils/jacobian.py:70-71; layout: rec_pack in align_kernel.cuh); the
//                         intensity rides along so that one 8-byte load per bilinear tap feeds the alignment kernel
//
// Plane layout: every level plane is [frame][h][pitch] with pitch a multiple of 16 elements; padding
// columns stay zero (the planes are cleared once at creation and kernels only write col < w), so a
// padded depth sample is "no depth".
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

// ---- a2 -----------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void cswap(T& a, T& b) {
    const T lo = a < b ? a : b;
    const T hi = a < b ? b : a;
    a = lo;
    b = hi;
}

// Median of nine with the classic 19 compare-exchange network.
template <typename T>
__device__ __forceinline__ T median9(T p0, T p1, T p2, T p3, T p4, T p5, T p6, T p7, T p8) {
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8);
    cswap(p0, p1); cswap(p3, p4); cswap(p6, p7);
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8);
    cswap(p0, p3); cswap(p5, p8); cswap(p4, p7);
    cswap(p3, p6); cswap(p1, p4); cswap(p2, p5);
    cswap(p4, p7); cswap(p4, p2); cswap(p6, p4);
    cswap(p4, p2);
    return p4;
}

// One thread per OUTPUT pixel: only the kept (even, even) medians are computed.
template <typename T>
__global__ void __launch_bounds__(256) median3_down_kernel(const T* __restrict__ src, T* __restrict__ dst, int sw,
                                                           int sh, int spitch, size_t splane, int dw, int dh,
                                                           int dpitch, size_t dplane) {
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y;
    const int frame = blockIdx.z;
    if (ox >= dw) return;
    const int cx = ox * 2, cy = oy * 2;
    const int x0 = max(cx - 1, 0), x2 = min(cx + 1, sw - 1);
    const int y0 = max(cy - 1, 0), y2 = min(cy + 1, sh - 1);
    const T* s = src + (size_t)frame * splane;
    const T* r0 = s + (size_t)y0 * spitch;
    const T* r1 = s + (size_t)cy * spitch;
    const T* r2 = s + (size_t)y2 * spitch;
    using W = int;
    const W m = median9<W>(__ldg(r0 + x0), __ldg(r0 + cx), __ldg(r0 + x2), __ldg(r1 + x0), __ldg(r1 + cx),
                           __ldg(r1 + x2), __ldg(r2 + x0), __ldg(r2 + cx), __ldg(r2 + x2));
    dst[(size_t)frame * dplane + (size_t)oy * dpitch + ox] = (T)m;
}

// ---- a9 -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sobel3_kernel(const uint8_t* __restrict__ gray, uint2* __restrict__ rec,
                                                     int w, int h, int pitch, size_t plane) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int frame = blockIdx.z;
    if (x >= w) return;
    const int xm = max(x - 1, 0), xp = min(x + 1, w - 1);
    const int ym = max(y - 1, 0), yp = min(y + 1, h - 1);
    const uint8_t* s = gray + (size_t)frame * plane;
    const uint8_t* r0 = s + (size_t)ym * pitch;
    const uint8_t* r1 = s + (size_t)y * pitch;
    const uint8_t* r2 = s + (size_t)yp * pitch;
    const int a00 = __ldg(r0 + xm), a01 = __ldg(r0 + x), a02 = __ldg(r0 + xp);
    const int a10 = __ldg(r1 + xm), a11 = __ldg(r1 + x), a12 = __ldg(r1 + xp);
    const int a20 = __ldg(r2 + xm), a21 = __ldg(r2 + x), a22 = __ldg(r2 + xp);
    const int gx = (a02 + 2 * a12 + a22) - (a00 + 2 * a10 + a20);
    const int gy = (a20 + 2 * a21 + a22) - (a00 + 2 * a01 + a02);
    rec[(size_t)frame * plane + (size_t)y * pitch + x] = rec_pack(gx, gy, a11);
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

}  // namespace dvo
