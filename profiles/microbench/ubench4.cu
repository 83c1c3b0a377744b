// mma.sync.m16n8k8 TF32 on B200: (1) issue rate per SM, alone and mixed with FFMA2, for the twelve-tile
// "diagonal" accumulation of align_kernel (Accum<1>); (2) numerics of the operand rounding: the tensor core
// ignores the low 13 mantissa bits, so A is rounded away from zero (+0x1FFF) and B left to truncate.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench4 ubench4.cu && ./ubench4
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ void mma_tf32(float* d, unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// MIX = number of independent FFMA2 per HMMA issued next to it
template <int MIX>
__global__ void rate_k(float* out, int iters) {
    float d[12][4];
    for (int i = 0; i < 12; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    unsigned a = threadIdx.x * 2654435761u, b = a ^ 0x9e3779b9u;
    float2 f[8];
    for (int i = 0; i < 8; ++i) f[i] = make_float2(1.0f + i, 2.0f + i);
    const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            mma_tf32(d[i], a, b, a ^ 1u, b ^ 1u, a, b);
#pragma unroll
            for (int k = 0; k < MIX; ++k) f[(i * MIX + k) & 7] = __ffma2_rn(f[(i * MIX + k) & 7], m, c);
        }
    }
    float s = 0.f;
    for (int i = 0; i < 12; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    for (int i = 0; i < 8; ++i) s += f[i].x + f[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// numerics: every lane holds n pixel pairs (x_a, x_b), (y_a, y_b); tile diag = sum x y over the lane group
__global__ void num_k(const float* x, const float* y, float* diag, int n, int round_a) {
    const int lane = threadIdx.x & 31;
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < n; ++i) {
        const float xa = x[(2 * i) * 32 + lane], xb = x[(2 * i + 1) * 32 + lane];
        const float ya = y[(2 * i) * 32 + lane], yb = y[(2 * i + 1) * 32 + lane];
        const unsigned add = round_a == 1 ? 0x1FFFu : (round_a == 2 ? 0x1000u : 0u);
        mma_tf32(d, __float_as_uint(xa) + add, 0u, __float_as_uint(xb) + add, 0u, __float_as_uint(ya), __float_as_uint(yb));
    }
    const int g = lane >> 2, t = lane & 3;
    if (t == (g >> 1)) diag[g] = (g & 1) ? d[1] : d[0];
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    const int sms = prop.multiProcessorCount;
    printf("%s, %d SMs, %.3f GHz nominal\n", prop.name, sms, ghz);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 1024 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
#define RUN(MIX, WARPS)                                                                                         \
    {                                                                                                           \
        rate_k<MIX><<<sms, 32 * WARPS>>>(out, 100);                                                             \
        cudaEventRecord(e0);                                                                                    \
        rate_k<MIX><<<sms, 32 * WARPS>>>(out, iters);                                                           \
        cudaEventRecord(e1);                                                                                    \
        cudaEventSynchronize(e1);                                                                               \
        float ms;                                                                                               \
        cudaEventElapsedTime(&ms, e0, e1);                                                                      \
        const double clk = ms * 1e-3 * ghz * 1e9;                                                               \
        const double hm = (double)iters * 12 * WARPS;                                                           \
        printf("HMMA.1688.F32.TF32 + %d FFMA2 each, %2d warps/SM: %.3f HMMA/clk/SM (%.2f clk per HMMA per SM), "  \
               "%.3f FFMA2/clk/SM\n", MIX, WARPS, hm / clk, clk / hm, hm * MIX / clk);                          \
    }
    RUN(0, 4) RUN(0, 8) RUN(0, 16)
    RUN(2, 8) RUN(4, 8) RUN(6, 8) RUN(8, 8)
    // numerics
    const int n = 4096;   // pixel pairs per lane -> 8 lanes x 2 x n products per diagonal entry
    std::vector<float> hx(64 * n), hy(64 * n);
    srand(1);
    for (size_t i = 0; i < hx.size(); ++i) {
        hx[i] = (float)((rand() / (double)RAND_MAX - 0.3) * 500.0);
        hy[i] = (float)((rand() / (double)RAND_MAX - 0.3) * 40.0);
    }
    float *dx, *dy, *dd;
    cudaMalloc(&dx, hx.size() * 4);
    cudaMalloc(&dy, hy.size() * 4);
    cudaMalloc(&dd, 8 * 4);
    cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dy, hy.data(), hy.size() * 4, cudaMemcpyHostToDevice);
    const char* names[3] = {"A, B truncated by the hardware", "A rounded away (+0x1FFF), B truncated", "A rounded to nearest (+0x1000), B truncated"};
    for (int mode = 0; mode < 3; ++mode) {
        num_k<<<1, 32>>>(dx, dy, dd, n, mode);
        float hd[8];
        cudaMemcpy(hd, dd, 32, cudaMemcpyDeviceToHost);
        double worst = 0, mean = 0;
        for (int g = 0; g < 8; ++g) {
            double ref = 0, refabs = 0;
            for (int i = 0; i < 2 * n; ++i)
                for (int t = 0; t < 4; ++t) {
                    const int lane = 4 * g + t;
                    ref += (double)hx[i * 32 + lane] * hy[i * 32 + lane];
                    refabs += fabs((double)hx[i * 32 + lane] * hy[i * 32 + lane]);
                }
            const double rel = (hd[g] - ref) / refabs;
            mean += rel / 8;
            if (fabs(rel) > fabs(worst)) worst = rel;
        }
        printf("sum of %d products, %s: mean relative error %+.2e, worst %+.2e (relative to sum |x y|)\n", 8 * 2 * n,
               names[mode], mean, worst);
    }
    return 0;
}
