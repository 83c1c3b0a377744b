// L1-resident load bandwidth per SM for the access shapes of align_kernel's tap gathers.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench3 ubench3.cu && ./ubench3
#include <cuda_runtime.h>
#include <cstdio>

template <int BYTES>
__global__ void l1_k(const char* __restrict__ base, float* out, int iters, int lane_stride) {
    // every warp re-reads its own 8 KB window: L1 hits after the first touch
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const char* p = base + (size_t)(blockIdx.x * 8 + warp) * 8192 + lane * lane_stride;
    float s = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const char* q = p + ((j * 1024 + it * 16) & 8191 & ~(BYTES - 1));
            if (BYTES == 16) { float4 v = __ldg((const float4*)q); s += v.x + v.w; }
            if (BYTES == 8) { float2 v = __ldg((const float2*)q); s += v.x + v.y; }
            if (BYTES == 4) { float v = __ldg((const float*)q); s += v; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int BYTES>
__global__ void lds_k(float* out, int iters, int lane_stride) {
    __shared__ __align__(16) char sm[8 * 4096];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 1024; i += blockDim.x) ((float*)sm)[i] = i;
    __syncthreads();
    const char* p = sm + warp * 4096 + lane * lane_stride;
    float s = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const char* q = sm + ((p - sm + j * 512 + it * 16) & (8 * 4096 - 1) & ~(BYTES - 1));
            if (BYTES == 16) { float4 v = *(const float4*)q; s += v.x + v.w; }
            if (BYTES == 8) { float2 v = *(const float2*)q; s += v.x + v.y; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    const int sms = prop.multiProcessorCount;
    char* buf;
    cudaMalloc(&buf, (size_t)sms * 8 * 8192 * 4 + 65536);
    cudaMemset(buf, 0, (size_t)sms * 8 * 8192 * 4 + 65536);
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 4 * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 4096;
#define RUN(NAME, LAUNCH, BYTES_PER_LANE, BLOCKS)                                                            \
    {                                                                                                        \
        LAUNCH;                                                                                              \
        cudaEventRecord(e0);                                                                                 \
        LAUNCH;                                                                                              \
        cudaEventRecord(e1);                                                                                 \
        cudaEventSynchronize(e1);                                                                            \
        float ms;                                                                                            \
        cudaEventElapsedTime(&ms, e0, e1);                                                                   \
        const double instr = (double)(BLOCKS) * 8 * iters * 8;                                               \
        const double clk = ms * 1e-3 * ghz * 1e9;                                                            \
        printf("%-44s %7.3f warp-loads/clk/SM  %7.1f B/clk/SM\n", NAME, instr / clk / sms,                   \
               instr * 32 * (BYTES_PER_LANE) / clk / sms);                                                   \
    }
    for (int b = 1; b <= 4; b *= 2) {
        char n[128];
        snprintf(n, 128, "LDG.128 L1-hit, lanes consecutive, %d warps/SM", 8 * b);
        RUN(n, (l1_k<16><<<sms * b, 256>>>(buf, out, iters, 16)), 16, sms * b);
        snprintf(n, 128, "LDG.64  L1-hit, lanes consecutive, %d warps/SM", 8 * b);
        RUN(n, (l1_k<8><<<sms * b, 256>>>(buf, out, iters, 8)), 8, sms * b);
        snprintf(n, 128, "LDG.32  L1-hit, lanes consecutive, %d warps/SM", 8 * b);
        RUN(n, (l1_k<4><<<sms * b, 256>>>(buf, out, iters, 4)), 4, sms * b);
        snprintf(n, 128, "LDS.128 lanes consecutive, %d warps/SM", 8 * b);
        RUN(n, (lds_k<16><<<sms * b, 256>>>(out, iters, 16)), 16, sms * b);
        snprintf(n, 128, "LDS.64  lanes consecutive, %d warps/SM", 8 * b);
        RUN(n, (lds_k<8><<<sms * b, 256>>>(out, iters, 8)), 8, sms * b);
    }
    return 0;
}
