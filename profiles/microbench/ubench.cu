// Instruction-throughput microbenchmark for the op mix of the alignment kernel (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cuda_runtime.h>
#include <cstdio>

constexpr int ITERS = 4096;

template <int OP>
__global__ void __launch_bounds__(256) k(float* out, float b, float c, int n) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
    float2 p[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = make_float2(a[2 * i], a[2 * i + 1]);
    const float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (OP == 0) a[i] = __fmaf_rn(a[i], b, c);
            if (OP == 2) a[i] = __fmul_rn(a[i], b);
            if (OP == 3) a[i] = floorf(a[i] * b);                      // FMUL + FRND
            if (OP == 4) a[i] = (float)__float2int_rd(a[i]) + c;       // F2I + I2F + FADD
            if (OP == 5) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a[i])); a[i] = r + c; }
            if (OP == 6) a[i] = __int_as_float(0x4B000000u | (__float_as_uint(a[i]) & 0xffffu)) - 8388608.0f;  // LOP3+FADD
            if (OP == 7) a[i] = (float)(__float_as_uint(a[i]) & 0xffu) + c;                                   // I2F.U8
            if (OP == 8) a[i] = (a[i] > b) ? a[i] * c : a[i] + c;     // FSETP + select
        }
        if (OP == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = __ffma2_rn(p[i], b2, c2);
        }
        if (OP == 9) {   // mixed: 8 FFMA2 + 8 scalar int ops (issue-slot sharing)
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = __ffma2_rn(p[i], b2, c2);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __int_as_float((__float_as_int(a[i]) * 3) ^ it);
        }
        if (OP == 10) {  // mixed: 16 FFMA + 8 scalar int ops
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i].x = __fmaf_rn(p[i].x, b, c); p[i].y = __fmaf_rn(p[i].y, b, c); }
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = __int_as_float((__float_as_int(a[i]) * 3) ^ it);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += p[i].x + p[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, double ops_per_iter, float* out, int sms, double ghz) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sms * 8;
    k<OP><<<blocks, 256>>>(out, 1.0001f, 0.5f, 16);
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(out, 1.0001f, 0.5f, ITERS);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)blocks * 8 * ITERS * ops_per_iter;
    const double per_sm_clk = warp_instr / (ms * 1e-3 * ghz * 1e9) / sms;
    printf("%-34s %8.3f ms  %6.3f warp-instr/clk/SM (%5.1f lanes/clk/SM)\n", name, ms, per_sm_clk, per_sm_clk * 32);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal (rates assume that clock)\n", p.name, p.multiProcessorCount, ghz);
    float* out;
    cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 8 * 256);
    run<0>("FFMA x16", 16, out, p.multiProcessorCount, ghz);
    run<1>("FFMA2 x8 (=16 fma/lane)", 8, out, p.multiProcessorCount, ghz);
    run<2>("FMUL x16", 16, out, p.multiProcessorCount, ghz);
    run<3>("FMUL+FRND.FLOOR x16", 32, out, p.multiProcessorCount, ghz);
    run<4>("F2I+I2F+FADD x16", 48, out, p.multiProcessorCount, ghz);
    run<5>("MUFU.RCP+FADD x16", 32, out, p.multiProcessorCount, ghz);
    run<6>("LOP3(x2)+FADD x16", 48, out, p.multiProcessorCount, ghz);
    run<7>("LOP3+I2F.U8+FADD x16", 48, out, p.multiProcessorCount, ghz);
    run<8>("FSETP+FMUL+FADD+SEL x16", 64, out, p.multiProcessorCount, ghz);
    run<9>("8 FFMA2 + 16 int (IMAD+LOP3)", 24, out, p.multiProcessorCount, ghz);
    run<10>("16 FFMA + 16 int (IMAD+LOP3)", 32, out, p.multiProcessorCount, ghz);
    return 0;
}
