// Issue-rate microbenchmark for the instruction mix of align_kernel (sm_100a), second edition:
// per-pipe rates of scalar vs packed FP32, ALU/FMA co-issue, conversion tricks, and how throughput
// depends on warps per scheduler and independent chains per warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench2 ubench2.cu && ./ubench2
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

typedef unsigned long long u64;

__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fadd2rm(u64 a, u64 b) { u64 d; asm volatile("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float faddrm(float a, float b) { float d; asm volatile("add.rm.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float rcp(float a) { float d; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a)); return d; }
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b) { unsigned d; asm volatile("xor.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned iadd(unsigned a, unsigned b) { unsigned d; asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned imad(unsigned a, unsigned b, unsigned c) { unsigned d; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned c) { unsigned d; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ float fsel(float a, float b, float c) { float d; asm volatile("{.reg .pred p; setp.gt.f32 p, %3, 0f00000000; selp.f32 %0, %1, %2, p;}" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ unsigned isel(unsigned a, unsigned b, unsigned c) { unsigned d; asm volatile("{.reg .pred p; setp.le.u32 p, %3, %2; selp.u32 %0, %1, %2, p;}" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

enum { FFMA, FFMA2, FMUL2, FADD2, FADD2RM, FADDRM, LOP, IADD, IMAD, PRMT, FSETP_SEL, ISETP_SEL, RCP, SHFL, F2I_FLOOR, I2F,
       MIX_FFMA2_LOP, MIX_FFMA2_FFMA, MIX_FFMA2_IMAD, MIX_FFMA_LOP, MIX_FFMA2_LOP_2to1, MIX_KERNELISH,
       MIX_FFMA2_I2F, MIX_FFMA2_PRMT, MIX_FFMA2_FSEL, MIX_FFMA2_IADD, MIX_FFMA2_RCP, MIX_FFMA2_IMADWIDE, MIX_FFMA2_SHF };

// N independent chains per thread; every op depends only on its own chain.
template <int OP, int N>
__global__ void k(float* out, int iters, float b, float c) {
    float a[N]; u64 p[N]; unsigned u[N];
    const u64 b2 = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b);
    const u64 c2 = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        a[i] = threadIdx.x * 1e-3f + i + 1.0f;
        p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 0.5f);
        u[i] = threadIdx.x * 77u + i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (OP == FFMA) a[i] = ffma(a[i], b, c);
            if (OP == FFMA2) p[i] = ffma2(p[i], b2, c2);
            if (OP == FMUL2) p[i] = fmul2(p[i], b2);
            if (OP == FADD2) p[i] = fadd2(p[i], c2);
            if (OP == FADD2RM) p[i] = fadd2rm(p[i], c2);
            if (OP == FADDRM) a[i] = faddrm(a[i], c);
            if (OP == LOP) u[i] = lop(u[i], 0x9e3779b9u);
            if (OP == IADD) u[i] = iadd(u[i], 0x9e3779b9u + it);
            if (OP == IMAD) u[i] = imad(u[i], 3u, (unsigned)it);
            if (OP == PRMT) u[i] = prmt(u[i], 0x4B000000u, 0x7440u + (it & 1));
            if (OP == FSETP_SEL) a[i] = fsel(a[i], c, a[i]);
            if (OP == ISETP_SEL) u[i] = isel(u[i], (unsigned)it, u[i]);
            if (OP == RCP) a[i] = rcp(a[i]);
            if (OP == SHFL) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1);
            if (OP == F2I_FLOOR) u[i] = (unsigned)__float2int_rd(__uint_as_float(u[i]));
            if (OP == I2F) u[i] = __float_as_uint((float)(int)u[i]);
            if (OP == MIX_FFMA2_LOP) { p[i] = ffma2(p[i], b2, c2); u[i] = lop(u[i], 0x9e3779b9u); }
            if (OP == MIX_FFMA2_FFMA) { p[i] = ffma2(p[i], b2, c2); a[i] = ffma(a[i], b, c); }
            if (OP == MIX_FFMA2_IMAD) { p[i] = ffma2(p[i], b2, c2); u[i] = imad(u[i], 3u, (unsigned)it); }
            if (OP == MIX_FFMA_LOP) { a[i] = ffma(a[i], b, c); u[i] = lop(u[i], 0x9e3779b9u); }
            if (OP == MIX_FFMA2_LOP_2to1) { p[i] = ffma2(p[i], b2, c2); p[i] = ffma2(p[i], c2, b2); u[i] = lop(u[i], 0x9e3779b9u); }
            if (OP == MIX_FFMA2_I2F) { p[i] = ffma2(p[i], b2, c2); u[i] = __float_as_uint((float)(int)(u[i] & 0xffffu)) + it; }
            if (OP == MIX_FFMA2_PRMT) { p[i] = ffma2(p[i], b2, c2); u[i] = prmt(u[i], 0x4B000000u, 0x7440u + (it & 1)); }
            if (OP == MIX_FFMA2_FSEL) { p[i] = ffma2(p[i], b2, c2); a[i] = fsel(a[i], c, a[i]); }
            if (OP == MIX_FFMA2_IADD) { p[i] = ffma2(p[i], b2, c2); u[i] = iadd(u[i], 0x9e3779b9u + it); }
            if (OP == MIX_FFMA2_RCP) { p[i] = ffma2(p[i], b2, c2); a[i] = rcp(a[i]); }
            if (OP == MIX_FFMA2_IMADWIDE) { p[i] = ffma2(p[i], b2, c2); u64 w; asm volatile("mad.wide.u32 %0, %1, 8, %2;" : "=l"(w) : "r"(u[i]), "l"((u64)it)); u[i] = (unsigned)w ^ (unsigned)(w >> 32); }
            if (OP == MIX_FFMA2_SHF) { p[i] = ffma2(p[i], b2, c2); unsigned d; asm volatile("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(u[i]), "r"(u[i]), "r"(it)); u[i] = d; }
            if (OP == MIX_KERNELISH) {  // per "pixel pair": 8 FFMA2/FMUL2, 3 FFMA, 3 ALU, 1 IMAD, 1 sel
                p[i] = ffma2(p[i], b2, c2); p[i] = fmul2(p[i], b2); p[i] = ffma2(p[i], b2, c2); p[i] = fadd2(p[i], c2);
                p[i] = ffma2(p[i], b2, c2); p[i] = fmul2(p[i], b2); p[i] = ffma2(p[i], b2, c2); p[i] = fadd2(p[i], c2);
                a[i] = ffma(a[i], b, c); a[i] = ffma(a[i], b, c); a[i] = ffma(a[i], b, c);
                u[i] = lop(u[i], 0x9e3779b9u); u[i] = iadd(u[i], 5u); u[i] = prmt(u[i], 0x4B000000u, 0x7440u);
                u[i] = imad(u[i], 3u, (unsigned)it); a[i] = fsel(a[i], c, a[i]);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float(u[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static int g_sms;
static double g_ghz;

template <int OP, int N>
void run(const char* name, double instr_per_chain_step, float* out, int warps_per_sm) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 2048;
    const int threads = warps_per_sm >= 8 ? 256 : warps_per_sm * 32;
    const int blocks = g_sms * (warps_per_sm * 32 / threads);
    k<OP, N><<<blocks, threads>>>(out, 16, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    k<OP, N><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)g_sms * warps_per_sm * iters * N * instr_per_chain_step;
    const double per_sm_clk = warp_instr / (ms * 1e-3 * g_ghz * 1e9) / g_sms;
    printf("%-40s warps/SM %2d chains %2d : %6.3f warp-instr/clk/SM\n", name, warps_per_sm, N, per_sm_clk);
}

// L1-resident 16-byte gathers: every lane reads a 16 B record near a per-warp moving base (neighbouring
// lanes -> neighbouring records), 8 independent loads in flight per lane.
__global__ void gather_k(const float4* __restrict__ rec, float* out, int iters, int span, int stride) {
    float s = 0;
    unsigned idx = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 4096 + (threadIdx.x & 31) * stride;
    for (int it = 0; it < iters; ++it) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(rec + ((idx + j * 640u + it * 32u) % span));
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[j].x + v[j].z;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    g_ghz = khz * 1e-6;
    g_sms = p.multiProcessorCount;
    printf("%s, %d SMs, %.3f GHz nominal (rates assume that clock)\n", p.name, g_sms, g_ghz);
    float* out;
    cudaMalloc(&out, sizeof(float) * g_sms * 64 * 32);
#define R(OP, N, IPS, W) run<OP, N>(#OP, IPS, out, W)
    R(FFMA, 8, 1, 32); R(FFMA2, 8, 1, 32); R(FMUL2, 8, 1, 32); R(FADD2, 8, 1, 32); R(FADD2RM, 8, 1, 32); R(FADDRM, 8, 1, 32);
    R(LOP, 8, 1, 32); R(IADD, 8, 1, 32); R(IMAD, 8, 1, 32); R(PRMT, 8, 1, 32); R(FSETP_SEL, 8, 2, 32); R(ISETP_SEL, 8, 2, 32);
    R(RCP, 8, 1, 32); R(SHFL, 8, 1, 32); R(F2I_FLOOR, 8, 1, 32); R(I2F, 8, 1, 32);
    R(MIX_FFMA2_LOP, 8, 2, 32); R(MIX_FFMA2_FFMA, 8, 2, 32); R(MIX_FFMA2_IMAD, 8, 2, 32); R(MIX_FFMA_LOP, 8, 2, 32);
    R(MIX_FFMA2_LOP_2to1, 8, 3, 32); R(MIX_KERNELISH, 4, 16, 32);
    // which pipe does an instruction share with FFMA2?  (2 instructions per chain step; ~3.6 = separate pipes, ~1.9 = same)
    R(MIX_FFMA2_I2F, 8, 2, 32); R(MIX_FFMA2_PRMT, 8, 2, 32); R(MIX_FFMA2_FSEL, 8, 2, 32); R(MIX_FFMA2_IADD, 8, 2, 32);
    R(MIX_FFMA2_RCP, 8, 2, 32); R(MIX_FFMA2_IMADWIDE, 8, 2, 32); R(MIX_FFMA2_SHF, 8, 2, 32);
    // occupancy / ILP sweep for the kernel-like mix and for FFMA2
    R(MIX_KERNELISH, 1, 16, 8); R(MIX_KERNELISH, 2, 16, 8); R(MIX_KERNELISH, 4, 16, 8);
    R(MIX_KERNELISH, 1, 16, 16); R(MIX_KERNELISH, 2, 16, 16); R(MIX_KERNELISH, 4, 16, 16);
    R(MIX_KERNELISH, 1, 16, 4); R(MIX_KERNELISH, 2, 16, 4); R(MIX_KERNELISH, 4, 16, 4);
    R(FFMA2, 1, 1, 8); R(FFMA2, 2, 1, 8); R(FFMA2, 4, 1, 8); R(FFMA2, 8, 1, 8);
    R(FFMA2, 1, 1, 16); R(FFMA2, 2, 1, 16); R(FFMA2, 4, 1, 16);
    // gathers
    {
        const int span = 1 << 20;  // 16 MB of records: L2-resident, L1 hit when lanes/iterations revisit lines
        float4* rec;
        cudaMalloc(&rec, sizeof(float4) * span);
        cudaMemset(rec, 0, sizeof(float4) * span);
        for (int stride = 1; stride <= 4; stride *= 2)
            for (int wps = 8; wps <= 32; wps *= 2) {
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0);
                cudaEventCreate(&e1);
                const int iters = 512, blocks = g_sms * wps / 8;
                gather_k<<<blocks, 256>>>(rec, out, 8, span, stride);
                cudaEventRecord(e0);
                gather_k<<<blocks, 256>>>(rec, out, iters, span, stride);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                const double ldg = (double)blocks * 8 * iters * 8;
                printf("LDG.128 gather lane-stride %d warps/SM %2d : %6.3f warp-LDG/clk/SM  (%.1f B/clk/SM)\n", stride, wps,
                       ldg / (ms * 1e-3 * g_ghz * 1e9) / g_sms, ldg * 512 / (ms * 1e-3 * g_ghz * 1e9) / g_sms);
            }
    }
    return 0;
}
