// Microbenchmark: issue cost of the integer multiply flavours on one B200 (sm_100a), cycles per warp instruction
// per SM sub-partition.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o imad imad.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

enum { WIDE_ACC, WIDE_NOACC, LO32, HI32, LO_PLUS_HI, IADD64, FFMA, WIDE_PLUS_FFMA, WIDE_PLUS_IADD, LO32_PLUS_FFMA };

template <int MODE>
__global__ void __launch_bounds__(256) k(u64* out, u32 a0, int iters) {
    u64 acc[8];
    u32 lo[8], hi[8];
    float f[8];
    u32 x = a0 + threadIdx.x;
    float fx = 1.0f + 1e-7f * threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = i + a0, lo[i] = i * 3 + a0, hi[i] = i * 7 + a0, f[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int n = (i + 1) % 8;
                if (MODE == WIDE_ACC || MODE == WIDE_PLUS_FFMA || MODE == WIDE_PLUS_IADD)
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(x), "r"((u32)acc[n]));
                if (MODE == WIDE_NOACC) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[i]) : "r"(x), "r"((u32)acc[n]));
                if (MODE == LO32 || MODE == LO_PLUS_HI || MODE == LO32_PLUS_FFMA)
                    asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(x), "r"(lo[n]));
                if (MODE == HI32 || MODE == LO_PLUS_HI) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[i]) : "r"(x), "r"(hi[n]));
                if (MODE == IADD64 || MODE == WIDE_PLUS_IADD) asm volatile("add.u64 %0, %0, %1;" : "+l"(acc[i]) : "l"(acc[n]));
                if (MODE == FFMA || MODE == WIDE_PLUS_FFMA || MODE == LO32_PLUS_FFMA)
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[i]) : "f"(fx), "f"(f[n]));
            }
        }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i] + lo[i] + hi[i] + (u64)f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, u64* out, double ghz) {
    const int iters = 2048, grid = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, 12345, 64);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, 12345, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double slots = (double)grid * 256 / 32 * iters * 4 * 8;   // loop-body slots executed (each = the MODE's instruction group)
    printf("%-28s %8.3f ms  %6.2f cycles per slot per SMSP (at %.2f GHz)\n", name, ms, ms * 1e6 * ghz * 592 / slots, ghz);
}

int main() {
    u64* out;
    cudaMalloc(&out, 148 * 8 * 256 * 8);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz * 1e-6;
    run<WIDE_ACC>("mad.wide.u32 (acc)", out, ghz);
    run<WIDE_NOACC>("mul.wide.u32", out, ghz);
    run<LO32>("mad.lo.u32", out, ghz);
    run<HI32>("mad.hi.u32", out, ghz);
    run<LO_PLUS_HI>("mad.lo + mad.hi", out, ghz);
    run<IADD64>("add.u64", out, ghz);
    run<FFMA>("fma.f32", out, ghz);
    run<WIDE_PLUS_FFMA>("mad.wide + fma.f32", out, ghz);
    run<WIDE_PLUS_IADD>("mad.wide + add.u64", out, ghz);
    run<LO32_PLUS_FFMA>("mad.lo + fma.f32", out, ghz);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
