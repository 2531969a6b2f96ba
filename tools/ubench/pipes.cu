// Microbenchmark: throughput of IMAD.WIDE (32x32+64), DFMA and their mix on one B200 (sm_100a).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

template <int NI, int ND>
__global__ void __launch_bounds__(256) k_mix(u64* out, double* dout, u32 a0, double d0, int iters) {
    u64 acc[8];
    double dac[8];
    u32 x = a0 + threadIdx.x;
    double y = d0 + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = i, dac[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < NI; i++)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(x), "r"((u32)acc[(i + 1) % 8]));
#pragma unroll
            for (int i = 0; i < ND; i++) dac[i] = fma(y, dac[(i + 1) % 8], dac[i]);
        }
    }
    u64 s = 0;
    double ds = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i], ds += dac[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    dout[blockIdx.x * blockDim.x + threadIdx.x] = ds;
}

template <int NI, int ND>
void run(const char* name, u64* out, double* dout) {
    const int iters = 4096, grid = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k_mix<NI, ND><<<grid, 256>>>(out, dout, 12345, 1.000001, 16);
    cudaEventRecord(e0);
    k_mix<NI, ND><<<grid, 256>>>(out, dout, 12345, 1.000001, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warps = (double)grid * 256 / 32, per = (double)iters * 4;
    double imad = warps * per * NI, dfma = warps * per * ND;
    // warp-instructions per ns per SM sub-partition (592 of them)
    printf("%-22s %8.3f ms  IMAD.WIDE %6.3f /ns/SMSP  DFMA %6.3f /ns/SMSP\n", name, ms, imad / (ms * 1e6) / 592, dfma / (ms * 1e6) / 592);
}

int main() {
    u64* out;
    double* dout;
    cudaMalloc(&out, 148 * 8 * 256 * 8), cudaMalloc(&dout, 148 * 8 * 256 * 8);
    run<8, 0>("imad.wide only", out, dout);
    run<0, 8>("dfma only", out, dout);
    run<8, 8>("8 imad + 8 dfma", out, dout);
    run<8, 4>("8 imad + 4 dfma", out, dout);
    run<4, 8>("4 imad + 8 dfma", out, dout);
    run<6, 8>("6 imad + 8 dfma", out, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
