// Microbenchmark: moving eight 64-bit values per thread between the lanes of a warp -- the data exchange between
// butterfly rounds of the NTT pass B -- through XOR-swizzled shared memory (what ntt_core.cuh does: 8 x STS.64,
// __syncwarp, 8 x LDS.64) against a register transpose by warp shuffles (three exchange stages of four 64-bit values:
// 24 SHFL.BFLY of 32-bit halves plus the selects).  BASELINE north_star mentions "warp-shuffle butterflies"; DESIGN.md section 5 cites this.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o exchange exchange.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;

__device__ __forceinline__ int swz(int x) { return x ^ (((x >> 4) & 7) | ((x >> 2) & 8)); }

template <int MODE>
__global__ void __launch_bounds__(128) k(u64* out, int iters) {
    __shared__ u64 sm[4][256];
    const int warp = threadIdx.x >> 5, j = threadIdx.x & 31;
    u64* s = sm[warp];
    u64 v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = (u64)threadIdx.x * 8 + i + blockIdx.x;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {   // layout (j + 32k) -> (8j + i) through shared memory
#pragma unroll
            for (int k2 = 0; k2 < 8; k2++) s[swz(j + 32 * k2)] = v[k2];
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = s[swz(8 * j + i)] + 1;
            __syncwarp();
        } else {           // a register transpose by shuffles: three exchange stages, each lane swaps four of its eight
                           // values with a partner lane (4 x 64-bit = 8 SHFL.BFLY) and re-sorts them with selects
#pragma unroll
            for (int st = 0; st < 3; st++) {
                const int d = 4 << st;              // partner distance: lane bits 2, 3, 4
                const bool up = (j & d) != 0;
                const int gap = 1 << st;            // register pairs (r, r + gap)
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    if (r & gap) continue;
                    const u64 send = up ? v[r] : v[r + gap];
                    const u64 recv = __shfl_xor_sync(0xffffffffu, send, d);
                    if (up) v[r] = recv; else v[r + gap] = recv;
                }
            }
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] += 1;
        }
    }
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, u64* out) {
    const int iters = 512, grid = 148 * 16;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k<MODE><<<grid, 128>>>(out, 8);
    cudaEventRecord(e0);
    k<MODE><<<grid, 128>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double rounds = (double)grid * 4 * iters;   // warp-rounds
    printf("%-44s %8.3f ms  %7.1f cycles per warp exchange of 8 x 64-bit per lane, per SMSP\n", name, ms,
           ms * 1e-3 * khz * 1e3 * 592 / rounds);
}

int main() {
    u64* out;
    cudaMalloc(&out, (size_t)148 * 16 * 128 * 8);
    run<0>("swizzled shared memory (8 STS.64 + 8 LDS.64)", out);
    run<1>("warp shuffles (3-stage register transpose)", out);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
