// Microbenchmark: ways to accumulate sum_j y_j * d_j of split-30 operands (the diagonal MAC inner loop) on sm_100a.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mac mac.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
constexpr int NG = 4, TERMS = 32;

__device__ __forceinline__ u64 mulw(u32 a, u32 b) {
    u64 r;
    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}

// acc += p + q with explicit carry chains on the 32-bit halves (keeps ptxas from folding the adds back into IMAD.WIDE)
__device__ __forceinline__ void add2(u64& acc, u64 p, u64 q) {
    asm volatile(
        "{\n\t.reg .u32 l1, h1, l2, h2, al, ah;\n\t"
        "mov.b64 {l1, h1}, %1;\n\tmov.b64 {l2, h2}, %2;\n\tmov.b64 {al, ah}, %0;\n\t"
        "add.cc.u32 al, al, l1;\n\taddc.u32 ah, ah, h1;\n\t"
        "add.cc.u32 al, al, l2;\n\taddc.u32 ah, ah, h2;\n\t"
        "mov.b64 %0, {al, ah};\n\t}"
        : "+l"(acc)
        : "l"(p), "l"(q));
}
// V0: Karatsuba with accumulating wide multiply-adds (what k_pmac_tma does today)
template <int V>
__global__ void __launch_bounds__(128) k(const u64* __restrict__ yin, const u64* __restrict__ din, u64* out, int iters) {
    __shared__ u64 ys[TERMS][128];
    __shared__ u64 ds[NG][TERMS][8];
    for (int j = 0; j < TERMS; j++) ys[j][threadIdx.x] = yin[j * 128 + threadIdx.x];
    for (int e = threadIdx.x; e < NG * TERMS * 8; e += 128) (&ds[0][0][0])[e] = din[e];
    __syncthreads();
    u64 tot = 0;
    const int w = threadIdx.x >> 4;
    for (int it = 0; it < iters; it++) {
        u64 s0[NG], s1[NG], s2[NG];
#pragma unroll
        for (int k = 0; k < NG; k++) s0[k] = s1[k] = s2[k] = 0;
        for (int b0 = 0; b0 < TERMS; b0 += 16) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const u64 ya = ys[b0 + j][threadIdx.x], yb = ys[b0 + j + 1][threadIdx.x];
                const u32 ya0 = (u32)ya, ya1 = (u32)(ya >> 32), yb0 = (u32)yb, yb1 = (u32)(yb >> 32);
#pragma unroll
                for (int k = 0; k < NG; k++) {
                    const u64 da = ds[k][b0 + j][w], db = ds[k][b0 + j + 1][w];
                    const u32 da0 = (u32)da, da1 = (u32)(da >> 32), db0 = (u32)db, db1 = (u32)(db >> 32);
                    if (V == 0) {
                        s0[k] = (u64)ya0 * da0 + s0[k], s2[k] = (u64)ya1 * da1 + s2[k];
                        s1[k] = (u64)(ya0 + ya1) * (da0 + da1) + s1[k];
                        s0[k] = (u64)yb0 * db0 + s0[k], s2[k] = (u64)yb1 * db1 + s2[k];
                        s1[k] = (u64)(yb0 + yb1) * (db0 + db1) + s1[k];
                    } else if (V == 1) {   // schoolbook, non-accumulating multiplies, three-input 64-bit adds
                        s0[k] += mulw(ya0, da0) + mulw(yb0, db0);
                        s1[k] += mulw(ya0, da1) + mulw(ya1, da0);
                        s1[k] += mulw(yb0, db1) + mulw(yb1, db0);
                        s2[k] += mulw(ya1, da1) + mulw(yb1, db1);
                    } else if (V == 3) {   // Karatsuba, explicit carry chains
                        add2(s0[k], mulw(ya0, da0), mulw(yb0, db0));
                        add2(s2[k], mulw(ya1, da1), mulw(yb1, db1));
                        add2(s1[k], mulw(ya0 + ya1, da0 + da1), mulw(yb0 + yb1, db0 + db1));
                    } else {               // Karatsuba, non-accumulating multiplies, three-input adds
                        s0[k] += mulw(ya0, da0) + mulw(yb0, db0);
                        s2[k] += mulw(ya1, da1) + mulw(yb1, db1);
                        s1[k] += mulw(ya0 + ya1, da0 + da1) + mulw(yb0 + yb1, db0 + db1);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NG; k++) tot += s0[k] ^ (s1[k] << 3) ^ (s2[k] << 7);
    }
    out[blockIdx.x * 128 + threadIdx.x] = tot;
}

template <int V>
void run(const char* name, const u64* y, const u64* d, u64* out, double ghz) {
    const int iters = 256, grid = 148 * 12;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k<V><<<grid, 128>>>(y, d, out, 4);
    cudaEventRecord(e0);
    k<V><<<grid, 128>>>(y, d, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warp_terms = (double)grid * 4 * iters * TERMS * NG;
    printf("%-44s %8.3f ms  %6.2f cycles per warp-term per SMSP (at %.2f GHz)\n", name, ms, ms * 1e6 * ghz * 592 / warp_terms, ghz);
}

int main() {
    u64 *y, *d, *out;
    cudaMalloc(&y, TERMS * 128 * 8), cudaMalloc(&d, NG * TERMS * 8 * 8), cudaMalloc(&out, 148 * 12 * 128 * 8);
    cudaMemset(y, 0x15, TERMS * 128 * 8), cudaMemset(d, 0x0b, NG * TERMS * 8 * 8);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz * 1e-6;
    run<0>("karatsuba, mad.wide accumulate (current)", y, d, out, ghz);
    run<1>("schoolbook, mul.wide + 3-input adds", y, d, out, ghz);
    run<2>("karatsuba, mul.wide + 3-input adds", y, d, out, ghz);
    run<3>("karatsuba, mul.wide + explicit carry chains", y, d, out, ghz);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
