// Microbenchmark 2: Karatsuba split-30 MAC with U products summed per accumulator update (sm_100a).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
constexpr int NG = 4, TERMS = 20;

__device__ __forceinline__ u64 mulw(u32 a, u32 b) {
    u64 r;
    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}

__device__ __forceinline__ void madw(u64& acc, u32 a, u32 b) { asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b)); }

// acc += p1 + p2 written as 32-bit carry chains: ptxas keeps the products as non-accumulating IMAD.WIDE (2 cycles on the
// multiplier pipe) and emits IADD3, IADD3, IADD3.X (three-input, two carry-ins) on the ALU pipe
__device__ __forceinline__ void add3(u64& acc, u64 p1, u64 p2) {
    asm("{ .reg .u32 al, ah, bl, bh, cl, ch;\n\t"
        "mov.b64 {al, ah}, %0; mov.b64 {bl, bh}, %1; mov.b64 {cl, ch}, %2;\n\t"
        "add.cc.u32 al, al, bl; addc.u32 ah, ah, bh; add.cc.u32 al, al, cl; addc.u32 ah, ah, ch;\n\t"
        "mov.b64 %0, {al, ah}; }"
        : "+l"(acc)
        : "l"(p1), "l"(p2));
}
__device__ __forceinline__ u64 mulnv(u32 a, u32 b) {
    u64 r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}

// MODE 0: every product pair through add3;  MODE 1: middle (Karatsuba) sum through accumulating IMAD.WIDE, outer sums add3
template <int MODE>
__global__ void __launch_bounds__(128) k2(const u64* __restrict__ yin, const u64* __restrict__ din, u64* out, int iters) {
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
#pragma unroll
        for (int j = 0; j < TERMS; j += 2) {
            const u64 ya = ys[j][threadIdx.x] + (u64)it, yb = ys[j + 1][threadIdx.x] + (u64)it;
            const u32 ya0 = (u32)ya, ya1 = (u32)(ya >> 32), yas = ya0 + ya1;
            const u32 yb0 = (u32)yb, yb1 = (u32)(yb >> 32), ybs = yb0 + yb1;
#pragma unroll
            for (int k = 0; k < NG; k++) {
                const u64 da = ds[k][j][w], db = ds[k][j + 1][w];
                const u32 da0 = (u32)da, da1 = (u32)(da >> 32), db0 = (u32)db, db1 = (u32)(db >> 32);
                add3(s0[k], mulnv(ya0, da0), mulnv(yb0, db0));
                add3(s2[k], mulnv(ya1, da1), mulnv(yb1, db1));
                if (MODE == 0) add3(s1[k], mulnv(yas, da0 + da1), mulnv(ybs, db0 + db1));
                else madw(s1[k], yas, da0 + da1), madw(s1[k], ybs, db0 + db1);
            }
        }
#pragma unroll
        for (int k = 0; k < NG; k++) tot += s0[k] ^ (s1[k] << 3) ^ (s2[k] << 7);
    }
    out[blockIdx.x * 128 + threadIdx.x] = tot;
}

template <int U, bool SCHOOL>
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
#pragma unroll
        for (int j = 0; j < TERMS; j += U) {
            u32 y0[U], y1[U], ysum[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const u64 y = ys[j + u][threadIdx.x] + (u64)it;
                y0[u] = (u32)y, y1[u] = (u32)(y >> 32), ysum[u] = y0[u] + y1[u];
            }
#pragma unroll
            for (int k = 0; k < NG; k++) {
                u64 a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const u64 d = ds[k][j + u][w];
                    const u32 d0 = (u32)d, d1 = (u32)(d >> 32);
                    if (U == 5) {   // the engine's mac_split: non-volatile mad.wide, accumulators updated directly
                        madw(s0[k], y0[u], d0), madw(s2[k], y1[u], d1), madw(s1[k], ysum[u], d0 + d1);
                        continue;
                    }
                    a0 += mulw(y0[u], d0), a2 += mulw(y1[u], d1);
                    if (SCHOOL) a1 += mulw(y0[u], d1) + mulw(y1[u], d0);
                    else a1 += mulw(ysum[u], d0 + d1);
                }
                s0[k] += a0, s1[k] += a1, s2[k] += a2;
            }
        }
#pragma unroll
        for (int k = 0; k < NG; k++) tot += s0[k] ^ (s1[k] << 3) ^ (s2[k] << 7);
    }
    out[blockIdx.x * 128 + threadIdx.x] = tot;
}

template <int MODE>
void run2(const char* name, const u64* y, const u64* d, u64* out, double ghz) {
    const int iters = 256, grid = 148 * 12;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k2<MODE><<<grid, 128>>>(y, d, out, 4);
    cudaEventRecord(e0);
    k2<MODE><<<grid, 128>>>(y, d, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warp_terms = (double)grid * 4 * iters * TERMS * NG;
    printf("%-46s %8.3f ms  %6.2f cycles per warp-term per SMSP (at %.2f GHz)\n", name, ms, ms * 1e6 * ghz * 592 / warp_terms, ghz);
}

template <int U, bool SCHOOL>
void run(const char* name, const u64* y, const u64* d, u64* out, double ghz) {
    const int iters = 256, grid = 148 * 12;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k<U, SCHOOL><<<grid, 128>>>(y, d, out, 4);
    cudaEventRecord(e0);
    k<U, SCHOOL><<<grid, 128>>>(y, d, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warp_terms = (double)grid * 4 * iters * TERMS * NG;
    printf("%-46s %8.3f ms  %6.2f cycles per warp-term per SMSP (at %.2f GHz)\n", name, ms, ms * 1e6 * ghz * 592 / warp_terms, ghz);
}

int main() {
    u64 *y, *d, *out;
    cudaMalloc(&y, TERMS * 128 * 8), cudaMalloc(&d, NG * TERMS * 8 * 8), cudaMalloc(&out, 148 * 12 * 128 * 8);
    cudaMemset(y, 0x15, TERMS * 128 * 8), cudaMemset(d, 0x0b, NG * TERMS * 8 * 8);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz * 1e-6;
    run<1, false>("karatsuba, mul.wide + add (fused by ptxas)", y, d, out, ghz);
    run<5, false>("karatsuba, mad.wide asm (engine mac_split)", y, d, out, ghz);
    run2<0>("karatsuba, term pairs: mul.wide + 3-input adds", y, d, out, ghz);
    run2<1>("karatsuba, pairs: outer sums split, middle fused", y, d, out, ghz);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
