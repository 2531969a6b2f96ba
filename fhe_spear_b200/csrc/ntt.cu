// ntt.cu -- 64-bit negacyclic NTT / INTT over RNS limbs for sm_100a.
//
// Same transform as the oracle (oracle/spear_oracle.c ntt_fwd_n / ntt_inv_n): Cooley-Tukey,
// natural order in, bit-reversed order out, twiddles psi^{bitrev(i)}; inverse is Gentleman-Sande.
// Harvey lazy butterflies keep values in [0,4q) (forward) / [0,2q) (inverse); outputs are fully
// reduced, so results are bit-identical to the oracle's eager arithmetic.
//
// Decomposition for n = 2^logn: the first sA = logn - sB stages act on "columns" (stride n >> sA)
// and are done by pass A on shared-memory tiles of 2^sA rows x 16 columns (128-byte row segments,
// coalesced); the last sB = min(logn, 8) stages act on contiguous 2^sB-element chunks (pass B,
// 2048 elements per CTA).  Each element crosses L2 twice per transform; twiddles are (w, shoup(w))
// pairs fetched with one 16-byte load.
#include "engine.h"

namespace {

constexpr int TPB = 256;
constexpr int COLS = 16;        // pass A tile width
constexpr int B_ELEMS = 2048;   // pass B elements per CTA

__device__ __forceinline__ void ct_butterfly(u64& x, u64& y, ulonglong2 w, u64 q, u64 q2) {
    // x, y in [0,4q) -> x + w*y, x - w*y in [0,4q)
    u64 u = x >= q2 ? x - q2 : x;
    u64 t = mul_shoup_lazy(y, w.x, w.y, q);
    x = u + t;
    y = u - t + q2;
}
__device__ __forceinline__ void gs_butterfly(u64& x, u64& y, ulonglong2 w, u64 q, u64 q2) {
    // x, y in [0,2q) -> x + y, (x - y)*w in [0,2q)
    u64 s = x + y;
    u64 d = x - y + q2;
    x = s >= q2 ? s - q2 : s;
    y = mul_shoup_lazy(d, w.x, w.y, q);
}

// ---- forward, pass A: stages 0..sA-1 on a [2^sA][COLS] tile --------------------------------
__global__ void __launch_bounds__(TPB) ntt_fwd_a(u64* __restrict__ data, RowMap rm, NttTab tb, int N, int n,
                                                  int sA, int skip_alpha) {
    extern __shared__ u64 sm[];
    const int row = blockIdx.y;
    const int limb = rm.limb(row);
    if (skip_alpha && limb < rm.L && limb / skip_alpha == row / rm.rpp) return;
    const u64 q = tb.q[limb], q2 = q << 1;
    const ulonglong2* __restrict__ tw = tb.psi + (size_t)limb * N;
    u64* base = data + (size_t)row * n;
    const int R = 1 << sA, S = n >> sA, c0 = blockIdx.x * COLS;
    for (int e = threadIdx.x; e < R * COLS; e += TPB) sm[e] = base[(size_t)(e / COLS) * S + c0 + (e % COLS)];
    __syncthreads();
    for (int s = 0; s < sA; s++) {
        const int m = 1 << s, tr = R >> (s + 1);
        for (int bf = threadIdx.x; bf < (R / 2) * COLS; bf += TPB) {
            int c = bf % COLS, k = bf / COLS;
            int i = k / tr, kk = k - i * tr;
            int r0 = 2 * i * tr + kk;
            u64 x = sm[r0 * COLS + c], y = sm[(r0 + tr) * COLS + c];
            ct_butterfly(x, y, tw[m + i], q, q2);
            sm[r0 * COLS + c] = x;
            sm[(r0 + tr) * COLS + c] = y;
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < R * COLS; e += TPB) base[(size_t)(e / COLS) * S + c0 + (e % COLS)] = sm[e];
}

// ---- forward, pass B: stages sA..logn-1 on contiguous chunks of M = 2^sB -------------------
__global__ void __launch_bounds__(TPB) ntt_fwd_b(u64* __restrict__ data, RowMap rm, NttTab tb, int N, int n,
                                                  int sA, int sB, int skip_alpha) {
    extern __shared__ u64 sm[];
    const int row = blockIdx.y;
    const int limb = rm.limb(row);
    if (skip_alpha && limb < rm.L && limb / skip_alpha == row / rm.rpp) return;
    const u64 q = tb.q[limb], q2 = q << 1;
    const ulonglong2* __restrict__ tw = tb.psi + (size_t)limb * N;
    const int M = 1 << sB;
    const int elems = n < B_ELEMS ? n : B_ELEMS;
    u64* base = data + (size_t)row * n + (size_t)blockIdx.x * elems;
    const int gc0 = blockIdx.x * (elems >> sB);   // first global chunk of this CTA
    for (int e = threadIdx.x; e < elems; e += TPB) sm[e] = base[e];
    __syncthreads();
    for (int s = 0; s < sB; s++) {
        const int m = 1 << (sA + s), t = M >> (s + 1);
        for (int bf = threadIdx.x; bf < elems / 2; bf += TPB) {
            int ch = bf >> (sB - 1), b = bf & (M / 2 - 1);
            int i = b / t, kk = b - i * t;
            int x0 = ch * M + 2 * i * t + kk;
            u64 x = sm[x0], y = sm[x0 + t];
            ct_butterfly(x, y, tw[m + ((gc0 + ch) << s) + i], q, q2);
            sm[x0] = x;
            sm[x0 + t] = y;
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < elems; e += TPB) {
        u64 v = sm[e];
        v = v >= q2 ? v - q2 : v;
        base[e] = v >= q ? v - q : v;
    }
}

// ---- inverse, pass B: stages with t = 1 .. M/2 on contiguous chunks ------------------------
__global__ void __launch_bounds__(TPB) ntt_inv_b(u64* __restrict__ data, RowMap rm, NttTab tb, int N, int n,
                                                  int sA, int sB, int logn) {
    extern __shared__ u64 sm[];
    const int row = blockIdx.y;
    const int limb = rm.limb(row);
    const u64 q = tb.q[limb], q2 = q << 1;
    const ulonglong2* __restrict__ tw = tb.ipsi + (size_t)limb * N;
    const int M = 1 << sB;
    const int elems = n < B_ELEMS ? n : B_ELEMS;
    u64* base = data + (size_t)row * n + (size_t)blockIdx.x * elems;
    const int gc0 = blockIdx.x * (elems >> sB);
    for (int e = threadIdx.x; e < elems; e += TPB) sm[e] = base[e];
    __syncthreads();
    // oracle loop: m = n, n/2, ...; h = m/2 groups, gap t doubles from 1.  Stage u (0..sB-1): t = 2^u,
    // h = n >> (u+1); group index of element x in global chunk gc: gc*(M/(2t)) + local
    for (int u = 0; u < sB; u++) {
        const int t = 1 << u, h = n >> (u + 1);
        for (int bf = threadIdx.x; bf < elems / 2; bf += TPB) {
            int ch = bf >> (sB - 1), b = bf & (M / 2 - 1);
            int i = b >> u, kk = b & (t - 1);
            int x0 = ch * M + 2 * i * t + kk;
            u64 x = sm[x0], y = sm[x0 + t];
            gs_butterfly(x, y, tw[h + (gc0 + ch) * (M >> (u + 1)) + i], q, q2);
            sm[x0] = x;
            sm[x0 + t] = y;
        }
        __syncthreads();
    }
    if (sA == 0) {
        ulonglong2 ninv = tb.invn[limb * 17 + logn];
        for (int e = threadIdx.x; e < elems; e += TPB) base[e] = mul_shoup(sm[e], ninv.x, ninv.y, q);
    } else {
        for (int e = threadIdx.x; e < elems; e += TPB) base[e] = sm[e];
    }
}

// ---- inverse, pass A: remaining sA stages on column tiles, then n^-1 ------------------------
__global__ void __launch_bounds__(TPB) ntt_inv_a(u64* __restrict__ data, RowMap rm, NttTab tb, int N, int n,
                                                  int sA, int logn) {
    extern __shared__ u64 sm[];
    const int row = blockIdx.y;
    const int limb = rm.limb(row);
    const u64 q = tb.q[limb], q2 = q << 1;
    const ulonglong2* __restrict__ tw = tb.ipsi + (size_t)limb * N;
    u64* base = data + (size_t)row * n;
    const int R = 1 << sA, S = n >> sA, c0 = blockIdx.x * COLS;
    for (int e = threadIdx.x; e < R * COLS; e += TPB) sm[e] = base[(size_t)(e / COLS) * S + c0 + (e % COLS)];
    __syncthreads();
    // remaining stages: row gap tr = 1, 2, ..., R/2 ; h = R / (2*tr) groups
    for (int u = 0; u < sA; u++) {
        const int tr = 1 << u, h = R >> (u + 1);
        for (int bf = threadIdx.x; bf < (R / 2) * COLS; bf += TPB) {
            int c = bf % COLS, k = bf / COLS;
            int i = k >> u, kk = k & (tr - 1);
            int r0 = 2 * i * tr + kk;
            u64 x = sm[r0 * COLS + c], y = sm[(r0 + tr) * COLS + c];
            gs_butterfly(x, y, tw[h + i], q, q2);
            sm[r0 * COLS + c] = x;
            sm[(r0 + tr) * COLS + c] = y;
        }
        __syncthreads();
    }
    ulonglong2 ninv = tb.invn[limb * 17 + logn];
    for (int e = threadIdx.x; e < R * COLS; e += TPB)
        base[(size_t)(e / COLS) * S + c0 + (e % COLS)] = mul_shoup(sm[e], ninv.x, ninv.y, q);
}

inline void split(int logn, int& sA, int& sB) {
    sB = logn < 8 ? logn : 8;
    sA = logn - sB;
}

}  // namespace

void ntt_forward(const Ctx* c, u64* data, int rows, RowMap rm, int n, cudaStream_t s, int skip_alpha) {
    int logn = 0;
    while ((1 << logn) < n) logn++;
    REQUIRE((1 << logn) == n && n <= c->N && logn <= 16 && n >= 2, "ntt: bad size %d", n);
    if (rows == 0) return;
    int sA, sB;
    split(logn, sA, sB);
    NttTab tb = c->ntttab();
    ProfScope ps(c, PROF_NTT, s);
    if (sA > 0) {
        dim3 grid((n >> sA) / COLS, rows);
        LAUNCH(ntt_fwd_a, grid, TPB, sizeof(u64) * COLS << sA, s)(data, rm, tb, c->N, n, sA, skip_alpha);
    }
    int elems = n < B_ELEMS ? n : B_ELEMS;
    dim3 grid(n / elems, rows);
    LAUNCH(ntt_fwd_b, grid, TPB, sizeof(u64) * elems, s)(data, rm, tb, c->N, n, sA, sB, skip_alpha);
    CUDA_CHECK(cudaGetLastError());
}

void ntt_inverse(const Ctx* c, u64* data, int rows, RowMap rm, int n, cudaStream_t s) {
    int logn = 0;
    while ((1 << logn) < n) logn++;
    REQUIRE((1 << logn) == n && n <= c->N && logn <= 16 && n >= 2, "intt: bad size %d", n);
    if (rows == 0) return;
    int sA, sB;
    split(logn, sA, sB);
    NttTab tb = c->ntttab();
    ProfScope ps(c, PROF_NTT, s);
    int elems = n < B_ELEMS ? n : B_ELEMS;
    dim3 grid(n / elems, rows);
    LAUNCH(ntt_inv_b, grid, TPB, sizeof(u64) * elems, s)(data, rm, tb, c->N, n, sA, sB, logn);
    if (sA > 0) {
        dim3 grida((n >> sA) / COLS, rows);
        LAUNCH(ntt_inv_a, grida, TPB, sizeof(u64) * COLS << sA, s)(data, rm, tb, c->N, n, sA, logn);
    }
    CUDA_CHECK(cudaGetLastError());
}
