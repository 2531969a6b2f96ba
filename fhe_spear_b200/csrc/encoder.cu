// encoder.cu -- CKKS encode / decode on the device (canonical embedding with generator 5).
// Mirrors oracle/spear_oracle.c orc_encode / orc_decode operation by operation: every double
// product and sum is rounded separately (__dmul_rn / __dadd_rn, no FMA contraction) and the
// root table comes from the same libm calls, so coefficients are bit-identical to the oracle's.
// Replaces PhantomCKKSEncoder::{encode,decode} (reference gpu/phantom_binding.cu:138-156) and the
// fork-only encode_*_vector_batch (reference scripts/bootstrap_generation.py:382,423).
#include <cstdlib>

#include "engine.h"
#include "ntt_core.cuh"
#include "ops.h"

namespace {

constexpr int TPB = 256;

__device__ __forceinline__ u32 pow5_mod(u32 j, u32 mask) {
    u32 r = 1, g = 5;
    for (; j; j >>= 1, g = (g * g) & mask)
        if (j & 1) r = (r * g) & mask;
    return r;
}
__device__ __forceinline__ u32 brev_n(u32 x, int bits) { return __brev(x) >> (32 - bits); }

// W[v][idx_j] = z_j, W[v][idxc_j] = conj(z_j)
__global__ void k_enc_scatter(const double2* __restrict__ vals, double2* __restrict__ W, int count, int n, int logn) {
    size_t total = (size_t)count * (n / 2);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        u32 j = (u32)(e % (n / 2));
        size_t v = e / (n / 2);
        u32 m = 2u * n, pos = pow5_mod(j, m - 1);
        double2 z = vals[e];
        double2* w = W + v * n;
        w[brev_n((pos - 1) >> 1, logn)] = z;
        w[brev_n((m - pos - 1) >> 1, logn)] = make_double2(z.x, -z.y);
    }
}

// one Gentleman-Sande stage of the inverse embedding; h groups, gap t; twiddle conj(zeta[h+i])
__global__ void k_fft_inv_stage(double2* __restrict__ W, const double2* __restrict__ zeta, int count, int n, int h, int t) {
    size_t total = (size_t)count * (n / 2);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        u32 b = (u32)(e % (n / 2));
        size_t v = e / (n / 2);
        u32 i = b / t, kk = b - i * t;
        double2* a = W + v * n + 2 * i * t + kk;
        double2 w = zeta[h + i], U = a[0], V = a[t];
        double wr = w.x, wi = -w.y;
        a[0] = make_double2(__dadd_rn(U.x, V.x), __dadd_rn(U.y, V.y));
        double dr = __dsub_rn(U.x, V.x), di = __dsub_rn(U.y, V.y);
        a[t] = make_double2(__dsub_rn(__dmul_rn(dr, wr), __dmul_rn(di, wi)),
                            __dadd_rn(__dmul_rn(dr, wi), __dmul_rn(di, wr)));
    }
}
// one Cooley-Tukey stage of the forward embedding; m groups, gap t; twiddle zeta[m+i]
__global__ void k_fft_fwd_stage(double2* __restrict__ W, const double2* __restrict__ zeta, int count, int n, int m, int t) {
    size_t total = (size_t)count * (n / 2);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        u32 b = (u32)(e % (n / 2));
        size_t v = e / (n / 2);
        u32 i = b / t, kk = b - i * t;
        double2* a = W + v * n + 2 * i * t + kk;
        double2 w = zeta[m + i], U = a[0], X = a[t];
        double vr = __dsub_rn(__dmul_rn(X.x, w.x), __dmul_rn(X.y, w.y));
        double vi = __dadd_rn(__dmul_rn(X.x, w.y), __dmul_rn(X.y, w.x));
        a[0] = make_double2(__dadd_rn(U.x, vr), __dadd_rn(U.y, vi));
        a[t] = make_double2(__dsub_rn(U.x, vr), __dsub_rn(U.y, vi));
    }
}

// coefficient = rint(re * fix); residues on every row
__global__ void k_enc_round(const double2* __restrict__ W, u64* __restrict__ out, int count, int n, int rows, RowMap rm,
                            double fix, ModTab mt, int* __restrict__ overflow) {
    size_t total = (size_t)count * n;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        u32 j = (u32)(e % n);
        size_t v = e / n;
        double x = rint(__dmul_rn(W[e].x, fix));
        bool neg = x < 0.0;
        double ax = fabs(x);
        if (!(ax < 0x1p126)) {
            *overflow = 1;
            ax = 0.0;
        }
        u64 lo = 0, hi = 0;
        if (ax >= 1.0) {
            long long bits = __double_as_longlong(ax);
            int ex = (int)(bits >> 52) - 1075;
            u64 mant = ((u64)bits & 0xFFFFFFFFFFFFFull) | (1ull << 52);
            if (ex <= 0) lo = mant >> (-ex);
            else if (ex < 64) lo = mant << ex, hi = mant >> (64 - ex);
            else hi = mant << (ex - 64);
        }
        for (int r = 0; r < rows; r++) {
            int t = rm.limb(r);
            u64 q = mt.q[t];
            u64 res = barrett128(lo, hi, q, mt.ratio0[t], mt.ratio1[t]);
            out[(v * rows + r) * n + j] = neg ? neg_mod(res, q) : res;
        }
    }
}

// x: [k][N] coefficient form of the first k (<= 3) limbs -> W[n] = (centred value / scale, 0)
__global__ void k_dec_garner(const u64* __restrict__ x, double2* __restrict__ W, int k, int N, double scale, ModTab mt,
                             const ulonglong2* __restrict__ gar) {
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        u64 v[3] = {0, 0, 0};
        for (int i = 0; i < k; i++) {
            u64 qi = mt.q[i], t = x[(size_t)i * N + n];
            for (int j = 0; j < i; j++) {
                ulonglong2 g = gar[i * 3 + j];
                t = mul_shoup(sub_mod(t, barrett64(v[j], qi, mt.ratio1[i]), qi), g.x, g.y, qi);
            }
            v[i] = t;
        }
        bool neg = false;
        for (int i = k - 1; i >= 0; i--) {
            u64 h = (mt.q[i] - 1) >> 1;
            if (v[i] > h) { neg = true; break; }
            if (v[i] < h) break;
        }
        double Wt[3];
        Wt[0] = 1.0;
        Wt[1] = k > 1 ? __ull2double_rn(mt.q[0]) : 0.0;
        Wt[2] = k > 2 ? __dmul_rn(__ull2double_rn(mt.q[0]), __ull2double_rn(mt.q[1])) : 0.0;
        double acc = 0.0;
        if (neg) {
            for (int i = k - 1; i >= 0; i--) acc = __dadd_rn(acc, __dmul_rn(__ull2double_rn(mt.q[i] - 1 - v[i]), Wt[i]));
            acc = -__dadd_rn(acc, 1.0);
        } else {
            for (int i = k - 1; i >= 0; i--) acc = __dadd_rn(acc, __dmul_rn(__ull2double_rn(v[i]), Wt[i]));
        }
        W[n] = make_double2(__ddiv_rn(acc, scale), 0.0);
    }
}

__global__ void k_dec_gather(const double2* __restrict__ W, double2* __restrict__ vals, int n, int logn) {
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < (u32)n / 2; j += gridDim.x * blockDim.x) {
        u32 pos = pow5_mod(j, 2u * n - 1);
        vals[j] = W[brev_n((pos - 1) >> 1, logn)];
    }
}

// ---- one kernel per encoding for rings that fit shared memory (sub-ring diagonals: n = 2D <= 8192) ----------------
// A CTA encodes one vector completely: scatter to the embedding order, the n-point inverse embedding (same
// Gentleman-Sande stages as k_fft_inv_stage, in shared memory), rounding to 128-bit integers, and for every limb the
// residues and the n-point forward NTT (same butterflies and twiddle prefix as ntt_forward on a ring of n), written out
// once, already in the split-30 storage form of the diagonal MAC.  Bit-identical to the staged path; it replaces
// 12 + 2 stage launches over count * n words each, a 2.6 GB round trip at C5 and a host synchronisation per batch
// (fully encrypted blocks encode four diagonal sets per block on the fly, reference test_fully_enc_bsgs.py:41-79).
__global__ void __launch_bounds__(256) k_encode_ring_fused(const double2* __restrict__ vals, u64* __restrict__ out, int n,
                                                           int logn, int rows, RowMap rm, double fix, ModTab mt,
                                                           NttTab tb, int N, const double2* __restrict__ zeta,
                                                           int split, int* __restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char smraw[];
    double2* W = reinterpret_cast<double2*>(smraw);                 // [n]   embedding values, then (lo, hi | sign) integers
    u64* buf = reinterpret_cast<u64*>(W + n);                       // [n]   one limb's residues
    const int T = blockDim.x, tid = threadIdx.x, half = n >> 1;
    const size_t v = blockIdx.x;
    const u32 m2 = 2u * n;
    for (int j = tid; j < half; j += T) {
        const u32 pos = pow5_mod((u32)j, m2 - 1);
        const double2 z = vals[v * half + j];
        W[brev_n((pos - 1) >> 1, logn)] = z;
        W[brev_n((m2 - pos - 1) >> 1, logn)] = make_double2(z.x, -z.y);
    }
    __syncthreads();
    for (int m = n, t = 1; m > 1; m >>= 1, t <<= 1) {               // inverse embedding: gap t doubles, twiddle conj(zeta[h + i])
        const int h = m >> 1;
        for (int b = tid; b < half; b += T) {
            const int i = b / t, kk = b - i * t;
            double2* a = W + 2 * i * t + kk;
            const double2 w = zeta[h + i], U = a[0], V = a[t];
            const double wr = w.x, wi = -w.y;
            a[0] = make_double2(__dadd_rn(U.x, V.x), __dadd_rn(U.y, V.y));
            const double dr = __dsub_rn(U.x, V.x), di = __dsub_rn(U.y, V.y);
            a[t] = make_double2(__dsub_rn(__dmul_rn(dr, wr), __dmul_rn(di, wi)), __dadd_rn(__dmul_rn(dr, wi), __dmul_rn(di, wr)));
        }
        __syncthreads();
    }
    u64* I = reinterpret_cast<u64*>(W);                             // I[2j] = low word, I[2j+1] = high word | sign << 63
    for (int j = tid; j < n; j += T) {
        const double x = rint(__dmul_rn(W[j].x, fix));
        const bool neg = x < 0.0;
        double ax = fabs(x);
        if (!(ax < 0x1p126)) {
            *overflow = 1;
            ax = 0.0;
        }
        u64 lo = 0, hi = 0;
        if (ax >= 1.0) {
            const long long bits = __double_as_longlong(ax);
            const int ex = (int)(bits >> 52) - 1075;
            const u64 mant = ((u64)bits & 0xFFFFFFFFFFFFFull) | (1ull << 52);
            if (ex <= 0) lo = mant >> (-ex);
            else if (ex < 64) lo = mant << ex, hi = mant >> (64 - ex);
            else hi = mant << (ex - 64);
        }
        I[2 * j] = lo, I[2 * j + 1] = hi | ((u64)neg << 63);      // |coefficient| < 2^126: bit 63 of the high word is free
    }
    __syncthreads();
    for (int r = 0; r < rows; r++) {
        const int lt = rm.limb(r);
        const u64 q = mt.q[lt], q2 = q << 1, r0 = mt.ratio0[lt], r1 = mt.ratio1[lt];
        const ulonglong2* __restrict__ tw = tb.psi + (size_t)lt * N;
        for (int j = tid; j < n; j += T) {
            const u64 hi = I[2 * j + 1];
            const u64 res = barrett128(I[2 * j], hi & ~(1ull << 63), q, r0, r1);
            buf[j] = (hi >> 63) ? neg_mod(res, q) : res;
        }
        __syncthreads();
        for (int m = 1, t = half; m < n; m <<= 1, t >>= 1) {        // forward NTT on the ring of n: twiddles psi[m + i]
            for (int b = tid; b < half; b += T) {
                const int i = b / t, kk = b - i * t, x0 = 2 * i * t + kk;
                u64 x = buf[x0], y = buf[x0 + t];
                ct_butterfly(x, y, tw[m + i], q, q2);
                buf[x0] = x, buf[x0 + t] = y;
            }
            __syncthreads();
        }
        u64* o = out + (v * rows + r) * n;
        for (int j = tid; j < n; j += T) {
            u64 x = buf[j];
            x = x >= q2 ? x - q2 : x;
            x = x >= q ? x - q : x;
            o[j] = split ? split30(x) : x;
        }
        __syncthreads();
    }
}

// The same with the per-limb NTT done by the register-tiled passes of ntt_core.cuh inside shared memory (n = 2^(SA+8):
// 2048 or 4096 points, n / 8 threads): pass A on [2^SA][16] column tiles of the residue buffer, pass B one 256-point
// chunk per warp, written straight to global memory -- four block barriers per limb instead of thirteen and the lazy
// butterflies instead of the exact ones.  The embedding itself (one transform of the l + P + 1) keeps the simple loop.
template <int SA>
__global__ void __launch_bounds__(32 << SA) k_encode_ring_tiled(const double2* __restrict__ vals, u64* __restrict__ out,
                                                                 int rows, RowMap rm, double fix, ModTab mt, NttTab tb, int N,
                                                                 const double2* __restrict__ zeta, int split,
                                                                 int* __restrict__ overflow) {
    using namespace nttc;
    constexpr int logn = SA + 8, n = 1 << logn, half = n >> 1, T = 32 << SA, S = n >> SA, TILE_T = 2 << SA;
    extern __shared__ __align__(16) unsigned char smraw[];
    double2* W = reinterpret_cast<double2*>(smraw);                 // [n]  embedding values, then (lo, hi | sign) integers
    u64* buf = reinterpret_cast<u64*>(W + n);                       // [n]  one limb's residues / pass-A output
    u64* xt = buf + n;                                              // [n]  exchange tiles of pass A, chunk buffers of pass B
    const int tid = threadIdx.x;
    const size_t v = blockIdx.x;
    const u32 m2 = 2u * n;
    for (int j = tid; j < half; j += T) {
        const u32 pos = pow5_mod((u32)j, m2 - 1);
        const double2 z = vals[v * half + j];
        W[brev_n((pos - 1) >> 1, logn)] = z;
        W[brev_n((m2 - pos - 1) >> 1, logn)] = make_double2(z.x, -z.y);
    }
    __syncthreads();
    for (int m = n, t = 1; m > 1; m >>= 1, t <<= 1) {
        const int h = m >> 1;
        for (int b = tid; b < half; b += T) {
            const int i = b / t, kk = b - i * t;
            double2* a = W + 2 * i * t + kk;
            const double2 w = zeta[h + i], U = a[0], V = a[t];
            const double wr = w.x, wi = -w.y;
            a[0] = make_double2(__dadd_rn(U.x, V.x), __dadd_rn(U.y, V.y));
            const double dr = __dsub_rn(U.x, V.x), di = __dsub_rn(U.y, V.y);
            a[t] = make_double2(__dsub_rn(__dmul_rn(dr, wr), __dmul_rn(di, wi)), __dadd_rn(__dmul_rn(dr, wi), __dmul_rn(di, wr)));
        }
        __syncthreads();
    }
    u64* I = reinterpret_cast<u64*>(W);
    for (int j = tid; j < n; j += T) {
        const double x = rint(__dmul_rn(W[j].x, fix));
        const bool neg = x < 0.0;
        double ax = fabs(x);
        if (!(ax < 0x1p126)) {
            *overflow = 1;
            ax = 0.0;
        }
        u64 lo = 0, hi = 0;
        if (ax >= 1.0) {
            const long long bits = __double_as_longlong(ax);
            const int ex = (int)(bits >> 52) - 1075;
            const u64 mant = ((u64)bits & 0xFFFFFFFFFFFFFull) | (1ull << 52);
            if (ex <= 0) lo = mant >> (-ex);
            else if (ex < 64) lo = mant << ex, hi = mant >> (64 - ex);
            else hi = mant << (ex - 64);
        }
        I[2 * j] = lo, I[2 * j + 1] = hi | ((u64)neg << 63);
    }
    __syncthreads();
    const int tile = tid / TILE_T, lt = tid % TILE_T, c = lt & (COLS - 1), g = lt >> 4;   // pass A: (tile, column, row group)
    const int warp = tid >> 5, lane = tid & 31;                                           // pass B: chunk = warp
    for (int r = 0; r < rows; r++) {
        const int lt_id = rm.limb(r);
        const u64 q = mt.q[lt_id], r0 = mt.ratio0[lt_id], r1 = mt.ratio1[lt_id];
        const ulonglong2* __restrict__ tw = tb.psi + (size_t)lt_id * N;
        const bool lazy = q < (1ull << 59) && q > (1ull << 33);
        for (int j = tid; j < n; j += T) {
            const u64 hi = I[2 * j + 1];
            const u64 res = barrett128(I[2 * j], hi & ~(1ull << 63), q, r0, r1);
            buf[j] = (hi >> 63) ? neg_mod(res, q) : res;
        }
        __syncthreads();
        u64 x8[8];
        if (lazy) fwd_a2_body<SA, true>(buf + tile * COLS + c, xt + tile * (COLS << SA), tw, q, S, c, g, x8);
        else fwd_a2_body<SA, false>(buf + tile * COLS + c, xt + tile * (COLS << SA), tw, q, S, c, g, x8);
        __syncthreads();
        u64* sw = xt + warp * 256;
        if (lazy) fwd_b2_body<true, false, const u64*>(buf + warp * 256, sw, tw, q, SA, warp, lane, split != 0);
        else fwd_b2_body<false, false, const u64*>(buf + warp * 256, sw, tw, q, SA, warp, lane, split != 0);
        __syncwarp();
        u64* o = out + (v * rows + r) * n + warp * 256;
#pragma unroll
        for (int k = 0; k < 8; k++) o[lane + 32 * k] = sw[swz(lane + 32 * k)];
        __syncthreads();
    }
}

int grid_for(const Ctx* c, size_t total) {
    size_t blocks = (total + TPB - 1) / TPB, cap = (size_t)c->sm_count * 16;
    return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

}  // namespace

namespace encoder {

void encode(const Ctx* c, const double2* vals, int count, int n, double scale, int l, bool ext, u64* out, cudaStream_t s) {
    int logn = 0;
    while ((1 << logn) < n) logn++;
    REQUIRE((1 << logn) == n && n >= 4 && n <= c->N, "encode: ring size %d invalid", n);
    const int rows = l + (ext ? c->P : 0);
    double2* W = (double2*)c->alloc((size_t)count * n * 2);
    int* flag = (int*)c->alloc(1);
    CUDA_CHECK(cudaMemsetAsync(W, 0, sizeof(double2) * count * n, s));
    CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), s));
    size_t half = (size_t)count * (n / 2);
    LAUNCH(k_enc_scatter, grid_for(c, half), TPB, 0, s)(vals, W, count, n, logn);
    for (int m = n, t = 1; m > 1; m >>= 1, t <<= 1)
        LAUNCH(k_fft_inv_stage, grid_for(c, half), TPB, 0, s)(W, c->d_zeta, count, n, m >> 1, t);
    RowMap rm{rows, l, c->L, 0};
    LAUNCH(k_enc_round, grid_for(c, (size_t)count * n), TPB, 0, s)(W, out, count, n, rows, rm, scale / (double)n,
                                                              c->modtab(), flag);
    ntt_forward(c, out, count * rows, rm, n, s);
    int h_flag = 0;
    CUDA_CHECK(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    c->free(W);
    c->free(flag);
    REQUIRE(!h_flag, "encode: scaled value too large (|coefficient| >= 2^126)");
}

// count vectors on a ring of n <= 8192, one kernel; out in split-30 form when `split30_out`.  `overflow`: sticky device
// flag (the caller checks it once per batch).  false: the ring does not fit shared memory.
bool encode_ring_fused(const Ctx* c, const double2* vals, int count, int n, double scale, int l, bool ext, u64* out,
                       bool split30_out, int* overflow, cudaStream_t s) {
    int logn = 0;
    while ((1 << logn) < n) logn++;
    const size_t smem = (size_t)n * (sizeof(double2) + sizeof(u64));
    if ((1 << logn) != n || n < 4 || n > c->N || smem > 200 * 1024 || count < 1) return false;
    const int rows = l + (ext ? c->P : 0);
    static const bool tiled = [] {
        const char* e = getenv("SPEAR_ENCODE_TILED");
        return !(e && e[0] == '0');
    }();
    if (tiled && (logn == 11 || logn == 12) && count <= 65535 * 16) {   // 2048 / 4096 points: register-tiled NTT passes in shared memory
        const size_t smem_t = (size_t)n * 32;
        auto go = [&](auto kern, int threads) {
            CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            LAUNCH(kern, count, threads, smem_t, s)(vals, out, rows, RowMap{rows, l, c->L, 0}, scale / (double)n, c->modtab(),
                                                    c->ntttab(), c->N, c->d_zeta, split30_out ? 1 : 0, overflow);
        };
        if (logn == 11) go(k_encode_ring_tiled<3>, 32 << 3);
        else go(k_encode_ring_tiled<4>, 32 << 4);
        CUDA_CHECK(cudaGetLastError());
        return true;
    }
    CUDA_CHECK(cudaFuncSetAttribute(k_encode_ring_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int v0 = 0; v0 < count; v0 += 65535 * 16) {   // grid.x is wide enough; kept as a loop for clarity of the bound
        const int nv = std::min(count - v0, 65535 * 16);
        LAUNCH(k_encode_ring_fused, nv, 256, smem, s)(vals + (size_t)v0 * (n / 2), out + (size_t)v0 * rows * n, n, logn, rows,
                                                     RowMap{rows, l, c->L, 0}, scale / (double)n, c->modtab(), c->ntttab(), c->N,
                                                     c->d_zeta, split30_out ? 1 : 0, overflow);
    }
    CUDA_CHECK(cudaGetLastError());
    return true;
}

void decode(const Ctx* c, const u64* pt, int l, double scale, double2* vals, cudaStream_t s) {
    const int N = c->N, k = l < 3 ? l : 3;
    u64* x = c->alloc((size_t)k * N);
    double2* W = (double2*)c->alloc((size_t)N * 2);
    CUDA_CHECK(cudaMemcpyAsync(x, pt, sizeof(u64) * k * N, cudaMemcpyDeviceToDevice, s));
    ntt_inverse(c, x, k, RowMap{k, k, c->L, 0}, N, s);
    LAUNCH(k_dec_garner, grid_for(c, N), TPB, 0, s)(x, W, k, N, scale, c->modtab(), c->d_garner);
    for (int m = 1, t = N >> 1; m < N; m <<= 1, t >>= 1)
        LAUNCH(k_fft_fwd_stage, grid_for(c, N / 2), TPB, 0, s)(W, c->d_zeta, 1, N, m, t);
    LAUNCH(k_dec_gather, grid_for(c, N / 2), TPB, 0, s)(W, vals, N, c->logn);
    CUDA_CHECK(cudaGetLastError());
    c->free(x);
    c->free(W);
}

}  // namespace encoder
