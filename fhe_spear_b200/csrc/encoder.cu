// encoder.cu -- CKKS encode / decode on the device (canonical embedding with generator 5).
// Mirrors oracle/spear_oracle.c orc_encode / orc_decode operation by operation: every double
// product and sum is rounded separately (__dmul_rn / __dadd_rn, no FMA contraction) and the
// root table comes from the same libm calls, so coefficients are bit-identical to the oracle's.
// Replaces PhantomCKKSEncoder::{encode,decode} (reference gpu/phantom_binding.cu:138-156) and the
// fork-only encode_*_vector_batch (reference scripts/bootstrap_generation.py:382,423).
#include "engine.h"
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
