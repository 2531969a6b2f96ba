// client.cu -- the client legs of a projection round trip in three launches each (SURVEY.md section 8 row f4;
// reference scripts/bootstrap_generation.py:119-147: encode + sk.encrypt_symmetric, sk.decrypt + decode; 192 of each per
// RWKV-7 token).
//
//   encode + encrypt (symmetric)                                  decrypt + decode
//   E1  slot gather + inverse embedding, gaps 1..128              D1  c0 + c1 s (+ c2 s^2) on the <= 3 limbs the decoder
//       (one warp per 256-point chunk)                                reads, fused with the inverse NTT's first pass
//   E2  inverse embedding, remaining stages on column tiles;      D2  inverse NTT's second pass of those limbs, n^-1,
//       rounding to 128-bit integers; + error e (ChaCha, in           Garner + centring to a double, forward embedding's
//       place); residues of one limb; forward NTT pass A              first stages on the same column tile
//   E3  forward NTT pass B fused with c1 = a (ChaCha, in place)   D3  forward embedding, last eight stages per chunk,
//       and c0 = NTT(m + e) - a s                                     gather of the wanted slots
//
// The staged path (encoder.cu + sampler.cu + ntt.cu: 15 single-stage FFT launches per transform, ~25 launches and a host
// synchronisation per encryption) stays as the general form; both produce the same limbs bit for bit -- every
// butterfly computes the same expression with the same operands (no FMA contraction), NTT(m) + NTT(e) = NTT(m + e) is
// exact modulo q, and the randomness comes from the same ChaCha words.
#include "chacha.cuh"
#include "engine.h"
#include "ntt_core.cuh"
#include "ops.h"

namespace {

using namespace nttc;

// ---- complex butterflies: exactly the expressions of encoder.cu k_fft_inv_stage / k_fft_fwd_stage ------------------
__device__ __forceinline__ void c_gs(double2& a, double2& b, double2 w) {   // (U, V) -> (U + V, (U - V) conj(w))
    const double wr = w.x, wi = -w.y;
    const double2 U = a, V = b;
    a = make_double2(__dadd_rn(U.x, V.x), __dadd_rn(U.y, V.y));
    const double dr = __dsub_rn(U.x, V.x), di = __dsub_rn(U.y, V.y);
    b = make_double2(__dsub_rn(__dmul_rn(dr, wr), __dmul_rn(di, wi)), __dadd_rn(__dmul_rn(dr, wi), __dmul_rn(di, wr)));
}
__device__ __forceinline__ void c_ct(double2& a, double2& b, double2 w) {   // (U, X) -> (U + X w, U - X w)
    const double2 U = a, X = b;
    const double vr = __dsub_rn(__dmul_rn(X.x, w.x), __dmul_rn(X.y, w.y));
    const double vi = __dadd_rn(__dmul_rn(X.x, w.y), __dmul_rn(X.y, w.x));
    a = make_double2(__dadd_rn(U.x, vr), __dadd_rn(U.y, vi));
    b = make_double2(__dsub_rn(U.x, vr), __dsub_rn(U.y, vi));
}

// Eight stages with gaps 1..128 (Gentleman-Sande, inverse embedding) of one 256-point chunk by one warp: the index
// structure of nttc::inv_b2_body with complex points.  Chunk in s (point x at s[swz(x)]); lane j ends with points j + 32k.
__device__ __forceinline__ void c_inv_b(double2* __restrict__ s, const double2* __restrict__ tw, int n, int gc, int j,
                                        double2 (&v)[8]) {
    {   // x = 8j + i ; t = 1, 2, 4
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = s[swz(8 * j + i)];
        int h = n >> 1;
#pragma unroll
        for (int k = 0; k < 4; k++) c_gs(v[2 * k], v[2 * k + 1], tw[h + 128 * gc + 4 * j + k]);
        h >>= 1;
        const double2 w0 = tw[h + 64 * gc + 2 * j], w1 = tw[h + 64 * gc + 2 * j + 1];
        c_gs(v[0], v[2], w0), c_gs(v[1], v[3], w0);
        c_gs(v[4], v[6], w1), c_gs(v[5], v[7], w1);
        h >>= 1;
        const double2 w = tw[h + 32 * gc + j];
#pragma unroll
        for (int k = 0; k < 4; k++) c_gs(v[k], v[k + 4], w);
#pragma unroll
        for (int i = 0; i < 8; i++) s[swz(8 * j + i)] = v[i];
    }
    __syncwarp();
    {   // x = 32*blk + b + 8i ; t = 8, 16
        const int bh = j >> 3, b = j & 7;
        const int h8 = n >> 4, h16 = n >> 5;
#pragma unroll
        for (int qd = 0; qd < 2; qd++) {
            const int blk = bh + 4 * qd;
            double2* e = v + 4 * qd;
#pragma unroll
            for (int i = 0; i < 4; i++) e[i] = s[swz(32 * blk + b + 8 * i)];
            const double2 wa = tw[h8 + 16 * gc + 2 * blk], wb = tw[h8 + 16 * gc + 2 * blk + 1];
            c_gs(e[0], e[1], wa), c_gs(e[2], e[3], wb);
            const double2 w = tw[h16 + 8 * gc + blk];
            c_gs(e[0], e[2], w), c_gs(e[1], e[3], w);
#pragma unroll
            for (int i = 0; i < 4; i++) s[swz(32 * blk + b + 8 * i)] = e[i];
        }
    }
    __syncwarp();
    {   // x = j + 32k ; t = 32, 64, 128
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = s[swz(j + 32 * k)];
        int h = n >> 6;
#pragma unroll
        for (int k = 0; k < 4; k++) c_gs(v[2 * k], v[2 * k + 1], tw[h + 4 * gc + k]);
        h >>= 1;
        const double2 w0 = tw[h + 2 * gc], w1 = tw[h + 2 * gc + 1];
        c_gs(v[0], v[2], w0), c_gs(v[1], v[3], w0);
        c_gs(v[4], v[6], w1), c_gs(v[5], v[7], w1);
        h >>= 1;
        const double2 w = tw[h + gc];
#pragma unroll
        for (int k = 0; k < 4; k++) c_gs(v[k], v[k + 4], w);
    }
}

// Eight stages with gaps 128..1 (Cooley-Tukey, forward embedding) of one chunk by one warp: the index structure of
// nttc::fwd_b2_body.  Lane j enters with points j + 32k in v and ends with points 8j + i in v[i].
__device__ __forceinline__ void c_fwd_b(double2* __restrict__ s, const double2* __restrict__ tw, int sA, int gc, int j,
                                        double2 (&v)[8]) {
    int m = 1 << sA;
    {   // x = j + 32k ; t = 128, 64, 32
        const double2 w = tw[m + gc];
#pragma unroll
        for (int k = 0; k < 4; k++) c_ct(v[k], v[k + 4], w);
        m <<= 1;
        const double2 w0 = tw[m + 2 * gc], w1 = tw[m + 2 * gc + 1];
        c_ct(v[0], v[2], w0), c_ct(v[1], v[3], w0);
        c_ct(v[4], v[6], w1), c_ct(v[5], v[7], w1);
        m <<= 1;
#pragma unroll
        for (int k = 0; k < 4; k++) c_ct(v[2 * k], v[2 * k + 1], tw[m + 4 * gc + k]);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) s[swz(j + 32 * k)] = v[k];
    __syncwarp();
    {   // x = 32*blk + b + 8i ; t = 16, 8
        const int bh = j >> 3, b = j & 7;
        m <<= 1;
#pragma unroll
        for (int qd = 0; qd < 2; qd++) {
            const int blk = bh + 4 * qd;
            double2* e = v + 4 * qd;
#pragma unroll
            for (int i = 0; i < 4; i++) e[i] = s[swz(32 * blk + b + 8 * i)];
            const double2 w = tw[m + 8 * gc + blk];
            c_ct(e[0], e[2], w), c_ct(e[1], e[3], w);
            const double2 wa = tw[2 * m + 16 * gc + 2 * blk], wb = tw[2 * m + 16 * gc + 2 * blk + 1];
            c_ct(e[0], e[1], wa), c_ct(e[2], e[3], wb);
#pragma unroll
            for (int i = 0; i < 4; i++) s[swz(32 * blk + b + 8 * i)] = e[i];
        }
        m <<= 1;
    }
    __syncwarp();
    {   // x = 8j + i ; t = 4, 2, 1
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = s[swz(8 * j + i)];
        m <<= 1;
        const double2 w = tw[m + 32 * gc + j];
#pragma unroll
        for (int k = 0; k < 4; k++) c_ct(v[k], v[k + 4], w);
        m <<= 1;
        const double2 w0 = tw[m + 64 * gc + 2 * j], w1 = tw[m + 64 * gc + 2 * j + 1];
        c_ct(v[0], v[2], w0), c_ct(v[1], v[3], w0);
        c_ct(v[4], v[6], w1), c_ct(v[5], v[7], w1);
        m <<= 1;
#pragma unroll
        for (int k = 0; k < 4; k++) c_ct(v[2 * k], v[2 * k + 1], tw[m + 128 * gc + 4 * j + k]);
    }
}

// The SA stages with the largest gaps on a [R = 2^SA][16] column tile, complex points -- the index structure of
// nttc::inv_a2_rounds (inverse: gaps grow) and nttc::fwd_a2_body (forward: gaps shrink).  2R threads, thread =
// (column c, row group g); the inverse ends / the forward starts with rows rowbase + k * 2^(SA-3) in v[k].
template <int SA>
__device__ __forceinline__ void c_inv_a(const double2* __restrict__ base, double2* __restrict__ sm,
                                        const double2* __restrict__ tw, int S, int c, int g, double2 (&v)[8]) {
    constexpr int R = 1 << SA;
    constexpr int NS0 = (SA % 3) ? (SA % 3) : 3;
#pragma unroll
    for (int u0 = 0; u0 < SA; u0 += (u0 == 0 ? NS0 : 3)) {
        const int ns = u0 == 0 ? NS0 : 3;
        const int gs = 1 << u0;
        const int rowbase = (g / gs) * (8 * gs) + (g % gs);
        if (u0 == 0) {
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = base[(size_t)(rowbase + k * gs) * S];
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = sm[(rowbase + k * gs) * COLS + c];
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {
            if (i < ns) {
                const int u = u0 + i, kgap = 1 << i;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (!(k & kgap)) {
                        const int r = rowbase + k * gs;
                        c_gs(v[k], v[k + kgap], tw[(R >> (u + 1)) + (r >> (u + 1))]);
                    }
                }
            }
        }
        if (u0 + ns < SA) {
#pragma unroll
            for (int k = 0; k < 8; k++) sm[(rowbase + k * gs) * COLS + c] = v[k];
            __syncthreads();
        }
    }
}
template <int SA>
__device__ __forceinline__ void c_fwd_a(double2* __restrict__ base, double2* __restrict__ sm,
                                        const double2* __restrict__ tw, int S, int c, int g, double2 (&v)[8]) {
    constexpr int R = 1 << SA;
#pragma unroll
    for (int s0 = 0; s0 < SA; s0 += 3) {
        const int ns = (SA - s0) < 3 ? (SA - s0) : 3;
        const int gbot = R >> (s0 + ns);
        const int rowbase = (g / gbot) * (8 * gbot) + (g % gbot);
        if (s0 != 0) {
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = sm[(rowbase + k * gbot) * COLS + c];
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {
            if (i < ns) {
                const int st = s0 + i, kgap = 1 << (ns - 1 - i);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (!(k & kgap)) {
                        const int r = rowbase + k * gbot;
                        c_ct(v[k], v[k + kgap], tw[(1 << st) + (r >> (SA - st))]);
                    }
                }
            }
        }
        if (s0 + 3 >= SA) {
#pragma unroll
            for (int k = 0; k < 8; k++) base[(size_t)(rowbase + k * gbot) * S] = v[k];
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) sm[(rowbase + k * gbot) * COLS + c] = v[k];
            __syncthreads();
        }
    }
}

// ---- E1: W[p] = slot value (or its conjugate) of position p, then the inverse embedding's first eight stages --------
// pos_slot[p]: slot index of embedding position p, bit 31 set where the position holds the conjugate.
__global__ void __launch_bounds__(WB * 32) k_enc_gather_fft_b(const double2* __restrict__ vals, int nvals, int replicate,
                                                               const u32* __restrict__ pos_slot,
                                                               const double2* __restrict__ zeta, double2* __restrict__ W,
                                                               int n) {
    __shared__ double2 smem[WB][256];
    const int warp = threadIdx.x >> 5, j = threadIdx.x & 31;
    const int gc = blockIdx.x * WB + warp;
    double2* s = smem[warp];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int x = j + 32 * k;
        const u32 ps = pos_slot[gc * 256 + x];
        u32 slot = ps & 0x7FFFFFFFu;
        double2 z = make_double2(0.0, 0.0);
        if (replicate) slot %= (u32)nvals;
        if (slot < (u32)nvals) z = vals[slot];
        if (ps >> 31) z.y = -z.y;
        s[swz(x)] = z;
    }
    __syncwarp();
    double2 v[8];
    c_inv_b(s, zeta, n, gc, j, v);
    double2* base = W + (size_t)gc * 256;
#pragma unroll
    for (int k = 0; k < 8; k++) base[j + 32 * k] = v[k];
}

// ---- E2: rest of the inverse embedding, rounding, error, residues of limb blockIdx.y, forward NTT pass A -----------
template <int SA>
__global__ void __launch_bounds__(2 << SA) k_enc_round_ntt_a(const double2* __restrict__ W, u64* __restrict__ c0, int N,
                                                             double fix, Seed seed, u64 nonce_e, ModTab mt, NttTab tb,
                                                             const double2* __restrict__ zeta) {
    constexpr int GS = 1 << (SA - 3);
    extern __shared__ __align__(16) double2 csm[];           // [2^SA][COLS] complex exchange tile (dynamic: 64 KB at SA = 8)
    u64* sm = reinterpret_cast<u64*>(csm);                   // the NTT's exchange tile reuses the complex one
    const int c = threadIdx.x & (COLS - 1), g = threadIdx.x >> 4, S = N >> SA;
    const int col = blockIdx.x * COLS + c, t = blockIdx.y;
    double2 v[8];
    c_inv_a<SA>(W + col, csm, zeta, S, c, g, v);
    const int rowbase = (g / GS) * (8 * GS) + (g % GS);
    const u64 q = mt.q[t], r0 = mt.ratio0[t], r1 = mt.ratio1[t];
    u64 x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const double xr = rint(__dmul_rn(v[k].x, fix));         // |coefficient| < 2^126 is checked by the host
        const bool neg = xr < 0.0;
        const double ax = fabs(xr);
        u64 lo = 0, hi = 0;
        if (ax >= 1.0) {
            const long long bits = __double_as_longlong(ax);
            const int ex = (int)(bits >> 52) - 1075;
            const u64 mant = ((u64)bits & 0xFFFFFFFFFFFFFull) | (1ull << 52);
            if (ex <= 0) lo = mant >> (-ex);
            else if (ex < 64) lo = mant << ex, hi = mant >> (64 - ex);
            else hi = mant << (ex - 64);
        }
        u64 m = barrett128(lo, hi, q, r0, r1);
        if (neg) m = neg_mod(m, q);
        const u32 jc = (u32)(rowbase + k * GS) * (u32)S + (u32)col;   // coefficient index
        u64 blk[8];
        chacha_block(seed, nonce_e, (u64)(jc >> 3), blk);
        const int e = cbd_of_word(blk[jc & 7]);
        x[k] = add_mod(m, e >= 0 ? (u64)e : q - (u64)(-e), q);
    }
    __syncthreads();                                           // every thread is done with the complex tile
    u64* base = c0 + (size_t)t * N + col;
    const ulonglong2* __restrict__ tw = tb.psi + (size_t)t * N;
    if (q < (1ull << 59) && q > (1ull << 33)) fwd_a2_body<SA, true, true>(base, sm, tw, q, S, c, g, x);
    else fwd_a2_body<SA, false, true>(base, sm, tw, q, S, c, g, x);
}

// ---- E3: forward NTT pass B of c0's rows, c1 = a (uniform, ChaCha), c0 = NTT(m + e) - a s ---------------------------
__global__ void __launch_bounds__(WB * 32) k_enc_ntt_b_combine(u64* __restrict__ c0, u64* __restrict__ c1,
                                                                const u64* __restrict__ sk, int N, int sA, Seed seed,
                                                                u64 nonce_a, ModTab mt, NttTab tb) {
    __shared__ u64 smem[WB][256];
    const int t = blockIdx.y;
    const u64 q = tb.q[t], r0 = mt.ratio0[t], r1 = mt.ratio1[t];
    const ulonglong2* __restrict__ tw = tb.psi + (size_t)t * N;
    const int warp = threadIdx.x >> 5, j = threadIdx.x & 31;
    const int gc = blockIdx.x * WB + warp;
    u64* s = smem[warp];
    u64* base = c0 + (size_t)t * N + (size_t)gc * 256;
    if (q < (1ull << 59) && q > (1ull << 33)) fwd_b2_body<true, false>(base, s, tw, q, sA, gc, j, false);
    else fwd_b2_body<false, false>(base, s, tw, q, sA, gc, j, false);
    // lane j owns coefficients 8j .. 8j + 7 of the chunk (it wrote them itself: no barrier needed)
    const size_t n0 = (size_t)gc * 256 + 8 * j, off = (size_t)t * N + n0;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        u64 blk[8];
        chacha_block(seed, nonce_a, ((u64)t * N + n0 + 4 * half) >> 2, blk);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = 4 * half + k;
            const u64 a = barrett128(blk[2 * k + 1], blk[2 * k], q, r0, r1);
            const u64 v = s[swz(8 * j + i)];
            c1[off + i] = a;
            c0[off + i] = sub_mod(v, mul_mod(a, sk[off + i], q, r0, r1), q);
        }
    }
}

// ---- D1: pt = c0 + c1 s (+ c2 s^2) on limb blockIdx.y, then the inverse NTT's first pass ------------------------------
__global__ void __launch_bounds__(WB * 32) k_dec_combine_inv_b(const u64* __restrict__ ct, int size, int l,
                                                                const u64* __restrict__ sk, u64* __restrict__ x, int N,
                                                                ModTab mt, NttTab tb) {
    __shared__ u64 smem[WB][256];
    const int t = blockIdx.y;
    const u64 q = tb.q[t], r0 = mt.ratio0[t], r1 = mt.ratio1[t];
    const ulonglong2* __restrict__ tw = tb.ipsi + (size_t)t * N;
    const int warp = threadIdx.x >> 5, j = threadIdx.x & 31;
    const int gc = blockIdx.x * WB + warp;
    u64* s = smem[warp];
    const size_t pw = (size_t)l * N, off = (size_t)t * N + (size_t)gc * 256;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const size_t e = off + j + 32 * k;
        const u64 sv = sk[e];
        u64 acc = ct[(size_t)(size - 1) * pw + e];
        for (int p = size - 2; p >= 0; p--) acc = add_mod(mul_mod(acc, sv, q, r0, r1), ct[(size_t)p * pw + e], q);
        s[swz(j + 32 * k)] = acc;
    }
    __syncwarp();
    u64 v[8];
    inv_b2_body(s, tw, q, N, gc, j, v);
    u64* base = x + off;
#pragma unroll
    for (int k = 0; k < 8; k++) base[j + 32 * k] = v[k];
}

// ---- D2: inverse NTT pass A of the k limbs, n^-1, Garner + centring, forward embedding's first SA stages ---------------
template <int SA>
__global__ void __launch_bounds__(2 << SA) k_dec_garner_fft_a(const u64* __restrict__ x, double2* __restrict__ W, int k,
                                                              int N, int logn, double scale, ModTab mt, NttTab tb,
                                                              const ulonglong2* __restrict__ gar,
                                                              const double2* __restrict__ zeta) {
    extern __shared__ __align__(16) double2 csm[];           // [2^SA][COLS]
    u64* sm = reinterpret_cast<u64*>(csm);
    const int c = threadIdx.x & (COLS - 1), g = threadIdx.x >> 4, S = N >> SA;
    const int col = blockIdx.x * COLS + c;
    u64 res[3][8];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (i < k) {
            const u64 q = tb.q[i];
            u64 v[8];
            inv_a2_rounds<SA>(x + (size_t)i * N + col, sm, tb.ipsi + (size_t)i * N, q, S, c, g, v);
            const ulonglong2 ninv = tb.invn[i * 17 + logn];
#pragma unroll
            for (int kk = 0; kk < 8; kk++) res[i][kk] = mul_shoup(v[kk], ninv.x, ninv.y, q);
            __syncthreads();
        }
    }
    double Wt[3];
    Wt[0] = 1.0;
    Wt[1] = k > 1 ? __ull2double_rn(mt.q[0]) : 0.0;
    Wt[2] = k > 2 ? __dmul_rn(__ull2double_rn(mt.q[0]), __ull2double_rn(mt.q[1])) : 0.0;
    double2 w[8];
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {                            // same arithmetic as encoder.cu k_dec_garner
        u64 v[3] = {0, 0, 0};
        for (int i = 0; i < k; i++) {
            const u64 qi = mt.q[i];
            u64 t = res[i][kk];
            for (int jj = 0; jj < i; jj++) {
                const ulonglong2 gg = gar[i * 3 + jj];
                t = mul_shoup(sub_mod(t, barrett64(v[jj], qi, mt.ratio1[i]), qi), gg.x, gg.y, qi);
            }
            v[i] = t;
        }
        bool neg = false;
        for (int i = k - 1; i >= 0; i--) {
            const u64 h = (mt.q[i] - 1) >> 1;
            if (v[i] > h) { neg = true; break; }
            if (v[i] < h) break;
        }
        double acc = 0.0;
        if (neg) {
            for (int i = k - 1; i >= 0; i--) acc = __dadd_rn(acc, __dmul_rn(__ull2double_rn(mt.q[i] - 1 - v[i]), Wt[i]));
            acc = -__dadd_rn(acc, 1.0);
        } else {
            for (int i = k - 1; i >= 0; i--) acc = __dadd_rn(acc, __dmul_rn(__ull2double_rn(v[i]), Wt[i]));
        }
        w[kk] = make_double2(__ddiv_rn(acc, scale), 0.0);
    }
    c_fwd_a<SA>(W + col, csm, zeta, S, c, g, w);
}

// ---- D3: forward embedding's last eight stages per chunk, then vals[slot] = W[position of slot] ----------------------
__global__ void __launch_bounds__(WB * 32) k_dec_fft_b_gather(const double2* __restrict__ W, const u32* __restrict__ pos_slot,
                                                               const double2* __restrict__ zeta, double2* __restrict__ vals,
                                                               int want, int sA) {
    __shared__ double2 smem[WB][256];
    const int warp = threadIdx.x >> 5, j = threadIdx.x & 31;
    const int gc = blockIdx.x * WB + warp;
    const double2* base = W + (size_t)gc * 256;
    double2 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = base[j + 32 * k];
    c_fwd_b(smem[warp], zeta, sA, gc, j, v);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const u32 ps = pos_slot[gc * 256 + 8 * j + i];
        if (!(ps >> 31) && ps < (u32)want) vals[ps] = v[i];
    }
}

}  // namespace

namespace client {

bool fused_applies(const Ctx* c) {
    static const bool enabled = [] {
        const char* e = getenv("SPEAR_FUSED_CLIENT");
        return !(e && e[0] == '0');
    }();
    return enabled && c->logn >= 11 && c->logn <= 16 && c->d_pos_slot != nullptr;
}

// vals: nvals complex slot values on the device; ct: [2][l][N] (l data limbs from limb 0); W: scratch [N] complex
void encode_encrypt(const Ctx* c, const double2* vals, int nvals, bool replicate, double scale, int l, const u32* seed,
                    u64 enc_id, const u64* sk, u64* ct, double2* W, cudaStream_t s) {
    const int N = c->N, sA = c->logn - 8;
    const Seed sd = make_seed(seed);
    LAUNCH(k_enc_gather_fft_b, N / (256 * WB), WB * 32, 0, s)(vals, nvals, replicate ? 1 : 0, c->d_pos_slot, c->d_zeta, W, N);
    const dim3 ga((N >> sA) / COLS, l);
    const double fix = scale / (double)N;
    const u64 ne = stream_id(DOM_ENC_E, enc_id), na = stream_id(DOM_ENC_A, enc_id);
    const size_t tile_bytes = sizeof(double2) * ((size_t)COLS << sA);
#define CL_CASE(SA_)                                                                                                        \
    case SA_:                                                                                                               \
        CUDA_CHECK(cudaFuncSetAttribute(k_enc_round_ntt_a<SA_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));   \
        LAUNCH(k_enc_round_ntt_a<SA_>, ga, 2 << SA_, tile_bytes, s)(W, ct, N, fix, sd, ne, c->modtab(), c->ntttab(), c->d_zeta); \
        break;
    switch (sA) { CL_CASE(3) CL_CASE(4) CL_CASE(5) CL_CASE(6) CL_CASE(7) default: CL_CASE(8) }
#undef CL_CASE
    LAUNCH(k_enc_ntt_b_combine, dim3(N / (256 * WB), l), WB * 32, 0, s)(ct, ct + (size_t)l * N, sk, N, sA, sd, na, c->modtab(),
                                                                       c->ntttab());
    CUDA_CHECK(cudaGetLastError());
}

// ct: [size][l][N]; vals_out: `want` complex slots on the device; x: scratch [3][N]; W: scratch [N] complex
void decrypt_decode(const Ctx* c, const u64* ct, int size, int l, double scale, const u64* sk, double2* vals_out, int want,
                    u64* x, double2* W, cudaStream_t s) {
    const int N = c->N, sA = c->logn - 8, k = l < 3 ? l : 3;
    LAUNCH(k_dec_combine_inv_b, dim3(N / (256 * WB), k), WB * 32, 0, s)(ct, size, l, sk, x, N, c->modtab(), c->ntttab());
    const int ga = (N >> sA) / COLS;
    const size_t tile_bytes = sizeof(double2) * ((size_t)COLS << sA);
#define CL_CASE(SA_)                                                                                                        \
    case SA_:                                                                                                               \
        CUDA_CHECK(cudaFuncSetAttribute(k_dec_garner_fft_a<SA_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));  \
        LAUNCH(k_dec_garner_fft_a<SA_>, ga, 2 << SA_, tile_bytes, s)(x, W, k, N, c->logn, scale, c->modtab(), c->ntttab(),  \
                                                                     c->d_garner, c->d_zeta);                               \
        break;
    switch (sA) { CL_CASE(3) CL_CASE(4) CL_CASE(5) CL_CASE(6) CL_CASE(7) default: CL_CASE(8) }
#undef CL_CASE
    LAUNCH(k_dec_fft_b_gather, N / (256 * WB), WB * 32, 0, s)(W, c->d_pos_slot, c->d_zeta, vals_out, want, sA);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace client
