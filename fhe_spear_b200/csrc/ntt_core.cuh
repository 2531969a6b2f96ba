// ntt_core.cuh -- device bodies of the register-tiled transform passes (n >= 2048), shared by the transform kernels
// (ntt.cu) and the fused client kernels (client.cu).  See ntt.cu for the decomposition into passes.
#pragma once
#include "common.cuh"

namespace nttc {

constexpr int COLS = 16;        // pass A tile width

// ---- lazy forward butterflies for q < 2^59 (31q < 2^64) -------------------------------------------------
// No conditional subtraction at all: the twiddle product uses an approximate Shoup quotient (three 32x32
// products instead of a full 64x64 high half, error <= 2, result in [0,4q)), so a value grows by at most 4q
// per stage; 7 such stages plus one exact-Shoup stage stay below 31q, and each pass ends with one cheap
// reduction (the Barrett ratio floor(2^64/q) fits 32 bits).
__device__ __forceinline__ u64 mul_shoup_apx(u64 a, u64 w, u64 wp, u64 q) {
    // Written out in PTX so that the instruction selection is exactly: 4 IMAD.WIDE (two of them only for their high
    // halves -- a non-accumulating IMAD.WIDE issues in 2 cycles on sm_100a, IMAD.HI in 5, tools/ubench/imad.cu),
    // 1 accumulating IMAD.WIDE, 4 IMAD and 4 carry adds:
    //   Q = a1*p1 + hi(a0*p1) + hi(a1*p0);   r = lo64(a*w) - lo64(Q*q)
    u64 r;
    asm("{\n\t"
        ".reg .u32 a0, a1, w0, w1, p0, p1, q0, q1, h1, h2, z, Q0, Q1, T0, T1, U0, U1;\n\t"
        ".reg .u64 t, Q, T, U;\n\t"
        "mov.b64 {a0, a1}, %1;\n\t"
        "mov.b64 {w0, w1}, %2;\n\t"
        "mov.b64 {p0, p1}, %3;\n\t"
        "mov.b64 {q0, q1}, %4;\n\t"
        "mul.wide.u32 t, a0, p1;\n\t"
        "mov.b64 {z, h1}, t;\n\t"
        "mul.wide.u32 t, a1, p0;\n\t"
        "mov.b64 {z, h2}, t;\n\t"
        "mov.u32 z, 0;\n\t"
        "mov.b64 t, {h1, z};\n\t"
        "mad.wide.u32 Q, a1, p1, t;\n\t"
        "mov.b64 {Q0, Q1}, Q;\n\t"
        "add.cc.u32 Q0, Q0, h2;\n\t"
        "addc.u32 Q1, Q1, 0;\n\t"
        "mul.wide.u32 T, a0, w0;\n\t"
        "mov.b64 {T0, T1}, T;\n\t"
        "mad.lo.u32 T1, a0, w1, T1;\n\t"
        "mad.lo.u32 T1, a1, w0, T1;\n\t"
        "mul.wide.u32 U, Q0, q0;\n\t"
        "mov.b64 {U0, U1}, U;\n\t"
        "mad.lo.u32 U1, Q0, q1, U1;\n\t"
        "mad.lo.u32 U1, Q1, q0, U1;\n\t"
        "sub.cc.u32 T0, T0, U0;\n\t"
        "subc.u32 T1, T1, U1;\n\t"
        "mov.b64 %0, {T0, T1};\n\t"
        "}"
        : "=l"(r)
        : "l"(a), "l"(w), "l"(wp), "l"(q));
    return r;
}
template <bool LAZY>
__device__ __forceinline__ void fwd_bf(u64& x, u64& y, ulonglong2 w, u64 q, u64 q2, u64 q4) {
    if (LAZY) {
        const u64 t = mul_shoup_apx(y, w.x, w.y, q);
        y = x - t + q4;
        x = x + t;
    } else {
        ct_butterfly(x, y, w, q, q2);
    }
}
// last stage of a lazy pass: exact Shoup product (grows by 2q only)
template <bool LAZY>
__device__ __forceinline__ void fwd_bf_last(u64& x, u64& y, ulonglong2 w, u64 q, u64 q2, u64 q4) {
    if (LAZY) {
        const u64 t = mul_shoup_lazy(y, w.x, w.y, q);
        y = x - t + q2;
        x = x + t;
    } else {
        ct_butterfly(x, y, w, q, q2);
    }
}
// x < 2^64 -> [0, q);  r1 = floor(2^64/q) < 2^32
__device__ __forceinline__ u64 reduce_small_ratio(u64 x, u64 q, u32 r1) {
    const u64 h = ((u64)(u32)(x >> 32) * r1 + (((u64)(u32)x * r1) >> 32)) >> 32;
    u64 r = x - h * q;            // in [0, 2q)
    return r >= q ? r - q : r;
}


constexpr int WB = 4;   // warps (= chunks) per pass-B CTA

__device__ __forceinline__ int swz(int x) { return x ^ (((x >> 4) & 7) | ((x >> 2) & 8)); }

// STORE = false leaves the chunk in shared memory (element x at s[swz(x)]) for a fused consumer
template <bool LAZY, bool STORE = true, typename PTR = u64*>
__device__ __forceinline__ void fwd_b2_body(PTR __restrict__ base, u64* __restrict__ s,
                                            const ulonglong2* __restrict__ tw, u64 q, int sA, int gc, int j,
                                            bool split) {
    const u64 q2 = q << 1, q4 = q << 2;
    u64 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = base[j + 32 * k];
    int m = 1 << sA;
    {   // x = j + 32k ; t = 128, 64, 32
        ulonglong2 w = tw[m + gc];
#pragma unroll
        for (int k = 0; k < 4; k++) fwd_bf<LAZY>(v[k], v[k + 4], w, q, q2, q4);
        m <<= 1;
        ulonglong2 w0 = tw[m + 2 * gc], w1 = tw[m + 2 * gc + 1];
        fwd_bf<LAZY>(v[0], v[2], w0, q, q2, q4), fwd_bf<LAZY>(v[1], v[3], w0, q, q2, q4);
        fwd_bf<LAZY>(v[4], v[6], w1, q, q2, q4), fwd_bf<LAZY>(v[5], v[7], w1, q, q2, q4);
        m <<= 1;
#pragma unroll
        for (int k = 0; k < 4; k++) fwd_bf<LAZY>(v[2 * k], v[2 * k + 1], tw[m + 4 * gc + k], q, q2, q4);
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
            u64* e = v + 4 * qd;
#pragma unroll
            for (int i = 0; i < 4; i++) e[i] = s[swz(32 * blk + b + 8 * i)];
            ulonglong2 w = tw[m + 8 * gc + blk];
            fwd_bf<LAZY>(e[0], e[2], w, q, q2, q4), fwd_bf<LAZY>(e[1], e[3], w, q, q2, q4);
            ulonglong2 wa = tw[2 * m + 16 * gc + 2 * blk], wb = tw[2 * m + 16 * gc + 2 * blk + 1];
            fwd_bf<LAZY>(e[0], e[1], wa, q, q2, q4), fwd_bf<LAZY>(e[2], e[3], wb, q, q2, q4);
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
        ulonglong2 w = tw[m + 32 * gc + j];
#pragma unroll
        for (int k = 0; k < 4; k++) fwd_bf<LAZY>(v[k], v[k + 4], w, q, q2, q4);
        m <<= 1;
        ulonglong2 w0 = tw[m + 64 * gc + 2 * j], w1 = tw[m + 64 * gc + 2 * j + 1];
        fwd_bf<LAZY>(v[0], v[2], w0, q, q2, q4), fwd_bf<LAZY>(v[1], v[3], w0, q, q2, q4);
        fwd_bf<LAZY>(v[4], v[6], w1, q, q2, q4), fwd_bf<LAZY>(v[5], v[7], w1, q, q2, q4);
        m <<= 1;
#pragma unroll
        for (int k = 0; k < 4; k++) fwd_bf_last<LAZY>(v[2 * k], v[2 * k + 1], tw[m + 128 * gc + 4 * j + k], q, q2, q4);
        if (LAZY) {
            const u32 r1 = (u32)(0xFFFFFFFFFFFFFFFFull / q);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const u64 x = reduce_small_ratio(v[i], q, r1);
                s[swz(8 * j + i)] = split ? split30(x) : x;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                u64 x = v[i];
                x = x >= q2 ? x - q2 : x;
                x = x >= q ? x - q : x;
                s[swz(8 * j + i)] = split ? split30(x) : x;
            }
        }
    }
    if constexpr (STORE) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) base[j + 32 * k] = s[swz(j + 32 * k)];
    }
}


// First eight stages of the inverse transform (gaps 1 .. 128) of one 256-coefficient chunk by one warp.  On entry the
// chunk sits in s (element x at s[swz(x)]); on exit lane j holds elements j + 32k in v[k], in [0, 2q).
__device__ __forceinline__ void inv_b2_body(u64* __restrict__ s, const ulonglong2* __restrict__ tw, u64 q, int n, int gc,
                                            int j, u64 (&v)[8]) {
    const u64 q2 = q << 1;
    {   // x = 8j + i ; t = 1, 2, 4
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = s[swz(8 * j + i)];
        int h = n >> 1;
#pragma unroll
        for (int k = 0; k < 4; k++) gs_butterfly(v[2 * k], v[2 * k + 1], tw[h + 128 * gc + 4 * j + k], q, q2);
        h >>= 1;
        ulonglong2 w0 = tw[h + 64 * gc + 2 * j], w1 = tw[h + 64 * gc + 2 * j + 1];
        gs_butterfly(v[0], v[2], w0, q, q2), gs_butterfly(v[1], v[3], w0, q, q2);
        gs_butterfly(v[4], v[6], w1, q, q2), gs_butterfly(v[5], v[7], w1, q, q2);
        h >>= 1;
        ulonglong2 w = tw[h + 32 * gc + j];
#pragma unroll
        for (int k = 0; k < 4; k++) gs_butterfly(v[k], v[k + 4], w, q, q2);
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
            u64* e = v + 4 * qd;
#pragma unroll
            for (int i = 0; i < 4; i++) e[i] = s[swz(32 * blk + b + 8 * i)];
            ulonglong2 wa = tw[h8 + 16 * gc + 2 * blk], wb = tw[h8 + 16 * gc + 2 * blk + 1];
            gs_butterfly(e[0], e[1], wa, q, q2), gs_butterfly(e[2], e[3], wb, q, q2);
            ulonglong2 w = tw[h16 + 8 * gc + blk];
            gs_butterfly(e[0], e[2], w, q, q2), gs_butterfly(e[1], e[3], w, q, q2);
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
        for (int k = 0; k < 4; k++) gs_butterfly(v[2 * k], v[2 * k + 1], tw[h + 4 * gc + k], q, q2);
        h >>= 1;
        ulonglong2 w0 = tw[h + 2 * gc], w1 = tw[h + 2 * gc + 1];
        gs_butterfly(v[0], v[2], w0, q, q2), gs_butterfly(v[1], v[3], w0, q, q2);
        gs_butterfly(v[4], v[6], w1, q, q2), gs_butterfly(v[5], v[7], w1, q, q2);
        h >>= 1;
        ulonglong2 w = tw[h + gc];
#pragma unroll
        for (int k = 0; k < 4; k++) gs_butterfly(v[k], v[k + 4], w, q, q2);
    }
}

// PRELOADED: v already holds the thread's eight first-round values (rows g % gbot + (g / gbot) * 8 gbot + k * gbot,
// gbot = R >> min(3, SA)) -- the fused ModUp kernel computes them in place instead of loading them.
template <int SA, bool LAZY, bool PRELOADED = false>
__device__ __forceinline__ void fwd_a2_body(u64* __restrict__ base, u64* __restrict__ sm,
                                            const ulonglong2* __restrict__ tw, u64 q, int S, int c, int g,
                                            u64 (&v)[8]) {
    constexpr int R = 1 << SA;
    const u64 q2 = q << 1, q4 = q << 2;
#pragma unroll
    for (int s0 = 0; s0 < SA; s0 += 3) {
        const int ns = (SA - s0) < 3 ? (SA - s0) : 3;
        const int gbot = R >> (s0 + ns);                 // smallest row gap of this round
        const int rowbase = (g / gbot) * (8 * gbot) + (g % gbot);
        if (s0 == 0) {
            if (!PRELOADED) {
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = base[(size_t)(rowbase + k * gbot) * S];
            }
        } else {
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
                        const ulonglong2 w = tw[(1 << st) + (r >> (SA - st))];
                        if (SA == 8 && st == SA - 1) fwd_bf_last<LAZY>(v[k], v[k + kgap], w, q, q2, q4);   // 8 stages: keep < 31q
                        else fwd_bf<LAZY>(v[k], v[k + kgap], w, q, q2, q4);
                    }
                }
            }
        }
        if (s0 + 3 >= SA) {
            if (LAZY) {   // hand pass B canonical residues
                const u32 r1 = (u32)(0xFFFFFFFFFFFFFFFFull / q);
#pragma unroll
                for (int k = 0; k < 8; k++) base[(size_t)(rowbase + k * gbot) * S] = reduce_small_ratio(v[k], q, r1);
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++) base[(size_t)(rowbase + k * gbot) * S] = v[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) sm[(rowbase + k * gbot) * COLS + c] = v[k];
            __syncthreads();
        }
    }
}

// The SA remaining stages of the inverse transform on a [R][16] column tile.  Returns with the thread's eight values of
// the LAST round in v (in [0, 2q), n^-1 not applied yet): rows rowbase + k * gs with gs = 2^(SA-3) -- the very rows the
// forward pass's first round starts from (gbot = R >> 3), which is what lets the fused ModUp kernel go straight on.
template <int SA>
__device__ __forceinline__ int inv_a2_rounds(const u64* __restrict__ base, u64* __restrict__ sm,
                                             const ulonglong2* __restrict__ tw, u64 q, int S, int c, int g, u64 (&v)[8]) {
    constexpr int R = 1 << SA;
    const u64 q2 = q << 1;
    // a partial round (SA % 3 stages) comes first, where the 8 rows of a thread are contiguous
    constexpr int NS0 = (SA % 3) ? (SA % 3) : 3;
    int last_rowbase = 0;
#pragma unroll
    for (int u0 = 0; u0 < SA; u0 += (u0 == 0 ? NS0 : 3)) {
        const int ns = u0 == 0 ? NS0 : 3;
        const int gs = 1 << u0;                           // smallest row gap of this round
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
                        gs_butterfly(v[k], v[k + kgap], tw[(R >> (u + 1)) + (r >> (u + 1))], q, q2);
                    }
                }
            }
        }
        if (u0 + ns >= SA) {
            last_rowbase = rowbase;
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) sm[(rowbase + k * gs) * COLS + c] = v[k];
            __syncthreads();
        }
    }
    return last_rowbase;
}


}  // namespace nttc
