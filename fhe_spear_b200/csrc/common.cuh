// common.cuh -- shared types and 64-bit modular arithmetic for the sm_100a CKKS engine.
//
// All residues are uint64 < q < 2^61.  Three multiplication flavours:
//   * Shoup (constant operand with precomputed floor(w*2^64/q))  -> NTT butterflies, scalar tables
//   * 128-bit lazy accumulate + one Barrett-128 reduction        -> key-switch inner product, PMAC, base conversion
//   * Barrett-64 for reducing a residue of one prime modulo another
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

typedef uint64_t u64;
typedef uint32_t u32;

#define SPEAR_MAX_LIMBS 64

// ---------------------------------------------------------------------------------------------
// error handling: every C-ABI entry point catches spear_error and returns a status code
// ---------------------------------------------------------------------------------------------
struct spear_error {
    int code;
    char msg[512];
};
void spear_throw(int code, const char* fmt, ...);

#define SPEAR_OK 0
#define SPEAR_ERR_INVALID 1
#define SPEAR_ERR_CUDA 2
#define SPEAR_ERR_OOM 3
#define SPEAR_ERR_NOKEY 4

#define CUDA_CHECK(expr)                                                                          \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            int code__ = (e__ == cudaErrorMemoryAllocation) ? SPEAR_ERR_OOM : SPEAR_ERR_CUDA;     \
            spear_throw(code__, "CUDA error: %s%s at %s:%d", cudaGetErrorString(e__),             \
                        code__ == SPEAR_ERR_OOM ? " (CUDA out of memory)" : "", __FILE__, __LINE__); \
        }                                                                                         \
    } while (0)

// every kernel launch goes through LAUNCH so spear_launch_count() reports real launches
extern unsigned long long g_spear_launches;
#define LAUNCH(kernel, ...) g_spear_launches++, kernel<<<__VA_ARGS__>>>

#define REQUIRE(cond, ...)                                   \
    do {                                                     \
        if (!(cond)) spear_throw(SPEAR_ERR_INVALID, __VA_ARGS__); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// device tables handed to kernels by value
// ---------------------------------------------------------------------------------------------
struct ModTab {
    const u64* q;          // [K]
    const u64* ratio0;     // [K] low  word of floor(2^128/q)
    const u64* ratio1;     // [K] high word of floor(2^128/q)
    const u64* rwide;      // [K] floor(2^(s+64)/q), s = bitlength(q) - 1   (reduce_wide)
};

struct NttTab {
    const ulonglong2* psi;   // [K][N]  (w, shoup(w)), w = psi^{bitrev(i)}
    const ulonglong2* ipsi;  // [K][N]  (w, shoup(w)), w = psi^{-bitrev(i)}
    const ulonglong2* invn;  // [K][17] (n^{-1}, shoup) for transform size 2^k, k = 0..16
    const u64* q;            // [K]
};

// Row r of a batch belongs to polynomial r / rpp; within the polynomial, rows < l are data limbs
// base+idx, rows >= l are the special limbs L + (idx - l).
struct RowMap {
    int rpp;   // rows per polynomial
    int l;     // data rows per polynomial
    int L;     // index of the first special limb
    int base;  // limb id of data row 0
    size_t pstride;   // words between consecutive polynomials (0: rows are contiguous, rpp * n)
    __host__ __device__ __forceinline__ int limb(int row) const {
        int idx = row % rpp;
        return idx < l ? base + idx : L + (idx - l);
    }
    // word offset of row `row` of a batch whose rows hold n coefficients
    __host__ __device__ __forceinline__ size_t offset(int row, int n) const {
        return pstride ? (size_t)(row / rpp) * pstride + (size_t)(row % rpp) * n : (size_t)row * n;
    }
};

// ---------------------------------------------------------------------------------------------
// arithmetic
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 add_mod(u64 a, u64 b, u64 q) {
    u64 s = a + b;
    return s >= q ? s - q : s;
}
__device__ __forceinline__ u64 sub_mod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
__device__ __forceinline__ u64 neg_mod(u64 a, u64 q) { return a ? q - a : 0; }

// a*w mod q in [0, 2q), wp = floor(w * 2^64 / q), any a < 2^64
__device__ __forceinline__ u64 mul_shoup_lazy(u64 a, u64 w, u64 wp, u64 q) {
    u64 h = __umul64hi(a, wp);
    return a * w - h * q;
}
__device__ __forceinline__ u64 mul_shoup(u64 a, u64 w, u64 wp, u64 q) {
    u64 r = mul_shoup_lazy(a, w, wp, q);
    return r >= q ? r - q : r;
}
// Harvey lazy butterflies shared by the transforms (ntt.cu) and the one-kernel ring encoder (encoder.cu)
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

// x mod q for any x < 2^64; r1 = floor(2^64 / q)
__device__ __forceinline__ u64 barrett64(u64 x, u64 q, u64 r1) {
    u64 h = __umul64hi(x, r1);
    u64 r = x - h * q;
    return r >= q ? r - q : r;
}
// (hi:lo) mod q, (r1:r0) = floor(2^128 / q)
__device__ __forceinline__ u64 barrett128(u64 lo, u64 hi, u64 q, u64 r0, u64 r1) {
    u64 carry = __umul64hi(lo, r0);
    u64 t2lo = lo * r1, t2hi = __umul64hi(lo, r1);
    u64 tmp1 = t2lo + carry;
    u64 tmp3 = t2hi + (tmp1 < t2lo);
    t2lo = hi * r0;
    t2hi = __umul64hi(hi, r0);
    tmp1 += t2lo;
    carry = t2hi + (tmp1 < t2lo);
    tmp1 = hi * r1 + tmp3 + carry;
    u64 r = lo - tmp1 * q;
    return r >= q ? r - q : r;
}
// (hi:lo) mod q for (hi:lo) < 2^(s+64), s = bitlength(q) - 1 -- e.g. a sum of up to 8 products of residues.
// One 64x64 high product on the top 64 significant bits (R = floor(2^(s+64)/q)): the quotient estimate is at
// most 2 short, so two conditional subtractions finish the job.  About half the instructions of barrett128.
__device__ __forceinline__ u64 reduce_wide(u64 lo, u64 hi, u64 q, u64 R) {
    const int s = 63 - __clzll((long long)q);
    const u64 top = (hi << (64 - s)) | (lo >> s);
    const u64 Q = __umul64hi(top, R);
    u64 r = lo - Q * q;
    const u64 q2 = q << 1;
    r = r >= q2 ? r - q2 : r;
    return r >= q ? r - q : r;
}
// same with s supplied (all moduli of a parameter set usually share their bit length: uniform shifts, no FLO)
__device__ __forceinline__ u64 reduce_wide_s(u64 lo, u64 hi, u64 q, u64 R, int s) {
    const u64 top = (hi << (64 - s)) | (lo >> s);
    const u64 Q = __umul64hi(top, R);
    u64 r = lo - Q * q;
    const u64 q2 = q << 1;
    r = r >= q2 ? r - q2 : r;
    return r >= q ? r - q : r;
}
// the same without the two conditional subtractions: a representative in [0, 3q) -- for consumers that take lazy inputs
// (the forward butterflies accept [0, 4q); the fully lazy passes of ntt_core.cuh have the head-room up to seven stages)
__device__ __forceinline__ u64 reduce_wide_lazy(u64 lo, u64 hi, u64 q, u64 R, int s) {
    const u64 top = (hi << (64 - s)) | (lo >> s);
    const u64 Q = __umul64hi(top, R);
    return lo - Q * q;
}
__device__ __forceinline__ u64 mul_mod(u64 a, u64 b, u64 q, u64 r0, u64 r1) {
    return barrett128(a * b, __umul64hi(a, b), q, r0, r1);
}
// (hi:lo) += a*b
__device__ __forceinline__ void mac128(u64& lo, u64& hi, u64 a, u64 b) {
    asm("mad.lo.cc.u64 %0, %2, %3, %0;\n\t"
        "madc.hi.u64 %1, %2, %3, %1;"
        : "+l"(lo), "+l"(hi)
        : "l"(a), "l"(b));
}
// (hi:lo) += x
__device__ __forceinline__ void add128(u64& lo, u64& hi, u64 x) {
    asm("add.cc.u64 %0, %0, %2;\n\t"
        "addc.u64 %1, %1, 0;"
        : "+l"(lo), "+l"(hi)
        : "l"(x));
}

// ---- split-30 operands: x = x1 * 2^30 + x0 stored as (x1 << 32) | x0, x0 < 2^30, x1 < 2^30 (q < 2^60) ----
// Products of 30-bit halves are < 2^60, so IMAD.WIDE can accumulate them in plain 64-bit registers
// with no carry handling: S0 += a0*b0, S1 += a0*b1 + a1*b0, S2 += a1*b1; the three sums are folded
// into a 128-bit accumulator every FOLD terms (8 for q < 2^60, 16 for q < 2^59).
__device__ __forceinline__ u64 split30(u64 x) { return ((x >> 30) << 32) | (x & 0x3FFFFFFFull); }
__device__ __forceinline__ u64 unsplit30(u64 s) { return ((s >> 32) << 30) | (s & 0xFFFFFFFFull); }
struct Acc3 {
    u64 s0, s1, s2;   // s1 holds the Karatsuba middle sum  sum (x0+x1)(y0+y1)  (mod 2^64)
};
// three 32x32+64 multiply-adds per term (IMAD.WIDE is the scarce pipe): xs = x0 + x1 is supplied by the caller
__device__ __forceinline__ u64 mul_wide(u32 a, u32 b) {
    u64 r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
// high word of a 32x32 product through the wide multiplier (kept as IMAD.WIDE by ptxas, not IMAD.HI)
__device__ __forceinline__ u32 hi_of_wide(u32 a, u32 b) {
    u32 hi;
    asm("{ .reg .u64 t; .reg .u32 lo; mul.wide.u32 t, %1, %2; mov.b64 {lo, %0}, t; }" : "=r"(hi) : "r"(a), "r"(b));
    return hi;
}
__device__ __forceinline__ void mac_split(Acc3& a, u64 x, u32 xs, u64 y) {
    // explicit PTX: written as (u64)x0 * y0 in C, the front end widens the products to 64-bit multiplies of masked
    // operands and ptxas then emits a stray 32-bit add per product; mul.wide + add folds to one IMAD.WIDE each
    u32 x0, x1, y0, y1;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(x0), "=r"(x1) : "l"(x));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(y0), "=r"(y1) : "l"(y));
    a.s0 += mul_wide(x0, y0);
    a.s2 += mul_wide(x1, y1);
    a.s1 += mul_wide(xs, y0 + y1);
}
// (hi:lo) += s0 + (s1 - s0 - s2) * 2^30 + s2 * 2^60 ; clears the partial sums.  The true middle sum is < 2^64,
// so the wrapped subtraction is exact.
__device__ __forceinline__ void fold_split(u64& lo, u64& hi, Acc3& a) {
    const u64 m = a.s1 - a.s0 - a.s2;
    asm("add.cc.u64 %0, %0, %2;\n\t"
        "addc.u64 %1, %1, 0;\n\t"
        "add.cc.u64 %0, %0, %3;\n\t"
        "addc.u64 %1, %1, %4;\n\t"
        "add.cc.u64 %0, %0, %5;\n\t"
        "addc.u64 %1, %1, %6;"
        : "+l"(lo), "+l"(hi)
        : "l"(a.s0), "l"(m << 30), "l"(m >> 34), "l"(a.s2 << 60), "l"(a.s2 >> 4));
    a.s0 = a.s1 = a.s2 = 0;
}

// Galois automorphism x -> x^elt in the bit-reversed NTT domain:
// out[i] = in[ bitrev( ((elt * (2*bitrev(i)+1) mod 2N) - 1) / 2 ) ].  Aligned blocks of 2^s
// consecutive indices map onto aligned blocks, so a warp's gather stays inside one 256-byte line pair.
__device__ __forceinline__ u32 galois_src(u32 i, u32 elt, int logn) {
    u32 br = __brev(i) >> (32 - logn);
    u32 k = (elt * (2u * br + 1u)) & ((2u << logn) - 1u);
    return __brev((k - 1u) >> 1) >> (32 - logn);
}

// streaming (read-once) loads: bypass L1 and mark the line evict-first in L2 so that rotation keys
// and diagonals flowing through do not displace the reused digits / baby ciphertexts
__device__ __forceinline__ u64 evict_first_policy() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ u64 ld_stream(const u64* p, u64 pol) {
    u64 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}
