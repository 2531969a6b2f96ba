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
#include <cstdlib>

#include "engine.h"
#include "ntt_core.cuh"
#include "ops.h"
#include "tma.cuh"

namespace {

using namespace nttc;

constexpr int TPB = 256;
constexpr int B_ELEMS = 2048;   // pass B elements per CTA

// Rows that ModUp already delivered in NTT form (the data limbs of a digit inside that digit's own block of rows) are
// skipped.  skip = alpha | (beta << 16): digit width, and -- for batches holding several decompositions back to back
// ([group][digit][row]) -- the digits per decomposition (0: the batch is a single decomposition).
__device__ __forceinline__ bool own_digit_row(int skip, int limb, int row, const RowMap& rm) {
    const int alpha = skip & 0xFFFF, beta = skip >> 16;
    if (!alpha || limb >= rm.L) return false;
    int digit = row / rm.rpp;
    if (beta) digit %= beta;
    return limb / alpha == digit;
}

// ---- forward, pass A: stages 0..sA-1 on a [2^sA][COLS] tile --------------------------------
__global__ void __launch_bounds__(TPB) ntt_fwd_a(u64* __restrict__ data, RowMap rm, NttTab tb, int N, int n,
                                                  int sA, int skip_alpha) {
    extern __shared__ u64 sm[];
    const int row = blockIdx.y;
    const int limb = rm.limb(row);
    if (own_digit_row(skip_alpha, limb, row, rm)) return;
    const u64 q = tb.q[limb], q2 = q << 1;
    const ulonglong2* __restrict__ tw = tb.psi + (size_t)limb * N;
    u64* base = data + rm.offset(row, n);
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
                                                  int sA, int sB, int skip_alpha, int split) {
    extern __shared__ u64 sm[];
    const int row = blockIdx.y;
    const int limb = rm.limb(row);
    if (own_digit_row(skip_alpha, limb, row, rm)) return;
    const u64 q = tb.q[limb], q2 = q << 1;
    const ulonglong2* __restrict__ tw = tb.psi + (size_t)limb * N;
    const int M = 1 << sB;
    const int elems = n < B_ELEMS ? n : B_ELEMS;
    u64* base = data + rm.offset(row, n) + (size_t)blockIdx.x * elems;
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
        v = v >= q ? v - q : v;
        base[e] = split ? split30(v) : v;
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
    u64* base = data + rm.offset(row, n) + (size_t)blockIdx.x * elems;
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
    u64* base = data + rm.offset(row, n);
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


// =============================================================================================
// v2 kernels (n >= 2048): every thread keeps 8 coefficients in registers and runs up to three
// butterfly stages per shared-memory exchange.
//   pass B: one warp owns one 256-coefficient chunk (8 stages = 3 + 2 + 3), all exchanges are
//           intra-warp (__syncwarp only) through an XOR-swizzled, bank-conflict-free layout;
//           global loads and stores are fully coalesced (lane j touches j + 32k).
//   pass A: a CTA owns a [2^SA rows][16 columns] tile; lanes run along the columns so every
//           shared/global access is a contiguous 128-byte row segment.
// =============================================================================================
__global__ void __launch_bounds__(WB * 32) ntt_fwd_b2(u64* __restrict__ data, RowMap rm, NttTab tb, int N, int n,
                                                       int sA, int skip_alpha, int split, int row_lo, int row_hi) {
    __shared__ u64 smem[WB][256];
    const int row = blockIdx.y;
    const int limb = rm.limb(row);
    if (own_digit_row(skip_alpha, limb, row, rm)) return;
    const int rr = row % rm.rpp;                 // row within its polynomial: launches restricted to a row range skip the rest
    if (rr < row_lo || rr >= row_hi) return;
    const u64 q = tb.q[limb];
    const ulonglong2* __restrict__ tw = tb.psi + (size_t)limb * N;
    const int warp = threadIdx.x >> 5, j = threadIdx.x & 31;
    const int gc = blockIdx.x * WB + warp;
    u64* base = data + rm.offset(row, n) + (size_t)gc * 256;
    if (q < (1ull << 59) && q > (1ull << 33)) fwd_b2_body<true>(base, smem[warp], tw, q, sA, gc, j, split != 0);
    else fwd_b2_body<false>(base, smem[warp], tw, q, sA, gc, j, split != 0);
}

__global__ void __launch_bounds__(WB * 32) ntt_inv_b2(u64* __restrict__ data, RowMap rm, NttTab tb, int N, int n,
                                                       int sA, int logn) {
    __shared__ u64 smem[WB][256];
    const int row = blockIdx.y;
    const int limb = rm.limb(row);
    const u64 q = tb.q[limb];
    const ulonglong2* __restrict__ tw = tb.ipsi + (size_t)limb * N;
    const int warp = threadIdx.x >> 5, j = threadIdx.x & 31;
    const int gc = blockIdx.x * WB + warp;
    u64* base = data + rm.offset(row, n) + (size_t)gc * 256;
    u64* s = smem[warp];
    u64 v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) s[swz(j + 32 * k)] = base[j + 32 * k];
    __syncwarp();
    inv_b2_body(s, tw, q, n, gc, j, v);
    if (sA == 0) {
        ulonglong2 ninv = tb.invn[limb * 17 + logn];
#pragma unroll
        for (int k = 0; k < 8; k++) base[j + 32 * k] = mul_shoup(v[k], ninv.x, ninv.y, q);
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) base[j + 32 * k] = v[k];
    }
}

template <int SA>
__global__ void __launch_bounds__(2 << SA) ntt_fwd_a2(u64* __restrict__ data, RowMap rm, NttTab tb, int N, int n,
                                                       int skip_alpha) {
    constexpr int R = 1 << SA;
    __shared__ u64 sm[R * COLS];
    const int row = blockIdx.y;
    const int limb = rm.limb(row);
    if (own_digit_row(skip_alpha, limb, row, rm)) return;
    const u64 q = tb.q[limb];
    const ulonglong2* __restrict__ tw = tb.psi + (size_t)limb * N;
    const int c = threadIdx.x & (COLS - 1), g = threadIdx.x >> 4;
    u64* base = data + rm.offset(row, n) + blockIdx.x * COLS + c;
    u64 v[8];
    if (q < (1ull << 59) && q > (1ull << 33)) fwd_a2_body<SA, true>(base, sm, tw, q, n >> SA, c, g, v);
    else fwd_a2_body<SA, false>(base, sm, tw, q, n >> SA, c, g, v);
}

// pass A, inverse: row gaps 1, 2, ..., R/2, then n^-1
template <int SA>
__global__ void __launch_bounds__(2 << SA) ntt_inv_a2(u64* __restrict__ data, RowMap rm, NttTab tb, int N, int n,
                                                       int logn) {
    constexpr int R = 1 << SA;
    __shared__ u64 sm[R * COLS];
    const int row = blockIdx.y;
    const int limb = rm.limb(row);
    const u64 q = tb.q[limb];
    const ulonglong2* __restrict__ tw = tb.ipsi + (size_t)limb * N;
    const int c = threadIdx.x & (COLS - 1), g = threadIdx.x >> 4;
    const int S = n >> SA;
    u64* base = data + rm.offset(row, n) + blockIdx.x * COLS + c;
    const ulonglong2 ninv = tb.invn[limb * 17 + logn];
    u64 v[8];
    const int rowbase = inv_a2_rounds<SA>(base, sm, tw, q, S, c, g, v);
    constexpr int GS = SA >= 3 ? (1 << (SA - 3)) : 1;
#pragma unroll
    for (int k = 0; k < 8; k++) base[(size_t)(rowbase + k * GS) * S] = mul_shoup(v[k], ninv.x, ninv.y, q);
}

// =============================================================================================
// Decomposition front end in ONE kernel: the last SA stages of the inverse transform of the alpha source limbs of a
// digit, the scaling by n^-1 * hatinv (one merged Shoup constant), ModUp to every other limb (fast base conversion,
// split-30 Karatsuba sums exactly as ops.cu k_modup) and the first SA stages of the forward transform of each ModUp'd
// row.  A CTA owns (decomposition z, digit j, a tile of 16 columns) and walks its share of the l + P - alpha target rows.
// The coefficient-form digits never exist in memory: per decomposition this removes a write and a read of
// beta * (l+P) * N words (57 MB at C3) and two launches ("cut the giant-step chain").
//   x   [count][l][N]  rows after inverse pass B (n^-1 pending)      cin [count][l][N]  the same polynomials, NTT form
//   E   [count][beta][l+P][N]: own-digit rows = split30(cin), the others = forward pass A done, canonical residues
// =============================================================================================
template <int SA, int A>
__global__ void __launch_bounds__(2 << SA) k_intt_modup_fwd_a(const u64* __restrict__ x, const u64* __restrict__ cin,
                                                               u64* __restrict__ E, int l, int N, int L, int P, int K,
                                                               NttTab tb, ModTab mt,
                                                               const ulonglong2* __restrict__ hatinv_n,
                                                               const u64* __restrict__ hat, int sbits, int rsplit, int row_lo,
                                                               int row_hi) {
    constexpr int R = 1 << SA, T = 2 << SA;
    extern __shared__ __align__(16) u64 dyn[];
    u64* sm = dyn;                                   // [R][COLS]      exchange tile of the transforms
    u64* ysm = sm + R * COLS;                        // [A][8][T]      scaled source values, split-30, thread-private slots
    u64* tab = ysm + (size_t)A * 8 * T;              // [rows][A + 2]  hat_0..hat_{A-1}, q, floor(2^(s+64)/q) per target row
    const int rows = l + P, beta = gridDim.y / rsplit;
    const int j = blockIdx.y / rsplit, part = blockIdx.y % rsplit, z = blockIdx.z;
    const int lo = j * P, hi = min(lo + P, l), a = hi - lo;
    const int c = threadIdx.x & (COLS - 1), g = threadIdx.x >> 4, S = N >> SA;
    x += (size_t)z * l * N, cin += (size_t)z * l * N;
    u64* Ej = E + ((size_t)z * beta + j) * rows * N;
    hatinv_n += (size_t)j * P;
    hat += (size_t)j * P * K;
    for (int e = threadIdx.x; e < rows * (A + 2); e += T) {
        const int r = e / (A + 2), cc = e % (A + 2), t = r < l ? r : L + (r - l);
        tab[e] = cc < A ? (cc < a ? hat[(size_t)cc * K + t] : 0) : cc == A ? mt.q[t] : mt.rwide[t];
    }
    constexpr int GS = 1 << (SA - 3);                // row gap of the values a thread ends / starts with
    const int rowbase = (g / GS) * (8 * GS) + (g % GS);
    const size_t col = (size_t)blockIdx.x * COLS + c;
    u64 v[8];
    // 1. inverse pass A of the digit's source limbs, scaled by n^-1 * hatinv
#pragma unroll
    for (int i = 0; i < A; i++) {
        if (i < a) {
            const int limb = lo + i;
            const u64 q = tb.q[limb];
            inv_a2_rounds<SA>(x + (size_t)limb * N + col, sm, tb.ipsi + (size_t)limb * N, q, S, c, g, v);
            const ulonglong2 h = hatinv_n[i];
#pragma unroll
            for (int k = 0; k < 8; k++) ysm[((size_t)i * 8 + k) * T + threadIdx.x] = split30(mul_shoup(v[k], h.x, h.y, q));
            __syncthreads();                         // the exchange tile is free again
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) ysm[((size_t)i * 8 + k) * T + threadIdx.x] = 0;
        }
    }
    __syncthreads();                                 // tab complete (also when the digit is empty of work above)
    // 2. every target row of this CTA: ModUp in registers, then forward pass A
    int idx = 0;
    for (int r = row_lo; r < row_hi; r++) {          // (all l + P rows unless the caller serves a row range only)
        const int t = r < l ? r : L + (r - l);
        if (t >= lo && t < hi) continue;             // own-digit rows: step 3
        if (rsplit > 1 && idx++ % rsplit != part) continue;
        const u64* tr = tab + r * (A + 2);
        const u64 q = tr[A], rw = tr[A + 1];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            Acc3 acc = {0, 0, 0};
#pragma unroll
            for (int i = 0; i < A; i++) {
                const u64 y = ysm[((size_t)i * 8 + k) * T + threadIdx.x];
                mac_split(acc, y, (u32)y + (u32)(y >> 32), tr[i]);   // rows i >= a hold zeros
            }
            u64 alo = 0, ahi = 0;
            fold_split(alo, ahi, acc);
            // SA <= 7: the value goes straight into the forward pass, whose butterflies take lazy inputs -- [0, 4q) for the
            // Harvey form, and 3q + 7 * 4q = 31q < 2^64 for the fully lazy form (q < 2^59) -- so the two conditional
            // subtractions are dropped (14 of ~62 instructions per output); the pass ends in canonical residues either way
            if (SA <= 7) v[k] = reduce_wide_lazy(alo, ahi, q, rw, sbits > 0 ? sbits : 63 - __clzll((long long)q));
            else v[k] = sbits > 0 ? reduce_wide_s(alo, ahi, q, rw, sbits) : reduce_wide(alo, ahi, q, rw);
        }
        u64* base = Ej + (size_t)r * N + col;
        const ulonglong2* __restrict__ tw = tb.psi + (size_t)t * N;
        if (q < (1ull << 59) && q > (1ull << 33)) fwd_a2_body<SA, true, true>(base, sm, tw, q, S, c, g, v);
        else fwd_a2_body<SA, false, true>(base, sm, tw, q, S, c, g, v);
        __syncthreads();                             // last-round readers of the exchange tile are done
    }
    // 3. own-digit rows arrive in NTT form already: the consumers want them split-30
    if (part == 0)
        for (int i = 0; i < a; i++) {
            if (lo + i < row_lo || lo + i >= row_hi) continue;
            const size_t off = (size_t)(lo + i) * N + col;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const size_t e = off + (size_t)(rowbase + k * GS) * S;
                Ej[e] = split30(cin[e]);
            }
        }
}

// =============================================================================================
// Forward pass B fused with the key-switch inner product (giant steps, rotate, relinearize).
// One CTA owns (row r, OUTPUT chunk of 256 coefficients).  The Galois permutation of the bit-reversed NTT domain maps
// aligned blocks onto aligned blocks (common.cuh galois_src), so the digits it needs are one 256-coefficient chunk of
// every digit: warp j runs the last eight butterfly stages of digit j's chunk and leaves it in shared memory -- the
// transformed digits never travel to L2/HBM and back -- while the TMA engine fetches the [2*beta][256] box of the
// rotation key, issued before the first butterfly, so the key stream (HBM) hides behind the transform (integer pipe).
// Then every thread forms  sum_j E_j[src(n)] * key[j][p][n]  for its coefficient exactly like ks_tile_body (ops.cu).
// =============================================================================================
constexpr int FK_CHUNK = 256, FK_MAXW = 8;

// FIX8: exactly 8 digits on 8 warps (C3, C2): every loop over digits / output coefficients is a single trip, known at
// compile time (no trip-count division by blockDim.x, no loop control in the prologue and epilogue of these small CTAs)
template <int FOLD, bool FIX8>
__device__ __forceinline__ void fk_body(const CUtensorMap* kmap_p, const KsArgs& a, const ModTab& mt, const NttTab& tb,
                                        const ulonglong2* __restrict__ pmod, int sA, int alpha, int wide_ok) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const int beta = FIX8 ? 8 : a.beta;
    const int nthreads = FIX8 ? 256 : (int)blockDim.x;
    u64* ksm = reinterpret_cast<u64*>(smraw);                     // [2*beta][256]  key box (TMA destination)
    u64* esm = ksm + (size_t)2 * beta * FK_CHUNK;                 // [beta][256]    transformed digits, swizzled per chunk
    uint64_t* full = reinterpret_cast<uint64_t*>(esm + (size_t)beta * FK_CHUNK);
    const int r = blockIdx.y, t = r < a.l ? r : a.L + (r - a.l);
    const u32 n0 = blockIdx.x * FK_CHUNK;
    if (threadIdx.x == 0) {
        mbar_init(full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(full, (u32)(2 * beta * FK_CHUNK * sizeof(u64)));
        tma_load_3d_hint(ksm, kmap_p, (int)n0, t, 0, full, evict_first_policy());
    }
    const u32 csrc = (a.elt ? galois_src(n0, a.elt, a.logn) : n0) / FK_CHUNK;   // chunk the output chunk is gathered from
    const bool has_add = a.addp && r < a.add_rows;
    // the epilogue's read-modify-write operands are pulled into L2 now, a whole transform ahead of their use
    for (int o = threadIdx.x; o < FK_CHUNK; o += nthreads) {
        if (a.accumulate) {
            const u64* o0 = a.out + (size_t)r * a.N + n0 + o;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(o0));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(o0 + (size_t)a.rows * a.N));
        }
        if (has_add) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.addp + (size_t)r * a.N + (size_t)csrc * FK_CHUNK + o));
    }
    const int own = r < a.l ? r / alpha : -1;   // the digit this row belongs to (ModUp wrote its NTT form already)
    const u64 q = tb.q[t];
    const ulonglong2* __restrict__ tw = tb.psi + (size_t)t * a.N;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = nthreads >> 5;
    const bool lazy = q < (1ull << 59) && q > (1ull << 33);
    for (int j = warp; j < beta; j += nwarps) {
        u64* s = esm + (size_t)j * FK_CHUNK;
        const u64* base = a.E + ((size_t)j * a.rows + r) * a.N + (size_t)csrc * FK_CHUNK;
        if (j == own) {   // own-digit row: ModUp already wrote the NTT form (split-30)
#pragma unroll
            for (int k = 0; k < 8; k++) s[swz(lane + 32 * k)] = base[lane + 32 * k];
        } else if (lazy) {
            fwd_b2_body<true, false, const u64*>(base, s, tw, q, sA, (int)csrc, lane, true);
        } else {
            fwd_b2_body<false, false, const u64*>(base, s, tw, q, sA, (int)csrc, lane, true);
        }
    }
    __syncthreads();   // digits complete; mbarrier initialised before anybody waits on it
    const size_t es = (size_t)a.rows * a.N;
    const u64 r0 = mt.ratio0[t], r1 = mt.ratio1[t], rw = mt.rwide[t];
    ulonglong2 pm = make_ulonglong2(0, 0);
    if (has_add && a.add_pscale) pm = pmod[t];
    mbar_wait(full, 0);
    for (int o = threadIdx.x; o < FK_CHUNK; o += nthreads) {
        const u32 n = n0 + o;
        const u32 src = a.elt ? galois_src(n, a.elt, a.logn) : n;
        const int sl = swz((int)(src & (FK_CHUNK - 1)));
        u64 c0add = has_add ? a.addp[(size_t)r * a.N + src] : 0;
        u64 lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
        Acc3 acc0 = {0, 0, 0}, acc1 = {0, 0, 0};
        const u64* kcol = ksm + o;
#pragma unroll(FIX8 ? 8 : 4)
        for (int j = 0; j < beta; j++) {
            const u64 dj = esm[(size_t)j * FK_CHUNK + sl];
            const u32 dsum = (u32)dj + (u32)(dj >> 32);
            mac_split(acc0, dj, dsum, kcol[(size_t)(2 * j) * FK_CHUNK]);
            mac_split(acc1, dj, dsum, kcol[(size_t)(2 * j + 1) * FK_CHUNK]);
            if ((j + 1) % FOLD == 0) {
                fold_split(lo0, hi0, acc0);
                fold_split(lo1, hi1, acc1);
            }
        }
        fold_split(lo0, hi0, acc0);
        fold_split(lo1, hi1, acc1);
        u64 v0, v1;
        if (wide_ok) {   // q < 2^59 and <= 8 digits: the sum is below 2^(s+64)
            v0 = reduce_wide(lo0, hi0, q, rw), v1 = reduce_wide(lo1, hi1, q, rw);
        } else {
            v0 = barrett128(lo0, hi0, q, r0, r1), v1 = barrett128(lo1, hi1, q, r0, r1);
        }
        if (has_add) {
            if (a.add_pscale) c0add = mul_shoup(c0add, pm.x, pm.y, q);
            v0 = add_mod(v0, c0add, q);
        }
        u64* o0 = a.out + (size_t)r * a.N + n;
        u64* o1 = o0 + es;
        if (a.accumulate) {
            v0 = add_mod(v0, *o0, q);
            v1 = add_mod(v1, *o1, q);
        }
        *o0 = v0;
        *o1 = v1;
    }
}

template <int FOLD, bool FIX8>
__global__ void __launch_bounds__(FK_MAXW * 32, 4) k_ntt_b_ks(const __grid_constant__ CUtensorMap kmap, KsArgs a, ModTab mt,
                                                            NttTab tb, const ulonglong2* __restrict__ pmod, int sA,
                                                            int alpha, int wide_ok) {
    fk_body<FOLD, FIX8>(&kmap, a, mt, tb, pmod, sA, alpha, wide_ok);
}

// All giant steps of a mat-vec in ONE launch (blockIdx.z = giant group): group z reads its own digits and rotation key,
// adds its own permuted c0 part and writes its own partial result; ops::sum_groups adds the partial results up.
constexpr int FK_MAX_GROUPS = 224;   // 224 x (128 B tensor map + element) = 29.6 KB of the 32 KB parameter space
struct GiantTab {
    CUtensorMap map[FK_MAX_GROUPS];
    u32 elt[FK_MAX_GROUPS];
};
template <int FOLD, bool FIX8>
__global__ void __launch_bounds__(FK_MAXW * 32, 4) k_ntt_b_ks_all(const __grid_constant__ GiantTab tab, KsArgs a, ModTab mt,
                                                                NttTab tb, const ulonglong2* __restrict__ pmod, int sA,
                                                                int alpha, int wide_ok, size_t e_stride, size_t out_stride,
                                                                size_t add_stride) {
    const int z = blockIdx.z;
    a.E += (size_t)z * e_stride, a.out += (size_t)z * out_stride, a.elt = tab.elt[z];
    if (a.addp) a.addp += (size_t)z * add_stride;
    fk_body<FOLD, FIX8>(&tab.map[z], a, mt, tb, pmod, sA, alpha, wide_ok);
}

template <int SA>
void launch_fwd_a2(u64* data, int rows, RowMap rm, NttTab tb, int N, int n, int skip_alpha, cudaStream_t s) {
    dim3 grid((n >> SA) / COLS, rows);
    LAUNCH(ntt_fwd_a2<SA>, grid, 2 << SA, 0, s)(data, rm, tb, N, n, skip_alpha);
}
template <int SA>
void launch_inv_a2(u64* data, int rows, RowMap rm, NttTab tb, int N, int n, int logn, cudaStream_t s) {
    dim3 grid((n >> SA) / COLS, rows);
    LAUNCH(ntt_inv_a2<SA>, grid, 2 << SA, 0, s)(data, rm, tb, N, n, logn);
}

inline void split(int logn, int& sA, int& sB) {
    sB = logn < 8 ? logn : 8;
    sA = logn - sB;
}

}  // namespace

// grid.y carries the row index: batches beyond 65535 rows are split at polynomial boundaries
static int max_rows_per_launch(const RowMap& rm) { return 65535 / rm.rpp * rm.rpp; }

void ntt_forward(const Ctx* c, u64* data, int rows, RowMap rm, int n, cudaStream_t s, int skip_alpha, bool split30_out,
                 bool pass_a_only, bool pass_b_only, int row_lo, int row_hi) {
    int logn = 0;
    while ((1 << logn) < n) logn++;
    REQUIRE((1 << logn) == n && n <= c->N && logn <= 16 && n >= 2, "ntt: bad size %d", n);
    if (rows == 0) return;
    if (rows > 65535) {
        const int step = max_rows_per_launch(rm);
        REQUIRE(step > 0 && !skip_alpha, "ntt: %d rows per polynomial do not fit one launch", rm.rpp);   // callers keep skip batches below 65536 rows
        for (int r0 = 0; r0 < rows; r0 += step)
            ntt_forward(c, data + rm.offset(r0, n), std::min(step, rows - r0), rm, n, s, skip_alpha, split30_out,
                        pass_a_only, pass_b_only, row_lo, row_hi);
        return;
    }
    int sA, sB;
    split(logn, sA, sB);
    NttTab tb = c->ntttab();
    REQUIRE(!(pass_a_only || pass_b_only) || sA >= 3, "ntt: a single pass needs n >= 2048");
    REQUIRE((row_lo == 0 && row_hi >= rm.rpp) || (pass_b_only && sA >= 3), "ntt: a row range needs the register-tiled second pass");
    if (sA >= 3) {   // n >= 2048: register-tiled kernels
        if (!pass_b_only) {
            ProfScope ps(c, PROF_NTT_FWD_A, s);
            switch (sA) {
                case 3: launch_fwd_a2<3>(data, rows, rm, tb, c->N, n, skip_alpha, s); break;
                case 4: launch_fwd_a2<4>(data, rows, rm, tb, c->N, n, skip_alpha, s); break;
                case 5: launch_fwd_a2<5>(data, rows, rm, tb, c->N, n, skip_alpha, s); break;
                case 6: launch_fwd_a2<6>(data, rows, rm, tb, c->N, n, skip_alpha, s); break;
                case 7: launch_fwd_a2<7>(data, rows, rm, tb, c->N, n, skip_alpha, s); break;
                default: launch_fwd_a2<8>(data, rows, rm, tb, c->N, n, skip_alpha, s); break;
            }
        }
        if (!pass_a_only) {
            ProfScope ps(c, PROF_NTT_FWD_B, s);
            LAUNCH(ntt_fwd_b2, dim3(n / (256 * WB), rows), WB * 32, 0, s)(data, rm, tb, c->N, n, sA, skip_alpha, split30_out ? 1 : 0,
                                                                          row_lo, row_hi);
        }
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    ProfScope ps(c, PROF_NTT_FWD_A, s);
    if (sA > 0) {
        dim3 grid((n >> sA) / COLS, rows);
        LAUNCH(ntt_fwd_a, grid, TPB, sizeof(u64) * COLS << sA, s)(data, rm, tb, c->N, n, sA, skip_alpha);
    }
    int elems = n < B_ELEMS ? n : B_ELEMS;
    dim3 grid(n / elems, rows);
    LAUNCH(ntt_fwd_b, grid, TPB, sizeof(u64) * elems, s)(data, rm, tb, c->N, n, sA, sB, skip_alpha, split30_out ? 1 : 0);
    CUDA_CHECK(cudaGetLastError());
}

// cin [count][l][N] (NTT form) -> E [count][beta][l+P][N]: own-digit rows split-30, the other rows ModUp'd with the
// forward transform's first pass done (pass B pending: ntt_ks_fused*(pass_a_done) or ntt_forward(pass_b_only)).
// x: scratch [count][l][N].  false: the fused front end does not apply (N < 2048, P > 4, or the tile does not fit).
static size_t decompose_a_smem(const Ctx* c, int l, int sA, int A) {
    return sizeof(u64) * (((size_t)COLS << sA) + (size_t)A * 8 * (2 << sA) + (size_t)(l + c->P) * (A + 2));
}
bool ntt_decompose_a_applies(const Ctx* c, int l) {
    static const bool enabled = [] {
        const char* e = getenv("SPEAR_FUSED_MODUP");
        return !(e && e[0] == '0');
    }();
    int sA, sB;
    split(c->logn, sA, sB);
    return enabled && sA >= 3 && c->P <= 4 && decompose_a_smem(c, l, sA, c->P) <= 200 * 1024;
}
bool ntt_decompose_a(const Ctx* c, const u64* cin, int l, u64* x, u64* E, int count, cudaStream_t s, int row_lo, int row_hi) {
    if (!ntt_decompose_a_applies(c, l)) return false;
    const int N = c->N, P = c->P, rows = l + P, beta = c->digits(l);
    if (row_hi < 0 || row_hi > rows) row_hi = rows;
    REQUIRE(row_lo >= 0 && row_lo <= row_hi, "decompose: bad row range");
    int sA, sB;
    split(c->logn, sA, sB);
    REQUIRE(count >= 1 && count <= 65535, "decompose: bad batch");
    CUDA_CHECK(cudaMemcpyAsync(x, cin, sizeof(u64) * count * l * N, cudaMemcpyDeviceToDevice, s));
    ntt_inverse(c, x, count * l, RowMap{l, l, c->L, 0}, N, s, /*pass_b_only=*/true);
    // target rows of a digit are dealt to `rsplit` CTAs so that small batches still fill the GPU (the inverse part is
    // repeated by each of them: alpha of l + P - alpha row transforms)
    const int tiles = (N >> sA) / COLS, targets = std::max(1, std::min(rows - std::min(P, l), row_hi - row_lo));
    int rsplit = 1;
    while (rsplit < targets && (size_t)tiles * beta * count * rsplit < (size_t)c->sm_count * 6) rsplit++;
    const size_t smem = decompose_a_smem(c, l, sA, P);
    ProfScope ps(c, PROF_MODUP, s);
    auto go = [&](auto kern) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        LAUNCH(kern, dim3(tiles, beta * rsplit, count), 2 << sA, smem, s)(
            x, cin, E, l, N, c->L, P, c->K, c->ntttab(), c->modtab(), c->d_up_hatinv_n + (size_t)l * c->beta * P,
            c->d_up_hat + (size_t)l * c->beta * P * c->K, c->sbits, rsplit, row_lo, row_hi);
    };
#define DA_CASE(SA_)                                              \
    case SA_:                                                     \
        if (P == 1) go(k_intt_modup_fwd_a<SA_, 1>);               \
        else if (P == 2) go(k_intt_modup_fwd_a<SA_, 2>);          \
        else if (P == 3) go(k_intt_modup_fwd_a<SA_, 3>);          \
        else go(k_intt_modup_fwd_a<SA_, 4>);                      \
        break;
    switch (sA) { DA_CASE(3) DA_CASE(4) DA_CASE(5) DA_CASE(6) DA_CASE(7) default: REQUIRE(sA == 8, "decompose: bad size"); DA_CASE(8) }
#undef DA_CASE
    CUDA_CHECK(cudaGetLastError());
    return true;
}

// E: [beta][l+P][N] as ModUp leaves it (own-digit rows in NTT split-30 form, the others in coefficient form).
// Runs the forward transform of the other rows and the key inner product (KsArgs semantics of ops::ks_inner) with
// the last eight stages fused into the product.  Returns false when the fused path does not apply (N < 2048).
bool ntt_ks_fused_applies(const Ctx* c, int l) {
    int sA, sB;
    split(c->logn, sA, sB);
    return sA >= 3 && (size_t)3 * c->digits(l) * FK_CHUNK * sizeof(u64) + 64 <= 200 * 1024;
}
// pass A of `groups` decompositions stored back to back ([group][beta][l+P][N]) in one launch
void ntt_pass_a_batch(const Ctx* c, u64* E, int groups, int l, cudaStream_t s) {
    const int beta = c->digits(l), rows = l + c->P;
    ntt_forward(c, E, groups * beta * rows, RowMap{rows, l, c->L, 0}, c->N, s, c->P | (beta << 16), true, /*pass_a_only=*/true);
}
bool ntt_ks_fused(const Ctx* c, u64* E, const u64* key, u64* out, int l, u32 elt, const u64* addp, int add_rows,
                  int add_pscale, int accumulate, cudaStream_t s, bool pass_a_done) {
    const int beta = c->digits(l), rows = l + c->P;
    const size_t smem = (size_t)3 * beta * FK_CHUNK * sizeof(u64) + 64;
    int sA, sB;
    split(c->logn, sA, sB);
    if (sA < 3 || smem > 200 * 1024) return false;
    if (!pass_a_done) ntt_forward(c, E, beta * rows, RowMap{rows, l, c->L, 0}, c->N, s, c->P, true, /*pass_a_only=*/true);
    KsArgs a;
    a.E = E, a.key = key, a.out = out, a.addp = addp, a.add_rows = add_rows, a.add_pscale = add_pscale;
    a.accumulate = accumulate, a.beta = beta, a.l = l, a.rows = rows, a.N = c->N, a.logn = c->logn;
    a.L = c->L, a.K = c->K, a.elt = elt;
    CUtensorMap kmap;
    ops::encode_key_map(c, key, FK_CHUNK, beta, &kmap);
    bool small = true;
    for (u64 qq : c->q) small = small && qq < (1ull << 59);
    const int threads = 32 * std::min(beta, FK_MAXW);
    ProfScope ps(c, PROF_NTT_KS, s);
    auto go = [&](auto kern) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        LAUNCH(kern, dim3(c->N / FK_CHUNK, rows), threads, smem, s)(kmap, a, c->modtab(), c->ntttab(), c->d_pmod, sA, c->P,
                                                                    small && beta <= 8 ? 1 : 0);
    };
    if (small && beta == 8) go(k_ntt_b_ks<16, true>);
    else if (small) go(k_ntt_b_ks<16, false>);
    else go(k_ntt_b_ks<8, false>);
    CUDA_CHECK(cudaGetLastError());
    return true;
}

// groups decompositions E [groups][beta][l+P][N] (pass A done) against keys[g] / elts[g]:
// out[g] = (pi_g(addp[g]) + <pi_g(E[g]), k0_g>, <pi_g(E[g]), k1_g>)   (out: [groups][2][l+P][N], addp stride add_stride words)
bool ntt_ks_fused_all(const Ctx* c, const u64* E, const u64* const* keys, const u32* elts, int groups, u64* out, int l,
                      const u64* addp, size_t add_stride, int add_rows, cudaStream_t s) {
    const int beta = c->digits(l), rows = l + c->P;
    const size_t smem = (size_t)3 * beta * FK_CHUNK * sizeof(u64) + 64;
    int sA, sB;
    split(c->logn, sA, sB);
    if (!ntt_ks_fused_applies(c, l) || groups < 1 || groups > FK_MAX_GROUPS) return false;
    static thread_local GiantTab tab;
    for (int g = 0; g < groups; g++) {
        ops::encode_key_map(c, keys[g], FK_CHUNK, beta, &tab.map[g]);
        tab.elt[g] = elts[g];
    }
    KsArgs a;
    a.E = E, a.key = nullptr, a.out = out, a.addp = addp, a.add_rows = add_rows, a.add_pscale = 0, a.accumulate = 0;
    a.beta = beta, a.l = l, a.rows = rows, a.N = c->N, a.logn = c->logn, a.L = c->L, a.K = c->K, a.elt = 0;
    bool small = true;
    for (u64 qq : c->q) small = small && qq < (1ull << 59);
    const int threads = 32 * std::min(beta, FK_MAXW);
    const size_t pw = (size_t)rows * c->N;
    ProfScope ps(c, PROF_NTT_KS, s);
    auto go = [&](auto kern) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        LAUNCH(kern, dim3(c->N / FK_CHUNK, rows, groups), threads, smem, s)(tab, a, c->modtab(), c->ntttab(), c->d_pmod, sA, c->P,
                                                                           small && beta <= 8 ? 1 : 0, (size_t)beta * pw, 2 * pw,
                                                                           add_stride);
    };
    if (small && beta == 8) go(k_ntt_b_ks_all<16, true>);
    else if (small) go(k_ntt_b_ks_all<16, false>);
    else go(k_ntt_b_ks_all<8, false>);
    CUDA_CHECK(cudaGetLastError());
    return true;
}

void ntt_inverse(const Ctx* c, u64* data, int rows, RowMap rm, int n, cudaStream_t s, bool pass_b_only) {
    int logn = 0;
    while ((1 << logn) < n) logn++;
    REQUIRE((1 << logn) == n && n <= c->N && logn <= 16 && n >= 2, "intt: bad size %d", n);
    if (rows == 0) return;
    if (rows > 65535) {
        const int step = max_rows_per_launch(rm);
        REQUIRE(step > 0, "intt: %d rows per polynomial do not fit one launch", rm.rpp);
        for (int r0 = 0; r0 < rows; r0 += step)
            ntt_inverse(c, data + rm.offset(r0, n), std::min(step, rows - r0), rm, n, s, pass_b_only);
        return;
    }
    int sA, sB;
    split(logn, sA, sB);
    NttTab tb = c->ntttab();
    REQUIRE(!pass_b_only || sA >= 3, "intt: a single pass needs n >= 2048");
    if (sA >= 3) {
        {
            ProfScope ps(c, PROF_NTT_INV_B, s);
            LAUNCH(ntt_inv_b2, dim3(n / (256 * WB), rows), WB * 32, 0, s)(data, rm, tb, c->N, n, sA, logn);
        }
        if (pass_b_only) {
            CUDA_CHECK(cudaGetLastError());
            return;
        }
        ProfScope ps(c, PROF_NTT_INV_A, s);
        switch (sA) {
            case 3: launch_inv_a2<3>(data, rows, rm, tb, c->N, n, logn, s); break;
            case 4: launch_inv_a2<4>(data, rows, rm, tb, c->N, n, logn, s); break;
            case 5: launch_inv_a2<5>(data, rows, rm, tb, c->N, n, logn, s); break;
            case 6: launch_inv_a2<6>(data, rows, rm, tb, c->N, n, logn, s); break;
            case 7: launch_inv_a2<7>(data, rows, rm, tb, c->N, n, logn, s); break;
            default: launch_inv_a2<8>(data, rows, rm, tb, c->N, n, logn, s); break;
        }
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    ProfScope ps(c, PROF_NTT_INV_B, s);
    int elems = n < B_ELEMS ? n : B_ELEMS;
    dim3 grid(n / elems, rows);
    LAUNCH(ntt_inv_b, grid, TPB, sizeof(u64) * elems, s)(data, rm, tb, c->N, n, sA, sB, logn);
    if (sA > 0) {
        dim3 grida((n >> sA) / COLS, rows);
        LAUNCH(ntt_inv_a, grida, TPB, sizeof(u64) * COLS << sA, s)(data, rm, tb, c->N, n, sA, logn);
    }
    CUDA_CHECK(cudaGetLastError());
}
