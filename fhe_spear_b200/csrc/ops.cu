// ops.cu -- RNS kernels of the key-switch / BSGS path: element-wise ops, Galois permutation,
// ModUp, key inner product, ModDown, rescale, plaintext-diagonal MAC.
//
// Every kernel is HBM/L2-bound integer work: one thread per coefficient (or coefficient pair),
// consecutive threads on consecutive coefficients (coalesced), read-once operands (rotation keys,
// diagonals) loaded with L1::no_allocate/L2::evict_first so the reused operands (decomposed digits,
// baby ciphertexts) stay L2-resident.  Sums of products are accumulated lazily in 128 bits and
// reduced once (Barrett-128).
#include <cstdlib>
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "engine.h"
#include "ops.h"
#include "tma.cuh"

namespace {

constexpr int TPB = 256;
constexpr int MAX_ALPHA = 8;


// ---- element-wise -------------------------------------------------------------------------
enum { EW_ADD = 0, EW_SUB = 1, EW_NEG = 2, EW_MUL = 3 };

// out[p][r][i] = a[p][r][i] (op) b[(bcast ? 0 : p)][r][i % bn]
template <int OP>
__global__ void k_ew(const u64* __restrict__ a, const u64* __restrict__ b, u64* __restrict__ out, int polys, int rows,
                     int n, RowMap rm, ModTab mt, int b_polys) {
    size_t total = (size_t)polys * rows * n;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        int i = (int)(e % n);
        size_t pr = e / n;
        int r = (int)(pr % rows), p = (int)(pr / rows);
        int limb = rm.limb(r);
        u64 q = mt.q[limb];
        u64 x = a[e], v;
        if (OP == EW_NEG) {
            v = neg_mod(x, q);
        } else {
            u64 y = b[((size_t)(p < b_polys ? p : 0) * rows + r) * n + i];
            if (OP == EW_ADD) v = add_mod(x, y, q);
            else if (OP == EW_SUB) v = sub_mod(x, y, q);
            else v = mul_mod(x, y, q, mt.ratio0[limb], mt.ratio1[limb]);
        }
        out[e] = v;
    }
}

// x[p][r][i] = x[p][r][i] mod q_r  for lazily summed residues (e.g. a u64 all-reduce of <= 16 shards)
__global__ void k_reduce(u64* __restrict__ x, int polys, int rows, int n, RowMap rm, ModTab mt) {
    size_t total = (size_t)polys * rows * n;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        int limb = rm.limb((int)((e / n) % rows));
        x[e] = barrett64(x[e], mt.q[limb], mt.ratio1[limb]);
    }
}

// out[e] = (first ? first[e] : 0) + sum_k parts[k][e]  (mod q of the row): the partial results of the giant steps
__global__ void k_sum_groups(const u64* __restrict__ parts, int nparts, size_t part_words, const u64* first, u64* out,
                             int rows, int n, RowMap rm, ModTab mt) {   // first may alias out (element-wise in place)
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < part_words; e += (size_t)gridDim.x * blockDim.x) {
        const u64 q = mt.q[rm.limb((int)((e / n) % rows))];
        u64 acc = first ? first[e] : 0;
        for (int k = 0; k < nparts; k++) acc = add_mod(acc, parts[(size_t)k * part_words + e], q);
        out[e] = acc;
    }
}

// tensor product of two size-2 ciphertexts -> size 3
__global__ void k_tensor(const u64* __restrict__ a, const u64* __restrict__ b, u64* __restrict__ out, int l, int n,
                         ModTab mt) {
    size_t total = (size_t)l * n, pw = total;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        int limb = (int)(e / n);
        u64 q = mt.q[limb], r0 = mt.ratio0[limb], r1 = mt.ratio1[limb];
        u64 a0 = a[e], a1 = a[pw + e], b0 = b[e], b1 = b[pw + e];
        out[e] = mul_mod(a0, b0, q, r0, r1);
        u64 lo = 0, hi = 0;
        mac128(lo, hi, a0, b1);
        mac128(lo, hi, a1, b0);
        out[pw + e] = barrett128(lo, hi, q, r0, r1);
        out[2 * pw + e] = mul_mod(a1, b1, q, r0, r1);
    }
}

// out[row][i] = in[row][galois_src(i)]
__global__ void k_galois(const u64* __restrict__ in, u64* __restrict__ out, int rows, int N, int logn, u32 elt) {
    size_t total = (size_t)rows * N;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        u32 i = (u32)(e & (N - 1));
        out[e] = in[e - i + galois_src(i, elt, logn)];
    }
}

// ---- ModUp --------------------------------------------------------------------------------
// x: [l][N] coefficient form; cin: [l][N] the same polynomial in NTT form (copied into own-digit rows)
// E: [beta][rows][N]; rows != own are left in coefficient form for the batched NTT that follows.
// hat tables are stored split-30, so the <= P products per target accumulate carry-free (mac_split).
template <int A>
__global__ void __launch_bounds__(TPB) k_modup(const u64* __restrict__ x, const u64* __restrict__ cin,
                                                u64* __restrict__ E, int l, int N, int L, int P, int K, ModTab mt,
                                                const ulonglong2* __restrict__ hatinv, const u64* __restrict__ hat,
                                                int sbits) {
    // per-row constants of this digit (hat_0..hat_{A-1}, q, floor(2^(s+64)/q)) staged once per CTA: the row loop
    // then issues shared-memory broadcasts instead of five global loads with 64-bit address arithmetic per row
    extern __shared__ u64 tab[];   // [rows][A + 2]
    const int j = blockIdx.y, n = blockIdx.x * TPB + threadIdx.x;
    const int rows = l + P, lo = j * P, hi = min(lo + P, l), a = hi - lo;
    // blockIdx.z: one of several decompositions laid out back to back (x, cin: [z][l][N]; E: [z][beta][rows][N])
    x += (size_t)blockIdx.z * l * N, cin += (size_t)blockIdx.z * l * N, E += (size_t)blockIdx.z * gridDim.y * rows * N;
    hatinv += (size_t)j * P;
    hat += (size_t)j * P * K;
    for (int e = threadIdx.x; e < rows * (A + 2); e += TPB) {
        const int r = e / (A + 2), c = e % (A + 2), t = r < l ? r : L + (r - l);
        tab[e] = c < A ? (c < a ? hat[(size_t)c * K + t] : 0) : c == A ? mt.q[t] : mt.rwide[t];
    }
    u64 y[A];
    u32 ys[A];
#pragma unroll
    for (int i = 0; i < A; i++) {
        y[i] = 0, ys[i] = 0;
        if (i < a) {
            ulonglong2 h = hatinv[i];
            y[i] = split30(mul_shoup(x[(size_t)(lo + i) * N + n], h.x, h.y, mt.q[lo + i]));
            ys[i] = (u32)y[i] + (u32)(y[i] >> 32);
        }
    }
    __syncthreads();
    u64* Ej = E + (size_t)j * rows * N + n;
    for (int r = 0; r < rows; r++) {
        const int t = r < l ? r : L + (r - l);
        if (t >= lo && t < hi) {
            Ej[(size_t)r * N] = split30(cin[(size_t)t * N + n]);   // digits are consumed in split-30 form
            continue;
        }
        const u64* tr = tab + r * (A + 2);
        Acc3 acc = {0, 0, 0};
#pragma unroll
        for (int i = 0; i < A; i++) mac_split(acc, y[i], ys[i], tr[i]);   // rows i >= a hold zeros
        u64 alo = 0, ahi = 0;
        fold_split(alo, ahi, acc);
        const u64 q = tr[A];
        Ej[(size_t)r * N] = sbits > 0 ? reduce_wide_s(alo, ahi, q, tr[A + 1], sbits) : reduce_wide(alo, ahi, q, tr[A + 1]);
    }
}

// ---- key-switch inner product -----------------------------------------------------------------
// out[p][r][n] (+)= sum_j E[j][r][src(n)] * key[j][p][limb(r)][n]  (+ addp[r][src(n)] (* P) into p = 0)
__global__ void __launch_bounds__(TPB) k_ks_inner(KsArgs a, ModTab mt, const ulonglong2* __restrict__ pmod) {
    const int r = blockIdx.y, n = blockIdx.x * TPB + threadIdx.x;
    const int t = r < a.l ? r : a.L + (r - a.l);
    const u32 src = a.elt ? galois_src((u32)n, a.elt, a.logn) : (u32)n;
    const u64 q = mt.q[t], r0 = mt.ratio0[t], r1 = mt.ratio1[t];
    const u64* e = a.E + (size_t)r * a.N + src;
    const u64* k = a.key + (size_t)t * a.N + n;
    const size_t es = (size_t)a.rows * a.N, ks = (size_t)a.K * a.N;
    u64 lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
    const u64 pol = evict_first_policy();
#pragma unroll 4
    for (int j = 0; j < a.beta; j++) {
        u64 d = unsplit30(e[j * es]);
        u64 k0 = unsplit30(ld_stream(k + (size_t)(2 * j) * ks, pol)), k1 = unsplit30(ld_stream(k + (size_t)(2 * j + 1) * ks, pol));
        mac128(lo0, hi0, d, k0);
        mac128(lo1, hi1, d, k1);
    }
    u64 v0 = barrett128(lo0, hi0, q, r0, r1), v1 = barrett128(lo1, hi1, q, r0, r1);
    if (a.addp && r < a.add_rows) {
        u64 c = a.addp[(size_t)r * a.N + src];
        if (a.add_pscale) {
            ulonglong2 pm = pmod[t];
            c = mul_shoup(c, pm.x, pm.y, q);
        }
        v0 = add_mod(v0, c, q);
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

// TMA-staged variant (keys stored split-30): one CTA = (row r, 256 coefficients).  The whole key tile of the
// row -- 2*beta rows of 2 KB, one 3-D box [2*beta][1][256] -- is fetched by a single UTMALDG with an
// L2 evict-first policy while the threads gather their digit values (L2-resident, evict-last); two to
// three CTAs per SM keep ~100 KB of key bytes in flight per SM.
constexpr int KS_TILE = 128;
// BETA > 0: digit count known at compile time (fully unrolled, no predication); BETA == 0: runtime loop.
// The digits E arrive in split-30 form (written that way by ModUp / the forward NTT), like the keys.
// One CTA walks KS_TPC consecutive 256-coefficient tiles of one (rotation, row) with a two-stage TMA ring:
// the key box of tile i+2 is in flight while tile i is multiplied, so every SM keeps several 32 KB boxes
// outstanding all the time (3 CTAs x 2 stages per SM).
constexpr int KS_TPC = 8;
template <int FOLD, int BETA>
__device__ __forceinline__ void ks_tile_body(const CUtensorMap* kmap, const KsArgs& a, const ModTab& mt,
                                             const ulonglong2* __restrict__ pmod, int r, int tile0, u32 elt,
                                             u64* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const int beta = BETA ? BETA : a.beta;
    const size_t stage_words = (size_t)2 * beta * KS_TILE;
    u64* ksm = reinterpret_cast<u64*>(smraw);                       // [2 stages][2*beta][KS_TILE]
    uint64_t* full = reinterpret_cast<uint64_t*>(ksm + 2 * stage_words);
    const int t = r < a.l ? r : a.L + (r - a.l);
    const int ntiles = min(KS_TPC, min(a.N / KS_TILE, a.tile1) - tile0);
    const u32 stage_bytes = (u32)(stage_words * sizeof(u64));
    u64 pol_first = 0;
    if (threadIdx.x == 0) {
        pol_first = evict_first_policy();
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < 2 && i < ntiles; i++) {
            mbar_expect_tx(&full[i], stage_bytes);
            tma_load_3d_hint(ksm + i * stage_words, kmap, (tile0 + i) * KS_TILE, t, 0, &full[i], pol_first);
        }
    }
    const u64 keep = evict_last_policy();
    const size_t es = (size_t)a.rows * a.N;
    const u64 q = mt.q[t], r0 = mt.ratio0[t], r1 = mt.ratio1[t], rw = mt.rwide[t];
    const bool has_add = a.addp && r < a.add_rows;
    ulonglong2 pm = make_ulonglong2(0, 0);
    if (has_add && a.add_pscale) pm = pmod[t];
    __syncthreads();   // barriers initialised before anybody waits on them
    // digits (and the c0 term) of tile i+1 are gathered from L2 while tile i is multiplied
    u64 dn[BETA ? BETA : 1];
    u64 c0n = 0;
    u32 srcn = 0;
    auto gather = [&](int i) {
        const int n = (tile0 + i) * KS_TILE + threadIdx.x;
        srcn = elt ? galois_src((u32)n, elt, a.logn) : (u32)n;
        const u64* e = a.E + (size_t)r * a.N + srcn;
        if (BETA) {
#pragma unroll
            for (int j = 0; j < BETA; j++) dn[j] = ld_keep(e + (size_t)j * es, keep);
        }
        c0n = has_add ? a.addp[(size_t)r * a.N + srcn] : 0;
    };
    gather(0);
    for (int i = 0; i < ntiles; i++) {
        const int n = (tile0 + i) * KS_TILE + threadIdx.x;
        const u32 src = srcn;
        const u64* e = a.E + (size_t)r * a.N + src;
        const u64* kcol = ksm + (i & 1) * stage_words + threadIdx.x;
        u64 c0add = c0n;
        u64 lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
        Acc3 acc0 = {0, 0, 0}, acc1 = {0, 0, 0};
        if (BETA) {
            u64 d[BETA ? BETA : 1];
#pragma unroll
            for (int j = 0; j < BETA; j++) d[j] = dn[j];
            if (i + 1 < ntiles) gather(i + 1);
            mbar_wait(&full[i & 1], (i >> 1) & 1);
#pragma unroll
            for (int j = 0; j < BETA; j++) {
                const u32 dsum = (u32)d[j] + (u32)(d[j] >> 32);
                mac_split(acc0, d[j], dsum, kcol[(2 * j) * KS_TILE]);
                mac_split(acc1, d[j], dsum, kcol[(2 * j + 1) * KS_TILE]);
                if ((j + 1) % FOLD == 0 && j + 1 < BETA) {
                    fold_split(lo0, hi0, acc0);
                    fold_split(lo1, hi1, acc1);
                }
            }
        } else {
            if (i + 1 < ntiles) gather(i + 1);
            mbar_wait(&full[i & 1], (i >> 1) & 1);
            for (int j = 0; j < beta; j++) {
                const u64 dj = ld_keep(e + (size_t)j * es, keep);
                const u32 dsum = (u32)dj + (u32)(dj >> 32);
                mac_split(acc0, dj, dsum, kcol[(size_t)(2 * j) * KS_TILE]);
                mac_split(acc1, dj, dsum, kcol[(size_t)(2 * j + 1) * KS_TILE]);
                if ((j + 1) % FOLD == 0) {
                    fold_split(lo0, hi0, acc0);
                    fold_split(lo1, hi1, acc1);
                }
            }
        }
        fold_split(lo0, hi0, acc0);
        fold_split(lo1, hi1, acc1);
        u64 v0, v1;
        if (FOLD == 16 && BETA != 0) {   // q < 2^59 and <= 8 digits: the sum is below 2^(s+64)
            v0 = reduce_wide(lo0, hi0, q, rw), v1 = reduce_wide(lo1, hi1, q, rw);
        } else {
            v0 = barrett128(lo0, hi0, q, r0, r1), v1 = barrett128(lo1, hi1, q, r0, r1);
        }
        if (has_add) {
            if (a.add_pscale) c0add = mul_shoup(c0add, pm.x, pm.y, q);
            v0 = add_mod(v0, c0add, q);
        }
        u64* o0 = out + (size_t)r * a.N + n;
        u64* o1 = o0 + es;
        if (a.accumulate) {
            v0 = add_mod(v0, *o0, q);
            v1 = add_mod(v1, *o1, q);
        }
        *o0 = v0;
        *o1 = v1;
        if (i + 2 < ntiles) {
            __syncthreads();   // every thread is done reading stage i & 1
            if (threadIdx.x == 0) {
                mbar_expect_tx(&full[i & 1], stage_bytes);
                tma_load_3d_hint(ksm + (i & 1) * stage_words, kmap, (tile0 + i + 2) * KS_TILE, t, 0, &full[i & 1],
                                 pol_first);
            }
        }
    }
}

template <int FOLD, int BETA>
__global__ void __launch_bounds__(KS_TILE, 6) k_ks_inner_tma(const __grid_constant__ CUtensorMap kmap, KsArgs a, ModTab mt,
                                                              const ulonglong2* __restrict__ pmod) {
    ks_tile_body<FOLD, BETA>(&kmap, a, mt, pmod, blockIdx.y, blockIdx.x * KS_TPC, a.elt, a.out);
}

// All hoisted baby steps in ONE launch: grid (tiles, babies, rows) is dispatched row-major, so the ~2 MB of
// digits of a row stay L2-resident while the (G-1) rotation keys stream past them exactly once.
constexpr int KS_MAX_BABY = 224;   // 224 x (128 B tensor map + element) = 29.6 KB of the 32 KB parameter space
struct BabyTab {
    CUtensorMap map[KS_MAX_BABY];   // key of baby step blockIdx.y + 1
    u32 elt[KS_MAX_BABY];
};
template <int FOLD, int BETA>
__global__ void __launch_bounds__(KS_TILE, 6) k_ks_baby_fused(const __grid_constant__ BabyTab tab, KsArgs a, ModTab mt,
                                                               const ulonglong2* __restrict__ pmod) {
    const int b = blockIdx.y;
    ks_tile_body<FOLD, BETA>(&tab.map[b], a, mt, pmod, blockIdx.z + a.row0, a.tile0 + blockIdx.x * KS_TPC, tab.elt[b],
                             a.out + (size_t)b * 2 * a.rows * a.N);
}

// y[p][r][n] = x[p][r][n] * (P mod q_r)  on data rows, 0 on special rows (the b = 0 baby "rotation")
__global__ void k_pscale(const u64* __restrict__ x, u64* __restrict__ y, int l, int rows, int N, ModTab mt,
                         const ulonglong2* __restrict__ pmod) {
    size_t total = (size_t)2 * rows * N;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        int n = (int)(e % N);
        size_t pr = e / N;
        int r = (int)(pr % rows), p = (int)(pr / rows);
        u64 v = 0;
        if (r < l) {
            ulonglong2 pm = pmod[r];
            v = mul_shoup(x[((size_t)p * l + r) * N + n], pm.x, pm.y, mt.q[r]);
        }
        y[e] = v;
    }
}

// ---- ModDown --------------------------------------------------------------------------------
// in: [polys][rows][N] with the P special rows already in coefficient form.
// tmp[p][i][n] = sum_k [ (sp_k + half_k) * hatinv_k ]_{p_k} * hat[k][i]  - half_i   (mod q_i), coefficient form
template <int A>
__global__ void __launch_bounds__(TPB) k_moddown_conv(const u64* __restrict__ in, u64* __restrict__ tmp, int l,
                                                       int N, int L, int P, int K, size_t in_pstride, ModTab mt,
                                                       const ulonglong2* __restrict__ hatinv,
                                                       const u64* __restrict__ half, const u64* __restrict__ hat) {
    const int p = blockIdx.y, n = blockIdx.x * TPB + threadIdx.x;
    const u64* sp = in + (size_t)p * in_pstride + (size_t)l * N + n;
    u64 y[A];
    u32 ys[A];
#pragma unroll
    for (int k = 0; k < A; k++) {
        y[k] = 0, ys[k] = 0;
        if (k < P) {
            u64 pk = mt.q[L + k];
            ulonglong2 h = hatinv[k];
            y[k] = split30(mul_shoup(add_mod(sp[(size_t)k * N], half[L + k], pk), h.x, h.y, pk));
            ys[k] = (u32)y[k] + (u32)(y[k] >> 32);
        }
    }
    u64* o = tmp + (size_t)p * l * N + n;
    for (int i = 0; i < l; i++) {
        Acc3 acc = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < A; k++) mac_split(acc, y[k], ys[k], hat[(size_t)k * K + i]);
        u64 alo = 0, ahi = 0;
        fold_split(alo, ahi, acc);
        u64 q = mt.q[i];
        o[(size_t)i * N] = sub_mod(reduce_wide(alo, ahi, q, mt.rwide[i]), half[i], q);
    }
}
// out[p][i][n] = (in[p][i][n] - tmp[p][i][n]) * P^-1  (+ add[p][i][n])
__global__ void k_moddown_final(const u64* __restrict__ in, const u64* __restrict__ tmp, const u64* __restrict__ add,
                                u64* __restrict__ out, int polys, int l, int N, size_t in_pstride, ModTab mt,
                                const ulonglong2* __restrict__ pinv) {
    size_t total = (size_t)polys * l * N;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        int n = (int)(e % N);
        size_t pr = e / N;
        int i = (int)(pr % l), p = (int)(pr / l);
        u64 q = mt.q[i];
        ulonglong2 pi = pinv[i];
        u64 v = mul_shoup(sub_mod(in[(size_t)p * in_pstride + (size_t)i * N + n], tmp[e], q), pi.x, pi.y, q);
        if (add) v = add_mod(v, add[e], q);
        out[e] = v;
    }
}

// ---- ModDown + rescale of a finished accumulator, in the coefficient domain ---------------------------------------
// R: [polys][l+P][N] with EVERY row in coefficient form.  Per coefficient: t_i = (R_i - conv_i) * P^-1 exactly as
// k_moddown_conv / k_moddown_final (linear, so the coefficient-domain result is the inverse transform of theirs), then the
// rescale by q_{l-1} exactly as k_rescale_conv / k_rescale_final:  out_i = (t_i - ([x]_{q_i} - [half]_{q_i})) * q_{l-1}^-1
// with x = t_{l-1} + half.  out: [polys][l-1][N], coefficient form.  One launch instead of two conversions, two
// finals and the forward transform of the l ModDown correction rows in between.
template <int A>
__global__ void __launch_bounds__(TPB) k_finish_conv(const u64* __restrict__ R, u64* __restrict__ out, int l, int N, int L,
                                                      int P, int K, ModTab mt, const ulonglong2* __restrict__ hatinv,
                                                      const u64* __restrict__ half, const u64* __restrict__ hat,
                                                      const ulonglong2* __restrict__ pinv,
                                                      const ulonglong2* __restrict__ rsinv) {
    const int p = blockIdx.y, n = blockIdx.x * TPB + threadIdx.x;
    const u64* in = R + (size_t)p * (l + P) * N + n;
    const u64* sp = in + (size_t)l * N;
    u64 y[A];
    u32 ys[A];
#pragma unroll
    for (int k = 0; k < A; k++) {
        y[k] = 0, ys[k] = 0;
        if (k < P) {
            u64 pk = mt.q[L + k];
            ulonglong2 h = hatinv[k];
            y[k] = split30(mul_shoup(add_mod(sp[(size_t)k * N], half[L + k], pk), h.x, h.y, pk));
            ys[k] = (u32)y[k] + (u32)(y[k] >> 32);
        }
    }
    auto moddown = [&](int i) {
        Acc3 acc = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < A; k++) mac_split(acc, y[k], ys[k], hat[(size_t)k * K + i]);
        u64 alo = 0, ahi = 0;
        fold_split(alo, ahi, acc);
        const u64 q = mt.q[i];
        const u64 conv = sub_mod(reduce_wide(alo, ahi, q, mt.rwide[i]), half[i], q);
        const ulonglong2 pi = pinv[i];
        return mul_shoup(sub_mod(in[(size_t)i * N], conv, q), pi.x, pi.y, q);
    };
    const u64 ql = mt.q[l - 1], hl = ql >> 1;
    const u64 x = add_mod(moddown(l - 1), hl, ql);
    u64* o = out + (size_t)p * (l - 1) * N + n;
    for (int i = 0; i < l - 1; i++) {
        const u64 q = mt.q[i], r1 = mt.ratio1[i];
        const u64 corr = sub_mod(barrett64(x, q, r1), barrett64(hl, q, r1), q);
        const ulonglong2 w = rsinv[i];
        o[(size_t)i * N] = mul_shoup(sub_mod(moddown(i), corr, q), w.x, w.y, q);
    }
}

// ---- ModRaise (bootstrapping) ---------------------------------------------------------------------
// x: [polys][N] coefficient form modulo q_0; out[p][i][n] = centred representative of x modulo q_i (coefficient form)
__global__ void __launch_bounds__(TPB) k_modraise(const u64* __restrict__ x, u64* __restrict__ out, int l, int N,
                                                   ModTab mt) {
    const int p = blockIdx.y, n = blockIdx.x * TPB + threadIdx.x;
    const u64 q0 = mt.q[0], half = q0 >> 1, v = x[(size_t)p * N + n];
    u64* o = out + (size_t)p * l * N + n;
    for (int i = 0; i < l; i++) {
        const u64 q = mt.q[i], r1 = mt.ratio1[i];
        u64 r;
        if (v > half) {   // negative representative v - q_0
            const u64 d = barrett64(q0 - v, q, r1);
            r = d ? q - d : 0;
        } else {
            r = barrett64(v, q, r1);
        }
        o[(size_t)i * N] = r;
    }
}

// ---- rescale ---------------------------------------------------------------------------------
// last: [polys][N] coefficient form of the dropped limb; tmp[p][i][n] = [(x + half)]_{q_i} - [half]_{q_i}
__global__ void __launch_bounds__(TPB) k_rescale_conv(const u64* __restrict__ last, u64* __restrict__ tmp, int l,
                                                       int N, ModTab mt) {
    const int p = blockIdx.y, n = blockIdx.x * TPB + threadIdx.x;
    const u64 ql = mt.q[l - 1], half = ql >> 1;
    u64 x = add_mod(last[(size_t)p * N + n], half, ql);
    u64* o = tmp + (size_t)p * (l - 1) * N + n;
    for (int i = 0; i < l - 1; i++) {
        u64 q = mt.q[i], r1 = mt.ratio1[i];
        o[(size_t)i * N] = sub_mod(barrett64(x, q, r1), barrett64(half, q, r1), q);
    }
}
__global__ void k_rescale_final(const u64* __restrict__ in, const u64* __restrict__ tmp, u64* __restrict__ out,
                                int polys, int l, int N, ModTab mt, const ulonglong2* __restrict__ inv) {
    size_t total = (size_t)polys * (l - 1) * N;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        int n = (int)(e % N);
        size_t pr = e / N;
        int i = (int)(pr % (l - 1)), p = (int)(pr / (l - 1));
        u64 q = mt.q[i];
        ulonglong2 w = inv[i];
        out[e] = mul_shoup(sub_mod(in[((size_t)p * l + i) * N + n], tmp[e], q), w.x, w.y, q);
    }
}

// ---- plaintext-diagonal multiply-accumulate (exact path) ------------------------------------------
// inner[p][i][n] = sum_{b < nb} baby[b][p][i][n] * pt[b][i][n]      (pt pointers: consecutive plaintexts)
struct PmacPtrs {
    const u64* baby[128];
    const u64* pt[128];
};
__global__ void __launch_bounds__(TPB) k_pmac_list(PmacPtrs ptrs, int nb, u64* __restrict__ out, int l, int N,
                                                    ModTab mt) {
    const int i = blockIdx.y, n = blockIdx.x * TPB + threadIdx.x;
    const size_t off = (size_t)i * N + n, pw = (size_t)l * N;
    u64 lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
    const u64 pol = evict_first_policy();
    for (int b = 0; b < nb; b++) {
        u64 d = ld_stream(ptrs.pt[b] + off, pol);
        mac128(lo0, hi0, ptrs.baby[b][off], d);
        mac128(lo1, hi1, ptrs.baby[b][pw + off], d);
    }
    u64 q = mt.q[i], r0 = mt.ratio0[i], r1 = mt.ratio1[i];
    out[off] = barrett128(lo0, hi0, q, r0, r1);
    out[pw + off] = barrett128(lo1, hi1, q, r0, r1);
}

// ---- plaintext-diagonal multiply-accumulate (hoisted path) ----------------------------------------
// A[g][p][r][n] = sum_{b : gG+b < D} Y[b][p][r][n] * diag[gG+b][r][n >> rshift]
// One CTA owns (row r, TILE coefficients): the G baby tiles (both polynomials) are staged once in
// shared memory and reused for all B giant groups, so Y is read from L2/HBM once and every diagonal
// element is read once.
constexpr int PM_TILE = 128;
__device__ __forceinline__ u64* pmac_dst(const PmacDst& d, int g, int p, size_t pw) {
    if (d.world == 1) return d.base[0] + (size_t)(g * 2 + p) * pw;   // (uniform branch: the one-GPU path pays no division)
    const int slot = (int)(((u32)g * ((65536u + d.world - 1) / d.world)) >> 16);   // g / world, exact for g < 8192
    return d.base[g - slot * d.world] + (size_t)(slot * 2 + p) * pw;
}
__global__ void __launch_bounds__(PM_TILE) k_pmac_hoisted(const u64* __restrict__ Y, const u64* __restrict__ diag,
                                                           PmacDst dst, int G, int B, int D, int l, int rows,
                                                           int N, int L, int rshift, int row0, int diag_rows, int col0,
                                                           int diag_cols, ModTab mt) {
    extern __shared__ u64 sm[];   // [G][2][PM_TILE]
    const int r = blockIdx.y + row0, n = col0 + blockIdx.x * PM_TILE + threadIdx.x;
    const int t = r < l ? r : L + (r - l);
    const size_t pw = (size_t)rows * N, off = (size_t)r * N + n;
    for (int b = 0; b < G; b++) {
        sm[(b * 2 + 0) * PM_TILE + threadIdx.x] = Y[(size_t)b * 2 * pw + off];
        sm[(b * 2 + 1) * PM_TILE + threadIdx.x] = Y[(size_t)b * 2 * pw + pw + off];
    }
    // each thread only re-reads its own column: no barrier needed
    const u64 q = mt.q[t], r0 = mt.ratio0[t], r1 = mt.ratio1[t];
    const int dn = diag_cols;
    const u64* dg = diag + (size_t)blockIdx.y * dn + ((n >> rshift) - (col0 >> rshift));   // diag: first row / column served
    const size_t dstride = (size_t)diag_rows * dn;
    const u64 pol = evict_first_policy();
    for (int g = 0; g < B; g++) {
        int nb = min(G, D - g * G);
        if (nb <= 0) break;
        u64 lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
        const u64* dgg = dg + (size_t)g * G * dstride;
#pragma unroll 4
        for (int b = 0; b < nb; b++) {
            u64 d = unsplit30(ld_stream(dgg + (size_t)b * dstride, pol));
            mac128(lo0, hi0, sm[(b * 2 + 0) * PM_TILE + threadIdx.x], d);
            mac128(lo1, hi1, sm[(b * 2 + 1) * PM_TILE + threadIdx.x], d);
        }
        pmac_dst(dst, g, 0, pw)[off] = barrett128(lo0, hi0, q, r0, r1);
        pmac_dst(dst, g, 1, pw)[off] = barrett128(lo1, hi1, q, r0, r1);
    }
    if (dst.world > 1) __threadfence_system();   // the stores crossed NVLink: visible before the next kernel posts its flag
}


// ---- plaintext-diagonal multiply-accumulate, TMA-staged (sub-ring compressed diagonals) -------------
// Same contraction as k_pmac_hoisted.  A CTA owns (row r, 128 coefficients); thread = (polynomial p,
// coefficient i).  The G baby tiles sit in shared memory for the whole kernel; for every giant group g
// one TMA box [G diagonals][1 row][128 >> rshift values] (UTMALDG, 3-stage mbarrier ring) brings the
// diagonal values, so each diagonal byte crosses HBM->SM once and the MAC loop only touches shared memory.
constexpr int PM_STAGES = 2;
constexpr int PM_GT = 4;        // giant groups per pipeline stage (register tile over g)
constexpr int PM_T2 = 64;       // coefficients per CTA

// A CTA owns (row r, PM_T2 coefficients); thread = (polynomial p, coefficient i) and carries PM_GT giant
// groups at once, so one shared-memory read of a baby value feeds PM_GT multiply-accumulates.
// One launch covers the baby steps [b0, b0 + nbc) of every group (Gc = chunk height, a multiple of 16): sets with
// many baby steps (hoisting-aware splits, G > 64) are walked in chunks so the shared-memory tile stays small
// enough for two to three CTAs per SM; later chunks add onto A.
// Shared memory: baby tile [Gc][2][PM_T2] in split-30 form (rows >= nbc zero), and a PM_STAGES-deep ring of
// diagonal boxes [PM_GT][Gc][W] filled by TMA (one 3-D box per giant group; a box may run past its group or
// past the end of the set -- those rows meet zero baby values / are zero-filled by the TMA unit).
// NG giant groups of one pipeline stage: A[g0 + k] (+)= sum_b y_b * d_{k,b}
template <int FOLD, int W, int NG>
__device__ __forceinline__ void pmac_groups(const u64* __restrict__ ycol, const u64* __restrict__ dg, size_t group_words,
                                            int Gc, const PmacDst& dst, const u64* Ain, int g0, int p, size_t pw,
                                            size_t off, u64 q, u64 r0, u64 r1) {
    Acc3 acc[NG];
    u64 lo[NG], hi[NG];
#pragma unroll
    for (int k = 0; k < NG; k++) acc[k].s0 = acc[k].s1 = acc[k].s2 = 0, lo[k] = hi[k] = 0;
    for (int b0 = 0; b0 < Gc; b0 += FOLD) {
#pragma unroll
        for (int j = 0; j < FOLD; j++) {
            const u64 y = ycol[(b0 + j) * 2 * PM_T2];
            const u32 ys = (u32)y + (u32)(y >> 32);
#pragma unroll
            for (int k = 0; k < NG; k++) mac_split(acc[k], y, ys, dg[k * group_words + (size_t)(b0 + j) * W]);
        }
#pragma unroll
        for (int k = 0; k < NG; k++) fold_split(lo[k], hi[k], acc[k]);
    }
#pragma unroll
    for (int k = 0; k < NG; k++) {
        u64 v = barrett128(lo[k], hi[k], q, r0, r1);
        if (Ain) v = add_mod(v, Ain[(size_t)((g0 + k) * 2 + p) * pw + off], q);   // earlier baby-step chunks (local)
        pmac_dst(dst, g0 + k, p, pw)[off] = v;
    }
}

// Threads: (h, p, i) -- polynomial p, coefficient i, and h in [0, PM_HS) takes the giant groups h*PM_GT/PM_HS ... of
// every stage: the baby column of (p, i) is shared by PM_HS threads, which doubles the warps per SM for the same
// shared-memory footprint (the MAC loop is latency-bound at 2-3 warps per scheduler).
constexpr int PM_HS = 2;
constexpr int PM_NG = PM_GT / PM_HS;   // groups per thread and stage
template <int FOLD, int RSH>
__global__ void __launch_bounds__(2 * PM_T2 * PM_HS) k_pmac_tma(const __grid_constant__ CUtensorMap tmap,
                                                                 const u64* __restrict__ Y, PmacDst dst,
                                                                 const u64* Ain, int G, int Gc, int b0, int nbc,
                                                                 int Beff, int l, int rows, int N, int L, int row0,
                                                                 int col0, ModTab mt) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr int W = PM_T2 >> RSH;
    u64* dsm = reinterpret_cast<u64*>(smraw);                         // [PM_STAGES][PM_GT][Gc][W]
    u64* ysm = dsm + (size_t)PM_STAGES * PM_GT * Gc * W;              // [Gc][2][PM_T2]
    uint64_t* full = reinterpret_cast<uint64_t*>(ysm + (size_t)Gc * 2 * PM_T2);
    const int tid = threadIdx.x, h = tid / (2 * PM_T2), p = (tid / PM_T2) & 1, i = tid % PM_T2;
    const int r = blockIdx.y + row0, n0 = col0 + blockIdx.x * PM_T2;   // the tensor map starts at (row0, col0): TMA row = blockIdx.y
    const int t = r < l ? r : L + (r - l);
    const int iters = (Beff + PM_GT - 1) / PM_GT;
    const size_t group_words = (size_t)Gc * W, stage_words = PM_GT * group_words;
    if (tid == 0) {
        for (int s = 0; s < PM_STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const size_t pw = (size_t)rows * N, off = (size_t)r * N + n0 + i;
    for (int b = h; b < Gc; b += PM_HS)
        ysm[(b * 2 + p) * PM_T2 + i] = b < nbc ? split30(Y[(size_t)((b0 + b) * 2 + p) * pw + off]) : 0;
    __syncthreads();
    auto issue = [&](int it) {   // only the groups that exist are fetched; expect_tx counts exactly those bytes
        const int s = it % PM_STAGES, ng = min(PM_GT, Beff - it * PM_GT);
        mbar_expect_tx(&full[s], (u32)(ng * Gc * W * sizeof(u64)));
        for (int k = 0; k < ng; k++)
            tma_load_3d(dsm + s * stage_words + k * group_words, &tmap, (int)(blockIdx.x * PM_T2) >> RSH, blockIdx.y,
                        (it * PM_GT + k) * G + b0, &full[s]);
    };
    if (tid == 0)
        for (int it = 0; it < PM_STAGES && it < iters; it++) issue(it);
    const u64 q = mt.q[t], r0 = mt.ratio0[t], r1 = mt.ratio1[t];
    const u64* ycol = ysm + p * PM_T2 + i;
    for (int it = 0; it < iters; it++) {
        const int s = it % PM_STAGES, ng = min(PM_NG, Beff - it * PM_GT - h * PM_NG);   // this thread's groups
        mbar_wait(&full[s], (it / PM_STAGES) & 1);
        const u64* dg = dsm + s * stage_words + (size_t)h * PM_NG * group_words + (i >> RSH);
        const int g0 = it * PM_GT + h * PM_NG;
        if (ng >= PM_NG) pmac_groups<FOLD, W, PM_NG>(ycol, dg, group_words, Gc, dst, Ain, g0, p, pw, off, q, r0, r1);
        else if (ng == 1) pmac_groups<FOLD, W, 1>(ycol, dg, group_words, Gc, dst, Ain, g0, p, pw, off, q, r0, r1);
        __syncthreads();   // every thread is done with stage s
        if (tid == 0 && it + PM_STAGES < iters) issue(it + PM_STAGES);
    }
    if (dst.world > 1) __threadfence_system();   // the stores crossed NVLink: visible before the next kernel posts its flag
}

// in-place conversion between canonical residues and the split-30 storage form
__global__ void k_split30(u64* __restrict__ x, size_t n, int unsplit) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
        x[e] = unsplit ? unsplit30(x[e]) : split30(x[e]);
}

}  // namespace

// ===============================================================================================
// host launchers
// ===============================================================================================
namespace ops {

// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult st;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st));
        REQUIRE(p && st == cudaDriverEntryPointSuccess, "CUDA driver does not provide cuTensorMapEncodeTiled");
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// tensor map over one switching key [2*beta_key][K][N] with boxes [2*beta][1][box_n] (shared with ntt.cu's fused kernel)
void encode_key_map(const Ctx* c, const u64* key, int box_n, int beta, CUtensorMap* out) {
    cuuint64_t dims[3] = {(cuuint64_t)c->N, (cuuint64_t)c->K, (cuuint64_t)(2 * c->beta)};
    cuuint64_t strides[2] = {(cuuint64_t)c->N * sizeof(u64), (cuuint64_t)c->K * c->N * sizeof(u64)};
    cuuint32_t box[3] = {(cuuint32_t)box_n, 1, (cuuint32_t)(2 * beta)};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult rc = encode_tiled()(out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void*)key, dims, strides, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
}

static int grid_for(const Ctx* c, size_t total) {
    size_t blocks = (total + TPB - 1) / TPB;
    size_t cap = (size_t)c->sm_count * 16;
    return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

void add(const Ctx* c, const u64* a, const u64* b, u64* out, int polys, int rows, int n, RowMap rm, int b_polys,
         cudaStream_t s) {
    LAUNCH(k_ew<EW_ADD>, grid_for(c, (size_t)polys * rows * n), TPB, 0, s)(a, b, out, polys, rows, n, rm, c->modtab(), b_polys);
}
void sub(const Ctx* c, const u64* a, const u64* b, u64* out, int polys, int rows, int n, RowMap rm, int b_polys,
         cudaStream_t s) {
    LAUNCH(k_ew<EW_SUB>, grid_for(c, (size_t)polys * rows * n), TPB, 0, s)(a, b, out, polys, rows, n, rm, c->modtab(), b_polys);
}
void neg(const Ctx* c, const u64* a, u64* out, int polys, int rows, int n, RowMap rm, cudaStream_t s) {
    LAUNCH(k_ew<EW_NEG>, grid_for(c, (size_t)polys * rows * n), TPB, 0, s)(a, nullptr, out, polys, rows, n, rm, c->modtab(), 0);
}
void mul(const Ctx* c, const u64* a, const u64* b, u64* out, int polys, int rows, int n, RowMap rm, int b_polys,
         cudaStream_t s) {
    LAUNCH(k_ew<EW_MUL>, grid_for(c, (size_t)polys * rows * n), TPB, 0, s)(a, b, out, polys, rows, n, rm, c->modtab(), b_polys);
}
void reduce_inplace(const Ctx* c, u64* x, int polys, int rows, RowMap rm, cudaStream_t s) {
    LAUNCH(k_reduce, grid_for(c, (size_t)polys * rows * c->N), TPB, 0, s)(x, polys, rows, c->N, rm, c->modtab());
    CUDA_CHECK(cudaGetLastError());
}
void sum_groups(const Ctx* c, const u64* parts, int nparts, const u64* first, u64* out, int l, cudaStream_t s) {
    const int rows = l + c->P;
    const size_t words = (size_t)2 * rows * c->N;
    ProfScope ps(c, PROF_SUM_GROUPS, s);
    LAUNCH(k_sum_groups, grid_for(c, words), TPB, 0, s)(parts, nparts, words, first, out, rows, c->N, RowMap{rows, l, c->L, 0},
                                                        c->modtab());
    CUDA_CHECK(cudaGetLastError());
}
void tensor(const Ctx* c, const u64* a, const u64* b, u64* out, int l, cudaStream_t s) {
    LAUNCH(k_tensor, grid_for(c, (size_t)l * c->N), TPB, 0, s)(a, b, out, l, c->N, c->modtab());
}
void galois(const Ctx* c, const u64* in, u64* out, int rows, u32 elt, cudaStream_t s) {
    LAUNCH(k_galois, grid_for(c, (size_t)rows * c->N), TPB, 0, s)(in, out, rows, c->N, c->logn, elt);
}

// cin [l][N] NTT form -> E [digits(l)][l+P][N] NTT form (scratch x: [l][N])
void decompose(const Ctx* c, const u64* cin, int l, u64* x, u64* E, cudaStream_t s, int row0, int nrows) {
    const int row1 = nrows < 0 ? l + c->P : row0 + nrows;
    if (ntt_decompose_a(c, cin, l, x, E, 1, s, row0, row1)) {   // fused front end; the second forward pass completes the digits
        ntt_forward(c, E, c->digits(l) * (l + c->P), RowMap{l + c->P, l, c->L, 0}, c->N, s, c->P, /*split30_out=*/true,
                    /*pass_a_only=*/false, /*pass_b_only=*/true, row0, row1);
        return;
    }
    // (the staged path below produces every row)
    CUDA_CHECK(cudaMemcpyAsync(x, cin, sizeof(u64) * l * c->N, cudaMemcpyDeviceToDevice, s));
    ntt_inverse(c, x, l, RowMap{l, l, c->L, 0}, c->N, s);
    decompose_from(c, cin, x, l, E, s);
}
// same, given both forms of the polynomial: cin (NTT) and x (coefficients)
void decompose_from(const Ctx* c, const u64* cin, const u64* x, int l, u64* E, cudaStream_t s, bool transform, int count) {
    const int N = c->N, P = c->P, rows = l + P, beta = c->digits(l);
    REQUIRE(P <= MAX_ALPHA, "special_modulus_size > %d not supported", MAX_ALPHA);
    REQUIRE(N % TPB == 0, "N must be a multiple of %d", TPB);
    {
        ProfScope ps(c, PROF_MODUP, s);
        const size_t tab_bytes = sizeof(u64) * rows * ((P <= 4 ? P : MAX_ALPHA) + 2);
        auto go = [&](auto kern) {
            LAUNCH(kern, dim3(N / TPB, beta, count), TPB, tab_bytes, s)(x, cin, E, l, N, c->L, P, c->K, c->modtab(),
                                                                 c->d_up_hatinv + (size_t)l * c->beta * P,
                                                                 c->d_up_hat + (size_t)l * c->beta * P * c->K, c->sbits);
        };
        if (P == 1) go(k_modup<1>);
        else if (P == 2) go(k_modup<2>);
        else if (P == 3) go(k_modup<3>);
        else if (P == 4) go(k_modup<4>);
        else go(k_modup<MAX_ALPHA>);
    }
    REQUIRE(count >= 1 && (count == 1 || !transform) && count <= 65535, "modup: bad batch");
    if (transform) ntt_forward(c, E, beta * rows, RowMap{rows, l, c->L, 0}, N, s, P, /*split30_out=*/true);
    CUDA_CHECK(cudaGetLastError());
}

// decompose + key inner product; the forward transform's last pass is fused into the product when it applies
// (SPEAR_FUSED_KS=0 keeps the two-kernel form, for A/B timing)
void decompose_ks(const Ctx* c, const u64* cin, u64* x, int l, u64* E, const u64* key, u64* out, u32 elt,
                  const u64* addp, int add_rows, int add_pscale, int accumulate, cudaStream_t s) {
    static const bool fused = [] {
        const char* e = getenv("SPEAR_FUSED_KS");
        return !(e && e[0] == '0');
    }();
    // x is scratch [l][N].  Fully fused form: one kernel for inverse pass A + ModUp + forward pass A, one for forward
    // pass B + key product
    if (fused && ntt_ks_fused_applies(c, l) && ntt_decompose_a(c, cin, l, x, E, 1, s) &&
        ntt_ks_fused(c, E, key, out, l, elt, addp, add_rows, add_pscale, accumulate, s, /*pass_a_done=*/true))
        return;
    CUDA_CHECK(cudaMemcpyAsync(x, cin, sizeof(u64) * l * c->N, cudaMemcpyDeviceToDevice, s));
    ntt_inverse(c, x, l, RowMap{l, l, c->L, 0}, c->N, s);
    decompose_from(c, cin, x, l, E, s, /*transform=*/false);
    if (fused && ntt_ks_fused(c, E, key, out, l, elt, addp, add_rows, add_pscale, accumulate, s)) return;
    ntt_forward(c, E, c->digits(l) * (l + c->P), RowMap{l + c->P, l, c->L, 0}, c->N, s, c->P, /*split30_out=*/true);
    ks_inner(c, E, key, out, l, elt, addp, add_rows, add_pscale, accumulate, s);
}

void ks_inner(const Ctx* c, const u64* E, const u64* key, u64* out, int l, u32 elt, const u64* addp, int add_rows,
              int add_pscale, int accumulate, cudaStream_t s) {
    KsArgs a;
    a.E = E, a.key = key, a.out = out, a.addp = addp, a.add_rows = add_rows, a.add_pscale = add_pscale;
    a.accumulate = accumulate, a.beta = c->digits(l), a.l = l, a.rows = l + c->P, a.N = c->N, a.logn = c->logn;
    a.L = c->L, a.K = c->K, a.elt = elt;
    ProfScope ps(c, PROF_KS_INNER, s);
    if (c->N % KS_TILE == 0 && 2 * a.beta <= 256 && (size_t)4 * a.beta * KS_TILE * sizeof(u64) + 64 <= 200 * 1024) {
        CUtensorMap kmap;
        cuuint64_t dims[3] = {(cuuint64_t)c->N, (cuuint64_t)c->K, (cuuint64_t)(2 * c->beta)};
        cuuint64_t strides[2] = {(cuuint64_t)c->N * sizeof(u64), (cuuint64_t)c->K * c->N * sizeof(u64)};
        cuuint32_t box[3] = {(cuuint32_t)KS_TILE, 1, (cuuint32_t)(2 * a.beta)};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult rc = encode_tiled()(&kmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void*)key, dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
        bool small = true;
        for (u64 qq : c->q) small = small && qq < (1ull << 59);
        const size_t smem = (size_t)4 * a.beta * KS_TILE * sizeof(u64) + 64;
        const int gx = (c->N / KS_TILE + KS_TPC - 1) / KS_TPC;
        auto go = [&](auto kern) {
            CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            LAUNCH(kern, dim3(gx, a.rows), KS_TILE, smem, s)(kmap, a, c->modtab(), c->d_pmod);
        };
#define KS_CASE(B)                                 \
    case B:                                        \
        if (small) go(k_ks_inner_tma<16, B>);      \
        else go(k_ks_inner_tma<8, B>);             \
        break;
        switch (a.beta) {
            KS_CASE(1) KS_CASE(2) KS_CASE(3) KS_CASE(4) KS_CASE(5) KS_CASE(6) KS_CASE(7) KS_CASE(8)
            default:
                if (small) go(k_ks_inner_tma<16, 0>);
                else go(k_ks_inner_tma<8, 0>);
        }
#undef KS_CASE
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    LAUNCH(k_ks_inner, dim3(c->N / TPB, a.rows), TPB, 0, s)(a, c->modtab(), c->d_pmod);
    CUDA_CHECK(cudaGetLastError());
}

// Y[b] for b = 1..nb (out points at Y[1]): all hoisted baby steps against their keys in one launch.
// Returns false when the fused path does not apply (caller falls back to one launch per baby step).
bool ks_baby_fused(const Ctx* c, const u64* E, const u64* const* keys, const u32* elts, int nb, u64* out, int l,
                   const u64* c0, cudaStream_t s, int row0, int nrows, int col0, int ncols) {
    const int beta = c->digits(l), rows = l + c->P;
    if (nrows < 0) nrows = rows - row0;
    if (ncols < 0) ncols = c->N - col0;
    REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= rows, "baby steps: bad row range");
    REQUIRE(col0 >= 0 && ncols >= 0 && col0 + ncols <= c->N && col0 % KS_TILE == 0 && ncols % KS_TILE == 0,
            "baby steps: bad column range");
    if (nrows == 0 || ncols == 0) return true;
    if (nb < 1 || nb > KS_MAX_BABY || beta > 8 || c->N % KS_TILE != 0) return false;
    static thread_local BabyTab tab;   // 12.6 KB by-value kernel parameter
    cuuint64_t dims[3] = {(cuuint64_t)c->N, (cuuint64_t)c->K, (cuuint64_t)(2 * c->beta)};
    cuuint64_t strides[2] = {(cuuint64_t)c->N * sizeof(u64), (cuuint64_t)c->K * c->N * sizeof(u64)};
    cuuint32_t box[3] = {(cuuint32_t)KS_TILE, 1, (cuuint32_t)(2 * beta)};
    cuuint32_t estr[3] = {1, 1, 1};
    for (int b = 0; b < nb; b++) {
        CUresult rc = encode_tiled()(&tab.map[b], CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void*)keys[b], dims, strides, box,
                                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
        tab.elt[b] = elts[b];
    }
    KsArgs a;
    a.E = E, a.key = nullptr, a.out = out, a.addp = c0, a.add_rows = l, a.add_pscale = 1, a.accumulate = 0;
    a.beta = beta, a.l = l, a.rows = rows, a.N = c->N, a.logn = c->logn, a.L = c->L, a.K = c->K, a.elt = 0;
    a.row0 = row0, a.tile0 = col0 / KS_TILE, a.tile1 = (col0 + ncols) / KS_TILE;
    bool small = true;
    for (u64 qq : c->q) small = small && qq < (1ull << 59);
    // SPEAR_BABY_SMEM_PAD (bytes, A/B switch): extra dynamic shared memory per CTA = fewer resident CTAs of this HBM-bound
    // kernel per SM, which leaves room for the CTAs of an issue-bound kernel running next to it on another stream
    static const size_t pad = getenv("SPEAR_BABY_SMEM_PAD") ? (size_t)atol(getenv("SPEAR_BABY_SMEM_PAD")) : 0;
    const size_t smem = std::min<size_t>((size_t)4 * beta * KS_TILE * sizeof(u64) + 64 + pad, 200 * 1024);
    const int gx = (ncols / KS_TILE + KS_TPC - 1) / KS_TPC;
    ProfScope ps(c, PROF_KS_BABY, s);
    auto go = [&](auto kern) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        LAUNCH(kern, dim3(gx, nb, nrows), KS_TILE, smem, s)(tab, a, c->modtab(), c->d_pmod);
    };
#define KB_CASE(B)                                  \
    case B:                                         \
        if (small) go(k_ks_baby_fused<16, B>);      \
        else go(k_ks_baby_fused<8, B>);             \
        break;
    switch (beta) { KB_CASE(1) KB_CASE(2) KB_CASE(3) KB_CASE(4) KB_CASE(5) KB_CASE(6) KB_CASE(7) KB_CASE(8) }
#undef KB_CASE
    CUDA_CHECK(cudaGetLastError());
    return true;
}

void pscale(const Ctx* c, const u64* x, u64* y, int l, cudaStream_t s) {
    int rows = l + c->P;
    LAUNCH(k_pscale, grid_for(c, (size_t)2 * rows * c->N), TPB, 0, s)(x, y, l, rows, c->N, c->modtab(), c->d_pmod);
}

// in: [polys] polynomials of (l+P) rows, polynomial stride in_pstride words; destroys the special rows of `in`.
// out[p][i] = ModDown(in[p])[i] (+ add[p][i]);  tmp: scratch [polys][l][N]
void moddown(const Ctx* c, u64* in, size_t in_pstride, int polys, int l, u64* tmp, const u64* add, u64* out,
             cudaStream_t s) {
    const int N = c->N, P = c->P;
    ntt_inverse(c, in + (size_t)l * N, polys * P, RowMap{P, 0, c->L, 0, in_pstride}, N, s);   // special rows of every polynomial
    {
        ProfScope ps(c, PROF_MODDOWN, s);
        auto go = [&](auto kern) {
            LAUNCH(kern, dim3(N / TPB, polys), TPB, 0, s)(in, tmp, l, N, c->L, P, c->K, in_pstride, c->modtab(),
                                                          c->d_dn_hatinv, c->d_dn_half, c->d_dn_hat);
        };
        if (P == 1) go(k_moddown_conv<1>);
        else if (P == 2) go(k_moddown_conv<2>);
        else if (P == 3) go(k_moddown_conv<3>);
        else if (P == 4) go(k_moddown_conv<4>);
        else go(k_moddown_conv<MAX_ALPHA>);
    }
    ntt_forward(c, tmp, polys * l, RowMap{l, l, c->L, 0}, N, s);
    ProfScope ps(c, PROF_MODDOWN, s);
    LAUNCH(k_moddown_final, grid_for(c, (size_t)polys * l * N), TPB, 0, s)(in, tmp, add, out, polys, l, N, in_pstride,
                                                                       c->modtab(), c->d_pinv);
    CUDA_CHECK(cudaGetLastError());
}

// R [polys][l+P][N] (NTT form, destroyed) -> out [polys][l-1][N] = rescale(ModDown(R)); false: not applicable
bool finish_fused(const Ctx* c, u64* R, int polys, int l, u64* out, cudaStream_t s) {
    static const bool enabled = [] {
        const char* e = getenv("SPEAR_FUSED_FINISH");
        return !(e && e[0] == '0');
    }();
    const int N = c->N, P = c->P, rows = l + P;
    if (!enabled || P > 4 || l < 2 || N % TPB != 0) return false;
    ntt_inverse(c, R, polys * rows, RowMap{rows, l, c->L, 0}, N, s);
    {
        ProfScope ps(c, PROF_MODDOWN, s);
        auto go = [&](auto kern) {
            LAUNCH(kern, dim3(N / TPB, polys), TPB, 0, s)(R, out, l, N, c->L, P, c->K, c->modtab(), c->d_dn_hatinv, c->d_dn_half,
                                                          c->d_dn_hat, c->d_pinv, c->d_rs_inv + (size_t)(l - 1) * c->K);
        };
        if (P == 1) go(k_finish_conv<1>);
        else if (P == 2) go(k_finish_conv<2>);
        else if (P == 3) go(k_finish_conv<3>);
        else go(k_finish_conv<4>);
    }
    ntt_forward(c, out, polys * (l - 1), RowMap{l - 1, l - 1, c->L, 0}, N, s);
    CUDA_CHECK(cudaGetLastError());
    return true;
}

// in [polys][l][N] -> out [polys][l-1][N]; scratch last [polys][N], tmp [polys][l-1][N]
void rescale(const Ctx* c, const u64* in, int polys, int l, u64* last, u64* tmp, u64* out, cudaStream_t s) {
    const int N = c->N;
    REQUIRE(l >= 2, "cannot rescale below one limb");
    CUDA_CHECK(cudaMemcpy2DAsync(last, sizeof(u64) * N, in + (size_t)(l - 1) * N, sizeof(u64) * l * N, sizeof(u64) * N,
                                 polys, cudaMemcpyDeviceToDevice, s));
    ntt_inverse(c, last, polys, RowMap{1, 1, c->L, l - 1}, N, s);
    LAUNCH(k_rescale_conv, dim3(N / TPB, polys), TPB, 0, s)(last, tmp, l, N, c->modtab());
    ntt_forward(c, tmp, polys * (l - 1), RowMap{l - 1, l - 1, c->L, 0}, N, s);
    LAUNCH(k_rescale_final, grid_for(c, (size_t)polys * (l - 1) * N), TPB, 0, s)(in, tmp, out, polys, l, N, c->modtab(),
                                                                             c->d_rs_inv + (size_t)(l - 1) * c->K);
    CUDA_CHECK(cudaGetLastError());
}

// in: [polys][1][N] (NTT form, modulus q_0) -> out: [polys][l][N] (NTT form); x: scratch [polys][N]
void mod_raise(const Ctx* c, const u64* in, int polys, int l, u64* x, u64* out, cudaStream_t s) {
    const int N = c->N;
    CUDA_CHECK(cudaMemcpyAsync(x, in, sizeof(u64) * polys * N, cudaMemcpyDeviceToDevice, s));
    ntt_inverse(c, x, polys, RowMap{1, 1, c->L, 0}, N, s);
    LAUNCH(k_modraise, dim3(N / TPB, polys), TPB, 0, s)(x, out, l, N, c->modtab());
    ntt_forward(c, out, polys * l, RowMap{l, l, c->L, 0}, N, s);
    CUDA_CHECK(cudaGetLastError());
}

void pmac_list(const Ctx* c, const u64* const* baby, const u64* const* pt, int nb, u64* out, int l, cudaStream_t s) {
    REQUIRE(nb <= 128, "at most 128 baby steps per giant group");
    PmacPtrs p;
    for (int b = 0; b < nb; b++) p.baby[b] = baby[b], p.pt[b] = pt[b];
    ProfScope ps(c, PROF_PMAC, s);
    LAUNCH(k_pmac_list, dim3(c->N / TPB, l), TPB, 0, s)(p, nb, out, l, c->N, c->modtab());
    CUDA_CHECK(cudaGetLastError());
}

void pmac_hoisted(const Ctx* c, const u64* Y, const u64* diag, u64* A, int G, int B, int D, int l, int rshift,
                  cudaStream_t s) {
    PmacDst dst = {};
    dst.base[0] = A, dst.world = 1;
    pmac_hoisted_rows(c, Y, diag, dst, A, G, B, D, l, rshift, 0, l + c->P, s);
}

void pmac_hoisted_rows(const Ctx* c, const u64* Y, const u64* diag, const PmacDst& dst, u64* tmp, int G, int B, int D, int l,
                       int rshift, int row0, int nrows, cudaStream_t s, int diag_rows, int col0, int ncols, int diag_cols) {
    const int rows = l + c->P, W = PM_T2 >> rshift;
    if (diag_rows < 0) diag_rows = nrows;
    if (ncols < 0) ncols = c->N - col0;
    if (diag_cols < 0) diag_cols = c->N >> rshift;
    const int dn = diag_cols;                     // stored values per (diagonal, row)
    REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= rows && dst.world >= 1 && dst.world <= 8, "diagonal MAC: bad row range");
    REQUIRE(col0 >= 0 && ncols >= 0 && col0 + ncols <= c->N && col0 % PM_TILE == 0 && ncols % PM_TILE == 0 &&
                (ncols >> rshift) <= diag_cols, "diagonal MAC: bad column range");
    if (nrows == 0 || ncols == 0) return;
    // baby steps are walked in chunks of Gc <= 64 rows (a multiple of 16) so that 2-3 CTAs fit per SM
    const int nchunks = (G + 63) / 64, Gc = ((G + nchunks - 1) / nchunks + 15) / 16 * 16;
    REQUIRE(c->N % PM_TILE == 0, "N must be a multiple of %d", PM_TILE);
    const size_t tma_smem =
        sizeof(u64) * ((size_t)PM_STAGES * PM_GT * Gc * W + (size_t)Gc * 2 * PM_T2) + 8 * PM_STAGES + 64;
    ProfScope ps(c, PROF_PMAC, s);
    if (rshift >= 1 && rshift <= 5 && W * sizeof(u64) >= 16 && ((size_t)Gc * W * sizeof(u64)) % 128 == 0 &&
        Gc <= D && tma_smem <= 227 * 1024) {
        CUtensorMap tmap;
        cuuint64_t dims[3] = {(cuuint64_t)(ncols >> rshift), (cuuint64_t)nrows, (cuuint64_t)D};
        cuuint64_t strides[2] = {(cuuint64_t)dn * sizeof(u64), (cuuint64_t)diag_rows * dn * sizeof(u64)};
        cuuint32_t box[3] = {(cuuint32_t)W, 1, (cuuint32_t)Gc};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult rc = encode_tiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void*)diag, dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
        // partial sums of 30x30-bit products: 16 terms fit 64 bits when every q < 2^59, else 8
        bool small = true;
        for (u64 qq : c->q) small = small && qq < (1ull << 59);
        dim3 grid(ncols / PM_T2, nrows);
        const ModTab mt = c->modtab();
        // several chunks: the earlier ones accumulate in the local array `tmp`, only the last one writes the destination
        // (which may be peer memory: no read-modify-write across NVLink)
        PmacDst local = {};
        local.base[0] = tmp, local.world = 1;
        auto go = [&](auto kern) {
            CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            for (int b0 = 0; b0 < G; b0 += Gc) {
                const bool last = b0 + Gc >= G;
                LAUNCH(kern, grid, 2 * PM_T2 * PM_HS, tma_smem, s)(tmap, Y, last ? dst : local, b0 > 0 ? tmp : nullptr, G, Gc, b0,
                                                                   std::min(Gc, G - b0), B, l, rows, c->N, c->L, row0, col0, mt);
            }
        };
        switch (rshift * 2 + (small ? 1 : 0)) {
            case 2: go(k_pmac_tma<8, 1>); break;
            case 3: go(k_pmac_tma<16, 1>); break;
            case 4: go(k_pmac_tma<8, 2>); break;
            case 5: go(k_pmac_tma<16, 2>); break;
            case 6: go(k_pmac_tma<8, 3>); break;
            case 7: go(k_pmac_tma<16, 3>); break;
            case 8: go(k_pmac_tma<8, 4>); break;
            case 9: go(k_pmac_tma<16, 4>); break;
            case 10: go(k_pmac_tma<8, 5>); break;
            default: go(k_pmac_tma<16, 5>); break;
        }
        CUDA_CHECK(cudaGetLastError());
        return;
    }
    size_t smem = sizeof(u64) * (size_t)G * 2 * PM_TILE;
    REQUIRE(smem <= 227 * 1024, "too many baby steps (%d) for the shared-memory tile", G);
    CUDA_CHECK(cudaFuncSetAttribute(k_pmac_hoisted, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));   // per device
    LAUNCH(k_pmac_hoisted, dim3(ncols / PM_TILE, nrows), PM_TILE, smem, s)(Y, diag, dst, G, B, D, l, rows, c->N, c->L, rshift,
                                                                       row0, diag_rows, col0, diag_cols, c->modtab());
    CUDA_CHECK(cudaGetLastError());
}

void split30_inplace(const Ctx* c, u64* x, size_t n, bool unsplit, cudaStream_t s) {
    LAUNCH(k_split30, grid_for(c, n), TPB, 0, s)(x, n, unsplit ? 1 : 0);
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace ops
