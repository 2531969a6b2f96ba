/*
 * spear_oracle.c -- CPU restatement of the CKKS BSGS diagonal mat-vec path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under fhe_spear_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg use it, and only as the checker / CPU
 * baseline.
 *
 * PARITY UNPINNED.  The arithmetic of the reference lives in the un-vendored,
 * un-pinned dependency mozendr/phantom-fhe (reference README.md:33-38), which
 * is absent from /root/reference; the reference holds no golden vectors for
 * this path.  What IS pinned by the reference and
 * followed here:
 *   - BSGS loop semantics           scripts/bootstrap_generation.py:464-484
 *   - diagonal pre-rotation/tiling  scripts/bootstrap_generation.py:361-378
 *   - Galois generator 5, conj 2N-1 scripts/bootstrap_generation.py:18-26
 *   - params [bits]*(L0+P), chain_index numbering
 *                                   scripts/bootstrap_generation.py:92-102
 *   - op surface                    gpu/phantom_binding.cu:165-205
 * Everything else follows the published SEAL 4.x / Phantom algorithms
 * (prime scan, minimal 2N-th root, Harvey NTT in bit-reversed order, hybrid
 * key switching with dnum = ceil(L/P) digits, rounding ModDown / rescale) and
 * is written down in DESIGN.md.
 *
 * Plain C, exact integer arithmetic with unsigned __int128; no laziness, every
 * intermediate fully reduced.  The CUDA path must reproduce every output limb
 * bit for bit.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef uint32_t u32;
typedef unsigned __int128 u128;

/* OpenMP team size, set explicitly by the timing legs of bench.py: launchers such as torch.distributed.run export
 * OMP_NUM_THREADS=1, which would silently turn the CPU baseline into a single-core number.  Returns the team size. */
int orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

#define ORC_MAX_LIMBS 64

/* ------------------------------------------------------------------ */
/* modular helpers                                                     */
/* ------------------------------------------------------------------ */
static inline u64 addmod(u64 a, u64 b, u64 q) { u64 s = a + b; return s >= q ? s - q : s; }
static inline u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
static inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }
static inline u64 negmod(u64 a, u64 q) { return a ? q - a : 0; }

static u64 powmod(u64 a, u64 e, u64 q) {
    u64 r = 1; a %= q;
    while (e) { if (e & 1) r = mulmod(r, a, q); a = mulmod(a, a, q); e >>= 1; }
    return r;
}
static u64 invmod(u64 a, u64 q) { return powmod(a % q, q - 2, q); }

static int is_prime_u64(u64 n) {
    static const u64 bases[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return 0;
    for (int i = 0; i < 12; i++) { if (n % bases[i] == 0) return n == bases[i]; }
    u64 d = n - 1; int r = 0;
    while ((d & 1) == 0) { d >>= 1; r++; }
    for (int i = 0; i < 12; i++) {
        u64 x = powmod(bases[i], d, n);
        if (x == 1 || x == n - 1) continue;
        int comp = 1;
        for (int k = 1; k < r; k++) { x = mulmod(x, x, n); if (x == n - 1) { comp = 0; break; } }
        if (comp) return 0;
    }
    return 1;
}

static inline u32 bitrev(u32 x, int bits) {
    u32 r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* ------------------------------------------------------------------ */
/* prime generation (SEAL get_primes / CoeffModulus::Create restated)  */
/* ------------------------------------------------------------------ */
/* descending scan from the largest value < 2^bits congruent 1 mod 2N */
int orc_gen_primes(u64 N, int bits, int count, u64 *out) {
    u64 factor = 2 * N;
    u64 v = ((((u64)1 << bits) - 1) / factor) * factor + 1;
    u64 lower = (u64)1 << (bits - 1);
    int got = 0;
    while (got < count && v > lower) {
        if (is_prime_u64(v)) out[got++] = v;
        v -= factor;
    }
    return got == count ? 0 : -1;
}

/* bit_sizes in request order; primes of one size are handed out smallest first */
int orc_create_coeff_modulus(u64 N, const int *bit_sizes, int n, u64 *out) {
    int cnt[65]; memset(cnt, 0, sizeof cnt);
    for (int i = 0; i < n; i++) { if (bit_sizes[i] < 2 || bit_sizes[i] > 61) return -1; cnt[bit_sizes[i]]++; }
    u64 *tab[65]; int left[65];
    for (int b = 0; b < 65; b++) {
        tab[b] = NULL; left[b] = cnt[b];
        if (cnt[b]) {
            tab[b] = (u64 *)malloc(sizeof(u64) * cnt[b]);
            if (orc_gen_primes(N, b, cnt[b], tab[b])) return -2;
        }
    }
    for (int i = 0; i < n; i++) { int b = bit_sizes[i]; out[i] = tab[b][--left[b]]; }
    for (int b = 0; b < 65; b++) free(tab[b]);
    return 0;
}

/* minimal primitive 2N-th root of unity mod q */
static u64 minimal_primitive_root(u64 N, u64 q) {
    u64 two_n = 2 * N, c = 0;
    for (u64 g = 2; g < q; g++) {
        c = powmod(g, (q - 1) / two_n, q);
        if (powmod(c, N, q) == q - 1) break;
    }
    u64 sq = mulmod(c, c, q), cur = c, best = c;
    for (u64 k = 1; k < N; k++) { cur = mulmod(cur, sq, q); if (cur < best) best = cur; }
    return best;
}

/* ------------------------------------------------------------------ */
/* context                                                             */
/* ------------------------------------------------------------------ */
typedef struct {
    u64 N; int logn; int K, P, L;          /* K = L + P limbs, specials last */
    u64 q[ORC_MAX_LIMBS];
    u64 *psi_rev[ORC_MAX_LIMBS];           /* psi^{bitrev(i)}   */
    u64 *psi_inv_rev[ORC_MAX_LIMBS];       /* psi^{-bitrev(i)}  */
    u64 inv_n[ORC_MAX_LIMBS];
    u64 psi[ORC_MAX_LIMBS];
    double *zr, *zi;                        /* zeta^{bitrev(i)}, zeta = exp(i*pi/N) */
} orc_ctx;

orc_ctx *orc_ctx_create(u64 N, int K, int P, const u64 *moduli) {
    if (K > ORC_MAX_LIMBS || P >= K) return NULL;
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
    c->N = N; c->K = K; c->P = P; c->L = K - P;
    c->logn = 0; while (((u64)1 << c->logn) < N) c->logn++;
    for (int i = 0; i < K; i++) {
        u64 q = moduli[i]; c->q[i] = q;
        u64 psi = minimal_primitive_root(N, q), ipsi = invmod(psi, q);
        c->psi[i] = psi;
        c->psi_rev[i] = (u64 *)malloc(sizeof(u64) * N);
        c->psi_inv_rev[i] = (u64 *)malloc(sizeof(u64) * N);
        u64 p = 1, ip = 1;
        for (u64 k = 0; k < N; k++) {
            u32 r = bitrev((u32)k, c->logn);
            c->psi_rev[i][r] = p; c->psi_inv_rev[i][r] = ip;
            p = mulmod(p, psi, q); ip = mulmod(ip, ipsi, q);
        }
        c->inv_n[i] = invmod(N % q, q);
    }
    c->zr = (double *)malloc(sizeof(double) * N);
    c->zi = (double *)malloc(sizeof(double) * N);
    for (u64 k = 0; k < N; k++) {
        u32 r = bitrev((u32)k, c->logn);
        double ang = M_PI * (double)k / (double)N;
        c->zr[r] = cos(ang); c->zi[r] = sin(ang);
    }
    return c;
}
void orc_ctx_free(orc_ctx *c) {
    if (!c) return;
    for (int i = 0; i < c->K; i++) { free(c->psi_rev[i]); free(c->psi_inv_rev[i]); }
    free(c->zr); free(c->zi); free(c);
}
u64 orc_ctx_psi(orc_ctx *c, int limb) { return c->psi[limb]; }

/* ------------------------------------------------------------------ */
/* negacyclic NTT, natural in -> bit-reversed out, and inverse         */
/* n may be a prefix size (n <= N, power of two): the size-n transform */
/* with root psi^(N/n) uses the first n table entries.                 */
/* ------------------------------------------------------------------ */
static void ntt_fwd_n(const orc_ctx *c, int limb, u64 *a, u64 n) {
    u64 q = c->q[limb]; const u64 *w = c->psi_rev[limb];
    u64 t = n;
    for (u64 m = 1; m < n; m <<= 1) {
        t >>= 1;
        for (u64 i = 0; i < m; i++) {
            u64 W = w[m + i], j1 = 2 * i * t;
            for (u64 j = j1; j < j1 + t; j++) {
                u64 U = a[j], V = mulmod(a[j + t], W, q);
                a[j] = addmod(U, V, q); a[j + t] = submod(U, V, q);
            }
        }
    }
}
static void ntt_inv_n(const orc_ctx *c, int limb, u64 *a, u64 n) {
    u64 q = c->q[limb]; const u64 *w = c->psi_inv_rev[limb];
    u64 t = 1;
    for (u64 m = n; m > 1; m >>= 1) {
        u64 h = m >> 1;
        for (u64 i = 0; i < h; i++) {
            u64 W = w[h + i], j1 = 2 * i * t;
            for (u64 j = j1; j < j1 + t; j++) {
                u64 U = a[j], V = a[j + t];
                a[j] = addmod(U, V, q); a[j + t] = mulmod(submod(U, V, q), W, q);
            }
        }
        t <<= 1;
    }
    u64 invn = invmod(n % q, q);
    for (u64 j = 0; j < n; j++) a[j] = mulmod(a[j], invn, q);
}
void orc_ntt_fwd(const orc_ctx *c, int limb, u64 *a) { ntt_fwd_n(c, limb, a, c->N); }
void orc_ntt_inv(const orc_ctx *c, int limb, u64 *a) { ntt_inv_n(c, limb, a, c->N); }
void orc_ntt_fwd_sub(const orc_ctx *c, int limb, u64 *a, u64 n) { ntt_fwd_n(c, limb, a, n); }

/* limb id of row r of an object that lives on l data limbs (+ P specials when ext) */
static inline int row_limb(const orc_ctx *c, int l, int r) { return r < l ? r : c->L + (r - l); }

/* ------------------------------------------------------------------ */
/* ChaCha20 counter-mode PRNG: key = 32-byte seed, 64-bit nonce = stream,*/
/* 64-bit block counter.  Word w of a stream = u64 #(w&7) of block w>>3. */
/* ------------------------------------------------------------------ */
#define ROTL32(v, n) (((v) << (n)) | ((v) >> (32 - (n))))
#define QR(a, b, c, d) \
    a += b; d ^= a; d = ROTL32(d, 16); c += d; b ^= c; b = ROTL32(b, 12); \
    a += b; d ^= a; d = ROTL32(d, 8);  c += d; b ^= c; b = ROTL32(b, 7);

static void chacha_block(const u32 key[8], u64 nonce, u64 counter, u64 out[8]) {
    u32 s[16], x[16];
    s[0] = 0x61707865; s[1] = 0x3320646e; s[2] = 0x79622d32; s[3] = 0x6b206574;
    for (int i = 0; i < 8; i++) s[4 + i] = key[i];
    s[12] = (u32)counter; s[13] = (u32)(counter >> 32);
    s[14] = (u32)nonce;   s[15] = (u32)(nonce >> 32);
    memcpy(x, s, sizeof s);
    for (int r = 0; r < 10; r++) {
        QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13])
        QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
        QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12])
        QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) x[i] += s[i];
    for (int i = 0; i < 8; i++) out[i] = (u64)x[2 * i] | ((u64)x[2 * i + 1] << 32);
}
static inline u64 prng_word(const u32 key[8], u64 nonce, u64 w) {
    u64 blk[8]; chacha_block(key, nonce, w >> 3, blk); return blk[w & 7];
}
void orc_prng_words(const u32 *key, u64 nonce, u64 first, u64 count, u64 *out) {
    for (u64 i = 0; i < count; i++) out[i] = prng_word(key, nonce, first + i);
}

/* stream ids: (domain << 56) | id  -- shared verbatim with the CUDA library */
enum { DOM_SK = 1, DOM_PK_A = 2, DOM_PK_E = 3, DOM_KSK_A = 4, DOM_KSK_E = 5,
       DOM_ENC_A = 6, DOM_ENC_E = 7, DOM_ASYM_U = 8, DOM_ASYM_E0 = 9, DOM_ASYM_E1 = 10, DOM_PK_SEED = 11 };
static inline u64 stream_id(int dom, u64 id) { return ((u64)dom << 56) | (id & 0x00FFFFFFFFFFFFFFull); }
/* the seed a public key carries for its encryption randomness: the first 32 bytes of the secret seed's stream
 * DOM_PK_SEED (one-way, so the public key does not reveal the secret seed) */
void orc_public_key_seed(const u32 *seed, u32 out[8]) {
    u64 blk[8]; chacha_block(seed, stream_id(DOM_PK_SEED, 0), 0, blk);
    memcpy(out, blk, 32);
}

/* uniform residue for (limb id, coefficient j): words 2*(limb*N+j), +1 -> 128 bit mod q */
static void sample_uniform_row(const orc_ctx *c, const u32 *key, u64 nonce, int limb, u64 *out) {
    u64 q = c->q[limb], N = c->N;
    for (u64 j = 0; j < N; j += 4) {
        u64 blk[8]; chacha_block(key, nonce, ((u64)limb * N + j) >> 2, blk);
        for (int k = 0; k < 4; k++) {
            u128 v = ((u128)blk[2 * k] << 64) | blk[2 * k + 1];
            out[j + k] = (u64)(v % q);
        }
    }
}
/* ternary in {-1,0,1}: floor(3*w / 2^64) - 1, word j */
static inline int sample_ternary(const u32 *key, u64 nonce, u64 j) {
    u64 w = prng_word(key, nonce, j);
    return (int)(((u128)w * 3) >> 64) - 1;
}
/* centred binomial, 21+21 bits of word j (variance 10.5, as SEAL's default noise) */
static inline int sample_cbd(const u32 *key, u64 nonce, u64 j) {
    u64 w = prng_word(key, nonce, j);
    return __builtin_popcountll(w & 0x1FFFFF) - __builtin_popcountll((w >> 21) & 0x1FFFFF);
}
static void small_poly_rows(const orc_ctx *c, const int *vals, int l, int ext, u64 *out /*[rows][N]*/) {
    int rows = l + (ext ? c->P : 0);
    #pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; r++) {
        int limb = row_limb(c, l, r); u64 q = c->q[limb];
        u64 *o = out + (u64)r * c->N;
        for (u64 j = 0; j < c->N; j++) o[j] = vals[j] >= 0 ? (u64)vals[j] : q - (u64)(-vals[j]);
        orc_ntt_fwd(c, limb, o);
    }
}

/* secret key: ternary, NTT form on all K limbs */
void orc_gen_secret(const orc_ctx *c, const u32 *seed, u64 *sk /*[K][N]*/) {
    int *v = (int *)malloc(sizeof(int) * c->N);
    u64 nonce = stream_id(DOM_SK, 0);
    for (u64 j = 0; j < c->N; j++) v[j] = sample_ternary(seed, nonce, j);
    small_poly_rows(c, v, c->L, 1, sk);
    free(v);
}

/* ------------------------------------------------------------------ */
/* Galois automorphism in the NTT (bit-reversed) domain                */
/* out[i] = in[ bitrev( ((elt*(2*bitrev(i)+1) mod 2N) - 1)/2 ) ]        */
/* ------------------------------------------------------------------ */
static inline u32 galois_src(const orc_ctx *c, u32 elt, u32 i) {
    u64 m = 2 * c->N;
    u64 k = ((u64)elt * (2 * (u64)bitrev(i, c->logn) + 1)) & (m - 1);
    return bitrev((u32)((k - 1) >> 1), c->logn);
}
void orc_apply_galois_ntt(const orc_ctx *c, u32 elt, const u64 *in, u64 *out) {
    for (u32 i = 0; i < c->N; i++) out[i] = in[galois_src(c, elt, i)];
}
u64 orc_galois_elt_from_step(u64 N, int step) {
    u64 m = 2 * N, half = N / 2;
    if (step == 0) return m - 1;                 /* conjugation, as SEAL */
    u64 e = step > 0 ? (u64)step % half : half - ((u64)(-step) % half);
    u64 r = 1, g = 5;
    while (e) { if (e & 1) r = (r * g) & (m - 1); g = (g * g) & (m - 1); e >>= 1; }
    return r;
}

/* ------------------------------------------------------------------ */
/* CKKS encode / decode                                                */
/* ------------------------------------------------------------------ */
static void cfft_inv(const orc_ctx *c, double *ar, double *ai, u64 n) {
    /* Gentleman-Sande with conj(zeta) powers; no 1/n (folded into fix) */
    u64 t = 1;
    for (u64 m = n; m > 1; m >>= 1) {
        u64 h = m >> 1;
        for (u64 i = 0; i < h; i++) {
            double wr = c->zr[h + i], wi = -c->zi[h + i];
            u64 j1 = 2 * i * t;
            for (u64 j = j1; j < j1 + t; j++) {
                double ur = ar[j], ui = ai[j], vr = ar[j + t], vi = ai[j + t];
                ar[j] = ur + vr; ai[j] = ui + vi;
                double dr = ur - vr, di = ui - vi;
                ar[j + t] = dr * wr - di * wi;
                ai[j + t] = dr * wi + di * wr;
            }
        }
        t <<= 1;
    }
}
static void cfft_fwd(const orc_ctx *c, double *ar, double *ai, u64 n) {
    u64 t = n;
    for (u64 m = 1; m < n; m <<= 1) {
        t >>= 1;
        for (u64 i = 0; i < m; i++) {
            double wr = c->zr[m + i], wi = c->zi[m + i];
            u64 j1 = 2 * i * t;
            for (u64 j = j1; j < j1 + t; j++) {
                double xr = ar[j + t], xi = ai[j + t];
                double vr = xr * wr - xi * wi, vi = xr * wi + xi * wr;
                double ur = ar[j], ui = ai[j];
                ar[j] = ur + vr; ai[j] = ui + vi;
                ar[j + t] = ur - vr; ai[j + t] = ui - vi;
            }
        }
    }
}
/* slot j of an n-coefficient ring sits at transform index bitrev((5^j mod 2n - 1)/2) */
static void slot_indices(u64 n, int logn, u32 *idx, u32 *idxc) {
    u64 m = 2 * n, pos = 1;
    for (u64 j = 0; j < n / 2; j++) {
        idx[j] = bitrev((u32)((pos - 1) >> 1), logn);
        idxc[j] = bitrev((u32)((m - pos - 1) >> 1), logn);
        pos = (pos * 5) & (m - 1);
    }
}
static inline u64 int_double_mod(double x, u64 q) {
    /* x is integer valued, |x| < 2^127 */
    int neg = x < 0; if (neg) x = -x;
    u128 v = (u128)x;
    u64 r = (u64)(v % q);
    return neg ? negmod(r, q) : r;
}
/*
 * Encode n/2 complex slot values into an n-coefficient plaintext (n = N for the
 * ordinary encoder; n = 2*D for the sub-ring encoder of period-D vectors, whose
 * coefficients are those of X^(N/n)).  rows: l data limbs (+P special if ext).
 * Output NTT form of size n per row (first n entries of each N-stride row when
 * n < N are written contiguously with stride n).
 */
int orc_encode(const orc_ctx *c, u64 n, const double *re, const double *im, double scale,
               int l, int ext, u64 *out) {
    int logn = 0; while (((u64)1 << logn) < n) logn++;
    double *ar = (double *)calloc(n, sizeof(double)), *ai = (double *)calloc(n, sizeof(double));
    u32 *idx = (u32 *)malloc(sizeof(u32) * n / 2), *idxc = (u32 *)malloc(sizeof(u32) * n / 2);
    slot_indices(n, logn, idx, idxc);
    for (u64 j = 0; j < n / 2; j++) {
        double i_ = im ? im[j] : 0.0;
        ar[idx[j]] = re[j]; ai[idx[j]] = i_;
        ar[idxc[j]] = re[j]; ai[idxc[j]] = -i_;
    }
    cfft_inv(c, ar, ai, n);
    double fix = scale / (double)n;
    int rows = l + (ext ? c->P : 0), rc = 0;
    for (u64 j = 0; j < n; j++) {
        ar[j] = rint(ar[j] * fix);
        if (!(fabs(ar[j]) < 0x1p126)) rc = -1;
    }
    if (!rc) for (int r = 0; r < rows; r++) {
        int limb = row_limb(c, l, r);
        u64 *o = out + (u64)r * n;
        for (u64 j = 0; j < n; j++) o[j] = int_double_mod(ar[j], c->q[limb]);
        ntt_fwd_n(c, limb, o, n);
    }
    free(ar); free(ai); free(idx); free(idxc);
    return rc;
}
/* decode: first k = min(l,3) limbs, Garner mixed radix, centred, /scale, forward FFT */
void orc_decode(const orc_ctx *c, const u64 *pt /*[l][N]*/, int l, double scale, double *re, double *im) {
    u64 N = c->N; int k = l < 3 ? l : 3;
    u64 *x = (u64 *)malloc(sizeof(u64) * N * k);
    for (int i = 0; i < k; i++) { memcpy(x + i * N, pt + (u64)i * N, sizeof(u64) * N); orc_ntt_inv(c, i, x + i * N); }
    u64 inv[3][3];
    for (int i = 0; i < k; i++) for (int j = 0; j < i; j++) inv[i][j] = invmod(c->q[j] % c->q[i], c->q[i]);
    double W[3]; W[0] = 1.0; if (k > 1) W[1] = (double)c->q[0]; if (k > 2) W[2] = (double)c->q[0] * (double)c->q[1];
    double *ar = (double *)malloc(sizeof(double) * N), *ai = (double *)calloc(N, sizeof(double));
    for (u64 n = 0; n < N; n++) {
        u64 v[3];
        for (int i = 0; i < k; i++) {
            u64 qi = c->q[i], t = x[i * N + n];
            for (int j = 0; j < i; j++) t = mulmod(submod(t, v[j] % qi, qi), inv[i][j], qi);
            v[i] = t;
        }
        int neg = 0;
        for (int i = k - 1; i >= 0; i--) {
            u64 h = (c->q[i] - 1) >> 1;
            if (v[i] > h) { neg = 1; break; }
            if (v[i] < h) break;
        }
        double acc = 0.0;
        if (neg) {
            for (int i = k - 1; i >= 0; i--) acc = acc + (double)(c->q[i] - 1 - v[i]) * W[i];
            acc = -(acc + 1.0);
        } else {
            for (int i = k - 1; i >= 0; i--) acc = acc + (double)v[i] * W[i];
        }
        ar[n] = acc / scale;
    }
    cfft_fwd(c, ar, ai, N);
    u32 *idx = (u32 *)malloc(sizeof(u32) * N / 2), *idxc = (u32 *)malloc(sizeof(u32) * N / 2);
    slot_indices(N, c->logn, idx, idxc);
    for (u64 j = 0; j < N / 2; j++) { re[j] = ar[idx[j]]; im[j] = ai[idx[j]]; }
    free(x); free(ar); free(ai); free(idx); free(idxc);
}

/* ------------------------------------------------------------------ */
/* encryption / decryption                                             */
/* ------------------------------------------------------------------ */
/* symmetric: c1 = a (uniform, NTT form), c0 = m + e - a*s ; enc_id picks the streams */
void orc_encrypt_symmetric(const orc_ctx *c, const u32 *seed, u64 enc_id, const u64 *sk,
                           const u64 *pt /*[l][N]*/, int l, u64 *ct /*[2][l][N]*/) {
    u64 N = c->N;
    int *ev = (int *)malloc(sizeof(int) * N);
    u64 ne = stream_id(DOM_ENC_E, enc_id), na = stream_id(DOM_ENC_A, enc_id);
    for (u64 j = 0; j < N; j++) ev[j] = sample_cbd(seed, ne, j);
    u64 *e = (u64 *)malloc(sizeof(u64) * N * l);
    small_poly_rows(c, ev, l, 0, e);
    for (int i = 0; i < l; i++) {
        u64 q = c->q[i]; u64 *c0 = ct + (u64)i * N, *c1 = ct + ((u64)l + i) * N;
        sample_uniform_row(c, seed, na, i, c1);
        for (u64 j = 0; j < N; j++)
            c0[j] = submod(addmod(pt[(u64)i * N + j], e[(u64)i * N + j], q), mulmod(c1[j], sk[(u64)i * N + j], q), q);
    }
    free(ev); free(e);
}
/* m = c0 + c1 s (+ c2 s^2) */
void orc_decrypt(const orc_ctx *c, const u64 *sk, const u64 *ct, int size, int l, u64 *pt) {
    u64 N = c->N;
    for (int i = 0; i < l; i++) {
        u64 q = c->q[i]; const u64 *s = sk + (u64)i * N;
        for (u64 j = 0; j < N; j++) {
            u64 acc = ct[((u64)(size - 1) * l + i) * N + j];
            for (int p = size - 2; p >= 0; p--) acc = addmod(mulmod(acc, s[j], q), ct[((u64)p * l + i) * N + j], q);
            pt[(u64)i * N + j] = acc;
        }
    }
}

/* ------------------------------------------------------------------ */
/* key switching keys                                                  */
/* key[j][poly][K][N], j < beta = ceil(L/P):                           */
/*   k1 = a_j, k0 = -a_j s + e_j + [P mod q_i] * snew on limbs i of digit j */
/* tag identifies the key (galois element, or 0 for the relin key)     */
/* ------------------------------------------------------------------ */
static int n_digits(const orc_ctx *c, int l) { return (l + c->P - 1) / c->P; }
int orc_num_digits(const orc_ctx *c, int l) { return n_digits(c, l); }

static u64 pmod(const orc_ctx *c, int limb) {
    u64 r = 1; for (int k = 0; k < c->P; k++) r = mulmod(r, c->q[c->L + k] % c->q[limb], c->q[limb]);
    return r;
}
void orc_gen_switch_key(const orc_ctx *c, const u32 *seed, u64 tag, const u64 *sk,
                        const u64 *snew /*[K][N] NTT*/, u64 *key) {
    u64 N = c->N; int K = c->K, beta = n_digits(c, c->L);
    int *ev = (int *)malloc(sizeof(int) * N);
    u64 *e = (u64 *)malloc(sizeof(u64) * N * K);
    for (int j = 0; j < beta; j++) {
        u64 id = (tag << 8) | (u64)j;
        u64 na = stream_id(DOM_KSK_A, id), ne = stream_id(DOM_KSK_E, id);
        for (u64 n = 0; n < N; n++) ev[n] = sample_cbd(seed, ne, n);
        small_poly_rows(c, ev, c->L, 1, e);
        u64 *k0 = key + ((u64)j * 2 + 0) * K * N, *k1 = key + ((u64)j * 2 + 1) * K * N;
        #pragma omp parallel for schedule(static)
        for (int i = 0; i < K; i++) {
            u64 q = c->q[i];
            sample_uniform_row(c, seed, na, i, k1 + (u64)i * N);
            int own = (i < c->L) && (i / c->P == j);
            u64 pm = own ? pmod(c, i) : 0;
            for (u64 n = 0; n < N; n++) {
                u64 v = submod(e[(u64)i * N + n], mulmod(k1[(u64)i * N + n], sk[(u64)i * N + n], q), q);
                if (own) v = addmod(v, mulmod(pm, snew[(u64)i * N + n], q), q);
                k0[(u64)i * N + n] = v;
            }
        }
    }
    free(ev); free(e);
}
void orc_gen_galois_key(const orc_ctx *c, const u32 *seed, u32 elt, const u64 *sk, u64 *key) {
    u64 N = c->N; u64 *sn = (u64 *)malloc(sizeof(u64) * N * c->K);
    for (int i = 0; i < c->K; i++) orc_apply_galois_ntt(c, elt, sk + (u64)i * N, sn + (u64)i * N);
    orc_gen_switch_key(c, seed, elt, sk, sn, key);
    free(sn);
}
void orc_gen_relin_key(const orc_ctx *c, const u32 *seed, const u64 *sk, u64 *key) {
    u64 N = c->N; u64 *sn = (u64 *)malloc(sizeof(u64) * N * c->K);
    for (int i = 0; i < c->K; i++) for (u64 n = 0; n < N; n++) sn[(u64)i * N + n] = mulmod(sk[(u64)i * N + n], sk[(u64)i * N + n], c->q[i]);
    orc_gen_switch_key(c, seed, 0, sk, sn, key);
    free(sn);
}
/* public key (pk0, pk1) = (e - a s, a) on all K limbs */
void orc_gen_public_key(const orc_ctx *c, const u32 *seed, const u64 *sk, u64 *pk /*[2][K][N]*/) {
    u64 N = c->N; int K = c->K;
    int *ev = (int *)malloc(sizeof(int) * N);
    u64 *e = (u64 *)malloc(sizeof(u64) * N * K);
    u64 na = stream_id(DOM_PK_A, 0), ne = stream_id(DOM_PK_E, 0);
    for (u64 n = 0; n < N; n++) ev[n] = sample_cbd(seed, ne, n);
    small_poly_rows(c, ev, c->L, 1, e);
    for (int i = 0; i < K; i++) {
        u64 q = c->q[i]; u64 *p0 = pk + (u64)i * N, *p1 = pk + ((u64)K + i) * N;
        sample_uniform_row(c, seed, na, i, p1);
        for (u64 n = 0; n < N; n++) p0[n] = submod(e[(u64)i * N + n], mulmod(p1[n], sk[(u64)i * N + n], q), q);
    }
    free(ev); free(e);
}

/* ------------------------------------------------------------------ */
/* RNS base conversion pieces                                          */
/* ------------------------------------------------------------------ */
/* digit j of level l: limbs [j*P, min((j+1)*P, l)) */
static void digit_range(const orc_ctx *c, int l, int j, int *lo, int *hi) {
    *lo = j * c->P; *hi = (j + 1) * c->P; if (*hi > l) *hi = l;
}
/*
 * ModUp: x[l][N] coefficient form -> E[beta][l+P][N] coefficient form.
 * own-digit rows are copied; others get sum_i [x_i * (Qj/q_i)^-1]_{q_i} * (Qj/q_i) mod t
 */
static void modup(const orc_ctx *c, int l, const u64 *x, u64 *E) {
    u64 N = c->N; int rows = l + c->P, beta = n_digits(c, l);
    for (int j = 0; j < beta; j++) {
        int lo, hi; digit_range(c, l, j, &lo, &hi);
        int a = hi - lo;
        u64 hatinv[ORC_MAX_LIMBS];
        for (int i = lo; i < hi; i++) {
            u64 qi = c->q[i], h = 1;
            for (int k = lo; k < hi; k++) if (k != i) h = mulmod(h, c->q[k] % qi, qi);
            hatinv[i - lo] = invmod(h, qi);
        }
        u64 *Ej = E + (u64)j * rows * N;
        #pragma omp parallel for schedule(static)
        for (int r = 0; r < rows; r++) {
            int t = row_limb(c, l, r);
            u64 *o = Ej + (u64)r * N;
            if (t >= lo && t < hi) { memcpy(o, x + (u64)t * N, sizeof(u64) * N); continue; }
            u64 qt = c->q[t], hat[ORC_MAX_LIMBS];
            for (int i = lo; i < hi; i++) {
                u64 h = 1;
                for (int k = lo; k < hi; k++) if (k != i) h = mulmod(h, c->q[k] % qt, qt);
                hat[i - lo] = h;
            }
            for (u64 n = 0; n < N; n++) {
                u64 acc = 0;
                for (int i = 0; i < a; i++) {
                    u64 y = mulmod(x[(u64)(lo + i) * N + n], hatinv[i], c->q[lo + i]);
                    acc = addmod(acc, mulmod(y % qt, hat[i], qt), qt);
                }
                o[n] = acc;
            }
        }
    }
}
/*
 * decompose c (NTT form, l limbs) into extended NTT-form digits E[beta][l+P][N]:
 * own-digit rows are the untouched NTT-form input limbs.
 */
void orc_decompose(const orc_ctx *c, int l, const u64 *cin, u64 *E) {
    u64 N = c->N; int rows = l + c->P, beta = n_digits(c, l);
    u64 *x = (u64 *)malloc(sizeof(u64) * N * l);
    memcpy(x, cin, sizeof(u64) * N * l);
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < l; i++) orc_ntt_inv(c, i, x + (u64)i * N);
    modup(c, l, x, E);
    #pragma omp parallel for schedule(static) collapse(2)
    for (int j = 0; j < beta; j++) for (int r = 0; r < rows; r++) {
        int lo, hi; digit_range(c, l, j, &lo, &hi);
        int t = row_limb(c, l, r);
        u64 *o = E + ((u64)j * rows + r) * N;
        if (t >= lo && t < hi) memcpy(o, cin + (u64)t * N, sizeof(u64) * N);
        else orc_ntt_fwd(c, t, o);
    }
    free(x);
}
/*
 * inner product of (optionally Galois-permuted) digits with a switching key:
 * out[poly][l+P][N] (+)= sum_j perm(E_j) * key_j[poly]   (elt = 0: no permutation)
 */
static void ks_inner(const orc_ctx *c, int l, const u64 *E, u32 elt, const u64 *key, u64 *out, int accumulate) {
    u64 N = c->N; int rows = l + c->P, beta = n_digits(c, l), K = c->K;
    #pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; r++) {
        int t = row_limb(c, l, r); u64 q = c->q[t];
        for (u64 n = 0; n < N; n++) {
            u64 src = elt ? galois_src(c, elt, (u32)n) : n;
            u64 a0 = accumulate ? out[((u64)0 * rows + r) * N + n] : 0;
            u64 a1 = accumulate ? out[((u64)1 * rows + r) * N + n] : 0;
            for (int j = 0; j < beta; j++) {
                u64 d = E[((u64)j * rows + r) * N + src];
                a0 = addmod(a0, mulmod(d, key[(((u64)j * 2 + 0) * K + t) * N + n], q), q);
                a1 = addmod(a1, mulmod(d, key[(((u64)j * 2 + 1) * K + t) * N + n], q), q);
            }
            out[((u64)0 * rows + r) * N + n] = a0;
            out[((u64)1 * rows + r) * N + n] = a1;
        }
    }
}
/*
 * ModDown one polynomial: in[l+P][N] NTT form over Q_l*P -> out[l][N] NTT form,
 * out = (in - [in + floor(P/2)]_P + floor(P/2)) / P    (rounded division by P)
 */
static void moddown_poly(const orc_ctx *c, int l, const u64 *in, u64 *out) {
    u64 N = c->N; int P = c->P, L = c->L;
    u64 *sp = (u64 *)malloc(sizeof(u64) * N * P);
    /* floor(P/2) mod each modulus: P odd -> (P-1)/2; compute via big number mod */
    /* half_mod(m) = ((P mod 2m) - 1)/2 mod m  (P mod 2m is odd) */
    u64 hatinv[ORC_MAX_LIMBS];
    for (int k = 0; k < P; k++) {
        u64 pk = c->q[L + k], h = 1;
        for (int k2 = 0; k2 < P; k2++) if (k2 != k) h = mulmod(h, c->q[L + k2] % pk, pk);
        hatinv[k] = invmod(h, pk);
    }
    #pragma omp parallel for schedule(static)
    for (int k = 0; k < P; k++) {
        u64 pk = c->q[L + k];
        memcpy(sp + (u64)k * N, in + (u64)(l + k) * N, sizeof(u64) * N);
        orc_ntt_inv(c, L + k, sp + (u64)k * N);
        /* half mod pk */
        u128 m2 = (u128)2 * pk; u128 pm = 1;
        for (int k2 = 0; k2 < P; k2++) pm = (pm * (c->q[L + k2] % m2)) % m2;
        u64 half = (u64)(((pm - 1) >> 1) % pk);
        for (u64 n = 0; n < N; n++)
            sp[(u64)k * N + n] = mulmod(addmod(sp[(u64)k * N + n], half, pk), hatinv[k], pk);
    }
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < l; i++) {
        u64 qi = c->q[i], hat[ORC_MAX_LIMBS];
        for (int k = 0; k < P; k++) {
            u64 h = 1;
            for (int k2 = 0; k2 < P; k2++) if (k2 != k) h = mulmod(h, c->q[L + k2] % qi, qi);
            hat[k] = h;
        }
        u128 m2 = (u128)2 * qi; u128 pm = 1;
        for (int k2 = 0; k2 < P; k2++) pm = (pm * (c->q[L + k2] % m2)) % m2;
        u64 half = (u64)(((pm - 1) >> 1) % qi);
        u64 pinv = invmod(pmod(c, i), qi);
        u64 *o = out + (u64)i * N;
        for (u64 n = 0; n < N; n++) {
            u64 acc = 0;
            for (int k = 0; k < P; k++) acc = addmod(acc, mulmod(sp[(u64)k * N + n] % qi, hat[k], qi), qi);
            o[n] = submod(acc, half, qi);
        }
        orc_ntt_fwd(c, i, o);
        for (u64 n = 0; n < N; n++) o[n] = mulmod(submod(in[(u64)i * N + n], o[n], qi), pinv, qi);
    }
    free(sp);
}
void orc_moddown(const orc_ctx *c, int l, const u64 *in, u64 *out) { moddown_poly(c, l, in, out); }

/* full key switch of one NTT-form polynomial: (o0, o1) = ModDown(sum_j E_j key_j) */
void orc_keyswitch(const orc_ctx *c, int l, const u64 *cin, const u64 *key, u64 *o0, u64 *o1) {
    u64 N = c->N; int rows = l + c->P, beta = n_digits(c, l);
    u64 *E = (u64 *)malloc(sizeof(u64) * N * rows * beta);
    u64 *acc = (u64 *)malloc(sizeof(u64) * N * rows * 2);
    orc_decompose(c, l, cin, E);
    ks_inner(c, l, E, 0, key, acc, 0);
    moddown_poly(c, l, acc, o0);
    moddown_poly(c, l, acc + (u64)rows * N, o1);
    free(E); free(acc);
}
/* apply_galois (reference op order): permute both polys, then key switch the permuted c1 */
void orc_apply_galois(const orc_ctx *c, int l, const u64 *ct, u32 elt, const u64 *key, u64 *out) {
    u64 N = c->N;
    u64 *p0 = (u64 *)malloc(sizeof(u64) * N * l), *p1 = (u64 *)malloc(sizeof(u64) * N * l);
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < l; i++) {
        orc_apply_galois_ntt(c, elt, ct + (u64)i * N, p0 + (u64)i * N);
        orc_apply_galois_ntt(c, elt, ct + ((u64)l + i) * N, p1 + (u64)i * N);
    }
    orc_keyswitch(c, l, p1, key, out, out + (u64)l * N);
    for (int i = 0; i < l; i++) for (u64 n = 0; n < N; n++)
        out[(u64)i * N + n] = addmod(out[(u64)i * N + n], p0[(u64)i * N + n], c->q[i]);
    free(p0); free(p1);
}
/* hoisted rotation (ph.hoisting): decompose c1 once, permute the digits, ModDown:
 * out = ModDown( P*pi(c0) + <pi(E),k0> , <pi(E),k1> ) */
void orc_hoisted_rotation(const orc_ctx *c, int l, const u64 *ct, u32 elt, const u64 *key, u64 *out) {
    u64 N = c->N; int rows = l + c->P, beta = n_digits(c, l);
    u64 *E = (u64 *)malloc(sizeof(u64) * N * rows * beta);
    u64 *acc = (u64 *)malloc(sizeof(u64) * N * rows * 2);
    orc_decompose(c, l, ct + (u64)l * N, E);
    ks_inner(c, l, E, elt, key, acc, 0);
    for (int i = 0; i < l; i++) { u64 q = c->q[i], pm = pmod(c, i);
        for (u64 n = 0; n < N; n++)
            acc[(u64)i * N + n] = addmod(acc[(u64)i * N + n], mulmod(pm, ct[(u64)i * N + galois_src(c, elt, (u32)n)], q), q); }
    moddown_poly(c, l, acc, out);
    moddown_poly(c, l, acc + (u64)rows * N, out + (u64)l * N);
    free(E); free(acc);
}
/* asymmetric encryption: (pk0*u + e0, pk1*u + e1) on Q_l*P, ModDown, + m */
void orc_encrypt_asymmetric(const orc_ctx *c, const u32 *seed, u64 enc_id, const u64 *pk /*[2][K][N]*/,
                            const u64 *pt, int l, u64 *ct) {
    u64 N = c->N; int rows = l + c->P, K = c->K;
    int *v = (int *)malloc(sizeof(int) * N);
    u64 *u = (u64 *)malloc(sizeof(u64) * N * rows), *e = (u64 *)malloc(sizeof(u64) * N * rows);
    u64 *t = (u64 *)malloc(sizeof(u64) * N * rows);
    for (u64 n = 0; n < N; n++) v[n] = sample_ternary(seed, stream_id(DOM_ASYM_U, enc_id), n);
    small_poly_rows(c, v, l, 1, u);
    for (int p = 0; p < 2; p++) {
        u64 ne = stream_id(p ? DOM_ASYM_E1 : DOM_ASYM_E0, enc_id);
        for (u64 n = 0; n < N; n++) v[n] = sample_cbd(seed, ne, n);
        small_poly_rows(c, v, l, 1, e);
        for (int r = 0; r < rows; r++) { int tl = row_limb(c, l, r); u64 q = c->q[tl];
            for (u64 n = 0; n < N; n++)
                t[(u64)r * N + n] = addmod(mulmod(pk[((u64)p * K + tl) * N + n], u[(u64)r * N + n], q), e[(u64)r * N + n], q); }
        moddown_poly(c, l, t, ct + (u64)p * l * N);
    }
    for (int i = 0; i < l; i++) for (u64 n = 0; n < N; n++)
        ct[(u64)i * N + n] = addmod(ct[(u64)i * N + n], pt[(u64)i * N + n], c->q[i]);
    free(v); free(u); free(e); free(t);
}
/* relinearize a size-3 ciphertext */
void orc_relinearize(const orc_ctx *c, int l, const u64 *ct3, const u64 *rlk, u64 *out) {
    u64 N = c->N;
    orc_keyswitch(c, l, ct3 + (u64)2 * l * N, rlk, out, out + (u64)l * N);
    for (int p = 0; p < 2; p++) for (int i = 0; i < l; i++) for (u64 n = 0; n < N; n++) {
        u64 k = ((u64)p * l + i) * N + n;
        out[k] = addmod(out[k], ct3[k], c->q[i]);
    }
}
/* tensor product of two size-2 ciphertexts */
void orc_multiply(const orc_ctx *c, int l, const u64 *a, const u64 *b, u64 *out3) {
    u64 N = c->N;
    for (int i = 0; i < l; i++) { u64 q = c->q[i];
        for (u64 n = 0; n < N; n++) {
            u64 a0 = a[(u64)i * N + n], a1 = a[((u64)l + i) * N + n];
            u64 b0 = b[(u64)i * N + n], b1 = b[((u64)l + i) * N + n];
            out3[(u64)i * N + n] = mulmod(a0, b0, q);
            out3[((u64)l + i) * N + n] = addmod(mulmod(a0, b1, q), mulmod(a1, b0, q), q);
            out3[((u64)2 * l + i) * N + n] = mulmod(a1, b1, q);
        }
    }
}
/* rescale: divide-and-round by q_{l-1}, drop that limb. ct[size][l][N] -> out[size][l-1][N] */
void orc_rescale(const orc_ctx *c, int l, int size, const u64 *ct, u64 *out) {
    u64 N = c->N; int last = l - 1; u64 ql = c->q[last], half = ql >> 1;
    u64 *x = (u64 *)malloc(sizeof(u64) * N);
    for (int p = 0; p < size; p++) {
        memcpy(x, ct + ((u64)p * l + last) * N, sizeof(u64) * N);
        orc_ntt_inv(c, last, x);
        for (u64 n = 0; n < N; n++) x[n] = addmod(x[n], half, ql);
        #pragma omp parallel for schedule(static)
        for (int i = 0; i < last; i++) {
            u64 qi = c->q[i], hq = half % qi, inv = invmod(ql % qi, qi);
            u64 *o = out + ((u64)p * last + i) * N;
            for (u64 n = 0; n < N; n++) o[n] = submod(x[n] % qi, hq, qi);
            orc_ntt_fwd(c, i, o);
            const u64 *ci = ct + ((u64)p * l + i) * N;
            for (u64 n = 0; n < N; n++) o[n] = mulmod(submod(ci[n], o[n], qi), inv, qi);
        }
    }
    free(x);
}
/* ModRaise (first step of CKKS bootstrapping): a ciphertext modulo q_0 alone, ct[size][1][N], is re-read modulo
 * q_0..q_{l-1}: every coefficient is lifted to its centred representative in (-q_0/2, q_0/2] and reduced modulo each
 * q_i.  The result decrypts to m + q_0 * I(X) with small integer I.  out[size][l][N], NTT form. */
void orc_mod_raise(const orc_ctx *c, int size, const u64 *ct, int l, u64 *out) {
    u64 N = c->N, q0 = c->q[0], half = q0 >> 1;
    u64 *x = (u64 *)malloc(sizeof(u64) * N);
    for (int p = 0; p < size; p++) {
        memcpy(x, ct + (u64)p * N, sizeof(u64) * N);
        orc_ntt_inv(c, 0, x);
        #pragma omp parallel for schedule(static)
        for (int i = 0; i < l; i++) {
            u64 qi = c->q[i];
            u64 *o = out + ((u64)p * l + i) * N;
            for (u64 n = 0; n < N; n++) {
                u64 v = x[n];
                if (v > half) { u64 d = (q0 - v) % qi; o[n] = d ? qi - d : 0; }   /* negative representative v - q0 */
                else o[n] = v % qi;
            }
            orc_ntt_fwd(c, i, o);
        }
    }
    free(x);
}
/* element-wise helpers on [rows][N] blocks whose row r has limb id r (data limbs) */
void orc_add(const orc_ctx *c, int rows_per_poly, int polys, const u64 *a, const u64 *b, u64 *o) {
    #pragma omp parallel for schedule(static) collapse(2)
    for (int p = 0; p < polys; p++) for (int i = 0; i < rows_per_poly; i++) for (u64 n = 0; n < c->N; n++) {
        u64 k = ((u64)p * rows_per_poly + i) * c->N + n; o[k] = addmod(a[k], b[k], c->q[i]); }
}
void orc_sub(const orc_ctx *c, int rows_per_poly, int polys, const u64 *a, const u64 *b, u64 *o) {
    #pragma omp parallel for schedule(static) collapse(2)
    for (int p = 0; p < polys; p++) for (int i = 0; i < rows_per_poly; i++) for (u64 n = 0; n < c->N; n++) {
        u64 k = ((u64)p * rows_per_poly + i) * c->N + n; o[k] = submod(a[k], b[k], c->q[i]); }
}
void orc_multiply_plain(const orc_ctx *c, int l, int polys, const u64 *ct, const u64 *pt, u64 *o) {
    #pragma omp parallel for schedule(static) collapse(2)
    for (int p = 0; p < polys; p++) for (int i = 0; i < l; i++) for (u64 n = 0; n < c->N; n++) {
        u64 k = ((u64)p * l + i) * c->N + n; o[k] = mulmod(ct[k], pt[(u64)i * c->N + n], c->q[i]); }
}

/* ------------------------------------------------------------------ */
/* BSGS, exact mode = reference op order                               */
/*   scripts/bootstrap_generation.py:464-484                            */
/* ct_baby[G][2][l][N] (already rotated by the caller, :215-220),      */
/* pts[D][l][N], giant keys gkeys[g] for g = 1..B-1 (step g*G).        */
/* out[2][l-1][N]                                                      */
/* ------------------------------------------------------------------ */
void orc_bsgs_exact(const orc_ctx *c, int l, const u64 *ct_baby, const u64 *pts, int G, int B, int D,
                    const u32 *giant_elts, const u64 *const *gkeys, u64 *out) {
    u64 N = c->N, ctw = (u64)2 * l * N;
    u64 *inner = (u64 *)malloc(sizeof(u64) * ctw), *res = (u64 *)calloc(ctw, sizeof(u64));
    u64 *rot = (u64 *)malloc(sizeof(u64) * ctw);
    int have = 0;
    for (int g = 0; g < B; g++) {
        int any = 0;
        memset(inner, 0, sizeof(u64) * ctw);
        for (int b = 0; b < G; b++) {
            int k = g * G + b; if (k >= D) continue;
            any = 1;
            const u64 *cb = ct_baby + (u64)b * ctw, *pt = pts + (u64)k * l * N;
            #pragma omp parallel for schedule(static) collapse(2)
            for (int p = 0; p < 2; p++) for (int i = 0; i < l; i++) { u64 q = c->q[i];
                for (u64 n = 0; n < N; n++) { u64 x = ((u64)p * l + i) * N + n;
                    inner[x] = addmod(inner[x], mulmod(cb[x], pt[(u64)i * N + n], q), q); } }
        }
        if (!any) continue;
        const u64 *term = inner;
        if (g > 0) { orc_apply_galois(c, l, inner, giant_elts[g], gkeys[g], rot); term = rot; }
        if (!have) { memcpy(res, term, sizeof(u64) * ctw); have = 1; }
        else orc_add(c, l, 2, res, term, res);
    }
    orc_rescale(c, l, 2, res, out);
    free(inner); free(res); free(rot);
}

/* ------------------------------------------------------------------ */
/* BSGS, hoisted mode (this build's fast path)          */
/*  1. decompose c1 once: E_j                                          */
/*  2. Y_0 = P*(c0,c1) ; Y_b = (P*pi_b(c0) + <pi_b(E),k0_b>, <pi_b(E),k1_b>) in basis Q_l*P */
/*  3. A_g = sum_b Y_b * pt_{gG+b}   (pt on l+P limbs, value of index n = diag[n >> rshift]) */
/*  4. R = A_0 ; g>=1: t = ModDown(A_g.1); F = decompose(t);            */
/*        R.0 += pi_g(A_g.0) + <pi_g(F),k0_g> ; R.1 += <pi_g(F),k1_g>  */
/*  5. out = rescale(ModDown(R))                                       */
/* diags[D][l+P][N >> rshift]; baby_elts[b], bkeys[b] for b = 1..G-1    */
/* ------------------------------------------------------------------ */
/* shard form: diags holds only the giant groups g = g_first + k*g_stride (k = 0,1,...) in that order;
 * R[2][l+P][N] is the shard's accumulator in basis Q_l*P (the sum over all shards mod q is the full one) */
void orc_bsgs_hoisted_partial(const orc_ctx *c, int l, const u64 *ct, const u64 *diags, int rshift,
                              int G, int B, int D, int g_first, int g_stride,
                              const u32 *baby_elts, const u64 *const *bkeys,
                              const u32 *giant_elts, const u64 *const *gkeys, u64 *R) {
    u64 N = c->N; int rows = l + c->P, beta = n_digits(c, l);
    u64 polyw = (u64)rows * N, dn = N >> rshift;
    u64 *E = (u64 *)malloc(sizeof(u64) * polyw * beta);
    u64 *Y = (u64 *)malloc(sizeof(u64) * polyw * 2 * G);
    u64 *A = (u64 *)malloc(sizeof(u64) * polyw * 2);
    u64 *t = (u64 *)malloc(sizeof(u64) * N * l);
    const u64 *c0 = ct, *c1 = ct + (u64)l * N;
    memset(R, 0, sizeof(u64) * 2 * polyw);
    orc_decompose(c, l, c1, E);
    for (int b = 0; b < G; b++) {
        u64 *Yb = Y + (u64)b * 2 * polyw;
        if (b == 0) memset(Yb, 0, sizeof(u64) * 2 * polyw);
        else ks_inner(c, l, E, baby_elts[b], bkeys[b], Yb, 0);
        for (int i = 0; i < l; i++) { u64 q = c->q[i], pm = pmod(c, i);
            for (u64 n = 0; n < N; n++) {
                u64 src = b ? galois_src(c, baby_elts[b], (u32)n) : n;
                Yb[(u64)i * N + n] = addmod(Yb[(u64)i * N + n], mulmod(pm, c0[(u64)i * N + src], q), q);
                if (b == 0) Yb[polyw + (u64)i * N + n] = mulmod(pm, c1[(u64)i * N + n], q);
            } }
    }
    u64 row0 = 0;   /* first stored diagonal of the current group */
    for (int g = g_first; g < B; g += g_stride) {
        memset(A, 0, sizeof(u64) * 2 * polyw);
        int any = 0, used = 0;
        for (int b = 0; b < G; b++) {
            int k = g * G + b; if (k >= D) continue;
            any = 1; used++;
            const u64 *Yb = Y + (u64)b * 2 * polyw, *pt = diags + (row0 + (u64)b) * rows * dn;
            #pragma omp parallel for schedule(static) collapse(2)
            for (int p = 0; p < 2; p++) for (int r = 0; r < rows; r++) { u64 q = c->q[row_limb(c, l, r)];
                for (u64 n = 0; n < N; n++) { u64 x = (u64)p * polyw + (u64)r * N + n;
                    A[x] = addmod(A[x], mulmod(Yb[x], pt[(u64)r * dn + (n >> rshift)], q), q); } }
        }
        row0 += (u64)used;
        if (!any) continue;
        if (g == 0) { for (u64 x = 0; x < 2 * polyw; x++) R[x] = addmod(R[x], A[x], c->q[row_limb(c, l, (int)((x / N) % rows))]); continue; }
        moddown_poly(c, l, A + polyw, t);
        orc_decompose(c, l, t, E);
        ks_inner(c, l, E, giant_elts[g], gkeys[g], R, 1);
        for (int r = 0; r < rows; r++) { u64 q = c->q[row_limb(c, l, r)];
            for (u64 n = 0; n < N; n++) {
                u64 src = galois_src(c, giant_elts[g], (u32)n);
                R[(u64)r * N + n] = addmod(R[(u64)r * N + n], A[(u64)r * N + src], q);
            } }
    }
    free(E); free(Y); free(A); free(t);
}
/* R[2][l+P][N] -> out[2][l-1][N]: ModDown both polynomials, rescale */
void orc_bsgs_finish(const orc_ctx *c, int l, const u64 *R, u64 *out) {
    u64 N = c->N, polyw = (u64)(l + c->P) * N;
    u64 *full = (u64 *)malloc(sizeof(u64) * 2 * l * N);
    moddown_poly(c, l, R, full);
    moddown_poly(c, l, R + polyw, full + (u64)l * N);
    orc_rescale(c, l, 2, full, out);
    free(full);
}
void orc_bsgs_hoisted(const orc_ctx *c, int l, const u64 *ct, const u64 *diags, int rshift,
                      int G, int B, int D,
                      const u32 *baby_elts, const u64 *const *bkeys,
                      const u32 *giant_elts, const u64 *const *gkeys, u64 *out) {
    u64 *R = (u64 *)malloc(sizeof(u64) * 2 * (u64)(l + c->P) * c->N);
    orc_bsgs_hoisted_partial(c, l, ct, diags, rshift, G, B, D, 0, 1, baby_elts, bkeys, giant_elts, gkeys, R);
    orc_bsgs_finish(c, l, R, out);
    free(R);
}
/* x mod q per row (rows follow row_limb), for lazily summed accumulators */
void orc_reduce_rows(const orc_ctx *c, int l, int ext, int polys, u64 *x) {
    int rows = l + (ext ? c->P : 0);
    for (int p = 0; p < polys; p++) for (int r = 0; r < rows; r++) { u64 q = c->q[row_limb(c, l, r)];
        for (u64 n = 0; n < c->N; n++) x[((u64)p * rows + r) * c->N + n] %= q; }
}
