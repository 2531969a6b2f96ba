// ops.h -- launchers of the RNS kernels in ops.cu / sampler.cu / encoder.cu (internal).
#pragma once
#include <cuda.h>   // CUtensorMap

#include "engine.h"

// key-switch inner product:  out[p][r][n] (+)= sum_j E[j][r][src(n)] * key[j][p][limb(r)][n]  (+ addp[r][src(n)] (* P) into p = 0)
struct KsArgs {
    const u64* E;      // [beta][rows][N]
    const u64* key;    // [beta_key][2][K][N]
    u64* out;          // [2][rows][N]
    const u64* addp;   // optional polynomial added (after the same permutation) to out poly 0
    int add_rows;      // rows of addp (l: data limbs only; rows: extended)
    int add_pscale;    // multiply addp by P mod q first
    int accumulate;    // out += instead of out =
    int beta, l, rows, N, logn, L, K;
    u32 elt;           // 0: identity
    int row0 = 0;      // first row served by a row-restricted launch (k_ks_baby_fused: grid.z rows from row0)
    int tile0 = 0, tile1 = 1 << 30;   // column range of such a launch, in tiles of 128 coefficients (k_ks_baby_fused)
};

// Destination of the diagonal MAC's giant-group accumulators.  Group g goes to  base[g % world] + (g / world) * 2 * pw :
// world = 1 is one local array [B][2][l+P][N]; world > 1 are the exchange windows of the ranks of a two-phase mat-vec
// (peer.cu), base[r] possibly a peer-mapped pointer -- the kernel's epilogue stores ARE the all-to-all over NVLink.
struct PmacDst {
    u64* base[8];
    int world;
};


namespace ops {
void add(const Ctx* c, const u64* a, const u64* b, u64* out, int polys, int rows, int n, RowMap rm, int b_polys, cudaStream_t s);
void sub(const Ctx* c, const u64* a, const u64* b, u64* out, int polys, int rows, int n, RowMap rm, int b_polys, cudaStream_t s);
void neg(const Ctx* c, const u64* a, u64* out, int polys, int rows, int n, RowMap rm, cudaStream_t s);
void mul(const Ctx* c, const u64* a, const u64* b, u64* out, int polys, int rows, int n, RowMap rm, int b_polys, cudaStream_t s);
void reduce_inplace(const Ctx* c, u64* x, int polys, int rows, RowMap rm, cudaStream_t s);
// out = first (optional) + sum of nparts accumulators [2][l+P][N], mod q
void sum_groups(const Ctx* c, const u64* parts, int nparts, const u64* first, u64* out, int l, cudaStream_t s);
void tensor(const Ctx* c, const u64* a, const u64* b, u64* out, int l, cudaStream_t s);
void galois(const Ctx* c, const u64* in, u64* out, int rows, u32 elt, cudaStream_t s);
// row0 / nrows: only the rows [row0, row0 + nrows) of every digit are produced (nrows < 0: all)
void decompose(const Ctx* c, const u64* cin, int l, u64* x, u64* E, cudaStream_t s, int row0 = 0, int nrows = -1);
// count > 1: `count` polynomials back to back (cin, x: [count][l][N]) -> E [count][beta][l+P][N], ModUp only (transform = false)
void decompose_from(const Ctx* c, const u64* cin, const u64* x, int l, u64* E, cudaStream_t s, bool transform = true,
                    int count = 1);
// x: scratch [l][N] (receives the coefficient form of cin on the way)
void decompose_ks(const Ctx* c, const u64* cin, u64* x, int l, u64* E, const u64* key, u64* out, u32 elt,
                  const u64* addp, int add_rows, int add_pscale, int accumulate, cudaStream_t s);
void ks_inner(const Ctx* c, const u64* E, const u64* key, u64* out, int l, u32 elt, const u64* addp, int add_rows,
              int add_pscale, int accumulate, cudaStream_t s);
// row0 / nrows: rows (of the l + P) this launch serves -- all of them by default
// col0 / ncols: columns (coefficients) served, multiples of 128 -- all of them by default
bool ks_baby_fused(const Ctx* c, const u64* E, const u64* const* keys, const u32* elts, int nb, u64* out, int l,
                   const u64* c0, cudaStream_t s, int row0 = 0, int nrows = -1, int col0 = 0, int ncols = -1);
void encode_key_map(const Ctx* c, const u64* key, int box_n, int beta, CUtensorMap* out);
void pscale(const Ctx* c, const u64* x, u64* y, int l, cudaStream_t s);
void moddown(const Ctx* c, u64* in, size_t in_pstride, int polys, int l, u64* tmp, const u64* add, u64* out, cudaStream_t s);
// R [polys][l+P][N] (NTT form, destroyed) -> out [polys][l-1][N] = rescale(ModDown(R)) in one conversion kernel between an
// inverse and a forward transform (bit-identical to moddown + rescale); false: not applicable, nothing done
bool finish_fused(const Ctx* c, u64* R, int polys, int l, u64* out, cudaStream_t s);
void mod_raise(const Ctx* c, const u64* in, int polys, int l, u64* x, u64* out, cudaStream_t s);
void rescale(const Ctx* c, const u64* in, int polys, int l, u64* last, u64* tmp, u64* out, cudaStream_t s);
void pmac_list(const Ctx* c, const u64* const* baby, const u64* const* pt, int nb, u64* out, int l, cudaStream_t s);
void split30_inplace(const Ctx* c, u64* x, size_t n, bool unsplit, cudaStream_t s);
void pmac_hoisted(const Ctx* c, const u64* Y, const u64* diag, u64* A, int G, int B, int D, int l, int rshift, cudaStream_t s);
// the same for the rows [row0, row0 + nrows) only -- diag points at the first of them, inside a set that stores diag_rows
// rows per diagonal ([D][diag_rows][diag_cols]; default: exactly the rows served, all N >> rshift columns) -- and for the
// columns [col0, col0 + ncols) only (multiples of 64; diag then points at the value of column col0) -- with the
// accumulators scattered to `dst`; tmp: local scratch [B][2][l+P][N] for sets walked in several baby-step chunks
void pmac_hoisted_rows(const Ctx* c, const u64* Y, const u64* diag, const PmacDst& dst, u64* tmp, int G, int B, int D, int l,
                       int rshift, int row0, int nrows, cudaStream_t s, int diag_rows = -1, int col0 = 0, int ncols = -1,
                       int diag_cols = -1);
}  // namespace ops

namespace sampler {
// rows: `nrows` rows whose limb ids follow rm; uniform residues, NTT domain by definition
void uniform(const Ctx* c, const u32* seed, u64 nonce, u64* out, int nrows, RowMap rm, cudaStream_t s);
// small polynomial (ternary or centred binomial) written to every row in coefficient form (caller runs the NTT)
void ternary(const Ctx* c, const u32* seed, u64 nonce, u64* out, int nrows, RowMap rm, cudaStream_t s);
void cbd(const Ctx* c, const u32* seed, u64 nonce, u64* out, int nrows, RowMap rm, cudaStream_t s);
// k0 = e - a*s (+ pmod * snew on the rows of digit j), one digit of a switching key; all [K][N]
void ksk_combine(const Ctx* c, const u64* a, const u64* e, const u64* sk, const u64* snew, int digit, u64* k0, cudaStream_t s);
// c0 = m + e - a*s on l data rows
void enc_combine(const Ctx* c, const u64* a, const u64* e, const u64* sk, const u64* m, u64* c0, int l, cudaStream_t s);
// pt = c0 + c1 s (+ c2 s^2)
void dec_combine(const Ctx* c, const u64* ct, int size, int l, const u64* sk, u64* pt, cudaStream_t s);
// out[p][r] = pk[p][limb(r)] * u[r] + e_p[r] on l+P rows
void asym_combine(const Ctx* c, const u64* pk, const u64* u, const u64* e0, const u64* e1, u64* out, int l, cudaStream_t s);
}  // namespace sampler

namespace encoder {
// values: host-resident doubles already on device as (re, im) pairs: [count][n/2]; out: [count][rows][n]
void encode(const Ctx* c, const double2* vals, int count, int n, double scale, int l, bool ext, u64* out, cudaStream_t s);
// the same in ONE kernel for rings that fit shared memory (sub-ring diagonals); overflow: sticky device flag
bool encode_ring_fused(const Ctx* c, const double2* vals, int count, int n, double scale, int l, bool ext, u64* out,
                       bool split30_out, int* overflow, cudaStream_t s);
// pt [l][N] -> vals [N/2] (re, im)
void decode(const Ctx* c, const u64* pt, int l, double scale, double2* vals, cudaStream_t s);
}  // namespace encoder

namespace client {
// the client legs of a projection round trip in three launches each (client.cu); false: use the staged path
bool fused_applies(const Ctx* c);
void encode_encrypt(const Ctx* c, const double2* vals, int nvals, bool replicate, double scale, int l, const u32* seed,
                    u64 enc_id, const u64* sk, u64* ct, double2* W, cudaStream_t s);
void decrypt_decode(const Ctx* c, const u64* ct, int size, int l, double scale, const u64* sk, double2* vals_out, int want,
                    u64* x, double2* W, cudaStream_t s);
}  // namespace client
