// bsgs.cu -- key switching, rotation and the two BSGS diagonal mat-vec drivers.
//
//  * bsgs_exact   : the reference's op order (scripts/bootstrap_generation.py:464-484): per giant
//                   group a plaintext-diagonal MAC over the caller's baby ciphertexts, a full
//                   rotate (permute -> decompose -> key inner product -> ModDown), sum, one rescale.
//  * bsgs_hoisted : this build's fast path (DESIGN.md "hoisted mode"): one decomposition of c1 shared
//                   by all baby steps, baby rotations kept in basis Q_l*P without ModDown, diagonal MAC in
//                   that basis, one ModDown of the c1 part per giant step, giant key switches accumulated
//                   in basis Q_l*P, one final ModDown, one rescale.
// Both are restated step for step by the oracle (orc_bsgs_exact / orc_bsgs_hoisted).
#include <cstdlib>

#include "engine.h"
#include "ops.h"

namespace eng {

struct Arena {   // bump allocator over the stream's grow-only workspace (hot path: no allocator calls)
    u64* p;
    size_t left;
    Arena(const Ctx* c, cudaStream_t s, size_t words) : p(c->workspace(s, words)), left(words) {}
    u64* get(size_t words) {
        words = (words + 31) & ~(size_t)31;   // keep 256-byte alignment
        REQUIRE(words <= left, "internal: workspace under-estimated");
        u64* r = p;
        p += words, left -= words;
        return r;
    }
};

struct Scratch {   // temporaries ordered on the stream that uses them, released on scope exit
    const Ctx* c;
    cudaStream_t s;
    std::vector<void*> ptrs;
    Scratch(const Ctx* c_, cudaStream_t s_) : c(c_), s(s_) {}
    u64* get(size_t words) {
        u64* p = c->alloc(words, s);
        ptrs.push_back(p);
        return p;
    }
    ~Scratch() {
        for (void* p : ptrs) c->free(p, s);
    }
};

// out[2][l][N] = ModDown(<decompose(cin), key>)  (+ add0[l][N] into polynomial 0)
void keyswitch(const Ctx* c, const u64* cin, int l, const u64* key, const u64* add0, u64* out, cudaStream_t s) {
    const size_t N = c->N, rows = l + c->P;
    Scratch sc(c, s);
    u64* x = sc.get(l * N);
    u64* E = sc.get(c->digits(l) * rows * N);
    u64* acc = sc.get(2 * rows * N);
    u64* tmp = sc.get(2 * l * N);
    ops::decompose_ks(c, cin, x, l, E, key, acc, 0, nullptr, 0, 0, 0, s);
    ops::moddown(c, acc, rows * N, 2, l, tmp, nullptr, out, s);
    if (add0) ops::add(c, out, add0, out, 1, l, (int)N, RowMap{l, l, c->L, 0}, 1, s);
}

// reference op order: permute both polynomials, key-switch the permuted c1, add the permuted c0
void apply_galois(const Ctx* c, const u64* ct, int l, u32 elt, const u64* key, u64* out, cudaStream_t s) {
    const size_t N = c->N;
    Scratch sc(c, s);
    u64* perm = sc.get(2 * l * N);
    ops::galois(c, ct, perm, 2 * l, elt, s);
    keyswitch(c, perm + l * N, l, key, perm, out, s);
}

// ct3 [3][l][N] -> out [2][l][N]
void relinearize(const Ctx* c, const u64* ct3, int l, const u64* rlk, u64* out, cudaStream_t s) {
    const size_t N = c->N;
    keyswitch(c, ct3 + 2 * l * N, l, rlk, nullptr, out, s);
    ops::add(c, out, ct3, out, 2, l, (int)N, RowMap{l, l, c->L, 0}, 2, s);
}

void rescale(const Ctx* c, const u64* in, int polys, int l, u64* out, cudaStream_t s) {
    const size_t N = c->N;
    Scratch sc(c, s);
    u64* last = sc.get(polys * N);
    u64* tmp = sc.get((size_t)polys * (l - 1) * N);
    ops::rescale(c, in, polys, l, last, tmp, out, s);
}

// baby[b]: [2][l][N] (b < G); pts[k]: [l][N] (k < D); gkey[g]: giant keys (g >= 1); out [2][l-1][N]
void bsgs_exact(const Ctx* c, const u64* const* baby, const u64* const* pts, int G, int B, int D, int l,
                const u32* gelt, const u64* const* gkey, u64* out, cudaStream_t s) {
    const size_t N = c->N, ctw = 2 * l * N;
    Scratch sc(c, s);
    u64* inner = sc.get(ctw);
    u64* rot = sc.get(ctw);
    u64* res = sc.get(ctw);
    bool have = false;
    for (int g = 0; g < B; g++) {
        int nb = std::min(G, D - g * G);
        if (nb <= 0) break;
        u64* dst = (g == 0) ? res : inner;
        ops::pmac_list(c, baby, pts + (size_t)g * G, nb, dst, l, s);
        if (g == 0) {
            have = true;
            continue;
        }
        apply_galois(c, inner, l, gelt[g], gkey[g], rot, s);
        ops::add(c, res, rot, res, 2, l, (int)N, RowMap{l, l, c->L, 0}, 2, s);
    }
    REQUIRE(have, "bsgs: empty diagonal set");
    rescale(c, res, 2, l, out, s);
}

// The same contraction with the D plaintexts in HOST memory ([D][pt_limbs][N], the reference's offload format,
// scripts/bootstrap_generation.py:336-358, 449): the diagonals of giant group g+1 cross PCIe into a two-slot device
// ring on a copy stream while group g is multiplied and rotated, so device memory holds 2 G plaintexts instead of D and
// the transfer overlaps the arithmetic (truly asynchronous when the host buffer is page-locked -- what
// pyPhantom.offload_plaintexts hands out -- and merely correct when it is pageable).  Same op order as bsgs_exact.
void bsgs_exact_from_host(const Ctx* c, const u64* const* baby, const u64* host_pts, int pt_limbs, int G, int B, int D,
                          int l, const u32* gelt, const u64* const* gkey, u64* out, cudaStream_t s) {
    const size_t N = c->N, ctw = 2 * l * N, ptw = (size_t)pt_limbs * N, slot_words = (size_t)G * ptw;
    Scratch sc(c, s);
    u64* inner = sc.get(ctw);
    u64* rot = sc.get(ctw);
    u64* res = sc.get(ctw);
    u64* ring = sc.get(2 * slot_words);
    cudaStream_t copy = c->aux[0];
    cudaEvent_t ready[2], freed[2], born;
    for (int i = 0; i < 2; i++) {
        CUDA_CHECK(cudaEventCreateWithFlags(&ready[i], cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&freed[i], cudaEventDisableTiming));
    }
    CUDA_CHECK(cudaEventCreateWithFlags(&born, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventRecord(born, s));            // the ring exists (stream-ordered allocation) before the copy stream touches it
    CUDA_CHECK(cudaStreamWaitEvent(copy, born, 0));
    bool have = false;
    std::vector<const u64*> pt(G);
    for (int g = 0; g < B; g++) {
        const int nb = std::min(G, D - g * G), slot = g & 1;
        if (nb <= 0) break;
        u64* dst_slot = ring + (size_t)slot * slot_words;
        if (g >= 2) CUDA_CHECK(cudaStreamWaitEvent(copy, freed[slot], 0));   // group g-2 is done with this slot
        CUDA_CHECK(cudaMemcpyAsync(dst_slot, host_pts + (size_t)g * G * ptw, sizeof(u64) * nb * ptw, cudaMemcpyHostToDevice, copy));
        CUDA_CHECK(cudaEventRecord(ready[slot], copy));
        CUDA_CHECK(cudaStreamWaitEvent(s, ready[slot], 0));
        for (int k = 0; k < nb; k++) pt[k] = dst_slot + (size_t)k * ptw;
        u64* dst = (g == 0) ? res : inner;
        ops::pmac_list(c, baby, pt.data(), nb, dst, l, s);
        CUDA_CHECK(cudaEventRecord(freed[slot], s));
        if (g == 0) {
            have = true;
            continue;
        }
        apply_galois(c, inner, l, gelt[g], gkey[g], rot, s);
        ops::add(c, res, rot, res, 2, l, (int)N, RowMap{l, l, c->L, 0}, 2, s);
    }
    for (int i = 0; i < 2; i++) cudaEventDestroy(ready[i]), cudaEventDestroy(freed[i]);
    cudaEventDestroy(born);
    REQUIRE(have, "bsgs: empty diagonal set");
    rescale(c, res, 2, l, out, s);
}

// Giant-step launch mode.  A mat-vec running alone on the engine stream takes everything in single launches over all
// giant groups (mode 1: ModUp, first transform pass, fused pass + key product; each small launch pays ~10 us of ramp-up
// and tail: 11.6 -> 10.8 ms at C3).  Mode 0 = one launch per group and stage, mode 2 = ModUp and first pass batched,
// product per group.  SPEAR_BATCH_GIANT / SPEAR_BATCH_GIANT_MULTI (mat-vecs on the auxiliary streams) override.
static int giant_mode(const Ctx* c, cudaStream_t s) {
    static const int mode_single = getenv("SPEAR_BATCH_GIANT") ? atoi(getenv("SPEAR_BATCH_GIANT")) : 1;
    static const int mode_multi = getenv("SPEAR_BATCH_GIANT_MULTI") ? atoi(getenv("SPEAR_BATCH_GIANT_MULTI")) : 1;
    return s == c->stream ? mode_single : mode_multi;
}
static bool giant_batched(const Ctx* c, int l, int nrot, cudaStream_t s) {
    const size_t pw = (size_t)(l + c->P) * c->N;
    const int beta = c->digits(l);
    return giant_mode(c, s) != 0 && nrot >= 2 && ntt_ks_fused_applies(c, l) && (size_t)nrot * beta * (l + c->P) <= 65535 &&
           (size_t)nrot * beta * pw * sizeof(u64) <= ((size_t)6 << 30);
}
static size_t phase1_words(const Ctx* c, int l, int G, int n_groups, bool local_a) {
    const size_t N = c->N, pw = (size_t)(l + c->P) * N;
    return l * N + c->digits(l) * pw + (size_t)G * 2 * pw + (local_a ? (size_t)n_groups * 2 * pw : 0) + 32 * 4;
}
static size_t phase2_words(const Ctx* c, int l, int nrot, cudaStream_t s) {
    const size_t N = c->N, pw = (size_t)(l + c->P) * N;
    const int beta = c->digits(l);
    return l * N + beta * pw + 3 * (size_t)nrot * l * N + 32 * 8 + (giant_batched(c, l, nrot, s) ? (size_t)nrot * (beta + 2) * pw + 64 : 0);
}

// Phase 1 of the hoisted mat-vec on the rows [row0, row0 + nrows) of the l + P: one decomposition of c1 shared by all
// baby steps, the baby rotations kept in basis Q_l*P, and the diagonal MAC of every giant group of the set, whose
// accumulators A_g go to `dst` (ops.h PmacDst: one local array, or the exchange windows of a rank group).
// diag [n_diags][nrows][N >> rshift]; x, E, Y: scratch ([l][N], [beta][l+P][N], [G][2][l+P][N]); tmp: [n_groups][2][l+P][N]
// when the set is walked in several baby-step chunks (G > 64), else unused.
static void bsgs_phase1(const Ctx* c, const u64* ct, int l, const u64* diag, int rshift, int G, int n_groups, int n_diags,
                        const u32* belt, const u64* const* bkey, u64* x, u64* E, u64* Y, const PmacDst& dst, u64* tmp,
                        int row0, int nrows, cudaStream_t s, int col0 = 0, int ncols = -1) {
    const size_t N = c->N, rows = l + c->P, pw = rows * N;
    const u64 *c0 = ct, *c1 = ct + l * N;
    if (ncols < 0) ncols = (int)N - col0;
    const int dcols = ncols >> rshift;   // a column-sliced set stores exactly its columns
    if (nrows == 0 || ncols == 0) return;   // more ranks than rows: this one only serves giant groups
    ops::decompose(c, c1, l, x, E, s, row0, nrows);   // only the rows this launch set serves
    ops::pscale(c, ct, Y, l, s);
    // A/B switch for the north-star item "diagonal MAC fused into the key-switch epilogue": SPEAR_PIPE_ROWS=k walks the rows
    // in k chunks, the MAC of chunk i on an auxiliary stream under the key stream of chunk i+1 -- the overlap of the HBM-bound
    // baby steps with the integer-bound MAC that a fused kernel could buy at best (profiles/r2_ns1_overlap.md)
    static const int pipe = getenv("SPEAR_PIPE_ROWS") ? atoi(getenv("SPEAR_PIPE_ROWS")) : 0;
    if (pipe > 1 && diag && s == c->stream && G > 1 && nrows >= pipe && ncols == (int)N) {
        cudaStream_t h = c->aux[0];
        std::vector<cudaEvent_t> ev(pipe + 1);
        for (auto& e : ev) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        bool ok = true;
        for (int ch = 0; ch < pipe && ok; ch++) {
            const int a = row0 + nrows * ch / pipe, b = row0 + nrows * (ch + 1) / pipe;
            ok = ops::ks_baby_fused(c, E, bkey + 1, belt + 1, G - 1, Y + 2 * pw, l, c0, s, a, b - a);
            CUDA_CHECK(cudaEventRecord(ev[ch], s));
            CUDA_CHECK(cudaStreamWaitEvent(h, ev[ch], 0));
            ops::pmac_hoisted_rows(c, Y, diag + (size_t)(a - row0) * (c->N >> rshift), dst, tmp, G, n_groups, n_diags, l, rshift, a,
                                   b - a, h, nrows);
        }
        REQUIRE(ok, "SPEAR_PIPE_ROWS: the fused baby-step kernel does not apply to this shape");
        CUDA_CHECK(cudaEventRecord(ev[pipe], h));
        CUDA_CHECK(cudaStreamWaitEvent(s, ev[pipe], 0));
        for (auto& e : ev) cudaEventDestroy(e);
        return;
    }
    if (G > 1 && !ops::ks_baby_fused(c, E, bkey + 1, belt + 1, G - 1, Y + 2 * pw, l, c0, s, row0, nrows, col0, ncols)) {
        REQUIRE(row0 == 0 && nrows == (int)rows && ncols == (int)N,
                "two-phase mat-vec: the fused baby-step kernel does not apply to this shape");
        for (int b = 1; b < G; b++)
            ops::ks_inner(c, E, bkey[b], Y + (size_t)b * 2 * pw, l, belt[b], c0, l, 1, 0, s);
    }
    if (diag) ops::pmac_hoisted_rows(c, Y, diag, dst, tmp, G, n_groups, n_diags, l, rshift, row0, nrows, s, nrows, col0, ncols, dcols);
}
// the baby steps alone (shared by several diagonal sets multiplying the same ciphertext): Y for the rows / columns served
static void bsgs_baby(const Ctx* c, const u64* ct, int l, int G, const u32* belt, const u64* const* bkey, u64* x, u64* E, u64* Y,
                      int row0, int nrows, cudaStream_t s, int col0 = 0, int ncols = -1) {
    PmacDst none = {};
    none.world = 1;
    bsgs_phase1(c, ct, l, nullptr, 0, G, 0, 0, belt, bkey, x, E, Y, none, nullptr, row0, nrows, s, col0, ncols);
}

// Phase 2: the giant steps of the n_groups accumulators A [n_groups][2][l+P][N] (destroyed: the ModDown transforms their
// special rows in place):  R = sum_k (pi_g(A_k.0) + <pi_g(F_k), k0_g>, <pi_g(F_k), k1_g>),  F_k = decompose(ModDown(A_k.1));
// a group with gelt[k] == 0 and gkey[k] == nullptr is giant group 0 (no rotation; it must come first).
// The ModDown of every A_k.1 and the INTT that starts its decomposition are batched over all groups.
static void bsgs_phase2(const Ctx* c, u64* A, int l, int n_groups, const u32* gelt, const u64* const* gkey, u64* R,
                        Arena& sc, cudaStream_t s) {
    const size_t N = c->N, rows = l + c->P, pw = rows * N;
    const int beta = c->digits(l);
    const int k0 = (n_groups > 0 && gkey[0] == nullptr) ? 1 : 0;   // group 0 (if owned) needs no rotation
    const int nrot = n_groups > k0 ? n_groups - k0 : 0;
    const int mode = giant_mode(c, s);
    const bool batch_e = giant_batched(c, l, nrot, s);
    bool have = false;
    if (k0) {
        CUDA_CHECK(cudaMemcpyAsync(R, A, sizeof(u64) * 2 * pw, cudaMemcpyDeviceToDevice, s));
        have = true;
    }
    if (nrot > 0) {
        u64* x = sc.get(l * N);
        u64* E = sc.get(beta * pw);
        u64* t_all = sc.get((size_t)nrot * l * N);      // ModDown(A_k.1), NTT form
        u64* x_all = sc.get((size_t)nrot * l * N);      // scratch: the same on its way to coefficient form
        u64* tmp = sc.get((size_t)nrot * l * N);
        (void)x;
        ops::moddown(c, A + (size_t)k0 * 2 * pw + pw, 2 * pw, nrot, l, tmp, nullptr, t_all, s);
        if (batch_e) {
            // The decomposition front end of ALL giant groups in one launch (inverse pass A + ModUp + forward pass A
            // fused: the coefficient-form digits are never written), then the fused pass B + key product
            u64* E_all = sc.get((size_t)nrot * beta * pw);
            if (!ntt_decompose_a(c, t_all, l, x_all, E_all, nrot, s)) {
                CUDA_CHECK(cudaMemcpyAsync(x_all, t_all, sizeof(u64) * nrot * l * N, cudaMemcpyDeviceToDevice, s));
                ntt_inverse(c, x_all, nrot * l, RowMap{l, l, c->L, 0}, (int)N, s);
                ops::decompose_from(c, t_all, x_all, l, E_all, s, /*transform=*/false, nrot);
                ntt_pass_a_batch(c, E_all, nrot, l, s);
            }
            u64* Rk = sc.get((size_t)nrot * 2 * pw);   // one partial result per giant group, summed below
            if (mode == 1 &&
                ntt_ks_fused_all(c, E_all, gkey + k0, gelt + k0, nrot, Rk, l, A + (size_t)k0 * 2 * pw, 2 * pw, (int)rows, s)) {
                ops::sum_groups(c, Rk, nrot, have ? R : nullptr, R, l, s);
                have = true;
            } else {
                for (int k = k0; k < n_groups; k++) {
                    u64* Ak = A + (size_t)k * 2 * pw;
                    ntt_ks_fused(c, E_all + (size_t)(k - k0) * beta * pw, gkey[k], R, l, gelt[k], Ak, (int)rows, 0, have ? 1 : 0,
                                 s, /*pass_a_done=*/true);
                    have = true;
                }
            }
        } else {
            for (int k = k0; k < n_groups; k++) {
                u64* Ak = A + (size_t)k * 2 * pw;
                const size_t o = (size_t)(k - k0) * l * N;
                ops::decompose_ks(c, t_all + o, x_all + o, l, E, gkey[k], R, gelt[k], Ak, (int)rows, 0, have ? 1 : 0, s);
                have = true;
            }
        }
    }
    if (!have) CUDA_CHECK(cudaMemsetAsync(R, 0, sizeof(u64) * 2 * pw, s));
}

// ct [2][l][N]; diag [n_diags][l+P][N >> rshift] holds the giant groups g_first + k*g_stride (k < n_groups);
// bkey[b] (1 <= b < G); gelt/gkey indexed by LOCAL group k (0 / nullptr where the global group is 0).
// R [2][l+P][N]: this shard's accumulator in basis Q_l*P (sum over shards, mod q, = the full accumulator).
void bsgs_hoisted_partial(const Ctx* c, const u64* ct, int l, const u64* diag, int rshift, int G, int n_groups,
                          int n_diags, int g_first, int g_stride, const u32* belt, const u64* const* bkey,
                          const u32* gelt, const u64* const* gkey, u64* R, cudaStream_t s) {
    const size_t N = c->N, rows = l + c->P, pw = rows * N;
    const int k0 = (g_first == 0) ? 1 : 0;
    const int nrot = n_groups > k0 ? n_groups - k0 : 0;
    (void)g_stride;
    Arena sc(c, s, phase1_words(c, l, G, n_groups, true) + phase2_words(c, l, nrot, s));
    u64* x = sc.get(l * N);
    u64* E = sc.get(c->digits(l) * pw);
    u64* Y = sc.get((size_t)G * 2 * pw);
    u64* A = sc.get((size_t)n_groups * 2 * pw);
    PmacDst dst = {};
    dst.base[0] = A, dst.world = 1;
    bsgs_phase1(c, ct, l, diag, rshift, G, n_groups, n_diags, belt, bkey, x, E, Y, dst, A, 0, (int)rows, s);
    bsgs_phase2(c, A, l, n_groups, gelt, gkey, R, sc, s);
}

// Several diagonal sets multiplying the SAME ciphertext (the chunk pairs of a D -> F projection: the reference computes
// the baby rotations once for all of them, scripts/bootstrap_generation.py:575-600): one decomposition and one set of baby
// steps, then per set the diagonal MAC and the giant steps.  Same limbs as `count` separate calls.
void bsgs_hoisted_shared(const Ctx* c, const u64* ct, int l, int G, const u32* belt, const u64* const* bkey,
                         const SharedSet* sets, int count, cudaStream_t s) {
    const size_t N = c->N, rows = l + c->P, pw = rows * N;
    int max_groups = 0, max_rot = 0;
    for (int i = 0; i < count; i++) {
        max_groups = std::max(max_groups, sets[i].n_groups);
        max_rot = std::max(max_rot, sets[i].n_groups);
    }
    Arena sc(c, s, phase1_words(c, l, G, max_groups, true) + phase2_words(c, l, max_rot, s));
    u64* x = sc.get(l * N);
    u64* E = sc.get(c->digits(l) * pw);
    u64* Y = sc.get((size_t)G * 2 * pw);
    u64* A = sc.get((size_t)max_groups * 2 * pw);
    bsgs_baby(c, ct, l, G, belt, bkey, x, E, Y, 0, (int)rows, s);
    const Arena mark = sc;
    for (int i = 0; i < count; i++) {
        PmacDst dst = {};
        dst.base[0] = A, dst.world = 1;
        ops::pmac_hoisted_rows(c, Y, sets[i].diag, dst, A, G, sets[i].n_groups, sets[i].n_diags, l, sets[i].rshift, 0, (int)rows, s);
        sc = mark;   // the giant-step scratch of the previous set is free again (stream order)
        bsgs_phase2(c, A, l, sets[i].n_groups, sets[i].gelt, sets[i].gkey, sets[i].R, sc, s);
    }
}

// ---- two-phase mat-vec over a rank group (peer.cu drives the exchange between the phases) -----------------------------
// Phase 1 on this rank's ROWS (and, in groups of more than four ranks, its half of the COLUMNS) for every giant group; the accumulators of group g land in dst.base[g % world]
// (slot (g / world)): the all-to-all that turns the row split into the giant-group split is the diagonal MAC's own
// epilogue.  diag [n_diags][nrows][N >> rshift] holds all B groups.
void bsgs_split_phase1(const Ctx* c, const u64* ct, int l, const u64* diag, int rshift, int G, int B, int n_diags, int row0,
                       int nrows, int col0, int ncols, const u32* belt, const u64* const* bkey, const PmacDst& dst,
                       cudaStream_t s) {
    const size_t N = c->N, pw = (size_t)(l + c->P) * N;
    const bool chunks = G > 64;
    Arena sc(c, s, std::max(phase1_words(c, l, G, B, chunks), phase2_words(c, l, (B + dst.world - 1) / dst.world, s)));
    u64* x = sc.get(l * N);
    u64* E = sc.get(c->digits(l) * pw);
    u64* Y = sc.get((size_t)G * 2 * pw);
    u64* tmp = chunks ? sc.get((size_t)B * 2 * pw) : nullptr;
    bsgs_phase1(c, ct, l, diag, rshift, G, B, n_diags, belt, bkey, x, E, Y, dst, tmp, row0, nrows, s, col0, ncols);
}
// The same in two steps for several sets multiplying one ciphertext: the baby steps once ...
void bsgs_split_baby(const Ctx* c, const u64* ct, int l, int G, int B, int row0, int nrows, int col0, int ncols,
                     const u32* belt, const u64* const* bkey, int world, cudaStream_t s) {
    const size_t N = c->N, pw = (size_t)(l + c->P) * N;
    Arena sc(c, s, std::max(phase1_words(c, l, G, B, G > 64), phase2_words(c, l, (B + world - 1) / world, s)));
    u64* x = sc.get(l * N);
    u64* E = sc.get(c->digits(l) * pw);
    u64* Y = sc.get((size_t)G * 2 * pw);
    bsgs_baby(c, ct, l, G, belt, bkey, x, E, Y, row0, nrows, s, col0, ncols);
}
// ... then the diagonal MAC of one set over the baby ciphertexts bsgs_split_baby left in the stream's workspace (every MAC
// of the sets comes before the first bsgs_split_phase2, which reuses that workspace)
void bsgs_split_mac(const Ctx* c, int l, const u64* diag, int rshift, int G, int B, int n_diags, int row0, int nrows, int col0,
                    int ncols, const PmacDst& dst, cudaStream_t s) {
    const size_t N = c->N, pw = (size_t)(l + c->P) * N;
    if (nrows == 0 || ncols == 0) return;
    const bool chunks = G > 64;
    Arena sc(c, s, std::max(phase1_words(c, l, G, B, chunks), phase2_words(c, l, (B + dst.world - 1) / dst.world, s)));
    sc.get(l * N);
    sc.get(c->digits(l) * pw);
    u64* Y = sc.get((size_t)G * 2 * pw);
    u64* tmp = chunks ? sc.get((size_t)B * 2 * pw) : nullptr;
    ops::pmac_hoisted_rows(c, Y, diag, dst, tmp, G, B, n_diags, l, rshift, row0, nrows, s, nrows, col0, ncols, ncols >> rshift);
}
// Phase 2 on this rank's giant groups, whose accumulators A [n_groups][2][l+P][N] every rank of the group has written.
void bsgs_split_phase2(const Ctx* c, u64* A, int l, int G, int B, int n_groups, const u32* gelt, const u64* const* gkey,
                       int world, u64* R, cudaStream_t s) {
    const int k0 = (n_groups > 0 && gkey[0] == nullptr) ? 1 : 0;
    Arena sc(c, s, std::max(phase1_words(c, l, G, B, G > 64), phase2_words(c, l, (B + world - 1) / world, s)));
    (void)k0;
    bsgs_phase2(c, A, l, n_groups, gelt, gkey, R, sc, s);
}

// R [2][l+P][N] (destroyed) -> out [2][l-1][N]: one ModDown, one rescale
void bsgs_finish(const Ctx* c, u64* R, int l, u64* out, cudaStream_t s) {
    const size_t N = c->N, pw = (l + c->P) * N;
    if (ops::finish_fused(c, R, 2, l, out, s)) return;
    Scratch sc(c, s);
    u64* tmp = sc.get(2 * l * N);
    u64* full = sc.get(2 * l * N);
    ops::moddown(c, R, pw, 2, l, tmp, nullptr, full, s);
    rescale(c, full, 2, l, out, s);
}

}  // namespace eng
