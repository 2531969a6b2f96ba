// engine.h -- host-side object model of the sm_100a CKKS engine (internal; the public surface
// is include/spear_b200.h).
#pragma once
#include <atomic>
#include <map>
#include <memory>
#include <vector>

#include "common.cuh"

struct Ctx {
    std::atomic<int> refs{1};   // the user's handle + one per live object (keys, plaintexts, ciphertexts, ...)
    int N = 0, logn = 0;
    int K = 0, P = 0, L = 0;   // K = L + P limbs, the P special primes last
    int beta = 0;              // digits at the key level: ceil(L / P)
    int device = 0;
    std::vector<u64> q;

    cudaStream_t stream = nullptr;
    cudaStream_t aux[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_main = nullptr;
    cudaEvent_t ev_aux[3] = {nullptr, nullptr, nullptr};
    cudaMemPool_t pool = nullptr;
    int sm_count = 148;
    int sbits = 0;   // bitlength(q) - 1 when every modulus has the same bit length, else 0

    // device tables
    u64 *d_q = nullptr, *d_ratio0 = nullptr, *d_ratio1 = nullptr, *d_rwide = nullptr;
    ulonglong2 *d_psi = nullptr, *d_ipsi = nullptr, *d_invn = nullptr;
    // per data limb i: P mod q_i, P^-1 mod q_i (with Shoup companions)
    ulonglong2 *d_pmod = nullptr, *d_pinv = nullptr;
    // ModUp tables, for level l (1..L) and digit j: hatinv[(l*beta + j)*P + a] (value, shoup),
    // hat[((l*beta + j)*P + a)*K + t] = (Q_j / q_a) mod q_t
    ulonglong2* d_up_hatinv = nullptr;
    ulonglong2* d_up_hatinv_n = nullptr;   // the same times N^-1: the fused decomposition kernel folds the INTT's scaling in
    u64* d_up_hat = nullptr;
    // ModDown tables: special prime k: (hatinv, shoup), half mod p_k ; data limb i: hat[k*K+i], half mod q_i
    ulonglong2* d_dn_hatinv = nullptr;
    u64 *d_dn_half = nullptr, *d_dn_hat = nullptr;
    // rescale tables for dropping limb `last`: inv[last*K + i] = q_last^-1 mod q_i (value, shoup)
    ulonglong2* d_rs_inv = nullptr;
    // decode (Garner) tables: inv[i*3+j] = q_j^-1 mod q_i, i<3
    ulonglong2* d_garner = nullptr;
    // encoder: complex roots zeta^{bitrev(i)}, slot index maps are computed in-kernel
    double2* d_zeta = nullptr;
    // embedding position p (index into the FFT array) -> slot index, bit 31 set where the position holds the conjugate
    u32* d_pos_slot = nullptr;

    ModTab modtab() const { return ModTab{d_q, d_ratio0, d_ratio1, d_rwide}; }
    NttTab ntttab() const { return NttTab{d_psi, d_ipsi, d_invn, d_q}; }
    int digits(int l) const { return (l + P - 1) / P; }
    int limbs_at(int chain_index) const { return chain_index == 0 ? K : L - (chain_index - 1); }

    // optional per-kernel-class timing (CUDA events on the launching stream), used by bench.py
    struct ProfRec {
        int cls;
        cudaEvent_t a, b;
    };
    mutable bool profiling = false;
    mutable std::vector<ProfRec> prof;

    // grow-only scratch arena of a stream (main or aux[i]) for the fused BSGS call: no allocator traffic on the hot path
    u64* workspace(cudaStream_t s, size_t words) const;
    mutable u64* ws_base[4] = {nullptr, nullptr, nullptr, nullptr};
    mutable size_t ws_cap[4] = {0, 0, 0, 0};
    // grow-only page-locked staging buffer for host -> device uploads of pageable caller memory (diagonal-set matrices):
    // rows are packed into it by the CPU and leave in ONE asynchronous copy; `staged` guards its reuse
    void* staging(size_t bytes) const;
    mutable void* stage_base = nullptr;
    mutable size_t stage_cap = 0;
    mutable cudaEvent_t staged = nullptr;
    u64* alloc(size_t n_u64, cudaStream_t s = nullptr) const;   // stream-ordered (default: the main stream)
    void free(void* p, cudaStream_t s = nullptr) const;
};

// Objects keep their context alive: destroying the context handle first (Python GC order is arbitrary)
// only drops the user's reference.
void ctx_retain(Ctx* c);
void ctx_release(Ctx* c);
struct CtxRef {
    Ctx* ctx = nullptr;
    void bind(Ctx* c) {
        ctx = c;
        ctx_retain(c);
    }
    ~CtxRef() {
        if (ctx) ctx_release(ctx);
    }
};

// plaintext (size 1) or ciphertext (size 2 or 3); rows per polynomial = l (+ P if ext)
struct Obj : CtxRef {
    int size = 0;      // polynomials
    int l = 0;         // data limbs
    bool ext = false;  // carries the P special limbs as well (basis Q_l * P)
    int n = 0;         // coefficients per row (N, or 2*D for a sub-ring diagonal)
    double scale = 1.0;
    u64* d = nullptr;

    int rows() const { return l + (ext ? ctx->P : 0); }
    size_t words() const { return (size_t)size * rows() * n; }
    u64* poly(int p) const { return d + (size_t)p * rows() * n; }
    ~Obj() {
        if (d && ctx) ctx->free(d);
    }
};

struct KSKey : CtxRef {
    u64* d = nullptr;  // [beta][2][K][N]
    ~KSKey() {
        if (d && ctx) ctx->free(d);
    }
};

struct GaloisKeys : CtxRef {
    std::map<u32, std::unique_ptr<KSKey>> keys;
};

struct SecretKey : CtxRef {
    u32 seed[8];
    u64* d = nullptr;  // [K][N] NTT form
    ~SecretKey() {
        if (d && ctx) ctx->free(d);
    }
};

struct PublicKey : CtxRef {
    u32 seed[8];
    u64* d = nullptr;  // [2][K][N]
    ~PublicKey() {
        if (d && ctx) ctx->free(d);
    }
};

// pre-encoded, pre-rotated BSGS diagonals in basis Q_l * P (hoisted path)
struct DiagSet : CtxRef {
    int D = 0, G = 0, B = 0;   // period (matrix dimension), baby steps per group, giant groups of the whole matrix
    int n_diags = 0;           // diagonals stored here: the groups g_first, g_first + g_stride, ... (< B)
    int g_first = 0, g_stride = 1;
    int l = 0;       // data limbs
    int n = 0;       // coefficients stored per row (N >> rshift)
    int rshift = 0;
    int row0 = 0, nrows = -1;  // row slice of a two-phase mat-vec: rows [row0, row0 + nrows) of the l + P (nrows < 0: all rows)
    int col0 = 0, ncols = -1;  // ... and its column slice, in STORED values (n = N >> rshift per row; ncols < 0: all)
    int stored_cols() const { return ncols >= 0 ? ncols : n; }
    double scale = 1.0;
    u64* d = nullptr;  // [n_diags][stored_rows()][n]
    int stored_rows() const { return nrows >= 0 ? nrows : l + ctx->P; }
    ~DiagSet() {
        if (d && ctx) ctx->free(d);
    }
};

// one class per kernel of the mat-vec path (bench.py picks the dominant one for its roofline line live)
enum { PROF_KS_INNER = 0, PROF_PMAC = 1, PROF_NTT_FWD_A = 2, PROF_MODUP = 3, PROF_MODDOWN = 4, PROF_RESCALE = 5,
       PROF_KS_BABY = 6, PROF_NTT_KS = 7, PROF_NTT_FWD_B = 8, PROF_NTT_INV_A = 9, PROF_NTT_INV_B = 10,
       PROF_SUM_GROUPS = 11, PROF_PEER_WAIT = 12, PROF_PEER_REDUCE = 13, PROF_CLASSES = 14 };
struct ProfScope {   // brackets the launches of one kernel class with an event pair when profiling is on
    const Ctx* c;
    cudaStream_t s;
    cudaEvent_t b = nullptr;
    ProfScope(const Ctx* c_, int cls, cudaStream_t s_) : c(c_), s(s_) {
        if (!c->profiling) return;
        cudaEvent_t a;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, s);
        c->prof.push_back({cls, a, b});
    }
    ~ProfScope() {
        if (b) cudaEventRecord(b, s);
    }
};

// ---- launchers (ntt.cu) --------------------------------------------------------------------
// rows x n in-place transforms; `n` may be a power-of-two prefix size (sub-ring) <= N
// row_lo / row_hi: only the rows [row_lo, row_hi) of every polynomial (second pass of the register-tiled path only)
void ntt_forward(const Ctx* c, u64* data, int rows, RowMap rm, int n, cudaStream_t s, int skip_digit_alpha = 0,
                 bool split30_out = false, bool pass_a_only = false, bool pass_b_only = false, int row_lo = 0,
                 int row_hi = 1 << 30);
// decomposition front end in one kernel (inverse pass A + n^-1 * hatinv + ModUp + forward pass A); false: not applicable
bool ntt_decompose_a_applies(const Ctx* c, int l);
// row_lo / row_hi: produce the rows [row_lo, row_hi) of every digit only (row_hi < 0: all l + P rows)
bool ntt_decompose_a(const Ctx* c, const u64* cin, int l, u64* x, u64* E, int count, cudaStream_t s, int row_lo = 0,
                     int row_hi = -1);
// forward transform of freshly ModUp'd digits fused with the key inner product (ntt.cu); false: not applicable.
// pass_a_done: the first pass already ran (ntt_pass_a_batch over several decompositions at once)
bool ntt_ks_fused_applies(const Ctx* c, int l);
void ntt_pass_a_batch(const Ctx* c, u64* E, int groups, int l, cudaStream_t s);
bool ntt_ks_fused(const Ctx* c, u64* E, const u64* key, u64* out, int l, u32 elt, const u64* addp, int add_rows,
                  int add_pscale, int accumulate, cudaStream_t s, bool pass_a_done = false);
// the fused pass + key product of `groups` decompositions (pass A done) in one launch, one partial result per group
bool ntt_ks_fused_all(const Ctx* c, const u64* E, const u64* const* keys, const u32* elts, int groups, u64* out, int l,
                      const u64* addp, size_t add_stride, int add_rows, cudaStream_t s);
void ntt_inverse(const Ctx* c, u64* data, int rows, RowMap rm, int n, cudaStream_t s, bool pass_b_only = false);

// one diagonal set of a shared-baby-step call (bsgs.cu bsgs_hoisted_shared)
struct SharedSet {
    const u64* diag;
    int rshift, n_groups, n_diags;
    const u32* gelt;
    const u64* const* gkey;
    u64* R;   // [2][l+P][N] accumulator of this set
};

// ---- two-phase mat-vec over a rank group: exchange hooks (peer.cu), phases (bsgs.cu) -----------------------------
// Phase-1 share of `rank`: up to four ranks split the rows of the RNS basis; larger even groups form world/2 row groups
// of two ranks that take one half of the columns each (27 rows over 8 ranks: 3.5 row-equivalents instead of 4).
struct SplitShare {
    int row0, nrows, col0, ncols;   // columns in coefficients of the full ring
};
inline SplitShare split_share(int rank, int world, int rows, int N) {
    const int halves = (world >= 5 && world % 2 == 0 && N % 256 == 0) ? 2 : 1, groups = world / halves;
    const int rg = rank / halves, h = rank % halves;
    SplitShare s;
    s.row0 = rg * rows / groups, s.nrows = (rg + 1) * rows / groups - s.row0;
    s.col0 = h * (N / halves), s.ncols = N / halves;
    return s;
}
struct PmacDst;
struct spear_peer_window;
namespace peer {
struct SplitView {
    int rank, world;
    u64* base[8];   // the slot of every rank of the group (base[rank]: this rank's own, the others peer-mapped)
};
void window_geometry(const spear_peer_window* win, int* rank, int* world);
// waits (on s) until every peer has consumed the slot's previous contents, opens a new epoch
SplitView split_begin(const Ctx* c, spear_peer_window* win, int slot, size_t need_words, cudaStream_t s);
// posts "my phase-1 stores are out" to every peer and waits for theirs (split_post / split_wait: the two halves, for
// callers that queue more work between them)
void split_exchange(spear_peer_window* win, int slot, cudaStream_t s);
void split_post(spear_peer_window* win, int slot, cudaStream_t s);
void split_wait(spear_peer_window* win, int slot, cudaStream_t s);
// posts "slot consumed" to every peer; R (words) is overwritten with all-ones words if a peer never arrived
void split_release(spear_peer_window* win, int slot, u64* R, size_t words, cudaStream_t s);
}  // namespace peer

// ---- stream ids shared with the oracle ------------------------------------------------------
enum { DOM_SK = 1, DOM_PK_A = 2, DOM_PK_E = 3, DOM_KSK_A = 4, DOM_KSK_E = 5,
       DOM_ENC_A = 6, DOM_ENC_E = 7, DOM_ASYM_U = 8, DOM_ASYM_E0 = 9, DOM_ASYM_E1 = 10, DOM_PK_SEED = 11 };
static inline u64 stream_id(int dom, u64 id) { return ((u64)dom << 56) | (id & 0x00FFFFFFFFFFFFFFull); }
