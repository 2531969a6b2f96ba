// api.cu -- extern "C" surface declared in include/spear_b200.h.
// Thin: argument checks, object allocation, and calls into the engine (bsgs.cu / ops.cu / ...).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>

#include "../../include/spear_b200.h"
#include "engine.h"
#include "ops.h"

unsigned long long g_spear_launches = 0;

int create_coeff_modulus(u64 N, const int* bits, int n, u64* out);
Ctx* ctx_create(u64 N, const u64* moduli, int K, int P, int device);
namespace eng {
void keyswitch(const Ctx* c, const u64* cin, int l, const u64* key, const u64* add0, u64* out, cudaStream_t s);
void apply_galois(const Ctx* c, const u64* ct, int l, u32 elt, const u64* key, u64* out, cudaStream_t s);
void relinearize(const Ctx* c, const u64* ct3, int l, const u64* rlk, u64* out, cudaStream_t s);
void rescale(const Ctx* c, const u64* in, int polys, int l, u64* out, cudaStream_t s);
void bsgs_exact(const Ctx* c, const u64* const* baby, const u64* const* pts, int G, int B, int D, int l,
                const u32* gelt, const u64* const* gkey, u64* out, cudaStream_t s);
void bsgs_exact_from_host(const Ctx* c, const u64* const* baby, const u64* host_pts, int pt_limbs, int G, int B, int D,
                          int l, const u32* gelt, const u64* const* gkey, u64* out, cudaStream_t s);
void bsgs_hoisted_partial(const Ctx* c, const u64* ct, int l, const u64* diag, int rshift, int G, int n_groups,
                          int n_diags, int g_first, int g_stride, const u32* belt, const u64* const* bkey,
                          const u32* gelt, const u64* const* gkey, u64* R, cudaStream_t s);
void bsgs_finish(const Ctx* c, u64* R, int l, u64* out, cudaStream_t s);
void bsgs_split_phase1(const Ctx* c, const u64* ct, int l, const u64* diag, int rshift, int G, int B, int n_diags, int row0,
                       int nrows, int col0, int ncols, const u32* belt, const u64* const* bkey, const PmacDst& dst,
                       cudaStream_t s);
void bsgs_split_phase2(const Ctx* c, u64* A, int l, int G, int B, int n_groups, const u32* gelt, const u64* const* gkey,
                       int world, u64* R, cudaStream_t s);
void bsgs_hoisted_shared(const Ctx* c, const u64* ct, int l, int G, const u32* belt, const u64* const* bkey,
                         const SharedSet* sets, int count, cudaStream_t s);
void bsgs_split_baby(const Ctx* c, const u64* ct, int l, int G, int B, int row0, int nrows, int col0, int ncols,
                     const u32* belt, const u64* const* bkey, int world, cudaStream_t s);
void bsgs_split_mac(const Ctx* c, int l, const u64* diag, int rshift, int G, int B, int n_diags, int row0, int nrows, int col0,
                    int ncols, const PmacDst& dst, cudaStream_t s);
}  // namespace eng

namespace {

thread_local std::string g_err;

#define API_BEGIN try {
#define API_END                                  \
    }                                            \
    catch (const spear_error& e) {               \
        g_err = e.msg;                           \
        return e.code ? e.code : SPEAR_ERR_INVALID; \
    }                                            \
    catch (const std::exception& e) {            \
        g_err = e.what();                        \
        return SPEAR_ERR_INVALID;                \
    }                                            \
    return SPEAR_OK;

inline Ctx* C_(spear_context* c) { return reinterpret_cast<Ctx*>(c); }
inline const Obj* O_(const spear_obj* o) { return reinterpret_cast<const Obj*>(o); }
inline spear_obj* H_(Obj* o) { return reinterpret_cast<spear_obj*>(o); }

Obj* new_obj(Ctx* c, int size, int l, bool ext, int n, double scale, cudaStream_t s = nullptr) {
    std::unique_ptr<Obj> o(new Obj);
    o->bind(c), o->size = size, o->l = l, o->ext = ext, o->n = n, o->scale = scale;
    o->d = c->alloc(o->words(), s);   // ordered on the stream that first writes it
    return o.release();
}
void use(Ctx* c) { CUDA_CHECK(cudaSetDevice(c->device)); }

// ChaCha20 block function on the host (same stream layout as sampler.cu / the oracle: key = seed, nonce = stream id,
// 64-bit block counter).  Only used to derive seeds from seeds, never on a data path.
void host_chacha_block(const u32 key[8], u64 nonce, u64 counter, u32 out[16]) {
    auto rotl = [](u32 v, int n) { return (v << n) | (v >> (32 - n)); };
    u32 in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    for (int i = 0; i < 8; i++) in[4 + i] = key[i];
    in[12] = (u32)counter, in[13] = (u32)(counter >> 32), in[14] = (u32)nonce, in[15] = (u32)(nonce >> 32);
    u32 x[16];
    memcpy(x, in, sizeof x);
    auto quarter = [&](int a, int b, int c, int d) {
        x[a] += x[b], x[d] = rotl(x[d] ^ x[a], 16);
        x[c] += x[d], x[b] = rotl(x[b] ^ x[c], 12);
        x[a] += x[b], x[d] = rotl(x[d] ^ x[a], 8);
        x[c] += x[d], x[b] = rotl(x[b] ^ x[c], 7);
    };
    for (int round = 0; round < 10; round++) {
        quarter(0, 4, 8, 12), quarter(1, 5, 9, 13), quarter(2, 6, 10, 14), quarter(3, 7, 11, 15);
        quarter(0, 5, 10, 15), quarter(1, 6, 11, 12), quarter(2, 7, 8, 13), quarter(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
}

void check_ct(const Obj* o, const char* what) {
    REQUIRE(o && o->size >= 2 && !o->ext && o->n == o->ctx->N, "%s: expected a ciphertext", what);
}
void check_pt(const Obj* o, const char* what) {
    REQUIRE(o && o->size == 1 && o->n == o->ctx->N, "%s: expected a plaintext", what);
}
RowMap data_rows(const Ctx* c, int l) { return RowMap{l, l, c->L, 0}; }

u64 elt_from_step(int step, u64 N) {
    u64 m = 2 * N, half = N / 2;
    if (step == 0) return m - 1;
    u64 e = step > 0 ? (u64)step % half : (half - ((u64)(-(long long)step) % half)) % half;
    u64 r = 1, g = 5;
    for (; e; e >>= 1, g = g * g & (m - 1))
        if (e & 1) r = r * g & (m - 1);
    return r;
}

// one switching key from `snew` (the key being switched away from), streams tagged by `tag`
KSKey* gen_switch_key(Ctx* c, const SecretKey* sk, u64 tag, const u64* snew) {
    const size_t N = c->N, K = c->K;
    std::unique_ptr<KSKey> key(new KSKey);
    key->bind(c);
    key->d = c->alloc((size_t)c->beta * 2 * K * N);
    u64* e = c->alloc(K * N);
    RowMap all{c->K, c->L, c->L, 0};
    for (int j = 0; j < c->beta; j++) {
        u64 id = (tag << 8) | (u64)j;
        u64* k0 = key->d + ((size_t)j * 2 + 0) * K * N;
        u64* k1 = key->d + ((size_t)j * 2 + 1) * K * N;
        sampler::uniform(c, sk->seed, stream_id(DOM_KSK_A, id), k1, c->K, all, c->stream);
        sampler::cbd(c, sk->seed, stream_id(DOM_KSK_E, id), e, c->K, all, c->stream);
        ntt_forward(c, e, c->K, all, c->N, c->stream);
        sampler::ksk_combine(c, k1, e, sk->d, snew, j, k0, c->stream);
    }
    c->free(e);
    // resident keys use the split-30 storage form consumed by the inner-product kernels (common.cuh)
    ops::split30_inplace(c, key->d, (size_t)c->beta * 2 * K * N, false, c->stream);
    return key.release();
}

KSKey* gen_galois_key(Ctx* c, const SecretKey* sk, u32 elt) {
    REQUIRE((elt & 1) && elt < 2u * c->N, "invalid Galois element %u", elt);
    u64* sn = c->alloc((size_t)c->K * c->N);
    ops::galois(c, sk->d, sn, c->K, elt, c->stream);
    KSKey* k = gen_switch_key(c, sk, elt, sn);
    c->free(sn);
    return k;
}

const KSKey* find_key(const GaloisKeys* gk, u32 elt) {
    auto it = gk->keys.find(elt);
    if (it == gk->keys.end()) spear_throw(SPEAR_ERR_NOKEY, "no Galois key for element %u", elt);
    return it->second.get();
}

// vals_full[v][j] = vals[v][j % D]  (replicate a period-D vector over all slots, as reference :371-378)
__global__ void k_tile(const double2* __restrict__ in, double2* __restrict__ out, int count, int D, int slots) {
    size_t total = (size_t)count * slots;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x)
        out[e] = in[(e / slots) * D + (e % slots) % D];
}

}  // namespace

void spear_set_last_error(const char* msg) { g_err = msg; }   // for the other translation units with entry points

extern "C" {

const char* spear_last_error(void) { return g_err.c_str(); }
const char* spear_version(void) { return "spear-b200 0.1 (sm_100a)"; }
uint64_t spear_launch_count(void) { return g_spear_launches; }

int spear_create_coeff_modulus(uint64_t N, const int* bits, int count, uint64_t* out) {
    API_BEGIN
    REQUIRE(N >= 8 && (N & (N - 1)) == 0 && count > 0, "create_coeff_modulus: bad arguments");
    create_coeff_modulus(N, bits, count, out);
    API_END
}
uint64_t spear_get_elt_from_step(int step, uint64_t N) { return elt_from_step(step, N); }

int spear_context_create(uint64_t N, const uint64_t* moduli, int count, int special, int device, spear_context** out) {
    API_BEGIN
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        spear_throw(SPEAR_ERR_CUDA, "CUDA device required (no CPU fallback): %s", cudaGetErrorString(e));
    REQUIRE(device >= 0 && device < ndev, "device %d out of range", device);
    *out = reinterpret_cast<spear_context*>(ctx_create(N, moduli, count, special, device));
    API_END
}
void spear_context_destroy(spear_context* ctx) { ctx_release(C_(ctx)); }
int spear_context_sync(spear_context* ctx) {
    API_BEGIN
    use(C_(ctx));
    CUDA_CHECK(cudaStreamSynchronize(C_(ctx)->stream));
    API_END
}
void* spear_context_stream(spear_context* ctx) { return (void*)C_(ctx)->stream; }

static thread_local cudaEvent_t t_ev0 = nullptr, t_ev1 = nullptr;
int spear_timer_start(spear_context* ctx) {
    API_BEGIN
    use(C_(ctx));
    if (!t_ev0) {
        CUDA_CHECK(cudaEventCreate(&t_ev0));
        CUDA_CHECK(cudaEventCreate(&t_ev1));
    }
    CUDA_CHECK(cudaEventRecord(t_ev0, C_(ctx)->stream));
    API_END
}
int spear_timer_stop(spear_context* ctx, float* ms) {
    API_BEGIN
    use(C_(ctx));
    REQUIRE(t_ev0, "timer not started");
    CUDA_CHECK(cudaEventRecord(t_ev1, C_(ctx)->stream));
    CUDA_CHECK(cudaEventSynchronize(t_ev1));
    CUDA_CHECK(cudaEventElapsedTime(ms, t_ev0, t_ev1));
    API_END
}
int spear_pinned_alloc(size_t bytes, void** out) {
    API_BEGIN
    CUDA_CHECK(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    API_END
}
void spear_pinned_free(void* p) { cudaFreeHost(p); }
int spear_mem_info(spear_context* ctx, uint64_t* used, uint64_t* reserved) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    unsigned long long u = 0, r = 0;
    CUDA_CHECK(cudaMemPoolGetAttribute(c->pool, cudaMemPoolAttrUsedMemCurrent, &u));
    CUDA_CHECK(cudaMemPoolGetAttribute(c->pool, cudaMemPoolAttrReservedMemCurrent, &r));
    *used = u, *reserved = r;
    API_END
}

int spear_mem_reserve(spear_context* ctx, uint64_t bytes) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    // one allocation of `bytes` handed straight back: the pool keeps the memory (release threshold = infinity) and later
    // requests are carved out of it instead of growing the pool by an OS-level mapping in the middle of a computation
    // (observed: 0.5 - 1.3 s stalls at random blocks of a fully encrypted run that creates 2.6 GB diagonal sets on the fly)
    void* p = nullptr;
    CUDA_CHECK(cudaMallocAsync(&p, bytes, c->stream));
    CUDA_CHECK(cudaFreeAsync(p, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    API_END
}

int spear_profile_enable(spear_context* ctx, int on) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    for (auto& r : c->prof) cudaEventDestroy(r.a), cudaEventDestroy(r.b);
    c->prof.clear();
    c->profiling = on != 0;
    API_END
}
int spear_profile_read(spear_context* ctx, double* ms, uint64_t* launches, int classes) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    REQUIRE(classes >= PROF_CLASSES, "profile_read: need room for %d classes", PROF_CLASSES);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < classes; i++) ms[i] = 0, launches[i] = 0;
    for (auto& r : c->prof) {
        float t = 0;
        CUDA_CHECK(cudaEventElapsedTime(&t, r.a, r.b));
        ms[r.cls] += t, launches[r.cls]++;
        cudaEventDestroy(r.a), cudaEventDestroy(r.b);
    }
    c->prof.clear();
    API_END
}

// ---- keys -------------------------------------------------------------------------------------
int spear_secret_key_create(spear_context* ctx, const uint8_t seed[32], spear_secret_key** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    std::unique_ptr<SecretKey> sk(new SecretKey);
    sk->bind(c);
    memcpy(sk->seed, seed, 32);
    sk->d = c->alloc((size_t)c->K * c->N);
    RowMap all{c->K, c->L, c->L, 0};
    sampler::ternary(c, sk->seed, stream_id(DOM_SK, 0), sk->d, c->K, all, c->stream);
    ntt_forward(c, sk->d, c->K, all, c->N, c->stream);
    *out = reinterpret_cast<spear_secret_key*>(sk.release());
    API_END
}
void spear_secret_key_destroy(spear_secret_key* sk) { delete reinterpret_cast<SecretKey*>(sk); }

int spear_gen_public_key(spear_context* ctx, const spear_secret_key* sk_, spear_public_key** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const SecretKey* sk = reinterpret_cast<const SecretKey*>(sk_);
    const size_t KN = (size_t)c->K * c->N;
    std::unique_ptr<PublicKey> pk(new PublicKey);
    pk->bind(c);
    // The public key carries its OWN seed for the encryption randomness (u, e0, e1): the first 32 bytes of the secret
    // seed's ChaCha stream DOM_PK_SEED.  One-way: whoever holds the public key cannot get back to the secret seed
    // (and through it to the secret key, which is ternary(seed, DOM_SK)).
    {
        u32 blk[16];
        host_chacha_block(sk->seed, stream_id(DOM_PK_SEED, 0), 0, blk);
        memcpy(pk->seed, blk, 32);
    }
    pk->d = c->alloc(2 * KN);
    u64* e = c->alloc(KN);
    RowMap all{c->K, c->L, c->L, 0};
    sampler::uniform(c, sk->seed, stream_id(DOM_PK_A, 0), pk->d + KN, c->K, all, c->stream);
    sampler::cbd(c, sk->seed, stream_id(DOM_PK_E, 0), e, c->K, all, c->stream);
    ntt_forward(c, e, c->K, all, c->N, c->stream);
    // pk0 = e - a*s  (same combine as a switching key digit that owns no limb)
    sampler::ksk_combine(c, pk->d + KN, e, sk->d, sk->d, -1, pk->d, c->stream);
    c->free(e);
    *out = reinterpret_cast<spear_public_key*>(pk.release());
    API_END
}
void spear_public_key_destroy(spear_public_key* pk) { delete reinterpret_cast<PublicKey*>(pk); }

int spear_gen_relin_key(spear_context* ctx, const spear_secret_key* sk_, spear_kswitch_key** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const SecretKey* sk = reinterpret_cast<const SecretKey*>(sk_);
    u64* s2 = c->alloc((size_t)c->K * c->N);
    RowMap all{c->K, c->L, c->L, 0};
    ops::mul(c, sk->d, sk->d, s2, 1, c->K, c->N, all, 1, c->stream);
    KSKey* k = gen_switch_key(c, sk, 0, s2);
    c->free(s2);
    *out = reinterpret_cast<spear_kswitch_key*>(k);
    API_END
}
void spear_kswitch_key_destroy(spear_kswitch_key* k) { delete reinterpret_cast<KSKey*>(k); }

int spear_galois_keys_add(spear_context* ctx, const spear_secret_key* sk_, spear_galois_keys* gk_, const uint32_t* elts,
                          int count) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    GaloisKeys* gk = reinterpret_cast<GaloisKeys*>(gk_);
    const SecretKey* sk = reinterpret_cast<const SecretKey*>(sk_);
    for (int i = 0; i < count; i++)
        if (!gk->keys.count(elts[i])) gk->keys[elts[i]].reset(gen_galois_key(c, sk, elts[i]));
    API_END
}
int spear_gen_galois_keys(spear_context* ctx, const spear_secret_key* sk, const uint32_t* elts, int count,
                          spear_galois_keys** out) {
    GaloisKeys* gk = new GaloisKeys;
    gk->bind(C_(ctx));
    int rc = spear_galois_keys_add(ctx, sk, reinterpret_cast<spear_galois_keys*>(gk), elts, count);
    if (rc) {
        delete gk;
        return rc;
    }
    *out = reinterpret_cast<spear_galois_keys*>(gk);
    return SPEAR_OK;
}
int spear_galois_keys_has(const spear_galois_keys* gk, uint32_t elt) {
    return (int)reinterpret_cast<const GaloisKeys*>(gk)->keys.count(elt);
}
void spear_galois_keys_destroy(spear_galois_keys* gk) { delete reinterpret_cast<GaloisKeys*>(gk); }

// ---- objects ------------------------------------------------------------------------------------
void spear_obj_destroy(spear_obj* o) { delete reinterpret_cast<Obj*>(o); }
int spear_obj_info(const spear_obj* o_, int* size, int* limbs, int* ext, int* ring_n, double* scale, int* chain_index) {
    API_BEGIN
    const Obj* o = O_(o_);
    REQUIRE(o, "null object");
    if (size) *size = o->size;
    if (limbs) *limbs = o->l;
    if (ext) *ext = o->ext;
    if (ring_n) *ring_n = o->n;
    if (scale) *scale = o->scale;
    if (chain_index) *chain_index = o->ctx->L - o->l + 1;
    API_END
}
int spear_obj_set_scale(spear_obj* o, double scale) {
    API_BEGIN
    REQUIRE(o && scale > 0, "set_scale: bad arguments");
    reinterpret_cast<Obj*>(o)->scale = scale;
    API_END
}
static int export_words(const Ctx* c, const u64* d, size_t have, uint64_t* host, size_t words) {
    API_BEGIN
    REQUIRE(words == have, "export: buffer holds %zu words, object has %zu", words, have);
    CUDA_CHECK(cudaSetDevice(c->device));
    CUDA_CHECK(cudaMemcpyAsync(host, d, sizeof(u64) * words, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    API_END
}
int spear_obj_export(const spear_obj* o, uint64_t* host, size_t words) {
    return export_words(O_(o)->ctx, O_(o)->d, O_(o)->words(), host, words);
}
int spear_obj_import(spear_context* ctx, const uint64_t* host, int size, int limbs, int ext, int ring_n, double scale,
                     spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    REQUIRE(size >= 1 && size <= 3 && limbs >= 1 && limbs <= c->L && ring_n >= 2 && ring_n <= c->N,
            "import: bad shape");
    std::unique_ptr<Obj> o(new_obj(c, size, limbs, ext != 0, ring_n, scale));
    CUDA_CHECK(cudaMemcpyAsync(o->d, host, sizeof(u64) * o->words(), cudaMemcpyHostToDevice, c->stream));
    *out = H_(o.release());
    API_END
}
int spear_secret_key_export(const spear_secret_key* sk_, uint64_t* host, size_t words) {
    const SecretKey* sk = reinterpret_cast<const SecretKey*>(sk_);
    return export_words(sk->ctx, sk->d, (size_t)sk->ctx->K * sk->ctx->N, host, words);
}
static int export_key(const KSKey* k, uint64_t* host, size_t words) {
    API_BEGIN
    Ctx* c = k->ctx;
    use(c);
    const size_t have = (size_t)c->beta * 2 * c->K * c->N;
    REQUIRE(words == have, "export: buffer holds %zu words, key has %zu", words, have);
    u64* tmp = c->alloc(have);   // canonical residues for the caller
    CUDA_CHECK(cudaMemcpyAsync(tmp, k->d, sizeof(u64) * have, cudaMemcpyDeviceToDevice, c->stream));
    ops::split30_inplace(c, tmp, have, true, c->stream);
    int rc = export_words(c, tmp, have, host, words);
    c->free(tmp);
    return rc;
    API_END
}
int spear_kswitch_key_export(const spear_kswitch_key* k_, uint64_t* host, size_t words) {
    return export_key(reinterpret_cast<const KSKey*>(k_), host, words);
}
int spear_galois_key_export(const spear_galois_keys* gk_, uint32_t elt, uint64_t* host, size_t words) {
    API_BEGIN
    const GaloisKeys* gk = reinterpret_cast<const GaloisKeys*>(gk_);
    return export_key(find_key(gk, elt), host, words);
    API_END
}
int spear_public_key_export(const spear_public_key* pk_, uint64_t* host, size_t words) {
    const PublicKey* pk = reinterpret_cast<const PublicKey*>(pk_);
    return export_words(pk->ctx, pk->d, (size_t)2 * pk->ctx->K * pk->ctx->N, host, words);
}

// ---- encoder ------------------------------------------------------------------------------------
int spear_encode(spear_context* ctx, const double* values, int count, int ring_n, double scale, int chain_index,
                 int ext, spear_obj** outs) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    REQUIRE(count >= 1 && chain_index >= 1 && chain_index <= c->L, "encode: chain_index %d out of range", chain_index);
    REQUIRE(scale > 0, "encode: scale must be positive");
    const int l = c->limbs_at(chain_index), rows = l + (ext ? c->P : 0);
    const int chunk = std::max(1, std::min(count, (int)((64u << 20) / ((size_t)rows * ring_n))));
    double2* dv = (double2*)c->alloc((size_t)chunk * ring_n);   // chunk * ring_n/2 double2 = chunk*ring_n words
    u64* buf = c->alloc((size_t)chunk * rows * ring_n);
    std::vector<std::unique_ptr<Obj>> made;
    for (int v0 = 0; v0 < count; v0 += chunk) {
        int nv = std::min(chunk, count - v0);
        CUDA_CHECK(cudaMemcpyAsync(dv, values + (size_t)v0 * ring_n, sizeof(double) * nv * ring_n,
                                   cudaMemcpyHostToDevice, c->stream));
        encoder::encode(c, dv, nv, ring_n, scale, l, ext != 0, buf, c->stream);
        for (int v = 0; v < nv; v++) {
            made.emplace_back(new_obj(c, 1, l, ext != 0, ring_n, scale));
            CUDA_CHECK(cudaMemcpyAsync(made.back()->d, buf + (size_t)v * rows * ring_n, sizeof(u64) * rows * ring_n,
                                       cudaMemcpyDeviceToDevice, c->stream));
        }
    }
    c->free(dv);
    c->free(buf);
    for (int v = 0; v < count; v++) outs[v] = H_(made[v].release());
    API_END
}
int spear_decode(spear_context* ctx, const spear_obj* pt_, double* out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const Obj* pt = O_(pt_);
    check_pt(pt, "decode");
    double2* dv = (double2*)c->alloc((size_t)c->N);
    encoder::decode(c, pt->d, pt->l, pt->scale, dv, c->stream);
    CUDA_CHECK(cudaMemcpyAsync(out, dv, sizeof(double2) * (c->N / 2), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->free(dv);
    API_END
}

// ---- encryption ---------------------------------------------------------------------------------
int spear_encrypt_symmetric(spear_context* ctx, const spear_secret_key* sk_, const spear_obj* pt_, uint64_t enc_id,
                            spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const SecretKey* sk = reinterpret_cast<const SecretKey*>(sk_);
    const Obj* pt = O_(pt_);
    check_pt(pt, "encrypt_symmetric");
    REQUIRE(!pt->ext, "encrypt_symmetric: plaintext carries special limbs");
    const int l = pt->l;
    const size_t N = c->N;
    std::unique_ptr<Obj> ct(new_obj(c, 2, l, false, c->N, pt->scale));
    u64* e = c->alloc(l * N);
    RowMap rm = data_rows(c, l);
    sampler::uniform(c, sk->seed, stream_id(DOM_ENC_A, enc_id), ct->poly(1), l, rm, c->stream);
    sampler::cbd(c, sk->seed, stream_id(DOM_ENC_E, enc_id), e, l, rm, c->stream);
    ntt_forward(c, e, l, rm, c->N, c->stream);
    sampler::enc_combine(c, ct->poly(1), e, sk->d, pt->d, ct->poly(0), l, c->stream);
    c->free(e);
    *out = H_(ct.release());
    API_END
}
int spear_encrypt_asymmetric(spear_context* ctx, const spear_public_key* pk_, const spear_obj* pt_, uint64_t enc_id,
                             spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const PublicKey* pk = reinterpret_cast<const PublicKey*>(pk_);
    const Obj* pt = O_(pt_);
    check_pt(pt, "encrypt_asymmetric");
    REQUIRE(!pt->ext, "encrypt_asymmetric: plaintext carries special limbs");
    const int l = pt->l, rows = l + c->P;
    const size_t N = c->N, pw = (size_t)rows * N;
    std::unique_ptr<Obj> ct(new_obj(c, 2, l, false, c->N, pt->scale));
    u64* u = c->alloc(3 * pw);
    u64 *e0 = u + pw, *e1 = u + 2 * pw;
    u64* t = c->alloc(2 * pw);
    u64* tmp = c->alloc(2 * l * N);
    RowMap rm{rows, l, c->L, 0};
    sampler::ternary(c, pk->seed, stream_id(DOM_ASYM_U, enc_id), u, rows, rm, c->stream);
    sampler::cbd(c, pk->seed, stream_id(DOM_ASYM_E0, enc_id), e0, rows, rm, c->stream);
    sampler::cbd(c, pk->seed, stream_id(DOM_ASYM_E1, enc_id), e1, rows, rm, c->stream);
    ntt_forward(c, u, 3 * rows, rm, c->N, c->stream);
    sampler::asym_combine(c, pk->d, u, e0, e1, t, l, c->stream);
    ops::moddown(c, t, pw, 2, l, tmp, nullptr, ct->d, c->stream);
    ops::add(c, ct->d, pt->d, ct->d, 1, l, c->N, data_rows(c, l), 1, c->stream);
    c->free(u);
    c->free(t);
    c->free(tmp);
    *out = H_(ct.release());
    API_END
}
int spear_decrypt(spear_context* ctx, const spear_secret_key* sk_, const spear_obj* ct_, spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const SecretKey* sk = reinterpret_cast<const SecretKey*>(sk_);
    const Obj* ct = O_(ct_);
    check_ct(ct, "decrypt");
    std::unique_ptr<Obj> pt(new_obj(c, 1, ct->l, false, c->N, ct->scale));
    sampler::dec_combine(c, ct->d, ct->size, ct->l, sk->d, pt->d, c->stream);
    *out = H_(pt.release());
    API_END
}

// ---- fused client legs (client.cu) ------------------------------------------------------------------
int spear_encrypt_vector(spear_context* ctx, const spear_secret_key* sk_, const double* values, int count, int replicate,
                         double scale, uint64_t enc_id, spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const SecretKey* sk = reinterpret_cast<const SecretKey*>(sk_);
    const int slots = c->N / 2, l = c->L;
    REQUIRE(values && count >= 1 && count <= slots, "encrypt_vector: 1..%d slot values expected", slots);
    REQUIRE(scale > 0, "encrypt_vector: scale must be positive");
    // |coefficient| <= max |z| * scale: the three-launch path needs no device-side overflow flag below 2^125
    double zmax = 0.0;
    for (int i = 0; i < 2 * count; i++) zmax = std::max(zmax, std::fabs(values[i]));
    if (client::fused_applies(c) && zmax * scale < 0x1p125) {
        std::unique_ptr<Obj> ct(new_obj(c, 2, l, false, c->N, scale));
        double2* dv = (double2*)c->alloc((size_t)2 * count);
        double2* W = (double2*)c->alloc((size_t)2 * c->N);
        CUDA_CHECK(cudaMemcpyAsync(dv, values, sizeof(double2) * count, cudaMemcpyHostToDevice, c->stream));
        client::encode_encrypt(c, dv, count, replicate != 0, scale, l, sk->seed, enc_id, sk->d, ct->d, W, c->stream);
        c->free(dv);
        c->free(W);
        *out = H_(ct.release());
        return SPEAR_OK;
    }
    // staged form: the full slot vector on the host, encode, encrypt
    std::vector<double> full((size_t)2 * slots, 0.0);
    for (int j = 0; j < (replicate ? slots : count); j++) {
        full[2 * j] = values[2 * (j % count)];
        full[2 * j + 1] = values[2 * (j % count) + 1];
    }
    spear_obj* pt = nullptr;
    int rc = spear_encode(ctx, full.data(), 1, c->N, scale, 1, 0, &pt);
    if (rc) return rc;
    rc = spear_encrypt_symmetric(ctx, sk_, pt, enc_id, out);
    spear_obj_destroy(pt);
    return rc;
    API_END
}
int spear_decrypt_decode(spear_context* ctx, const spear_secret_key* sk_, const spear_obj* ct_, double* out, int want) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const SecretKey* sk = reinterpret_cast<const SecretKey*>(sk_);
    const Obj* ct = O_(ct_);
    check_ct(ct, "decrypt_decode");
    REQUIRE(out && want >= 1 && want <= c->N / 2, "decrypt_decode: 1..%d slots", c->N / 2);
    double2* dv = (double2*)c->alloc((size_t)c->N);
    if (client::fused_applies(c)) {
        u64* x = c->alloc((size_t)3 * c->N);
        double2* W = (double2*)c->alloc((size_t)2 * c->N);
        client::decrypt_decode(c, ct->d, ct->size, ct->l, ct->scale, sk->d, dv, want, x, W, c->stream);
        c->free(x);
        c->free(W);
    } else {
        std::unique_ptr<Obj> pt(new_obj(c, 1, ct->l, false, c->N, ct->scale));
        sampler::dec_combine(c, ct->d, ct->size, ct->l, sk->d, pt->d, c->stream);
        encoder::decode(c, pt->d, pt->l, pt->scale, dv, c->stream);
    }
    CUDA_CHECK(cudaMemcpyAsync(out, dv, sizeof(double2) * want, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->free(dv);
    API_END
}

// ---- evaluator ------------------------------------------------------------------------------------
int spear_negate(spear_context* ctx, const spear_obj* a_, spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const Obj* a = O_(a_);
    check_ct(a, "negate");
    std::unique_ptr<Obj> o(new_obj(c, a->size, a->l, false, c->N, a->scale));
    ops::neg(c, a->d, o->d, a->size, a->l, c->N, data_rows(c, a->l), c->stream);
    *out = H_(o.release());
    API_END
}
static int addsub(spear_context* ctx, const spear_obj* a_, const spear_obj* b_, spear_obj** out, bool is_sub) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const Obj *a = O_(a_), *b = O_(b_);
    REQUIRE(a && b && a->size >= 2 && b->size >= 2 && a->n == c->N && b->n == c->N && a->ext == b->ext,
            "add/sub: expected two ciphertexts in the same basis");
    REQUIRE(a->l == b->l, "add/sub: chain_index mismatch (%d vs %d limbs)", a->l, b->l);
    // a is the larger one
    const bool swap = b->size > a->size;
    const Obj *big = swap ? b : a, *small = swap ? a : b;
    std::unique_ptr<Obj> o(new_obj(c, big->size, a->l, a->ext, c->N, a->scale));
    const int nr = a->rows();
    const size_t pw = (size_t)nr * c->N;
    RowMap rm{nr, a->l, c->L, 0};
    if (!is_sub) {
        ops::add(c, big->d, small->d, o->d, small->size, nr, c->N, rm, small->size, c->stream);
        if (big->size > small->size)
            CUDA_CHECK(cudaMemcpyAsync(o->d + small->size * pw, big->d + small->size * pw,
                                       sizeof(u64) * (big->size - small->size) * pw, cudaMemcpyDeviceToDevice, c->stream));
    } else {
        ops::sub(c, a->d, b->d, o->d, small->size, nr, c->N, rm, small->size, c->stream);
        if (a->size > b->size)
            CUDA_CHECK(cudaMemcpyAsync(o->d + small->size * pw, a->d + small->size * pw,
                                       sizeof(u64) * (a->size - b->size) * pw, cudaMemcpyDeviceToDevice, c->stream));
        else if (b->size > a->size)
            ops::neg(c, b->d + small->size * pw, o->d + small->size * pw, b->size - a->size, nr, c->N, rm, c->stream);
    }
    *out = H_(o.release());
    API_END
}
int spear_add(spear_context* ctx, const spear_obj* a, const spear_obj* b, spear_obj** out) { return addsub(ctx, a, b, out, false); }
int spear_sub(spear_context* ctx, const spear_obj* a, const spear_obj* b, spear_obj** out) { return addsub(ctx, a, b, out, true); }

static int addsub_plain(spear_context* ctx, const spear_obj* ct_, const spear_obj* pt_, spear_obj** out, bool is_sub) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const Obj *ct = O_(ct_), *pt = O_(pt_);
    check_ct(ct, "add_plain");
    check_pt(pt, "add_plain");
    REQUIRE(pt->l >= ct->l && !pt->ext, "add_plain: plaintext has fewer limbs than the ciphertext");
    std::unique_ptr<Obj> o(new_obj(c, ct->size, ct->l, false, c->N, ct->scale));
    CUDA_CHECK(cudaMemcpyAsync(o->d, ct->d, sizeof(u64) * ct->words(), cudaMemcpyDeviceToDevice, c->stream));
    RowMap rm = data_rows(c, ct->l);
    if (is_sub) ops::sub(c, ct->d, pt->d, o->d, 1, ct->l, c->N, rm, 1, c->stream);
    else ops::add(c, ct->d, pt->d, o->d, 1, ct->l, c->N, rm, 1, c->stream);
    *out = H_(o.release());
    API_END
}
int spear_add_plain(spear_context* ctx, const spear_obj* ct, const spear_obj* pt, spear_obj** out) { return addsub_plain(ctx, ct, pt, out, false); }
int spear_sub_plain(spear_context* ctx, const spear_obj* ct, const spear_obj* pt, spear_obj** out) { return addsub_plain(ctx, ct, pt, out, true); }

int spear_multiply(spear_context* ctx, const spear_obj* a_, const spear_obj* b_, spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const Obj *a = O_(a_), *b = O_(b_);
    check_ct(a, "multiply");
    check_ct(b, "multiply");
    REQUIRE(a->size == 2 && b->size == 2, "multiply: operands must be size-2 ciphertexts (relinearize first)");
    REQUIRE(a->l == b->l, "multiply: chain_index mismatch");
    std::unique_ptr<Obj> o(new_obj(c, 3, a->l, false, c->N, a->scale * b->scale));
    ops::tensor(c, a->d, b->d, o->d, a->l, c->stream);
    *out = H_(o.release());
    API_END
}
int spear_multiply_plain(spear_context* ctx, const spear_obj* ct_, const spear_obj* pt_, spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const Obj *ct = O_(ct_), *pt = O_(pt_);
    check_ct(ct, "multiply_plain");
    check_pt(pt, "multiply_plain");
    REQUIRE(pt->l >= ct->l && !pt->ext, "multiply_plain: plaintext has fewer limbs than the ciphertext");
    std::unique_ptr<Obj> o(new_obj(c, ct->size, ct->l, false, c->N, ct->scale * pt->scale));
    ops::mul(c, ct->d, pt->d, o->d, ct->size, ct->l, c->N, data_rows(c, ct->l), 1, c->stream);
    *out = H_(o.release());
    API_END
}
int spear_relinearize(spear_context* ctx, const spear_obj* ct_, const spear_kswitch_key* rlk_, spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const Obj* ct = O_(ct_);
    check_ct(ct, "relinearize");
    const KSKey* rlk = reinterpret_cast<const KSKey*>(rlk_);
    std::unique_ptr<Obj> o(new_obj(c, 2, ct->l, false, c->N, ct->scale));
    if (ct->size == 2) CUDA_CHECK(cudaMemcpyAsync(o->d, ct->d, sizeof(u64) * ct->words(), cudaMemcpyDeviceToDevice, c->stream));
    else eng::relinearize(c, ct->d, ct->l, rlk->d, o->d, c->stream);
    *out = H_(o.release());
    API_END
}
int spear_rescale_to_next(spear_context* ctx, const spear_obj* ct_, spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const Obj* ct = O_(ct_);
    check_ct(ct, "rescale_to_next");
    REQUIRE(ct->l >= 2, "rescale_to_next: already at the last level");
    std::unique_ptr<Obj> o(new_obj(c, ct->size, ct->l - 1, false, c->N, ct->scale / (double)c->q[ct->l - 1]));
    eng::rescale(c, ct->d, ct->size, ct->l, o->d, c->stream);
    *out = H_(o.release());
    API_END
}
int spear_mod_raise(spear_context* ctx, const spear_obj* ct_, int chain_index, spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const Obj* ct = O_(ct_);
    check_ct(ct, "mod_raise");
    REQUIRE(ct->l == 1, "mod_raise: the ciphertext must be at the last level (one limb), has %d", ct->l);
    const int l = c->L - chain_index + 1;
    REQUIRE(chain_index >= 1 && l >= 1, "mod_raise: chain_index %d out of range", chain_index);
    std::unique_ptr<Obj> o(new_obj(c, ct->size, l, false, c->N, ct->scale));
    u64* x = c->alloc((size_t)ct->size * c->N);
    ops::mod_raise(c, ct->d, ct->size, l, x, o->d, c->stream);
    c->free(x);
    *out = H_(o.release());
    API_END
}
int spear_mod_switch_to_next(spear_context* ctx, const spear_obj* a_, spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const Obj* a = O_(a_);
    REQUIRE(a && !a->ext && a->n == c->N, "mod_switch_to_next: bad operand");
    REQUIRE(a->l >= 2, "mod_switch_to_next: already at the last level");
    std::unique_ptr<Obj> o(new_obj(c, a->size, a->l - 1, false, c->N, a->scale));
    const size_t N = c->N;
    CUDA_CHECK(cudaMemcpy2DAsync(o->d, sizeof(u64) * (a->l - 1) * N, a->d, sizeof(u64) * a->l * N,
                                 sizeof(u64) * (a->l - 1) * N, a->size, cudaMemcpyDeviceToDevice, c->stream));
    *out = H_(o.release());
    API_END
}
int spear_apply_galois(spear_context* ctx, const spear_obj* ct_, uint32_t elt, const spear_galois_keys* gk_,
                       spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const Obj* ct = O_(ct_);
    check_ct(ct, "apply_galois");
    REQUIRE(ct->size == 2, "apply_galois: relinearize first");
    const KSKey* key = find_key(reinterpret_cast<const GaloisKeys*>(gk_), elt);
    std::unique_ptr<Obj> o(new_obj(c, 2, ct->l, false, c->N, ct->scale));
    eng::apply_galois(c, ct->d, ct->l, elt, key->d, o->d, c->stream);
    *out = H_(o.release());
    API_END
}
int spear_hoisted_rotations(spear_context* ctx, const spear_obj* ct_, const uint32_t* elts, int count,
                            const spear_galois_keys* gk_, spear_obj** outs) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const Obj* ct = O_(ct_);
    check_ct(ct, "hoisting");
    REQUIRE(ct->size == 2, "hoisting: relinearize first");
    const GaloisKeys* gk = reinterpret_cast<const GaloisKeys*>(gk_);
    const int l = ct->l, rows = l + c->P;
    const size_t N = c->N, pw = (size_t)rows * N;
    for (int i = 0; i < count; i++) find_key(gk, elts[i]);
    u64* x = c->alloc(l * N);
    u64* E = c->alloc((size_t)c->digits(l) * pw);
    u64* acc = c->alloc(2 * pw);
    u64* tmp = c->alloc(2 * l * N);
    ops::decompose(c, ct->poly(1), l, x, E, c->stream);
    std::vector<std::unique_ptr<Obj>> made;
    for (int i = 0; i < count; i++) {
        made.emplace_back(new_obj(c, 2, l, false, c->N, ct->scale));
        // (P*pi(c0) + <pi(E), k0>, <pi(E), k1>) then ModDown
        ops::ks_inner(c, E, find_key(gk, elts[i])->d, acc, l, elts[i], ct->poly(0), l, 1, 0, c->stream);
        ops::moddown(c, acc, pw, 2, l, tmp, nullptr, made.back()->d, c->stream);
    }
    c->free(x), c->free(E), c->free(acc), c->free(tmp);
    for (int i = 0; i < count; i++) outs[i] = H_(made[i].release());
    API_END
}

// ---- BSGS ---------------------------------------------------------------------------------------
int spear_bsgs_multiply_accumulate(spear_context* ctx, spear_obj* const* ct_baby, int n_baby, spear_obj* const* pts,
                                   int n_pts, int G, int B, int D, const spear_galois_keys* gk_, spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    REQUIRE(G >= 1 && B >= 1 && D >= 1 && n_baby >= std::min(G, D) && n_pts >= D && (size_t)G * B >= (size_t)D,
            "bsgs: need G baby ciphertexts, D plaintexts and G*B >= D");
    const GaloisKeys* gk = reinterpret_cast<const GaloisKeys*>(gk_);
    const Obj* b0 = O_(ct_baby[0]);
    check_ct(b0, "bsgs");
    const int l = b0->l;
    std::vector<const u64*> baby(G, nullptr), pt(D);
    for (int b = 0; b < std::min(G, n_baby); b++) {
        const Obj* o = O_(ct_baby[b]);
        check_ct(o, "bsgs");
        REQUIRE(o->size == 2 && o->l == l, "bsgs: baby ciphertext %d has a different level", b);
        baby[b] = o->d;
    }
    for (int k = 0; k < D; k++) {
        const Obj* o = O_(pts[k]);
        check_pt(o, "bsgs");
        REQUIRE(o->l >= l && !o->ext, "bsgs: diagonal %d has fewer limbs than the ciphertext", k);
        pt[k] = o->d;
    }
    std::vector<u32> gelt(B, 0);
    std::vector<const u64*> gkey(B, nullptr);
    for (int g = 1; g < B && g * G < D; g++) {
        gelt[g] = (u32)elt_from_step(g * G, c->N);
        gkey[g] = find_key(gk, gelt[g])->d;
    }
    double scale = b0->scale * O_(pts[0])->scale / (double)c->q[l - 1];
    REQUIRE(l >= 2, "bsgs: no level left for the final rescale");
    std::unique_ptr<Obj> o(new_obj(c, 2, l - 1, false, c->N, scale));
    eng::bsgs_exact(c, baby.data(), pt.data(), G, B, D, l, gelt.data(), gkey.data(), o->d, c->stream);
    *out = H_(o.release());
    API_END
}

int spear_bsgs_from_host(spear_context* ctx, spear_obj* const* ct_baby, int n_baby, const uint64_t* host_pts, int n_pts,
                         int pt_limbs, double pt_scale, int G, int B, int D, const spear_galois_keys* gk_, spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    REQUIRE(G >= 1 && B >= 1 && D >= 1 && n_baby >= std::min(G, D) && n_pts >= D && (size_t)G * B >= (size_t)D && host_pts,
            "bsgs: need G baby ciphertexts, D plaintexts and G*B >= D");
    const GaloisKeys* gk = reinterpret_cast<const GaloisKeys*>(gk_);
    const Obj* b0 = O_(ct_baby[0]);
    check_ct(b0, "bsgs");
    const int l = b0->l;
    REQUIRE(pt_limbs >= l && pt_limbs <= c->L, "bsgs: plaintexts have %d limbs, the ciphertext %d", pt_limbs, l);
    REQUIRE(l >= 2, "bsgs: no level left for the final rescale");
    std::vector<const u64*> baby(G, nullptr);
    for (int b = 0; b < std::min(G, n_baby); b++) {
        const Obj* o = O_(ct_baby[b]);
        check_ct(o, "bsgs");
        REQUIRE(o->size == 2 && o->l == l, "bsgs: baby ciphertext %d has a different level", b);
        baby[b] = o->d;
    }
    std::vector<u32> gelt(B, 0);
    std::vector<const u64*> gkey(B, nullptr);
    for (int g = 1; g < B && g * G < D; g++) {
        gelt[g] = (u32)elt_from_step(g * G, c->N);
        gkey[g] = find_key(gk, gelt[g])->d;
    }
    std::unique_ptr<Obj> o(new_obj(c, 2, l - 1, false, c->N, b0->scale * pt_scale / (double)c->q[l - 1]));
    eng::bsgs_exact_from_host(c, baby.data(), host_pts, pt_limbs, G, B, D, l, gelt.data(), gkey.data(), o->d, c->stream);
    // the ring is recycled by the pool once the stream has passed it; the host buffer must stay untouched until then
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    *out = H_(o.release());
    API_END
}
// count objects of one shape -> one host buffer [count][words_each]; one synchronisation for the whole batch
int spear_objs_export(spear_context* ctx, spear_obj* const* objs, int count, uint64_t* host, size_t words_each) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    for (int i = 0; i < count; i++) {
        const Obj* o = O_(objs[i]);
        REQUIRE(o && o->words() == words_each, "export: object %d has %zu words, expected %zu", i, o ? o->words() : 0, words_each);
        CUDA_CHECK(cudaMemcpyAsync(host + (size_t)i * words_each, o->d, sizeof(u64) * words_each, cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    API_END
}

// dv: [n_diags][D] complex rows on the device (pre-rotated diagonals of the shard's groups, in storage order), or
// nullptr with `host` pointing at the same rows in host memory
static DiagSet* diagset_build(Ctx* c, const double* host, double2* dv_in, int n_diags, int D, int G, int B, int g_first,
                              int g_stride, double scale, int chain_index, int compress) {
    REQUIRE(D >= 1 && G >= 1 && B >= 1 && (size_t)G * B >= (size_t)D && D <= c->N / 2, "diagset: bad D/G/B");
    REQUIRE(g_first >= 0 && g_stride >= 1 && n_diags >= 0 && n_diags <= D, "diagset: bad shard");
    REQUIRE(chain_index >= 1 && chain_index <= c->L, "diagset: chain_index out of range");
    {   // the shard must hold exactly the diagonals of groups g_first, g_first + g_stride, ...
        int expect = 0;
        for (int g = g_first; g < B && g * G < D; g += g_stride) expect += std::min(G, D - g * G);
        REQUIRE(expect == n_diags, "diagset: shard (first %d, stride %d) holds %d diagonals, got %d", g_first, g_stride,
                expect, n_diags);
    }
    const int slots = c->N / 2, l = c->limbs_at(chain_index), rows = l + c->P;
    const bool pow2 = (D & (D - 1)) == 0;
    REQUIRE(!compress || pow2, "diagset: sub-ring compression needs D to be a power of two");
    const int n = (compress && pow2 && D >= 2) ? 2 * D : c->N;
    std::unique_ptr<DiagSet> ds(new DiagSet);
    ds->bind(c), ds->D = D, ds->G = G, ds->B = B, ds->l = l, ds->n = n, ds->scale = scale;
    ds->n_diags = n_diags, ds->g_first = g_first, ds->g_stride = g_stride;
    ds->rshift = 0;
    while ((n << ds->rshift) < c->N) ds->rshift++;
    // capacity of a set at the top level whatever `l` is: sets encoded on the fly at falling levels (fully encrypted
    // blocks) then recycle the same pool block instead of growing the pool with a new size every time
    ds->d = c->alloc((size_t)std::max(n_diags, 1) * (c->L + c->P) * n);
    if (n_diags > 0) {
        double2* dv = dv_in;
        if (!dv) {
            dv = (double2*)c->alloc((size_t)n_diags * D * 2);
            CUDA_CHECK(cudaMemcpyAsync(dv, host, sizeof(double2) * n_diags * D, cudaMemcpyHostToDevice, c->stream));
        }
        static const bool fused_enc = [] {
            const char* e = getenv("SPEAR_FUSED_ENCODE");
            return !(e && e[0] == '0');
        }();
        if (fused_enc && n == 2 * D) {   // sub-ring diagonals: one kernel per set, one synchronisation for the overflow flag
            int* flag = (int*)c->alloc(1);
            CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
            if (encoder::encode_ring_fused(c, dv, n_diags, n, scale, l, true, ds->d, /*split30_out=*/true, flag, c->stream)) {
                int h_flag = 0;
                CUDA_CHECK(cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
                CUDA_CHECK(cudaStreamSynchronize(c->stream));
                c->free(flag);
                if (!dv_in) c->free(dv);
                REQUIRE(!h_flag, "encode: scaled value too large (|coefficient| >= 2^126)");
                return ds.release();
            }
            c->free(flag);
        }
        const int chunk = std::max(1, std::min(n_diags, (int)((32u << 20) / ((size_t)n))));
        double2* full = n == 2 * D ? nullptr : (double2*)c->alloc((size_t)chunk * slots * 2);
        for (int v0 = 0; v0 < n_diags; v0 += chunk) {
            int nv = std::min(chunk, n_diags - v0);
            const double2* src = dv + (size_t)v0 * D;
            if (full) {   // replicate the period-D vector over all slots (reference :371-378)
                LAUNCH(k_tile, c->sm_count * 8, 256, 0, c->stream)(src, full, nv, D, slots);
                src = full;
            }
            encoder::encode(c, src, nv, n, scale, l, true, ds->d + (size_t)v0 * rows * n, c->stream);
        }
        if (full) c->free(full);
        if (!dv_in) c->free(dv);
        // the diagonal MAC kernels consume the split-30 storage form (common.cuh)
        ops::split30_inplace(c, ds->d, (size_t)n_diags * rows * n, false, c->stream);
    }
    return ds.release();
}
int spear_diagset_encode_shard(spear_context* ctx, const double* diags, int n_diags, int D, int G, int B, int g_first,
                               int g_stride, double scale, int chain_index, int compress, spear_diagset** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    *out = reinterpret_cast<spear_diagset*>(
        diagset_build(c, diags, nullptr, n_diags, D, G, B, g_first, g_stride, scale, chain_index, compress));
    API_END
}
// rows of the shard's giant groups straight from the matrix: row (g, b), k = gG + b:
//   out[row][t] = M[j][(j + k) mod D],  j = (t - gG) mod D     (diagonal k of y = M x, rolled right by gG)
// The device copy of the matrix keeps the host layout: element (i, j) of M sits at  i * ld + j  (or  j * ld + i  when
// `transposed`), and is zero outside [0, rv) x [0, cv) -- so callers hand over views such as W[:, lo:hi].T without a
// host-side transposed copy (33 MB and ~35 ms per chunk at D = 2048).
__global__ void k_diag_rows(const double* __restrict__ mre, const double* __restrict__ mim, double2* __restrict__ out,
                            int n_diags, int D, int G, int B, int g_first, int g_stride, int ld, int transposed, int rv,
                            int cv) {
    const size_t total = (size_t)n_diags * D;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(e / D), t = (int)(e % D);
        // full groups hold G diagonals; only the last group of the matrix can be shorter, and it is last in the shard
        const int gi = row / G, b = row % G, g = g_first + gi * g_stride, k = g * G + b;
        int j = t - (g * G) % D;
        if (j < 0) j += D;
        const int col = (j + k) % D;
        double re = 0.0, im = 0.0;
        if (j < rv && col < cv) {
            const size_t at = transposed ? (size_t)col * ld + j : (size_t)j * ld + col;
            re = mre[at];
            if (mim) im = mim[at];
        }
        out[e] = make_double2(re, im);
    }
}
int spear_diagset_encode_matrix(spear_context* ctx, const double* m_re, const double* m_im, int D, int G, int B,
                                int g_first, int g_stride, double scale, int chain_index, int compress,
                                spear_diagset** out) {
    return spear_diagset_encode_matrix_view(ctx, m_re, m_im, D, D, D, (size_t)D, 0, G, B, g_first, g_stride, scale, chain_index,
                                            compress, out);
}
int spear_diagset_encode_matrix_view(spear_context* ctx, const double* m_re, const double* m_im, int D, int rows_valid,
                                     int cols_valid, size_t pitch, int transposed, int G, int B, int g_first, int g_stride,
                                     double scale, int chain_index, int compress, spear_diagset** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    REQUIRE(m_re && D >= 1 && G >= 1 && B >= 1 && g_first >= 0 && g_stride >= 1, "diagset: bad matrix arguments");
    REQUIRE(rows_valid >= 0 && rows_valid <= D && cols_valid >= 0 && cols_valid <= D, "diagset: view larger than D x D");
    // host lines: `lines` runs of `run` contiguous doubles, `pitch` doubles apart; compacted on the device (ld = run)
    const int lines = transposed ? cols_valid : rows_valid, run = transposed ? rows_valid : cols_valid;
    REQUIRE(pitch >= (size_t)run, "diagset: pitch %zu shorter than a line of %d", pitch, run);
    int n_diags = 0;
    for (int g = g_first; g < B && g * G < D; g += g_stride) n_diags += std::min(G, D - g * G);
    const size_t mw = (size_t)std::max(lines, 1) * std::max(run, 1);
    double* dm = (double*)c->alloc(mw * (m_im ? 2 : 1));
    if (lines > 0 && run > 0) {
        // pageable caller memory: the lines are packed into the context's page-locked staging buffer (plain memcpy, one
        // line at a time) and cross the bus in one asynchronous copy
        const int parts = m_im ? 2 : 1;
        auto page_locked = [](const void* p) {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
                cudaGetLastError();
                return false;
            }
            return at.type == cudaMemoryTypeHost;
        };
        if (page_locked(m_re) && (!m_im || page_locked(m_im))) {
            // weights already in page-locked memory (spear_pinned_alloc / cudaHostRegister): straight DMA of the view
            for (int part = 0; part < parts; part++)
                CUDA_CHECK(cudaMemcpy2DAsync(dm + (size_t)part * mw, sizeof(double) * run, part ? m_im : m_re,
                                             sizeof(double) * pitch, sizeof(double) * run, lines, cudaMemcpyHostToDevice, c->stream));
        } else {
            double* st = (double*)c->staging(sizeof(double) * mw * parts);
            for (int part = 0; part < parts; part++) {
                const double* src = part ? m_im : m_re;
                double* dst = st + (size_t)part * mw;
                if (pitch == (size_t)run) memcpy(dst, src, sizeof(double) * mw);
                else
                    for (int i = 0; i < lines; i++) memcpy(dst + (size_t)i * run, src + (size_t)i * pitch, sizeof(double) * run);
            }
            CUDA_CHECK(cudaMemcpyAsync(dm, st, sizeof(double) * mw * parts, cudaMemcpyHostToDevice, c->stream));
            CUDA_CHECK(cudaEventRecord(c->staged, c->stream));
        }
    }
    double2* dv = (double2*)c->alloc((size_t)std::max(n_diags, 1) * D * 2);
    if (n_diags > 0)
        LAUNCH(k_diag_rows, c->sm_count * 8, 256, 0, c->stream)(dm, m_im ? dm + mw : nullptr, dv, n_diags, D, G, B, g_first,
                                                              g_stride, run, transposed ? 1 : 0, rows_valid, cols_valid);
    DiagSet* ds = nullptr;
    try {
        ds = diagset_build(c, nullptr, dv, n_diags, D, G, B, g_first, g_stride, scale, chain_index, compress);
    } catch (...) {
        c->free(dv), c->free(dm);
        throw;
    }
    c->free(dv), c->free(dm);
    *out = reinterpret_cast<spear_diagset*>(ds);
    API_END
}
int spear_diagset_encode(spear_context* ctx, const double* diags, int D, int G, int B, double scale, int chain_index,
                         int compress, spear_diagset** out) {
    return spear_diagset_encode_shard(ctx, diags, D, D, G, B, 0, 1, scale, chain_index, compress, out);
}
void spear_diagset_destroy(spear_diagset* d) { delete reinterpret_cast<DiagSet*>(d); }
int spear_diagset_info(const spear_diagset* d_, int* D, int* G, int* B, int* limbs, int* ring_n, double* scale,
                       uint64_t* bytes) {
    API_BEGIN
    const DiagSet* d = reinterpret_cast<const DiagSet*>(d_);
    REQUIRE(d, "null diagset");
    if (D) *D = d->D;
    if (G) *G = d->G;
    if (B) *B = d->B;
    if (limbs) *limbs = d->l;
    if (ring_n) *ring_n = d->n;
    if (scale) *scale = d->scale;
    if (bytes) *bytes = sizeof(u64) * (size_t)d->n_diags * d->stored_rows() * d->stored_cols();
    API_END
}
int spear_diagset_export(const spear_diagset* d_, uint64_t* host, size_t words) {
    API_BEGIN
    const DiagSet* d = reinterpret_cast<const DiagSet*>(d_);
    Ctx* c = d->ctx;
    use(c);
    const size_t have = (size_t)d->n_diags * d->stored_rows() * d->stored_cols();
    REQUIRE(words == have, "export: buffer holds %zu words, diagonal set has %zu", words, have);
    u64* tmp = c->alloc(have);   // canonical residues for the caller; the resident copy stays split-30
    CUDA_CHECK(cudaMemcpyAsync(tmp, d->d, sizeof(u64) * have, cudaMemcpyDeviceToDevice, c->stream));
    ops::split30_inplace(c, tmp, have, true, c->stream);
    int rc = export_words(c, tmp, have, host, words);
    c->free(tmp);
    return rc;
    API_END
}

// out[d][r][c] = in[d][row0 + r][col0 + c]
__global__ void k_diag_slice(const u64* __restrict__ in, u64* __restrict__ out, int n_diags, int rows, int n, int row0,
                             int nrows, int col0, int ncols) {
    const size_t total = (size_t)n_diags * nrows * ncols;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int cc = (int)(e % ncols);
        const size_t dr = e / ncols;
        const int r = (int)(dr % nrows);
        const size_t d = dr / nrows;
        out[e] = in[(d * rows + row0 + r) * n + col0 + cc];
    }
}
// rows [row0, row0 + nrows) x stored columns [col0, col0 + ncols) of a full set, as a set of its own: the diagonals a
// rank holds in a two-phase mat-vec
static DiagSet* diagset_slice(Ctx* c, const DiagSet* f, int row0, int nrows, int col0, int ncols) {
    REQUIRE(f && f->nrows < 0 && f->ncols < 0 && f->g_first == 0 && f->g_stride == 1, "slice: expected a full diagonal set");
    const int rows = f->l + c->P;
    REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= rows, "slice: rows [%d, %d) of %d", row0, row0 + nrows, rows);
    REQUIRE(col0 >= 0 && ncols >= 0 && col0 + ncols <= f->n, "slice: columns [%d, %d) of %d", col0, col0 + ncols, f->n);
    std::unique_ptr<DiagSet> ds(new DiagSet);
    ds->bind(c), ds->D = f->D, ds->G = f->G, ds->B = f->B, ds->l = f->l, ds->n = f->n, ds->scale = f->scale;
    ds->n_diags = f->n_diags, ds->g_first = 0, ds->g_stride = 1, ds->rshift = f->rshift;
    ds->row0 = row0, ds->nrows = nrows, ds->col0 = col0, ds->ncols = ncols;
    const size_t words = (size_t)std::max(f->n_diags, 1) * std::max(nrows, 1) * std::max(ncols, 1);
    ds->d = c->alloc(words);
    if (f->n_diags > 0 && nrows > 0 && ncols > 0)
        LAUNCH(k_diag_slice, c->sm_count * 8, 256, 0, c->stream)(f->d, ds->d, f->n_diags, rows, f->n, row0, nrows, col0, ncols);
    CUDA_CHECK(cudaGetLastError());
    return ds.release();
}
int spear_diagset_slice_rows(spear_context* ctx, const spear_diagset* full_, int row0, int nrows, spear_diagset** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const DiagSet* f = reinterpret_cast<const DiagSet*>(full_);
    REQUIRE(f, "slice_rows: null set");
    *out = reinterpret_cast<spear_diagset*>(diagset_slice(c, f, row0, nrows, 0, f->n));
    API_END
}
// the phase-1 share of `rank` in a group of `world` (engine.h split_share): rows, and for groups of more than four ranks
// one half of the columns
int spear_diagset_slice_share(spear_context* ctx, const spear_diagset* full_, int rank, int world, spear_diagset** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const DiagSet* f = reinterpret_cast<const DiagSet*>(full_);
    REQUIRE(f && world >= 1 && world <= 8 && rank >= 0 && rank < world, "slice_share: rank %d of %d", rank, world);
    const SplitShare sh = split_share(rank, world, f->l + c->P, c->N);
    *out = reinterpret_cast<spear_diagset*>(diagset_slice(c, f, sh.row0, sh.nrows, sh.col0 >> f->rshift, sh.ncols >> f->rshift));
    API_END
}
int spear_split_share(int rank, int world, int rows, int N, int* row0, int* nrows, int* col0, int* ncols) {
    API_BEGIN
    REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world && rows >= 1 && N >= 2, "split_share: bad arguments");
    const SplitShare sh = split_share(rank, world, rows, N);
    *row0 = sh.row0, *nrows = sh.nrows, *col0 = sh.col0, *ncols = sh.ncols;
    API_END
}

static Obj* bsgs_partial(Ctx* c, const Obj* ct, const DiagSet* ds, const GaloisKeys* gk, cudaStream_t s = nullptr) {
    if (!s) s = c->stream;
    check_ct(ct, "bsgs_hoisted");
    REQUIRE(ds->nrows < 0 && ds->ncols < 0, "bsgs_hoisted: sliced diagonal set (use bsgs_split)");
    REQUIRE(ct->size == 2, "bsgs_hoisted: relinearize first");
    REQUIRE(ds->l == ct->l, "bsgs_hoisted: diagonals encoded for %d limbs, ciphertext has %d", ds->l, ct->l);
    const int G = std::min(ds->G, ds->D), B = ds->B, D = ds->D, l = ct->l;
    std::vector<u32> belt(G, 0), gelt;
    std::vector<const u64*> bkey(G, nullptr), gkey;
    for (int b = 1; b < G; b++) {
        belt[b] = (u32)elt_from_step(b, c->N);
        bkey[b] = find_key(gk, belt[b])->d;
    }
    for (int g = ds->g_first; g < B && g * ds->G < D; g += ds->g_stride) {
        gelt.push_back(g ? (u32)elt_from_step(g * ds->G, c->N) : 0);
        gkey.push_back(g ? find_key(gk, gelt.back())->d : nullptr);
    }
    std::unique_ptr<Obj> R(new_obj(c, 2, l, true, c->N, ct->scale * ds->scale, s));
    eng::bsgs_hoisted_partial(c, ct->d, l, ds->d, ds->rshift, G, (int)gelt.size(), ds->n_diags, ds->g_first,
                              ds->g_stride, belt.data(), bkey.data(), gelt.data(), gkey.data(), R->d, s);
    return R.release();
}
static Obj* bsgs_finish(Ctx* c, Obj* R, cudaStream_t s = nullptr) {
    if (!s) s = c->stream;
    REQUIRE(R && R->size == 2 && R->ext && R->n == c->N, "bsgs_finish: expected an accumulator in basis Q_l*P");
    REQUIRE(R->l >= 2, "bsgs: no level left for the final rescale");
    std::unique_ptr<Obj> o(new_obj(c, 2, R->l - 1, false, c->N, R->scale / (double)c->q[R->l - 1], s));
    eng::bsgs_finish(c, R->d, R->l, o->d, s);
    return o.release();
}

// Joins the auxiliary streams back into the main stream when it goes out of scope -- also when an item of a batch
// throws: the accumulators of the earlier items are then released on the main stream only after the kernels still
// running on the auxiliary streams are ordered before it.  Declare it AFTER the objects it protects.
struct AuxJoin {
    Ctx* c;
    int used;
    ~AuxJoin() {
        for (int k = 0; k < used; k++)
            if (cudaEventRecord(c->ev_aux[k], c->aux[k]) == cudaSuccess) cudaStreamWaitEvent(c->stream, c->ev_aux[k], 0);
    }
};

int spear_bsgs_hoisted(spear_context* ctx, const spear_obj* ct_, const spear_diagset* ds_, const spear_galois_keys* gk_,
                       spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    const DiagSet* ds = reinterpret_cast<const DiagSet*>(ds_);
    REQUIRE(ds->g_first == 0 && ds->g_stride == 1, "bsgs_hoisted: sharded diagonal set (use bsgs_hoisted_partial)");
    REQUIRE(O_(ct_)->l >= 2, "bsgs_hoisted: no level left for the final rescale");
    std::unique_ptr<Obj> R(bsgs_partial(c, O_(ct_), ds, reinterpret_cast<const GaloisKeys*>(gk_)));
    *out = H_(bsgs_finish(c, R.get()));
    API_END
}
int spear_bsgs_hoisted_batch(spear_context* ctx, spear_obj* const* cts, spear_diagset* const* dss, int count,
                             const spear_galois_keys* gk_, spear_obj** outs) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    REQUIRE(count >= 1, "bsgs_hoisted_batch: empty batch");
    const GaloisKeys* gk = reinterpret_cast<const GaloisKeys*>(gk_);
    for (int i = 0; i < count; i++) {   // validate everything before anything is queued
        const DiagSet* ds = reinterpret_cast<const DiagSet*>(dss[i]);
        REQUIRE(ds && cts[i], "bsgs_hoisted_batch: null item %d", i);
        REQUIRE(ds->g_first == 0 && ds->g_stride == 1, "bsgs_hoisted_batch: sharded diagonal set");
        REQUIRE(O_(cts[i])->l >= 2, "bsgs_hoisted_batch: no level left for the final rescale");
    }
    std::vector<std::unique_ptr<Obj>> acc(count), res(count);
    // independent mat-vecs on the auxiliary streams: the HBM-bound key streams of one overlap the
    // integer-bound NTT / MAC phases of the others
    CUDA_CHECK(cudaEventRecord(c->ev_main, c->stream));
    {
        AuxJoin join{c, count > 1 ? std::min(count, 3) : 0};
        for (int i = 0; i < count; i++) {
            cudaStream_t s = count == 1 ? c->stream : c->aux[i % 3];
            if (count > 1 && i < 3) CUDA_CHECK(cudaStreamWaitEvent(s, c->ev_main, 0));
            acc[i].reset(bsgs_partial(c, O_(cts[i]), reinterpret_cast<const DiagSet*>(dss[i]), gk, s));
            res[i].reset(bsgs_finish(c, acc[i].get(), s));
        }
    }
    for (int i = 0; i < count; i++) outs[i] = H_(res[i].release());
    API_END
}
// `count` diagonal sets (same D, G, level) multiplying ONE ciphertext -- the chunk pairs of a D -> F projection
// [ref: scripts/bootstrap_generation.py:575-600, where the baby rotations are computed once for all chunks]: one
// decomposition and one set of hoisted baby steps, then the diagonal MAC, the giant steps and the finish per set.
int spear_bsgs_hoisted_shared(spear_context* ctx, const spear_obj* ct_, spear_diagset* const* dss, int count,
                              const spear_galois_keys* gk_, spear_obj** outs) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    REQUIRE(count >= 1, "bsgs_hoisted_shared: empty batch");
    const Obj* ct = O_(ct_);
    const GaloisKeys* gk = reinterpret_cast<const GaloisKeys*>(gk_);
    check_ct(ct, "bsgs_hoisted_shared");
    REQUIRE(ct->size == 2 && ct->l >= 2, "bsgs_hoisted_shared: a fresh size-2 ciphertext with a level to spare expected");
    const DiagSet* d0 = reinterpret_cast<const DiagSet*>(dss[0]);
    REQUIRE(d0, "bsgs_hoisted_shared: null set");
    const int G = std::min(d0->G, d0->D), l = ct->l;
    std::vector<u32> belt(G, 0);
    std::vector<const u64*> bkey(G, nullptr);
    for (int b = 1; b < G; b++) {
        belt[b] = (u32)elt_from_step(b, c->N);
        bkey[b] = find_key(gk, belt[b])->d;
    }
    std::vector<std::vector<u32>> gelt(count);
    std::vector<std::vector<const u64*>> gkey(count);
    std::vector<SharedSet> sets(count);
    std::vector<std::unique_ptr<Obj>> acc(count), res(count);
    cudaStream_t s = c->stream;
    for (int i = 0; i < count; i++) {
        const DiagSet* ds = reinterpret_cast<const DiagSet*>(dss[i]);
        REQUIRE(ds && ds->g_first == 0 && ds->g_stride == 1 && ds->nrows < 0 && ds->ncols < 0, "bsgs_hoisted_shared: full sets expected");
        REQUIRE(ds->l == l && ds->G == d0->G && ds->D == d0->D, "bsgs_hoisted_shared: the sets must share D, G and the level");
        for (int g = 0; g < ds->B && g * ds->G < ds->D; g++) {
            gelt[i].push_back(g ? (u32)elt_from_step(g * ds->G, c->N) : 0);
            gkey[i].push_back(g ? find_key(gk, gelt[i].back())->d : nullptr);
        }
        acc[i].reset(new_obj(c, 2, l, true, c->N, ct->scale * ds->scale, s));
        sets[i] = SharedSet{ds->d, ds->rshift, (int)gelt[i].size(), ds->n_diags, gelt[i].data(), gkey[i].data(), acc[i]->d};
    }
    eng::bsgs_hoisted_shared(c, ct->d, l, G, belt.data(), bkey.data(), sets.data(), count, s);
    for (int i = 0; i < count; i++) res[i].reset(bsgs_finish(c, acc[i].get(), s));
    for (int i = 0; i < count; i++) outs[i] = H_(res[i].release());
    API_END
}
// Serving form of the batch: ciphertexts arrive in and leave to HOST memory.  Item i is uploaded, multiplied and
// downloaded on auxiliary stream i % 3, so the PCIe legs of one item run under the arithmetic of the others.
int spear_bsgs_hoisted_batch_host(spear_context* ctx, const uint64_t* const* in, int limbs, double scale,
                                  spear_diagset* const* dss, int count, const spear_galois_keys* gk_, uint64_t* const* out,
                                  double* out_scale) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    REQUIRE(count >= 1 && limbs >= 2 && limbs <= c->L, "bsgs_hoisted_batch_host: bad batch");
    const GaloisKeys* gk = reinterpret_cast<const GaloisKeys*>(gk_);
    for (int i = 0; i < count; i++) {   // validate everything before anything is queued
        const DiagSet* ds = reinterpret_cast<const DiagSet*>(dss[i]);
        REQUIRE(ds && in[i] && out[i], "bsgs_hoisted_batch_host: null item %d", i);
        REQUIRE(ds->g_first == 0 && ds->g_stride == 1 && ds->nrows < 0 && ds->ncols < 0, "bsgs_hoisted_batch_host: sharded diagonal set");
        REQUIRE(ds->l == limbs, "bsgs_hoisted_batch_host: diagonals encoded for %d limbs, ciphertexts have %d", ds->l, limbs);
    }
    std::vector<std::unique_ptr<Obj>> ct(count), acc(count), res(count);
    const int used = std::min(count, 3);
    CUDA_CHECK(cudaEventRecord(c->ev_main, c->stream));
    {
        AuxJoin join{c, used};
        for (int i = 0; i < count; i++) {
            cudaStream_t s = c->aux[i % 3];
            if (i < 3) CUDA_CHECK(cudaStreamWaitEvent(s, c->ev_main, 0));
            ct[i].reset(new_obj(c, 2, limbs, false, c->N, scale, s));
            CUDA_CHECK(cudaMemcpyAsync(ct[i]->d, in[i], sizeof(u64) * ct[i]->words(), cudaMemcpyHostToDevice, s));
            acc[i].reset(bsgs_partial(c, ct[i].get(), reinterpret_cast<const DiagSet*>(dss[i]), gk, s));
            res[i].reset(bsgs_finish(c, acc[i].get(), s));
            CUDA_CHECK(cudaMemcpyAsync(out[i], res[i]->d, sizeof(u64) * res[i]->words(), cudaMemcpyDeviceToHost, s));
            if (out_scale) out_scale[i] = res[i]->scale;
        }
    }
    CUDA_CHECK(cudaStreamSynchronize(c->stream));   // the auxiliary streams were joined into it: every result has landed
    API_END
}
int spear_bsgs_hoisted_partial(spear_context* ctx, const spear_obj* ct_, const spear_diagset* ds_,
                               const spear_galois_keys* gk_, spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    *out = H_(bsgs_partial(c, O_(ct_), reinterpret_cast<const DiagSet*>(ds_), reinterpret_cast<const GaloisKeys*>(gk_)));
    API_END
}
int spear_bsgs_hoisted_partial_batch(spear_context* ctx, spear_obj* const* cts, spear_diagset* const* dss, int count,
                                     const spear_galois_keys* gk_, spear_obj** outs) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    REQUIRE(count >= 1, "bsgs_hoisted_partial_batch: empty batch");
    const GaloisKeys* gk = reinterpret_cast<const GaloisKeys*>(gk_);
    for (int i = 0; i < count; i++) REQUIRE(dss[i] && cts[i], "bsgs_hoisted_partial_batch: null item %d", i);
    std::vector<std::unique_ptr<Obj>> acc(count);
    CUDA_CHECK(cudaEventRecord(c->ev_main, c->stream));
    {
        AuxJoin join{c, count > 1 ? std::min(count, 3) : 0};
        for (int i = 0; i < count; i++) {
            cudaStream_t s = count == 1 ? c->stream : c->aux[i % 3];
            if (count > 1 && i < 3) CUDA_CHECK(cudaStreamWaitEvent(s, c->ev_main, 0));
            acc[i].reset(bsgs_partial(c, O_(cts[i]), reinterpret_cast<const DiagSet*>(dss[i]), gk, s));
        }
    }
    for (int i = 0; i < count; i++) outs[i] = H_(acc[i].release());
    API_END
}
// ---- two-phase mat-vec over a rank group ---------------------------------------------------------------------------
// Phase 1 (baby steps + diagonal MAC) split by ROWS, phase 2 (giant steps) split by GIANT GROUP; in between every rank's
// MAC epilogue scatters the accumulators of group g into the window of rank g % world (csrc/peer.cu, bsgs.cu).
struct SplitPlan {
    int G, nB, l;
    std::vector<u32> belt;
    std::vector<const u64*> bkey;
    size_t slot_words(const Ctx* c, int world) const { return (size_t)((nB + world - 1) / world) * 2 * (l + c->P) * c->N; }
};
static SplitPlan split_plan(Ctx* c, const Obj* ct, const DiagSet* ds, const GaloisKeys* gk) {
    check_ct(ct, "bsgs_split");
    REQUIRE(ct->size == 2, "bsgs_split: relinearize first");
    REQUIRE(ds->l == ct->l, "bsgs_split: diagonals encoded for %d limbs, ciphertext has %d", ds->l, ct->l);
    REQUIRE(ds->g_first == 0 && ds->g_stride == 1, "bsgs_split: the set must hold every giant group (rows are sliced instead)");
    SplitPlan p;
    p.G = std::min(ds->G, ds->D), p.l = ct->l;
    p.nB = (ds->D + ds->G - 1) / ds->G;
    p.belt.assign(p.G, 0), p.bkey.assign(p.G, nullptr);
    for (int b = 1; b < p.G; b++) {
        p.belt[b] = (u32)elt_from_step(b, c->N);
        p.bkey[b] = find_key(gk, p.belt[b])->d;
    }
    return p;
}
// the giant groups rank serves: g = rank, rank + world, ... (Galois elements / keys; group 0: none)
static void split_groups(Ctx* c, const DiagSet* ds, const GaloisKeys* gk, int nB, int rank, int world, std::vector<u32>& gelt,
                         std::vector<const u64*>& gkey) {
    for (int g = rank; g < nB; g += world) {
        gelt.push_back(g ? (u32)elt_from_step(g * ds->G, c->N) : 0);
        gkey.push_back(g ? find_key(gk, gelt.back())->d : nullptr);
    }
}
static SplitShare split_check_rows(const Ctx* c, const DiagSet* ds, int l, int rank, int world) {
    const SplitShare sh = split_share(rank, world, l + c->P, c->N);
    REQUIRE(ds->row0 == sh.row0 && ds->stored_rows() == sh.nrows && ds->col0 == (sh.col0 >> ds->rshift) &&
                ds->stored_cols() == (sh.ncols >> ds->rshift),
            "bsgs_split: rank %d of %d serves rows [%d, %d) x columns [%d, %d); the set holds rows [%d, %d) x stored columns [%d, %d)",
            rank, world, sh.row0, sh.row0 + sh.nrows, sh.col0, sh.col0 + sh.ncols, ds->row0, ds->row0 + ds->stored_rows(), ds->col0,
            ds->col0 + ds->stored_cols());
    return sh;
}
static Obj* bsgs_split(Ctx* c, const Obj* ct, const DiagSet* ds, const GaloisKeys* gk, spear_peer_window* win, int slot,
                       cudaStream_t s) {
    SplitPlan p = split_plan(c, ct, ds, gk);
    int rank = 0, world = 1;
    peer::window_geometry(win, &rank, &world);
    const SplitShare sh = split_check_rows(c, ds, p.l, rank, world);
    std::vector<u32> gelt;
    std::vector<const u64*> gkey;
    split_groups(c, ds, gk, p.nB, rank, world, gelt, gkey);
    // everything that can throw comes before the first kernel a peer will wait for
    std::unique_ptr<Obj> R(new_obj(c, 2, p.l, true, c->N, ct->scale * ds->scale, s));
    const peer::SplitView v = peer::split_begin(c, win, slot, p.slot_words(c, world), s);
    PmacDst dst = {};
    for (int r = 0; r < 8; r++) dst.base[r] = v.base[r];
    dst.world = world;
    eng::bsgs_split_phase1(c, ct->d, p.l, ds->d, ds->rshift, p.G, p.nB, ds->n_diags, sh.row0, sh.nrows, sh.col0, sh.ncols,
                           p.belt.data(), p.bkey.data(), dst, s);
    peer::split_exchange(win, slot, s);
    eng::bsgs_split_phase2(c, v.base[rank], p.l, p.G, p.nB, (int)gelt.size(), gelt.data(), gkey.data(), world, R->d, s);
    peer::split_release(win, slot, R->d, R->words(), s);
    return R.release();
}
int spear_bsgs_split(spear_context* ctx, const spear_obj* ct_, const spear_diagset* ds_, const spear_galois_keys* gk_,
                     spear_peer_window* win, int slot, spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    *out = H_(bsgs_split(c, O_(ct_), reinterpret_cast<const DiagSet*>(ds_), reinterpret_cast<const GaloisKeys*>(gk_), win, slot,
                         c->stream));
    API_END
}
// `count` row-sliced sets multiplying ONE ciphertext over a rank group: this rank's share of the baby steps once, the
// diagonal MAC of set i scattered into window slot slot0 + i, then -- after every set's stores are posted -- the giant
// steps of this rank's groups per set.  outs[i]: this rank's accumulator of set i.
int spear_bsgs_split_shared(spear_context* ctx, const spear_obj* ct_, spear_diagset* const* dss, int count,
                            const spear_galois_keys* gk_, spear_peer_window* win, int slot0, spear_obj** outs) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    REQUIRE(count >= 1, "bsgs_split_shared: empty batch");
    const Obj* ct = O_(ct_);
    const GaloisKeys* gk = reinterpret_cast<const GaloisKeys*>(gk_);
    int rank = 0, world = 1;
    peer::window_geometry(win, &rank, &world);
    std::vector<SplitPlan> plans;
    std::vector<std::vector<u32>> gelt(count);
    std::vector<std::vector<const u64*>> gkey(count);
    for (int i = 0; i < count; i++) {   // validate everything before anything is queued
        const DiagSet* ds = reinterpret_cast<const DiagSet*>(dss[i]);
        REQUIRE(ds, "bsgs_split_shared: null set %d", i);
        plans.push_back(split_plan(c, ct, ds, gk));
        REQUIRE(plans[i].G == plans[0].G && ds->D == reinterpret_cast<const DiagSet*>(dss[0])->D,
                "bsgs_split_shared: the sets must share D and G");
        split_check_rows(c, ds, plans[i].l, rank, world);
        split_groups(c, ds, gk, plans[i].nB, rank, world, gelt[i], gkey[i]);
    }
    const SplitPlan& p0 = plans[0];
    const SplitShare sh = split_share(rank, world, p0.l + c->P, c->N);
    cudaStream_t s = c->stream;
    std::vector<std::unique_ptr<Obj>> R(count);
    for (int i = 0; i < count; i++)
        R[i].reset(new_obj(c, 2, p0.l, true, c->N, ct->scale * reinterpret_cast<const DiagSet*>(dss[i])->scale, s));
    std::vector<peer::SplitView> v(count);
    int maxB = 0;
    for (int i = 0; i < count; i++) {
        v[i] = peer::split_begin(c, win, slot0 + i, plans[i].slot_words(c, world), s);
        maxB = std::max(maxB, plans[i].nB);
    }
    eng::bsgs_split_baby(c, ct->d, p0.l, p0.G, maxB, sh.row0, sh.nrows, sh.col0, sh.ncols, p0.belt.data(), p0.bkey.data(), world, s);
    for (int i = 0; i < count; i++) {
        const DiagSet* ds = reinterpret_cast<const DiagSet*>(dss[i]);
        PmacDst dst = {};
        for (int r = 0; r < 8; r++) dst.base[r] = v[i].base[r];
        dst.world = world;
        eng::bsgs_split_mac(c, p0.l, ds->d, ds->rshift, p0.G, plans[i].nB, ds->n_diags, sh.row0, sh.nrows, sh.col0, sh.ncols, dst, s);
        peer::split_post(win, slot0 + i, s);
    }
    for (int i = 0; i < count; i++) {
        peer::split_wait(win, slot0 + i, s);
        eng::bsgs_split_phase2(c, v[i].base[rank], p0.l, p0.G, plans[i].nB, (int)gelt[i].size(), gelt[i].data(), gkey[i].data(), world,
                               R[i]->d, s);
        peer::split_release(win, slot0 + i, R[i]->d, R[i]->words(), s);
    }
    for (int i = 0; i < count; i++) outs[i] = H_(R[i].release());
    API_END
}
// item i runs on auxiliary stream i % 3 and exchanges through window slot slot0 + i
int spear_bsgs_split_batch(spear_context* ctx, spear_obj* const* cts, spear_diagset* const* dss, int count,
                           const spear_galois_keys* gk_, spear_peer_window* win, int slot0, spear_obj** outs) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    REQUIRE(count >= 1, "bsgs_split_batch: empty batch");
    const GaloisKeys* gk = reinterpret_cast<const GaloisKeys*>(gk_);
    for (int i = 0; i < count; i++) {   // validate everything before anything is queued
        REQUIRE(dss[i] && cts[i], "bsgs_split_batch: null item %d", i);
        split_plan(c, O_(cts[i]), reinterpret_cast<const DiagSet*>(dss[i]), gk);
    }
    std::vector<std::unique_ptr<Obj>> acc(count);
    CUDA_CHECK(cudaEventRecord(c->ev_main, c->stream));
    {
        AuxJoin join{c, count > 1 ? std::min(count, 3) : 0};
        for (int i = 0; i < count; i++) {
            cudaStream_t s = count == 1 ? c->stream : c->aux[i % 3];
            if (count > 1 && i < 3) CUDA_CHECK(cudaStreamWaitEvent(s, c->ev_main, 0));
            acc[i].reset(bsgs_split(c, O_(cts[i]), reinterpret_cast<const DiagSet*>(dss[i]), gk, win, slot0 + i, s));
        }
    }
    for (int i = 0; i < count; i++) outs[i] = H_(acc[i].release());
    API_END
}
// Test hook (one GPU, one process): the two phases of all `world` ranks run one after the other over local stand-ins
// for the exchange windows; dss[r] = the row slice of rank r.  *out = the summed accumulator (bsgs_finish completes it).
int spear_bsgs_split_selftest(spear_context* ctx, const spear_obj* ct_, spear_diagset* const* dss, int world,
                              const spear_galois_keys* gk_, spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    REQUIRE(world >= 1 && world <= 8, "bsgs_split_selftest: 1..8 ranks");
    const Obj* ct = O_(ct_);
    const GaloisKeys* gk = reinterpret_cast<const GaloisKeys*>(gk_);
    cudaStream_t s = c->stream;
    SplitPlan p = split_plan(c, ct, reinterpret_cast<const DiagSet*>(dss[0]), gk);
    const size_t words = std::max<size_t>(p.slot_words(c, world), 32);
    PmacDst dst = {};
    dst.world = world;
    std::vector<u64*> win(world);
    for (int r = 0; r < world; r++) win[r] = c->alloc(words, s);
    for (int r = 0; r < 8; r++) dst.base[r] = win[r < world ? r : 0];
    for (int r = 0; r < world; r++) {
        const DiagSet* ds = reinterpret_cast<const DiagSet*>(dss[r]);
        split_plan(c, ct, ds, gk);
        const SplitShare sh = split_check_rows(c, ds, p.l, r, world);
        eng::bsgs_split_phase1(c, ct->d, p.l, ds->d, ds->rshift, p.G, p.nB, ds->n_diags, sh.row0, sh.nrows, sh.col0, sh.ncols,
                               p.belt.data(), p.bkey.data(), dst, s);
    }
    std::unique_ptr<Obj> sum;
    for (int r = 0; r < world; r++) {
        const DiagSet* ds = reinterpret_cast<const DiagSet*>(dss[r]);
        std::vector<u32> gelt;
        std::vector<const u64*> gkey;
        split_groups(c, ds, gk, p.nB, r, world, gelt, gkey);
        std::unique_ptr<Obj> R(new_obj(c, 2, p.l, true, c->N, ct->scale * ds->scale, s));
        eng::bsgs_split_phase2(c, win[r], p.l, p.G, p.nB, (int)gelt.size(), gelt.data(), gkey.data(), world, R->d, s);
        if (!sum) sum = std::move(R);
        else ops::add(c, sum->d, R->d, sum->d, 2, sum->rows(), c->N, RowMap{sum->rows(), sum->l, c->L, 0}, 2, s);
    }
    for (int r = 0; r < world; r++) c->free(win[r], s);
    *out = H_(sum.release());
    API_END
}

int spear_bsgs_finish(spear_context* ctx, spear_obj* acc, spear_obj** out) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    *out = H_(bsgs_finish(c, reinterpret_cast<Obj*>(acc)));
    API_END
}
int spear_obj_reduce(spear_context* ctx, spear_obj* o_) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    Obj* o = reinterpret_cast<Obj*>(o_);
    REQUIRE(o && o->n == c->N, "reduce: bad operand");
    ops::reduce_inplace(c, o->d, o->size, o->rows(), RowMap{o->rows(), o->l, c->L, 0}, c->stream);
    API_END
}
void* spear_obj_device_ptr(spear_obj* o) { return reinterpret_cast<Obj*>(o)->d; }

int spear_ntt_host(spear_context* ctx, uint64_t* data, int rows, const int* limb_ids, int ring_n, int inverse) {
    API_BEGIN
    Ctx* c = C_(ctx);
    use(c);
    u64* d = c->alloc((size_t)rows * ring_n);
    CUDA_CHECK(cudaMemcpyAsync(d, data, sizeof(u64) * rows * ring_n, cudaMemcpyHostToDevice, c->stream));
    for (int r = 0; r < rows; r++) {
        REQUIRE(limb_ids[r] >= 0 && limb_ids[r] < c->K, "ntt: limb id out of range");
        RowMap rm{1, 0, limb_ids[r], 0};
        if (inverse) ntt_inverse(c, d + (size_t)r * ring_n, 1, rm, ring_n, c->stream);
        else ntt_forward(c, d + (size_t)r * ring_n, 1, rm, ring_n, c->stream);
    }
    CUDA_CHECK(cudaMemcpyAsync(data, d, sizeof(u64) * rows * ring_n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->free(d);
    API_END
}

}  // extern "C"
