// context.cu -- parameter set, prime/root search and device tables (host code).
// Replaces what phantom::EncryptionParameters / PhantomContext provide behind
// gpu/phantom_binding.cu:81-98 of the reference (create_coeff_modulus, params, context).
#include <cmath>
#include <cstdarg>
#include <algorithm>
#include <cstring>

#include "engine.h"

typedef unsigned __int128 u128;

void spear_throw(int code, const char* fmt, ...) {
    spear_error e;
    e.code = code;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(e.msg, sizeof e.msg, fmt, ap);
    va_end(ap);
    throw e;
}

namespace host {

u64 mulm(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }
u64 powm(u64 a, u64 e, u64 q) {
    u64 r = 1;
    for (a %= q; e; e >>= 1, a = mulm(a, a, q))
        if (e & 1) r = mulm(r, a, q);
    return r;
}
u64 invm(u64 a, u64 q) { return powm(a, q - 2, q); }
u64 shoup(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
ulonglong2 with_shoup(u64 w, u64 q) { return make_ulonglong2(w, shoup(w, q)); }

// deterministic Miller-Rabin for 64-bit integers
bool is_prime(u64 n) {
    if (n < 4) return n == 2 || n == 3;
    if (!(n & 1)) return false;
    u64 d = n - 1;
    int s = 0;
    while (!(d & 1)) d >>= 1, s++;
    for (u64 a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        if (a % n == 0) continue;
        u64 x = powm(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool witness = true;
        for (int r = 1; r < s && witness; r++) {
            x = mulm(x, x, n);
            if (x == n - 1) witness = false;
        }
        if (witness) return false;
    }
    return true;
}

// NTT-friendly primes of exactly `bits` bits, largest first (p = 1 mod 2N)
std::vector<u64> primes_below(u64 N, int bits, int count) {
    std::vector<u64> out;
    const u64 step = 2 * N, floor_ = 1ull << (bits - 1);
    for (u64 p = ((1ull << bits) - 1) / step * step + 1; p > floor_ && (int)out.size() < count; p -= step)
        if (is_prime(p)) out.push_back(p);
    REQUIRE((int)out.size() == count, "not enough %d-bit primes for N=%llu", bits, (unsigned long long)N);
    return out;
}

// smallest primitive 2N-th root of unity
u64 min_root(u64 N, u64 q) {
    u64 root = 0;
    for (u64 g = 2; !root; g++) {
        u64 c = powm(g, (q - 1) / (2 * N), q);
        if (powm(c, N, q) == q - 1) root = c;
    }
    u64 step = mulm(root, root, q), best = root;
    for (u64 k = 1, cur = root; k < N; k++) {
        cur = mulm(cur, step, q);
        best = cur < best ? cur : best;
    }
    return best;
}

u32 brev(u32 x, int bits) {
    u32 r = 0;
    for (int i = 0; i < bits; i++, x >>= 1) r = (r << 1) | (x & 1);
    return r;
}

template <class T>
T* upload(const std::vector<T>& v) {
    T* d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, sizeof(T) * (v.size() ? v.size() : 1)));
    CUDA_CHECK(cudaMemcpy(d, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return d;
}

}  // namespace host

using namespace host;

int create_coeff_modulus(u64 N, const int* bits, int n, u64* out) {
    std::map<int, int> want;
    for (int i = 0; i < n; i++) {
        REQUIRE(bits[i] >= 20 && bits[i] <= 60, "prime size %d outside [20,60]", bits[i]);
        want[bits[i]]++;
    }
    std::map<int, std::vector<u64>> pool;
    for (auto& kv : want) pool[kv.first] = primes_below(N, kv.first, kv.second);
    // as SEAL's CoeffModulus::Create: each request takes the smallest prime left of its size
    for (int i = 0; i < n; i++) {
        out[i] = pool[bits[i]].back();
        pool[bits[i]].pop_back();
    }
    return 0;
}

u64* Ctx::alloc(size_t n_u64, cudaStream_t s) const {
    void* p = nullptr;
    CUDA_CHECK(cudaMallocAsync(&p, sizeof(u64) * (n_u64 ? n_u64 : 1), s ? s : stream));
    return (u64*)p;
}
u64* Ctx::workspace(cudaStream_t s, size_t words) const {
    int slot = 0;
    for (int i = 0; i < 3; i++)
        if (s == aux[i]) slot = i + 1;
    if (ws_cap[slot] < words) {   // rare: first call, or a larger problem than any before on this stream
        cudaStream_t st = slot ? aux[slot - 1] : stream;
        CUDA_CHECK(cudaStreamSynchronize(st));
        if (ws_base[slot]) CUDA_CHECK(cudaFree(ws_base[slot]));
        ws_base[slot] = nullptr, ws_cap[slot] = 0;
        void* p = nullptr;
        CUDA_CHECK(cudaMalloc(&p, sizeof(u64) * words));
        ws_base[slot] = (u64*)p, ws_cap[slot] = words;
    }
    return ws_base[slot];
}
void* Ctx::staging(size_t bytes) const {
    if (!staged) CUDA_CHECK(cudaEventCreateWithFlags(&staged, cudaEventDisableTiming));
    else CUDA_CHECK(cudaEventSynchronize(staged));   // the previous upload has left the buffer
    if (stage_cap < bytes) {
        if (stage_base) CUDA_CHECK(cudaFreeHost(stage_base));
        stage_base = nullptr, stage_cap = 0;
        CUDA_CHECK(cudaMallocHost(&stage_base, bytes));
        stage_cap = bytes;
    }
    return stage_base;
}
void Ctx::free(void* p, cudaStream_t s) const {
    if (p) cudaFreeAsync(p, s ? s : stream);
}

Ctx* ctx_create(u64 N, const u64* moduli, int K, int P, int device) {
    REQUIRE(N >= 8 && N <= 65536 && (N & (N - 1)) == 0, "poly_modulus_degree must be a power of two in [8, 65536]");
    REQUIRE(K >= 2 && K <= SPEAR_MAX_LIMBS, "coeff_modulus size must be in [2, %d]", SPEAR_MAX_LIMBS);
    REQUIRE(P >= 1 && P < K, "special_modulus_size must be in [1, %d)", K);
    std::unique_ptr<Ctx> c(new Ctx);
    c->N = (int)N;
    while ((1u << c->logn) < N) c->logn++;
    c->K = K, c->P = P, c->L = K - P;
    c->beta = c->digits(c->L);
    c->device = device;
    c->q.assign(moduli, moduli + K);
    for (int i = 0; i < K; i++) {
        REQUIRE(c->q[i] < (1ull << 60) && (c->q[i] - 1) % (2 * N) == 0 && is_prime(c->q[i]),
                "modulus %d is not an NTT-friendly prime below 2^60", i);
        for (int j = 0; j < i; j++) REQUIRE(c->q[i] != c->q[j], "duplicate modulus");
    }
    CUDA_CHECK(cudaSetDevice(device));
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    for (int i = 0; i < 3; i++) {
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_aux[i], cudaEventDisableTiming));
    }
    CUDA_CHECK(cudaDeviceGetDefaultMemPool(&c->pool, device));
    unsigned long long keep = ~0ull;   // cache freed blocks: ~4k temporaries per mat-vec in the op-by-op path
    CUDA_CHECK(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep));
    CUDA_CHECK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));

    const int L = c->L, beta = c->beta;
    std::vector<u64> r0(K), r1(K), rw(K);
    std::vector<ulonglong2> psi((size_t)K * N), ipsi((size_t)K * N), invn((size_t)K * 17);
    for (int i = 0; i < K; i++) {
        u64 q = c->q[i];
        u128 ratio = ~(u128)0 / q;
        r0[i] = (u64)ratio, r1[i] = (u64)(ratio >> 64);
        {
            int sbits = 63 - __builtin_clzll(q);
            rw[i] = (u64)((((u128)1 << (sbits + 64)) - 1) / q);   // q odd, so this equals floor(2^(s+64)/q)
        }
        u64 w = min_root(N, q), iw = invm(w, q), p = 1, ip = 1;
        for (u64 k = 0; k < N; k++) {
            u32 r = brev((u32)k, c->logn);
            psi[(size_t)i * N + r] = with_shoup(p, q);
            ipsi[(size_t)i * N + r] = with_shoup(ip, q);
            p = mulm(p, w, q), ip = mulm(ip, iw, q);
        }
        for (int k = 0; k <= 16; k++) invn[i * 17 + k] = with_shoup(invm((1ull << k) % q, q), q);
    }
    c->d_q = upload(c->q), c->d_ratio0 = upload(r0), c->d_ratio1 = upload(r1), c->d_rwide = upload(rw);
    {
        auto bits = [](u64 v) { int b = 0; while (v) b++, v >>= 1; return b; };
        c->sbits = bits(c->q[0]) - 1;
        for (u64 qq : c->q) if (bits(qq) - 1 != c->sbits) c->sbits = 0;
    }
    c->d_psi = upload(psi), c->d_ipsi = upload(ipsi), c->d_invn = upload(invn);

    // P mod q_i and its inverse
    std::vector<ulonglong2> pmod(K), pinv(K);
    for (int i = 0; i < K; i++) {
        u64 q = c->q[i], pm = 1;
        for (int k = 0; k < P; k++) pm = mulm(pm, c->q[L + k] % q, q);
        pmod[i] = with_shoup(pm, q);
        pinv[i] = i < L ? with_shoup(invm(pm, q), q) : make_ulonglong2(0, 0);
    }
    c->d_pmod = upload(pmod), c->d_pinv = upload(pinv);

    // ModUp: level l in 1..L, digit j: limbs [jP, min((j+1)P, l))
    std::vector<ulonglong2> up_hatinv((size_t)(L + 1) * beta * P, make_ulonglong2(0, 0));
    std::vector<ulonglong2> up_hatinv_n((size_t)(L + 1) * beta * P, make_ulonglong2(0, 0));
    std::vector<u64> up_hat((size_t)(L + 1) * beta * P * K, 0);
    for (int l = 1; l <= L; l++)
        for (int j = 0; j < c->digits(l); j++) {
            int lo = j * P, hi = std::min((j + 1) * P, l);
            for (int a = lo; a < hi; a++) {
                size_t e = ((size_t)l * beta + j) * P + (a - lo);
                u64 qa = c->q[a], h = 1;
                for (int b = lo; b < hi; b++)
                    if (b != a) h = mulm(h, c->q[b] % qa, qa);
                up_hatinv[e] = with_shoup(invm(h, qa), qa);
                up_hatinv_n[e] = with_shoup(mulm(invm(h, qa), invm(N % qa, qa), qa), qa);
                for (int t = 0; t < K; t++) {
                    u64 qt = c->q[t], ht = 1;
                    for (int b = lo; b < hi; b++)
                        if (b != a) ht = mulm(ht, c->q[b] % qt, qt);
                    up_hat[e * K + t] = ((ht >> 30) << 32) | (ht & 0x3FFFFFFFull);   // split-30 (common.cuh)
                }
            }
        }
    c->d_up_hatinv = upload(up_hatinv), c->d_up_hat = upload(up_hat), c->d_up_hatinv_n = upload(up_hatinv_n);

    // ModDown: from the special primes to every data limb, with the rounding constant floor(P/2)
    auto half_mod = [&](u64 m) {   // ((P mod 2m) - 1) / 2 mod m, P = prod of special primes (odd)
        u128 m2 = (u128)2 * m, pm = 1;
        for (int k = 0; k < P; k++) pm = pm * (c->q[L + k] % m2) % m2;
        return (u64)(((pm - 1) >> 1) % m);
    };
    std::vector<ulonglong2> dn_hatinv(P);
    std::vector<u64> dn_half(K), dn_hat((size_t)P * K);
    for (int k = 0; k < P; k++) {
        u64 pk = c->q[L + k], h = 1;
        for (int k2 = 0; k2 < P; k2++)
            if (k2 != k) h = mulm(h, c->q[L + k2] % pk, pk);
        dn_hatinv[k] = with_shoup(invm(h, pk), pk);
        for (int i = 0; i < K; i++) {
            u64 qi = c->q[i], hi = 1;
            for (int k2 = 0; k2 < P; k2++)
                if (k2 != k) hi = mulm(hi, c->q[L + k2] % qi, qi);
            dn_hat[(size_t)k * K + i] = ((hi >> 30) << 32) | (hi & 0x3FFFFFFFull);   // split-30
        }
    }
    for (int i = 0; i < K; i++) dn_half[i] = half_mod(c->q[i]);
    c->d_dn_hatinv = upload(dn_hatinv), c->d_dn_half = upload(dn_half), c->d_dn_hat = upload(dn_hat);

    // rescale: q_last^-1 mod q_i
    std::vector<ulonglong2> rs((size_t)K * K, make_ulonglong2(0, 0));
    for (int last = 0; last < K; last++)
        for (int i = 0; i < K; i++)
            if (i != last) rs[(size_t)last * K + i] = with_shoup(invm(c->q[last] % c->q[i], c->q[i]), c->q[i]);
    c->d_rs_inv = upload(rs);

    // decode: Garner inverses for the first (up to) three limbs
    std::vector<ulonglong2> gar(9, make_ulonglong2(0, 0));
    for (int i = 0; i < 3 && i < K; i++)
        for (int j = 0; j < i; j++) gar[i * 3 + j] = with_shoup(invm(c->q[j] % c->q[i], c->q[i]), c->q[i]);
    c->d_garner = upload(gar);

    // encoder roots zeta^{bitrev(k)}, zeta = exp(i*pi/N)  (same libm calls as the oracle)
    std::vector<double2> zeta(N);
    for (u64 k = 0; k < N; k++) {
        double ang = M_PI * (double)k / (double)N;
        zeta[brev((u32)k, c->logn)] = make_double2(cos(ang), sin(ang));
    }
    c->d_zeta = upload(zeta);
    // slot j sits at position bitrev((5^j mod 2N - 1) / 2), its conjugate at bitrev((2N - 5^j mod 2N - 1) / 2)
    std::vector<u32> pos_slot(N);
    {
        const u64 m = 2 * N;
        u64 pos = 1;
        for (u64 j = 0; j < N / 2; j++) {
            pos_slot[brev((u32)((pos - 1) >> 1), c->logn)] = (u32)j;
            pos_slot[brev((u32)((m - pos - 1) >> 1), c->logn)] = (u32)j | 0x80000000u;
            pos = pos * 5 % m;
        }
    }
    c->d_pos_slot = upload(pos_slot);
    return c.release();
}

void ctx_retain(Ctx* c) { c->refs.fetch_add(1); }
void ctx_destroy(Ctx* c);
void ctx_release(Ctx* c) {
    if (c && c->refs.fetch_sub(1) == 1) ctx_destroy(c);
}

void ctx_destroy(Ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (void* p : {(void*)c->d_q, (void*)c->d_ratio0, (void*)c->d_ratio1, (void*)c->d_rwide, (void*)c->d_psi, (void*)c->d_ipsi,
                    (void*)c->d_invn, (void*)c->d_pmod, (void*)c->d_pinv, (void*)c->d_up_hatinv, (void*)c->d_up_hat, (void*)c->d_up_hatinv_n,
                    (void*)c->d_dn_hatinv, (void*)c->d_dn_half, (void*)c->d_dn_hat, (void*)c->d_rs_inv,
                    (void*)c->d_garner, (void*)c->d_zeta, (void*)c->d_pos_slot})
        cudaFree(p);
    for (int i = 0; i < 4; i++) cudaFree(c->ws_base[i]);
    if (c->stage_base) cudaFreeHost(c->stage_base);
    if (c->staged) cudaEventDestroy(c->staged);
    for (int i = 0; i < 3; i++) {
        cudaStreamDestroy(c->aux[i]);
        cudaEventDestroy(c->ev_aux[i]);
    }
    cudaEventDestroy(c->ev_main);
    cudaStreamDestroy(c->stream);
    delete c;
}
