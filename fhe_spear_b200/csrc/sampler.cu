// sampler.cu -- ChaCha20 counter-mode sampling on the device and the key / encryption combines.
// Streams, word layout and distributions are those of the oracle (oracle/spear_oracle.c
// chacha_block / sample_uniform_row / sample_ternary / sample_cbd), so keys and ciphertexts
// generated here are bit-identical to the oracle's for the same 32-byte seed.
// Client-side surface replaced: PhantomSecretKey::{gen_*, encrypt_symmetric, decrypt},
// PhantomPublicKey::encrypt_asymmetric (reference gpu/phantom_binding.cu:100-116).
#include "chacha.cuh"
#include "engine.h"
#include "ops.h"

namespace {

constexpr int TPB = 256;

// 4 coefficients per thread: coefficient j of limb t uses words 2*(t*N + j), +1 as (hi, lo) of a 128-bit value
__global__ void __launch_bounds__(TPB) k_uniform(Seed seed, u64 nonce, u64* __restrict__ out, int N, RowMap rm, ModTab mt) {
    const int row = blockIdx.y, j = (blockIdx.x * TPB + threadIdx.x) * 4;
    if (j >= N) return;
    const int t = rm.limb(row);
    u64 blk[8];
    chacha_block(seed, nonce, ((u64)t * N + j) >> 2, blk);
    const u64 q = mt.q[t], r0 = mt.ratio0[t], r1 = mt.ratio1[t];
    u64* o = out + (size_t)row * N + j;
#pragma unroll
    for (int k = 0; k < 4; k++) o[k] = barrett128(blk[2 * k + 1], blk[2 * k], q, r0, r1);
}

// 8 coefficients per thread, one value for all rows. KIND 0: ternary, 1: centred binomial (21 + 21 bits)
template <int KIND>
__global__ void __launch_bounds__(TPB) k_small(Seed seed, u64 nonce, u64* __restrict__ out, int N, int nrows, RowMap rm,
                                                ModTab mt) {
    const int j = (blockIdx.x * TPB + threadIdx.x) * 8;
    if (j >= N) return;
    u64 blk[8];
    chacha_block(seed, nonce, (u64)j >> 3, blk);
    int v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        u64 w = blk[k];
        v[k] = KIND == 0 ? ternary_of_word(w) : cbd_of_word(w);
    }
    for (int r = 0; r < nrows; r++) {
        const u64 q = mt.q[rm.limb(r)];
        u64* o = out + (size_t)r * N + j;
#pragma unroll
        for (int k = 0; k < 8; k++) o[k] = v[k] >= 0 ? (u64)v[k] : q - (u64)(-v[k]);
    }
}

__global__ void k_ksk_combine(const u64* __restrict__ a, const u64* __restrict__ e, const u64* __restrict__ sk,
                              const u64* __restrict__ snew, u64* __restrict__ k0, int K, int N, int L, int P, int digit,
                              ModTab mt, const ulonglong2* __restrict__ pmod) {
    size_t total = (size_t)K * N;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
        int t = (int)(x / N);
        u64 q = mt.q[t], r0 = mt.ratio0[t], r1 = mt.ratio1[t];
        u64 v = sub_mod(e[x], mul_mod(a[x], sk[x], q, r0, r1), q);
        if (t < L && t / P == digit) {
            ulonglong2 pm = pmod[t];
            v = add_mod(v, mul_shoup(snew[x], pm.x, pm.y, q), q);
        }
        k0[x] = v;
    }
}

__global__ void k_enc_combine(const u64* __restrict__ a, const u64* __restrict__ e, const u64* __restrict__ sk,
                              const u64* __restrict__ m, u64* __restrict__ c0, int l, int N, ModTab mt) {
    size_t total = (size_t)l * N;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
        int t = (int)(x / N);
        u64 q = mt.q[t];
        u64 v = add_mod(m ? m[x] : 0, e[x], q);
        c0[x] = sub_mod(v, mul_mod(a[x], sk[x], q, mt.ratio0[t], mt.ratio1[t]), q);
    }
}

__global__ void k_dec_combine(const u64* __restrict__ ct, int size, int l, int N, const u64* __restrict__ sk,
                              u64* __restrict__ pt, ModTab mt) {
    size_t total = (size_t)l * N;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
        int t = (int)(x / N);
        u64 q = mt.q[t], r0 = mt.ratio0[t], r1 = mt.ratio1[t], s = sk[x];
        u64 acc = ct[(size_t)(size - 1) * total + x];
        for (int p = size - 2; p >= 0; p--) acc = add_mod(mul_mod(acc, s, q, r0, r1), ct[(size_t)p * total + x], q);
        pt[x] = acc;
    }
}

__global__ void k_asym_combine(const u64* __restrict__ pk, const u64* __restrict__ u, const u64* __restrict__ e0,
                               const u64* __restrict__ e1, u64* __restrict__ out, int l, int rows, int N, int L, int K,
                               ModTab mt) {
    size_t total = (size_t)rows * N;
    for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
        int r = (int)(x / N), n = (int)(x % N);
        int t = r < l ? r : L + (r - l);
        u64 q = mt.q[t], r0 = mt.ratio0[t], r1 = mt.ratio1[t];
        u64 uu = u[x];
        out[x] = add_mod(mul_mod(pk[(size_t)t * N + n], uu, q, r0, r1), e0[x], q);
        out[total + x] = add_mod(mul_mod(pk[((size_t)K + t) * N + n], uu, q, r0, r1), e1[x], q);
    }
}

Seed mk(const u32* s) { return make_seed(s); }
int grid_for(const Ctx* c, size_t total) {
    size_t blocks = (total + TPB - 1) / TPB, cap = (size_t)c->sm_count * 16;
    return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

}  // namespace

namespace sampler {

void uniform(const Ctx* c, const u32* seed, u64 nonce, u64* out, int nrows, RowMap rm, cudaStream_t s) {
    int threads = c->N / 4;
    LAUNCH(k_uniform, dim3((threads + TPB - 1) / TPB, nrows), TPB, 0, s)(mk(seed), nonce, out, c->N, rm, c->modtab());
    CUDA_CHECK(cudaGetLastError());
}
void ternary(const Ctx* c, const u32* seed, u64 nonce, u64* out, int nrows, RowMap rm, cudaStream_t s) {
    int threads = c->N / 8;
    LAUNCH(k_small<0>, (threads + TPB - 1) / TPB, TPB, 0, s)(mk(seed), nonce, out, c->N, nrows, rm, c->modtab());
    CUDA_CHECK(cudaGetLastError());
}
void cbd(const Ctx* c, const u32* seed, u64 nonce, u64* out, int nrows, RowMap rm, cudaStream_t s) {
    int threads = c->N / 8;
    LAUNCH(k_small<1>, (threads + TPB - 1) / TPB, TPB, 0, s)(mk(seed), nonce, out, c->N, nrows, rm, c->modtab());
    CUDA_CHECK(cudaGetLastError());
}
void ksk_combine(const Ctx* c, const u64* a, const u64* e, const u64* sk, const u64* snew, int digit, u64* k0,
                 cudaStream_t s) {
    LAUNCH(k_ksk_combine, grid_for(c, (size_t)c->K * c->N), TPB, 0, s)(a, e, sk, snew, k0, c->K, c->N, c->L, c->P, digit,
                                                                   c->modtab(), c->d_pmod);
    CUDA_CHECK(cudaGetLastError());
}
void enc_combine(const Ctx* c, const u64* a, const u64* e, const u64* sk, const u64* m, u64* c0, int l, cudaStream_t s) {
    LAUNCH(k_enc_combine, grid_for(c, (size_t)l * c->N), TPB, 0, s)(a, e, sk, m, c0, l, c->N, c->modtab());
    CUDA_CHECK(cudaGetLastError());
}
void dec_combine(const Ctx* c, const u64* ct, int size, int l, const u64* sk, u64* pt, cudaStream_t s) {
    LAUNCH(k_dec_combine, grid_for(c, (size_t)l * c->N), TPB, 0, s)(ct, size, l, c->N, sk, pt, c->modtab());
    CUDA_CHECK(cudaGetLastError());
}
void asym_combine(const Ctx* c, const u64* pk, const u64* u, const u64* e0, const u64* e1, u64* out, int l,
                  cudaStream_t s) {
    int rows = l + c->P;
    LAUNCH(k_asym_combine, grid_for(c, (size_t)rows * c->N), TPB, 0, s)(pk, u, e0, e1, out, l, rows, c->N, c->L, c->K,
                                                                    c->modtab());
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace sampler
