// chacha.cuh -- ChaCha20 counter-mode block function on the device, shared by the samplers (sampler.cu) and the fused
// client kernels (client.cu).  Streams, word layout and distributions are those of the oracle (oracle/spear_oracle.c
// chacha_block / sample_uniform_row / sample_ternary / sample_cbd).
#pragma once
#include "common.cuh"

struct Seed {
    u32 k[8];
};

static __device__ __forceinline__ u32 chacha_rotl(u32 v, int n) { return (v << n) | (v >> (32 - n)); }
#define CHACHA_QR(a, b, c, d)                                                          \
    a += b; d ^= a; d = chacha_rotl(d, 16); c += d; b ^= c; b = chacha_rotl(b, 12);   \
    a += b; d ^= a; d = chacha_rotl(d, 8);  c += d; b ^= c; b = chacha_rotl(b, 7);

// key = 32-byte seed, 64-bit nonce = stream id, 64-bit block counter; out: the block as eight 64-bit words
static __device__ void chacha_block(const Seed& key, u64 nonce, u64 counter, u64 out[8]) {
    u32 s[16], x[16];
    s[0] = 0x61707865u, s[1] = 0x3320646eu, s[2] = 0x79622d32u, s[3] = 0x6b206574u;
#pragma unroll
    for (int i = 0; i < 8; i++) s[4 + i] = key.k[i];
    s[12] = (u32)counter, s[13] = (u32)(counter >> 32);
    s[14] = (u32)nonce, s[15] = (u32)(nonce >> 32);
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = s[i];
#pragma unroll
    for (int r = 0; r < 10; r++) {
        CHACHA_QR(x[0], x[4], x[8], x[12]) CHACHA_QR(x[1], x[5], x[9], x[13])
        CHACHA_QR(x[2], x[6], x[10], x[14]) CHACHA_QR(x[3], x[7], x[11], x[15])
        CHACHA_QR(x[0], x[5], x[10], x[15]) CHACHA_QR(x[1], x[6], x[11], x[12])
        CHACHA_QR(x[2], x[7], x[8], x[13]) CHACHA_QR(x[3], x[4], x[9], x[14])
    }
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = (u64)(x[2 * i] + s[2 * i]) | ((u64)(x[2 * i + 1] + s[2 * i + 1]) << 32);
}
// centred binomial error (21 + 21 bits of a word) and ternary value of a word, as the oracle's sample_cbd / sample_ternary
static __device__ __forceinline__ int cbd_of_word(u64 w) { return __popcll(w & 0x1FFFFFull) - __popcll((w >> 21) & 0x1FFFFFull); }
static __device__ __forceinline__ int ternary_of_word(u64 w) { return (int)__umul64hi(w, 3ull) - 1; }

static inline Seed make_seed(const u32* s) {
    Seed k;
    for (int i = 0; i < 8; i++) k.k[i] = s[i];
    return k;
}
