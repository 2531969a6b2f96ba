/*
 * spear_b200.h -- C ABI of libspear_b200.so, the sm_100a CKKS BSGS diagonal mat-vec engine.
 *
 * This is the drop-in boundary for the hot path of mozendr/FHE-SPEAR: it replaces what the
 * reference binds through pybind11 in gpu/phantom_binding.cu (module `pyPhantom`) plus the
 * fork-only symbols its scripts call.  Each entry point cites the reference
 * interface it stands in for as  [ref: file:line].  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returning int returns 0 on success, non-zero on failure; the message is
 *     available from spear_last_error() (thread-local).  Out-of-memory messages contain both
 *     "CUDA" and "out of memory" (the reference string-matches them, scripts/bootstrap_generation.py:1164).
 *   - handles are opaque; each *_create / operation result is owned by the caller and released
 *     with the matching *_destroy.  Operations never mutate their inputs (reference style:
 *     every evaluator call returns a new object) except spear_obj_set_scale.
 *   - all work is queued on the context's CUDA stream; calls that return host data synchronise it.
 *   - chain_index: 0 = key level (all L+P limbs), 1 = fresh ciphertext (L limbs), +1 per rescale /
 *     mod-switch  [ref: fhe_rwkv_inference.py:218, scripts/bootstrap_generation.py:272].
 *   - polynomials are uint64 residues, layout [poly][limb][coefficient], NTT (bit-reversed) form.
 *   - there is no CPU fallback: without a CUDA device spear_context_create fails.
 */
#ifndef SPEAR_B200_H
#define SPEAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct spear_context spear_context;         /* [ref: phantom_binding.cu:97  PhantomContext]     */
typedef struct spear_secret_key spear_secret_key;   /* [ref: phantom_binding.cu:100 PhantomSecretKey]   */
typedef struct spear_public_key spear_public_key;   /* [ref: phantom_binding.cu:112 PhantomPublicKey]   */
typedef struct spear_kswitch_key spear_kswitch_key; /* [ref: phantom_binding.cu:118 PhantomRelinKey]    */
typedef struct spear_galois_keys spear_galois_keys; /* [ref: phantom_binding.cu:121 PhantomGaloisKey]   */
typedef struct spear_obj spear_obj;                 /* [ref: phantom_binding.cu:158-163 plaintext / ciphertext] */
typedef struct spear_diagset spear_diagset;         /* pre-encoded BSGS diagonals [ref: bootstrap_generation.py:252-262] */

const char* spear_last_error(void);
const char* spear_version(void);
uint64_t spear_launch_count(void);   /* kernels launched by this library so far (bench.py gpu_launches) */

/* ---- parameters ------------------------------------------------------------------------------- */
/* [ref: phantom_binding.cu:81 create_coeff_modulus] NTT-friendly primes for the requested bit sizes */
int spear_create_coeff_modulus(uint64_t poly_degree, const int* bit_sizes, int count, uint64_t* out);
/* [ref: phantom_binding.cu:124-126 get_elt_from_step / get_elts_from_steps] generator 5, step 0 = conjugation */
uint64_t spear_get_elt_from_step(int step, uint64_t poly_degree);

/* ---- context ---------------------------------------------------------------------------------- */
/* [ref: phantom_binding.cu:85-98 params + context] special primes are the last `special` moduli */
int spear_context_create(uint64_t poly_degree, const uint64_t* moduli, int count, int special, int device,
                         spear_context** out);
void spear_context_destroy(spear_context* ctx);
int spear_context_sync(spear_context* ctx);
void* spear_context_stream(spear_context* ctx);   /* cudaStream_t */
/* device-side timing on the context stream (CUDA events) */
int spear_timer_start(spear_context* ctx);
int spear_timer_stop(spear_context* ctx, float* elapsed_ms);
/* per-kernel-class device timing (event pair around every launch of the class) for the roofline line:
 * classes 0 key inner product (one rotation per launch), 1 diagonal MAC, 2 forward NTT pass A, 3 ModUp (stand-alone, or
 * the fused inverse pass A + ModUp + forward pass A front end), 4 ModDown, 5 rescale, 6 fused hoisted baby-step key inner
 * product (G-1 rotations per launch), 7 forward pass B fused with the key product (giant steps), 8 forward pass B,
 * 9 inverse pass A, 10 inverse pass B, 11 sum of the giant groups' partial results, 12 peer-exchange flag kernels (post +
 * wait: the time this rank waits for its peers), 13 peer reduce-scatter + Barrett + all-gather kernel and its copies */
int spear_profile_enable(spear_context* ctx, int on);
int spear_profile_read(spear_context* ctx, double* ms, uint64_t* launches, int classes);
/* pinned host buffers for the host<->device legs of the end-to-end path */
int spear_pinned_alloc(size_t bytes, void** out);
void spear_pinned_free(void* p);
/* device memory currently held by the stream-ordered pool (bytes) */
int spear_mem_info(spear_context* ctx, uint64_t* used, uint64_t* reserved);
/* grow the pool to at least `bytes` now (one allocation, freed at once: the pool keeps it), so that a computation
 * that creates large temporaries on the fly never waits for the pool to grow in the middle */
int spear_mem_reserve(spear_context* ctx, uint64_t bytes);

/* ---- keys ------------------------------------------------------------------------------------- */
/* [ref: phantom_binding.cu:100-101 secret_key(ctx)] ternary secret from a 32-byte seed (ChaCha20 streams) */
int spear_secret_key_create(spear_context* ctx, const uint8_t seed[32], spear_secret_key** out);
void spear_secret_key_destroy(spear_secret_key* sk);
/* [ref: phantom_binding.cu:102 gen_publickey] */
int spear_gen_public_key(spear_context* ctx, const spear_secret_key* sk, spear_public_key** out);
void spear_public_key_destroy(spear_public_key* pk);
/* [ref: phantom_binding.cu:103 gen_relinkey] */
int spear_gen_relin_key(spear_context* ctx, const spear_secret_key* sk, spear_kswitch_key** out);
void spear_kswitch_key_destroy(spear_kswitch_key* k);
/* [ref: phantom_binding.cu:104 create_galois_keys; :91 set_galois_elts] one hybrid switching key per element */
int spear_gen_galois_keys(spear_context* ctx, const spear_secret_key* sk, const uint32_t* elts, int count,
                          spear_galois_keys** out);
int spear_galois_keys_add(spear_context* ctx, const spear_secret_key* sk, spear_galois_keys* gk, const uint32_t* elts,
                          int count);
int spear_galois_keys_has(const spear_galois_keys* gk, uint32_t elt);
void spear_galois_keys_destroy(spear_galois_keys* gk);

/* ---- plaintext / ciphertext objects -------------------------------------------------------------- */
void spear_obj_destroy(spear_obj* o);
/* [ref: fork-only ct.chain_index() / ct.scale() / ct.coeff_modulus_size(), bootstrap_generation.py:152,178,272] */
int spear_obj_info(const spear_obj* o, int* size, int* limbs, int* ext, int* ring_n, double* scale, int* chain_index);
/* [ref: phantom_binding.cu:163 ciphertext.set_scale] */
int spear_obj_set_scale(spear_obj* o, double scale);
/* parity hooks and the host-buffer path: raw limbs [size][limbs(+P)][ring_n] */
int spear_obj_export(const spear_obj* o, uint64_t* host, size_t words);
int spear_obj_import(spear_context* ctx, const uint64_t* host, int size, int limbs, int ext, int ring_n, double scale,
                     spear_obj** out);
int spear_secret_key_export(const spear_secret_key* sk, uint64_t* host, size_t words);             /* [K][N] */
int spear_kswitch_key_export(const spear_kswitch_key* k, uint64_t* host, size_t words);            /* [beta][2][K][N] */
int spear_galois_key_export(const spear_galois_keys* gk, uint32_t elt, uint64_t* host, size_t words);
int spear_public_key_export(const spear_public_key* pk, uint64_t* host, size_t words);             /* [2][K][N] */

/* ---- encoder ------------------------------------------------------------------------------------ */
/* [ref: phantom_binding.cu:138-156 ckks_encoder.encode_*; fork-only encode_*_vector_batch,
 *  bootstrap_generation.py:382,423]  values: `count` vectors of ring_n/2 complex slots, interleaved
 *  (re, im) doubles on the host.  ring_n = poly_degree for ordinary plaintexts.  ext != 0 also emits
 *  the special limbs (basis Q_l * P).  outs receives `count` handles. */
int spear_encode(spear_context* ctx, const double* values, int count, int ring_n, double scale, int chain_index,
                 int ext, spear_obj** outs);
/* [ref: phantom_binding.cu:149-156 decode_*] out: poly_degree/2 complex slots, interleaved (re, im) */
int spear_decode(spear_context* ctx, const spear_obj* pt, double* out);

/* ---- encryption --------------------------------------------------------------------------------- */
/* [ref: phantom_binding.cu:105-107 encrypt_symmetric] enc_id selects the randomness streams (parity hook) */
int spear_encrypt_symmetric(spear_context* ctx, const spear_secret_key* sk, const spear_obj* pt, uint64_t enc_id,
                            spear_obj** out);
/* [ref: phantom_binding.cu:113-116 encrypt_asymmetric] */
int spear_encrypt_asymmetric(spear_context* ctx, const spear_public_key* pk, const spear_obj* pt, uint64_t enc_id,
                             spear_obj** out);
/* [ref: phantom_binding.cu:108-110 decrypt] */
int spear_decrypt(spear_context* ctx, const spear_secret_key* sk, const spear_obj* ct, spear_obj** out);

/* The client legs of a projection round trip, three kernel launches each [ref: CKKSBootstrapContext.encrypt /
 * encrypt_replicated / encrypt_replicated_complex and decrypt_vec / decrypt_vec_complex / decrypt_slot0,
 * scripts/bootstrap_generation.py:119-147 -- encode_*_vector + sk.encrypt_symmetric, sk.decrypt + decode_*_vector].
 * values: `count` complex slot values (interleaved re, im) on the host; replicate != 0 tiles them over all slots
 * (replicate_vector, :53-58), else the remaining slots are zero.  Fresh ciphertext (chain_index 1), bit-identical to
 * spear_encode + spear_encrypt_symmetric with the same enc_id.  out of spear_decrypt_decode: the first `want` slots
 * (interleaved re, im), bit-identical to spear_decrypt + spear_decode. */
int spear_encrypt_vector(spear_context* ctx, const spear_secret_key* sk, const double* values, int count, int replicate,
                         double scale, uint64_t enc_id, spear_obj** out);
int spear_decrypt_decode(spear_context* ctx, const spear_secret_key* sk, const spear_obj* ct, double* out, int want);

/* ---- evaluator  [ref: phantom_binding.cu:165-205] ----------------------------------------------------- */
int spear_negate(spear_context* ctx, const spear_obj* a, spear_obj** out);
int spear_add(spear_context* ctx, const spear_obj* a, const spear_obj* b, spear_obj** out);
int spear_sub(spear_context* ctx, const spear_obj* a, const spear_obj* b, spear_obj** out);
int spear_add_plain(spear_context* ctx, const spear_obj* ct, const spear_obj* pt, spear_obj** out);
int spear_sub_plain(spear_context* ctx, const spear_obj* ct, const spear_obj* pt, spear_obj** out);
int spear_multiply(spear_context* ctx, const spear_obj* a, const spear_obj* b, spear_obj** out);
int spear_multiply_plain(spear_context* ctx, const spear_obj* ct, const spear_obj* pt, spear_obj** out);
int spear_relinearize(spear_context* ctx, const spear_obj* ct3, const spear_kswitch_key* rlk, spear_obj** out);
int spear_rescale_to_next(spear_context* ctx, const spear_obj* ct, spear_obj** out);
/* ModRaise, the entry of CKKS bootstrapping [ref: inside the fork-only ckks_bootstrapper.bootstrap,
 * scripts/bootstrap_generation.py:149-154]: a one-limb ciphertext re-read modulo the primes of `chain_index`
 * (centred lift of every coefficient); it then decrypts to m + q_0 * I(X). */
int spear_mod_raise(spear_context* ctx, const spear_obj* ct, int chain_index, spear_obj** out);
int spear_mod_switch_to_next(spear_context* ctx, const spear_obj* o, spear_obj** out);   /* ct or pt */
int spear_apply_galois(spear_context* ctx, const spear_obj* ct, uint32_t elt, const spear_galois_keys* gk,
                       spear_obj** out);
/* [ref: phantom_binding.cu:205 hoisting] rotations of one ciphertext by several elements sharing one
 * decomposition; outs receives `count` ciphertexts */
int spear_hoisted_rotations(spear_context* ctx, const spear_obj* ct, const uint32_t* elts, int count,
                            const spear_galois_keys* gk, spear_obj** outs);

/* ---- BSGS diagonal mat-vec ------------------------------------------------------------------------ */
/* [ref: fork-only ph.bsgs_multiply_accumulate(ctx, ct_baby, pts, G, B, D, gk), bootstrap_generation.py:459;
 *  executable spec = the Python loop :464-484]  exact mode: reference op order, result rescaled. */
int spear_bsgs_multiply_accumulate(spear_context* ctx, spear_obj* const* ct_baby, int n_baby, spear_obj* const* pts,
                                   int n_pts, int G, int B, int D, const spear_galois_keys* gk, spear_obj** out);
/* [ref: fork-only ph.bsgs_from_cpu(ctx, ct_baby, data, ci, sc, cms, pmd, G, B, D, gk), bootstrap_generation.py:449, over the
 *  offload format of ph.offload_plaintexts, :336-358]  host_pts: [n_pts][pt_limbs][N] NTT-form plaintext residues in host
 *  memory.  Same result as spear_bsgs_multiply_accumulate; the diagonals stream through a two-slot device ring, the copy of
 *  giant group g+1 overlapping the arithmetic of group g (page-locked host memory makes the copies asynchronous). */
int spear_bsgs_from_host(spear_context* ctx, spear_obj* const* ct_baby, int n_baby, const uint64_t* host_pts, int n_pts,
                         int pt_limbs, double pt_scale, int G, int B, int D, const spear_galois_keys* gk, spear_obj** out);
/* [ref: fork-only ph.offload_plaintexts(list[pt]), :339]  count objects of one shape -> host [count][words_each] */
int spear_objs_export(spear_context* ctx, spear_obj* const* objs, int count, uint64_t* host, size_t words_each);
/* [ref: pre_encode_real_diags / pre_encode_complex_diags, bootstrap_generation.py:252-262, 361-432]
 *  diags: D period-D complex vectors (already pre-rotated by +gG per giant group, :365-369), interleaved
 *  (re, im).  compress != 0 stores the sub-ring form (ring 2D, N/(2D)-fold smaller). */
int spear_diagset_encode(spear_context* ctx, const double* diags, int D, int G, int B, double scale, int chain_index,
                         int compress, spear_diagset** out);
/* one shard of the same set for giant-step sharding over GPUs: only the giant groups
 * g_first, g_first + g_stride, ... are stored; `diags` holds their n_diags rows in that order. */
int spear_diagset_encode_shard(spear_context* ctx, const double* diags, int n_diags, int D, int G, int B, int g_first,
                               int g_stride, double scale, int chain_index, int compress, spear_diagset** out);
/* the same from the matrix itself (row-major D x D, y = M x; m_im optional second matrix for the complex packing
 * M_re + i M_im): diagonal extraction, the +gG pre-rotation and the slot tiling of _extract_diagonals /
 * _batch_encode_diags_* [ref: bootstrap_generation.py:198-203, 361-432] run on the device. */
int spear_diagset_encode_matrix(spear_context* ctx, const double* m_re, const double* m_im, int D, int G, int B,
                                int g_first, int g_stride, double scale, int chain_index, int compress,
                                spear_diagset** out);
/* the same from a VIEW of host memory, so that chunks of a larger weight matrix need no host-side copy
 * [ref: the chunk extraction of fhe_projection_bsgs :575-591, :630-642 and test_fully_enc_bsgs.py:41-79]:
 * M[i][j] = m[i * pitch + j] (transposed = 0) or m[j * pitch + i] (transposed != 0) for i < rows_valid, j < cols_valid,
 * zero elsewhere in the D x D matrix; pitch in doubles. */
int spear_diagset_encode_matrix_view(spear_context* ctx, const double* m_re, const double* m_im, int D, int rows_valid,
                                     int cols_valid, size_t pitch, int transposed, int G, int B, int g_first, int g_stride,
                                     double scale, int chain_index, int compress, spear_diagset** out);
void spear_diagset_destroy(spear_diagset* d);
int spear_diagset_info(const spear_diagset* d, int* D, int* G, int* B, int* limbs, int* ring_n, double* scale,
                       uint64_t* bytes);
int spear_diagset_export(const spear_diagset* d, uint64_t* host, size_t words);   /* [D][limbs+P][ring_n] */
/* hoisted mode (this build's fast path): baby steps, diagonal MAC and giant steps in one call
 * [ref: fork-only ph.bsgs_complete_from_cpu(ctx, ct_x, ...), bootstrap_generation.py:242] */
int spear_bsgs_hoisted(spear_context* ctx, const spear_obj* ct, const spear_diagset* diags,
                       const spear_galois_keys* gk, spear_obj** out);

/* `count` independent mat-vecs (e.g. the r, k, v projections of one block: same keys, different inputs and
 * diagonal sets, reference bootstrap_generation.py:784-792) issued on separate streams so that key streaming of
 * one overlaps the transforms of the others.  outs receives `count` ciphertexts. */
int spear_bsgs_hoisted_batch(spear_context* ctx, spear_obj* const* cts, spear_diagset* const* diags, int count,
                             const spear_galois_keys* gk, spear_obj** outs);
/* `count` diagonal sets (same D, G and level) multiplying ONE ciphertext -- the chunk pairs of a D -> F projection, for
 * which the reference computes the baby rotations once [ref: scripts/bootstrap_generation.py:575-600]: one decomposition
 * and one set of hoisted baby steps, then diagonal MAC, giant steps, ModDown + rescale per set.  outs[i] has the limbs of
 * spear_bsgs_hoisted(ct, diags[i]). */
int spear_bsgs_hoisted_shared(spear_context* ctx, const spear_obj* ct, spear_diagset* const* diags, int count,
                              const spear_galois_keys* gk, spear_obj** outs);
/* Serving form of the batch (the reference keeps client and server in one process and hands ciphertexts over in host
 * memory when it offloads them, scripts/bootstrap_generation.py:336-358, 545-556): in[i] = [2][limbs][N] ciphertext limbs
 * in host memory (page-locked for asynchronous copies), out[i] = [2][limbs-1][N].  Item i is uploaded, multiplied and
 * downloaded on auxiliary stream i % 3: the PCIe legs of one item run under the arithmetic of the others.  Returns when
 * every result has landed; out_scale[i] (optional) = scale of result i. */
int spear_bsgs_hoisted_batch_host(spear_context* ctx, const uint64_t* const* in, int limbs, double scale,
                                  spear_diagset* const* diags, int count, const spear_galois_keys* gk, uint64_t* const* out,
                                  double* out_scale);
/* Sharded form: the shard's accumulator in basis Q_l*P (size 2, ext).  Accumulators of all shards are summed
 * (spear_add, or an integer all-reduce over spear_obj_device_ptr followed by spear_obj_reduce) and finished once. */
int spear_bsgs_hoisted_partial(spear_context* ctx, const spear_obj* ct, const spear_diagset* shard,
                               const spear_galois_keys* gk, spear_obj** out);
/* `count` shard accumulators at once, on separate streams like spear_bsgs_hoisted_batch (the independent
 * projections of one block phase); their all-reduces are then issued back to back. */
int spear_bsgs_hoisted_partial_batch(spear_context* ctx, spear_obj* const* cts, spear_diagset* const* shards, int count,
                                     const spear_galois_keys* gk, spear_obj** outs);
int spear_bsgs_finish(spear_context* ctx, spear_obj* acc, spear_obj** out);   /* ModDown + rescale; clobbers acc */
int spear_obj_reduce(spear_context* ctx, spear_obj* o);                       /* every residue mod its modulus, in place */
void* spear_obj_device_ptr(spear_obj* o);                                     /* device address of the limbs */

/* ---- peer-memory exchange of shard accumulators (one process per GPU) ---------------
 * [ref: the reference has no multi-GPU path; BASELINE north_star: "partial ciphertexts are combined with ... P2P
 *  over NVLink plus a modular-add kernel"]  A window is device memory of this rank that the other ranks of its
 * group map through CUDA IPC.  create -> exchange the 64-byte handles by any host channel (torch.distributed
 * all_gather in fhe_spear_b200/sharding.py) -> connect(handles of all ranks, rank-major).  spear_peer_allreduce then
 * replaces acc (size 2, basis Q_l*P, same shape on every rank) by the sum over ranks mod q, in place, with one fused
 * reduce-scatter + Barrett + all-gather kernel over NVLink, ordered on the context's stream (no host sync).
 * Every rank of the group must call it with the same slot in the same order.  A peer that does not arrive within
 * 20 s raises the window status (non-zero) instead of hanging: the reduce step is skipped, acc is filled with all-ones
 * words (no valid residue), and later calls on the window fail.  Check spear_peer_window_status after the next
 * synchronisation (fhe_spear_b200.sharding does, before it hands a result out). */
typedef struct spear_peer_window spear_peer_window;
#define SPEAR_IPC_HANDLE_BYTES 64
int spear_peer_window_create(spear_context* ctx, int rank, int world, uint64_t slot_bytes, int slots,
                             uint8_t* handle_out /* [SPEAR_IPC_HANDLE_BYTES] */, spear_peer_window** out);
int spear_peer_window_connect(spear_context* ctx, spear_peer_window* w, const uint8_t* handles /* [world][64] */);
int spear_peer_allreduce(spear_context* ctx, spear_peer_window* w, int slot, spear_obj* acc);
int spear_peer_window_status(const spear_peer_window* w);   /* 0 = healthy; 1 + r = peer r timed out */
/* Test hook for one-GPU boxes: runs the reduce kernels of all `world` ranks one after the other over local windows
 * holding accs[0..world) (no epoch waits: kernels that wait on one another must not share a GPU); afterwards every
 * accs[r] holds the sum of all of them mod q, as after a real exchange. */
int spear_peer_selftest(spear_context* ctx, spear_obj* const* accs, int world);
void spear_peer_window_destroy(spear_peer_window* w);

/* ---- two-phase mat-vec over a rank group ("baby steps sharded, partial ciphertexts combined over
 *      NVLink"; reference loop scripts/bootstrap_generation.py:435-485) -------------------------------------------
 * Phase 1 -- the hoisted baby steps and the diagonal multiply-accumulate -- is split by ROWS of the l + P RNS limbs:
 * up to four ranks take rows [r*(l+P)/w, (r+1)*(l+P)/w) each; larger even groups form w/2 row groups of two ranks that
 * take one half of the coefficient columns each (27 rows over 8 ranks: 3.5 row-equivalents per rank instead of 4).  A rank
 * serves its share for EVERY giant group and holds just that share of the diagonals (spear_split_share,
 * spear_diagset_slice_share).  Phase 2 -- the giant steps -- is split by giant group g = r, r + w, ...  In between, the
 * MAC kernel's own epilogue stores scatter the accumulators of group g into the window of rank g % w over NVLink peer
 * memory (the all-to-all is fused into the compute kernel), epoch flags order the phases, and the call returns this
 * rank's accumulator in basis Q_l*P: spear_peer_allreduce (on a SECOND window) sums them, spear_bsgs_finish completes.
 * `w` must have slots of at least ceil(B/world) * 2 * (l+P) * N * 8 bytes and must not be used for all-reduces. */
int spear_split_share(int rank, int world, int rows, int N, int* row0, int* nrows, int* col0, int* ncols);
int spear_diagset_slice_share(spear_context* ctx, const spear_diagset* full, int rank, int world, spear_diagset** out);
int spear_diagset_slice_rows(spear_context* ctx, const spear_diagset* full, int row0, int nrows, spear_diagset** out);
int spear_bsgs_split(spear_context* ctx, const spear_obj* ct, const spear_diagset* rows, const spear_galois_keys* gk,
                     spear_peer_window* w, int slot, spear_obj** out);
/* item i on auxiliary stream i % 3, exchanging through slot slot0 + i */
int spear_bsgs_split_batch(spear_context* ctx, spear_obj* const* cts, spear_diagset* const* rows, int count,
                           const spear_galois_keys* gk, spear_peer_window* w, int slot0, spear_obj** outs);
/* the same for `count` sets multiplying ONE ciphertext (shared baby steps, see spear_bsgs_hoisted_shared); set i goes
 * through slot slot0 + i; outs[i] = this rank's accumulator of set i */
int spear_bsgs_split_shared(spear_context* ctx, const spear_obj* ct, spear_diagset* const* rows, int count,
                            const spear_galois_keys* gk, spear_peer_window* w, int slot0, spear_obj** outs);
/* Test hook for one-GPU boxes: both phases of all `world` ranks, one after the other, over local stand-ins for the
 * windows; rows[r] = the slice of rank r.  *out = the summed accumulator (spear_bsgs_finish completes it). */
int spear_bsgs_split_selftest(spear_context* ctx, const spear_obj* ct, spear_diagset* const* rows, int world,
                              const spear_galois_keys* gk, spear_obj** out);

/* ---- raw transforms (tests / profiling) ------------------------------------------------------------- */
/* in-place on a host buffer of `rows` x ring_n residues whose row r uses modulus limb_ids[r] */
int spear_ntt_host(spear_context* ctx, uint64_t* data, int rows, const int* limb_ids, int ring_n, int inverse);

#ifdef __cplusplus
}
#endif
#endif
