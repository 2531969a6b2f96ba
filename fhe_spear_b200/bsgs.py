"""Host-side mirror of the reference's BSGS layer (scripts/bootstrap_generation.py:18-660).

Same function names, argument meaning and return values as the reference, so callers such as
client_aided_block (:756) and fully_encrypted_ffn_block (test_fully_enc_bsgs.py:26) work unchanged
when they import from here instead.  Differences, all additive:

  * `pre_encode_real_diags` / `pre_encode_complex_diags` return a `ph.diagonal_set` (basis Q_l*P,
    sub-ring compressed when D is a power of two) instead of a list of plaintexts, unless
    `as_plaintexts=True`;
  * `fhe_matmul_bsgs*` given a diagonal_set (or nothing pre-encoded) run the hoisted one-call path
    `ph.bsgs_hoisted` on the GPU; given a list of plaintexts they run `ph.bsgs_multiply_accumulate`
    (reference op order), exactly like the reference's fork-only fast path (:458-462).
"""
import numpy as np

from . import pyPhantom as ph


# ---- planning  [ref: :18-58] ----------------------------------------------------------------------
def compute_rotation_galois_elements(poly_degree, max_dim):
    m = 2 * poly_degree
    elts = {m - 1}
    s = 1
    while s <= max_dim:
        elts.add(pow(5, s, m))
        s <<= 1
    return list(elts)


def compute_bsgs_params(D, baby_weight=1.0):
    """(G baby steps, B giant steps).  baby_weight = 1 is the reference's split G = ceil(sqrt(D)) [ref: :29-32].
    Hoisted baby steps only stream a key while every giant step pays a full decomposition, so the hoisted
    path is faster with more baby steps: baby_weight = 2 gives G = ceil(sqrt(2 D)) (D=2048: 64 x 32)."""
    G = int(np.ceil(np.sqrt(D * baby_weight)))
    return G, int(np.ceil(D / G))


def hoisting_weight(world=1):
    """baby_weight that minimises the hoisted path's time on B200: a hoisted baby step only streams its key
    (~19 us at C3) while a giant step pays ModUp + NTT + key product (~200 us), and with giant-step sharding the
    baby steps are replicated on every rank while the giant steps are divided: weight ~ 8 / world
    (D=2048: 128 x 16 on one GPU, 91 x 23 on two, 64 x 32 on four, the reference's 46 x 45 on eight)."""
    return max(1.0, 8.0 / max(1, world))


def bsgs_steps(D, baby_weight=1.0):
    G, B = compute_bsgs_params(D, baby_weight)
    return list(range(1, G)) + [g * G for g in range(1, B)]


def compute_bsgs_galois_elements(poly_degree, D, baby_weight=1.0):
    return ph.get_elts_from_steps(bsgs_steps(D, baby_weight), poly_degree)


def compute_diagonals(W, D):
    return [np.array([W[j, (j + k) % D] for j in range(D)]) for k in range(D)]


def replicate_vector(vec, slots):
    reps, rem = divmod(slots, len(vec))
    return list(vec) * reps + list(vec[:rem])


def _replicate_to_slots(vec, slots):
    reps, rem = divmod(slots, len(vec))
    return np.concatenate([np.tile(vec, reps), vec[:rem]])


# ---- context wrapper  [ref: :61-154] -----------------------------------------------------------------
class CKKSBootstrapContext:
    def __init__(self, poly_degree=32768, L0=24, prime_bits=59, special_mod_size=3, level_budget=None,
                 max_rot_dim=256, bsgs_dim=0, skip_bootstrap=False, seed=None, device=None, verbose=True,
                 baby_weights=(1.0,)):
        if level_budget is None:
            level_budget = [2, 2]
        say = print if verbose else (lambda *a, **k: None)
        say(f"[CKKS] Setting up: N={poly_degree}, L0={L0}, bits={prime_bits}, P={special_mod_size}"
            + ("" if skip_bootstrap else f", budget={level_budget}"))
        boot_elts = [] if skip_bootstrap else ph.ckks_bootstrapper.get_galois_elements(poly_degree, 0, level_budget)
        rot_elts = compute_rotation_galois_elements(poly_degree, max_dim=max_rot_dim)
        dims = bsgs_dim if isinstance(bsgs_dim, (list, tuple)) else [bsgs_dim]
        bsgs_elts = set()
        for d in sorted({d for d in dims if d > 0}):
            for w in baby_weights:          # extra (G, B) splits only add the keys they need
                elts = compute_bsgs_galois_elements(poly_degree, d, w)
                bsgs_elts.update(elts)
                G, B = compute_bsgs_params(d, w)
                say(f"[CKKS] BSGS: D={d}, G={G} baby, B={B} giant, {len(elts)} galois elements")
        all_elts = sorted(set(boot_elts) | set(rot_elts) | bsgs_elts)
        say(f"[CKKS] Galois elements: {len(boot_elts)} boot + {len(rot_elts)} rot + {len(bsgs_elts)} bsgs = {len(all_elts)} total")

        parms = ph.params(ph.scheme_type.ckks)
        parms.set_poly_modulus_degree(poly_degree)
        parms.set_special_modulus_size(special_mod_size)
        parms.set_galois_elts(all_elts)
        parms.set_coeff_modulus(ph.create_coeff_modulus(poly_degree, [prime_bits] * (L0 + special_mod_size)))

        self.ctx = ph.context(parms, device=device)
        self.sk = ph.secret_key(self.ctx, seed=seed)
        self.encoder = ph.ckks_encoder(self.ctx)
        self.scale = 2.0 ** prime_bits
        # L0<=2: half-scale diagonals so the product still fits the remaining modulus  [ref: :104]
        self.diag_scale = 2.0 ** (prime_bits // 2) if L0 <= 2 else self.scale
        self.slots = self.encoder.slot_count()
        self.L0 = L0
        self.rlk = self.sk.gen_relinkey(self.ctx)
        self.gk = self.sk.create_galois_keys(self.ctx)
        self.bt = None
        if not skip_bootstrap:                      # [ref: :110-116]
            self.bt = ph.ckks_bootstrapper(self.encoder)
            self.bt.setup(self.ctx, level_budget)
            self.bt.impl.gk, self.bt.impl.rlk = self.gk, self.rlk     # the context's keys already cover the bootstrap elements
            bt_depth = ph.ckks_bootstrapper.get_bootstrap_depth(level_budget, poly_degree)
            say(f"[CKKS] Bootstrap depth={bt_depth}, post-bootstrap levels={L0 - bt_depth - 1}")
        say(f"[CKKS] Slots={self.slots}")

    # The client legs go through the one-call forms of the drop-in (encode + encrypt / decrypt + decode in three launches
    # each, csrc/client.cu) when the secret key offers them; results are bit-identical to the reference's two-step form
    # below them [ref: :119-147], which stays as the path for a reference-style pyPhantom.
    def encrypt(self, vec):
        if hasattr(self.sk, "encrypt_vector"):
            return self.sk.encrypt_vector(self.ctx, np.asarray(vec, dtype=np.float64), self.scale, replicate=False)
        padded = np.zeros(self.slots)
        padded[:len(vec)] = vec
        return self.sk.encrypt_symmetric(self.ctx, self.encoder.encode_double_vector(self.ctx, padded, self.scale))

    def encrypt_replicated(self, vec):
        if hasattr(self.sk, "encrypt_vector"):
            return self.sk.encrypt_vector(self.ctx, np.asarray(vec, dtype=np.float64), self.scale, replicate=True)
        rep = _replicate_to_slots(np.asarray(vec, dtype=np.float64), self.slots)
        return self.sk.encrypt_symmetric(self.ctx, self.encoder.encode_double_vector(self.ctx, rep, self.scale))

    def encrypt_replicated_complex(self, vec_real, vec_imag):
        if hasattr(self.sk, "encrypt_vector"):
            return self.sk.encrypt_vector(self.ctx, np.asarray(vec_real) + 1j * np.asarray(vec_imag), self.scale, replicate=True)
        rep = _replicate_to_slots(np.asarray(vec_real) + 1j * np.asarray(vec_imag), self.slots)
        return self.sk.encrypt_symmetric(self.ctx, self.encoder.encode_complex_vector(self.ctx, rep, self.scale))

    def _decode(self, pt, dim):
        """first `dim` slots as a complex array; list-returning decode of a reference-style encoder as the fallback"""
        fast = getattr(self.encoder, "decode_array", None)
        if fast is not None:
            return fast(self.ctx, pt)[:dim].copy()
        return np.array(self.encoder.decode_complex_vector(self.ctx, pt)[:dim])

    def _decrypt_decode(self, ct, dim):
        if hasattr(self.sk, "decrypt_decode"):
            return self.sk.decrypt_decode(self.ctx, ct, dim)
        return self._decode(self.sk.decrypt(self.ctx, ct), dim)

    def decrypt_vec(self, ct, dim):
        return self._decrypt_decode(ct, dim).real.copy()

    def decrypt_vec_complex(self, ct, dim):
        return self._decrypt_decode(ct, dim)

    def decrypt_slot0(self, ct):
        return float(self._decrypt_decode(ct, 1).real[0])

    def bootstrap(self, ct):
        """[ref: :149-154] mod-switch down to two limbs, then ckks_bootstrapper.bootstrap"""
        if self.bt is None:
            raise RuntimeError("Bootstrap not available (skip_bootstrap=True)")
        while ct.coeff_modulus_size() > 2:
            ct = ph.mod_switch_to_next(self.ctx, ct)
        return self.bt.bootstrap(self.ctx, ct)


# ---- diagonals  [ref: :198-203, :361-432] --------------------------------------------------------------
def _extract_diagonals(W, D):
    """diags[k][j] = W[j, (j + k) mod D]"""
    j = np.arange(D)
    return W[j[None, :], (j[None, :] + j[:, None]) % D]


def _pre_rotate(diags, D, G):
    """rows of giant group g (k in [gG, (g+1)G)) rolled right by gG  [ref: :365-369]"""
    out = np.array(diags, copy=True)
    for g in range(1, (D + G - 1) // G):
        s, e = g * G, min((g + 1) * G, D)
        out[s:e] = np.roll(diags[s:e], g * G, axis=1)
    return out


def _tile_rows(rows, slots):
    D = rows.shape[1]
    reps, rem = divmod(slots, D)
    return np.concatenate([np.tile(rows, (1, reps)), rows[:, :rem]], axis=1)


def _batch_encode_diags_real(ckks, diags, D, G, slots, level):
    vecs = _tile_rows(_pre_rotate(diags, D, G), slots)
    return ckks.encoder.encode_double_vector_batch(ckks.ctx, vecs, ckks.diag_scale, chain_index=level)


def _batch_encode_diags_complex(ckks, diags1, diags2, D, G, slots, level):
    vecs = _tile_rows(_pre_rotate(diags1, D, G) + 1j * _pre_rotate(diags2, D, G), slots)
    return ckks.encoder.encode_complex_vector_batch(ckks.ctx, vecs, ckks.diag_scale, chain_index=level)


def pre_encode_real_diags(ckks, W, D, G, B, level, as_plaintexts=False, compress=True, shard=(0, 1)):
    """shard = (rank, world): keep only this rank's giant groups (giant-step sharding over GPUs)."""
    W = np.asarray(W, dtype=np.float64)
    if as_plaintexts:
        return _batch_encode_diags_real(ckks, _extract_diagonals(_padded(W, D), D), D, G, ckks.slots, level)
    # diagonal extraction and pre-rotation run on the device (the host gather alone took 0.19 s at D = 2048); W may be a
    # view of a larger weight matrix (chunks, transposes) and smaller than D x D: it is read as it lies in host memory
    return ph.diagonal_set.from_matrix(ckks.ctx, W[:D, :D], G, B, ckks.diag_scale, chain_index=level, compress=compress,
                                       shard=shard, D=D)


def pre_encode_complex_diags(ckks, W1, W2, D, G, B, level, as_plaintexts=False, compress=True, shard=(0, 1)):
    W1, W2 = np.asarray(W1, dtype=np.float64), np.asarray(W2, dtype=np.float64)
    if as_plaintexts:
        return _batch_encode_diags_complex(ckks, _extract_diagonals(_padded(W1, D), D), _extract_diagonals(_padded(W2, D), D), D, G,
                                           ckks.slots, level)
    return ph.diagonal_set.from_matrix(ckks.ctx, W1[:D, :D], G, B, ckks.diag_scale, chain_index=level, compress=compress,
                                       shard=shard, M_imag=W2[:D, :D], D=D)


def _padded(W, D):
    """W zero-padded to D x D (host copy; only the plaintext-list forms need it)"""
    if W.shape == (D, D):
        return W
    M = np.zeros((D, D))
    M[:min(D, W.shape[0]), :min(D, W.shape[1])] = W[:D, :D]
    return M


def _chunk_pairs(F, D):
    """chunks of D taken two at a time: [(c, c+1 or None), ...]  [ref: :283-305]"""
    n = int(np.ceil(F / D))
    return [(c, c + 1 if c + 1 < n else None) for c in range(0, n, 2)]


def _key_chunk(W, c, D, F):
    """rows = outputs of chunk c of a (D, F) matrix used as x @ W  [ref: :287-291]"""
    lo, hi = c * D, min((c + 1) * D, F)
    return W[:, lo:hi].T          # a (hi - lo, D) view: the device reads it in place and zero-pads to D x D


def _val_chunk(W, c, D, F, sign=1.0):
    """columns = inputs of chunk c of an (F, D) matrix used as x @ W  [ref: :313-317]"""
    lo, hi = c * D, min((c + 1) * D, F)
    return W[lo:hi, :].T if sign == 1.0 else (sign * W[lo:hi, :]).T      # a (D, hi - lo) view (one scaled copy for sign = -1)


def pre_encode_block(ckks, block, D, F, G=None, B=None, as_plaintexts=False, shard=(0, 1)):
    """All 8 diagonal sets of one RWKV-7 block at the level of a fresh ciphertext  [ref: :265-333]"""
    if G is None or B is None:
        G, B = compute_bsgs_params(D)
    level = ckks.encrypt_replicated(np.zeros(1)).chain_index()
    kw = dict(as_plaintexts=as_plaintexts, shard=shard)
    pe = {name: pre_encode_real_diags(ckks, W.T, D, G, B, level, **kw)
          for name, W in (("r", block.W_r), ("k", block.W_k), ("v", block.W_v), ("o", block.W_o))}
    pe["ffn_key"] = []
    for c, c2 in _chunk_pairs(F, D):
        M1 = _key_chunk(block.W_key_ffn, c, D, F)
        pe["ffn_key"].append(pre_encode_real_diags(ckks, M1, D, G, B, level, **kw) if c2 is None else
                             pre_encode_complex_diags(ckks, M1, _key_chunk(block.W_key_ffn, c2, D, F), D, G, B, level, **kw))
    pe["ffn_val"] = []
    for c, c2 in _chunk_pairs(F, D):
        M0 = _val_chunk(block.W_val_ffn, c, D, F)
        pe["ffn_val"].append(pre_encode_real_diags(ckks, M0, D, G, B, level, **kw) if c2 is None else
                             pre_encode_complex_diags(ckks, M0, _val_chunk(block.W_val_ffn, c2, D, F, -1.0), D, G, B, level, **kw))
    return pe


def offload_block_plaintexts(pe_block):
    """[ref: :336-342] only meaningful for blocks pre-encoded with as_plaintexts=True"""
    cpu = {k: ph.offload_plaintexts(pe_block[k]) for k in ("r", "k", "v", "o")}
    cpu["ffn_key"] = [ph.offload_plaintexts(p) for p in pe_block["ffn_key"]]
    cpu["ffn_val"] = [ph.offload_plaintexts(p) for p in pe_block["ffn_val"]]
    return cpu


def upload_block_plaintexts(cpu_block):
    pe = {k: ph.upload_plaintexts(*cpu_block[k]) for k in ("r", "k", "v", "o")}
    pe["ffn_key"] = [ph.upload_plaintexts(*item) for item in cpu_block["ffn_key"]]
    pe["ffn_val"] = [ph.upload_plaintexts(*item) for item in cpu_block["ffn_val"]]
    return pe


# ---- the mat-vec  [ref: :215-220, :435-542] ---------------------------------------------------------------
def _compute_baby_rotations(ckks, ct_x_rep, G):
    return [ct_x_rep] + [ph.rotate(ckks.ctx, ct_x_rep, b, ckks.gk) for b in range(1, G)]


def _matmul(ckks, ct_x_rep, make_set, make_pts, D, G, B, ct_baby, preencoded, cpu_offloaded):
    if G is None or B is None:
        G, B = compute_bsgs_params(D)
    level = ct_x_rep.chain_index()
    if cpu_offloaded is not None:
        preencoded = ph.upload_plaintexts(*cpu_offloaded, ctx=ckks.ctx)
    if preencoded is None:
        preencoded = make_set(level) if ct_baby is None else make_pts(level)
    if isinstance(preencoded, ph.diagonal_set):
        if preencoded.shard[1] > 1:      # giant-step shard: partial accumulator, all-reduce over the ranks, finish
            from .sharding import sharded_matvec
            return sharded_matvec(ckks, ct_x_rep, preencoded)
        return ph.bsgs_hoisted(ckks.ctx, ct_x_rep, preencoded, ckks.gk)
    if ct_baby is None:
        ct_baby = _compute_baby_rotations(ckks, ct_x_rep, G)
    return ph.bsgs_multiply_accumulate(ckks.ctx, ct_baby, preencoded, G, B, D, ckks.gk)


def fhe_matmul_bsgs(ckks, ct_x_rep, W, D, G=None, B=None, ct_baby=None, preencoded=None, cpu_offloaded=None):
    """Enc(x replicated) -> Enc(W @ x), one level consumed  [ref: :435-485]"""
    g, b = (G, B) if G and B else compute_bsgs_params(D)
    return _matmul(ckks, ct_x_rep,
                   lambda lvl: pre_encode_real_diags(ckks, W, D, g, b, lvl),
                   lambda lvl: pre_encode_real_diags(ckks, W, D, g, b, lvl, as_plaintexts=True),
                   D, G, B, ct_baby, preencoded, cpu_offloaded)


def fhe_matmul_bsgs_complex(ckks, ct_x_rep, W1, W2, D, G=None, B=None, ct_baby=None, preencoded=None,
                            cpu_offloaded=None):
    """Enc(x) -> Enc(W1 @ x + i W2 @ x)  [ref: :488-542]"""
    g, b = (G, B) if G and B else compute_bsgs_params(D)
    return _matmul(ckks, ct_x_rep,
                   lambda lvl: pre_encode_complex_diags(ckks, W1, W2, D, g, b, lvl),
                   lambda lvl: pre_encode_complex_diags(ckks, W1, W2, D, g, b, lvl, as_plaintexts=True),
                   D, G, B, ct_baby, preencoded, cpu_offloaded)


def fhe_projection_bsgs(ckks, x, W, D_in, D_out, label="", preencoded_diags=None, cpu_offloaded_diags=None):
    """x @ W for D_in == D_out, D_out > D_in (complex-packed output chunks) and D_out < D_in
    (conjugate-packed input chunks); encrypts x, decrypts the result  [ref: :545-659]"""
    pick = lambda lst, i: lst[i] if lst else None
    x = np.asarray(x, dtype=np.float64)
    if D_in == D_out:
        ct_y = fhe_matmul_bsgs(ckks, ckks.encrypt_replicated(x), W.T, D_in, preencoded=pick(preencoded_diags, 0),
                               cpu_offloaded=pick(cpu_offloaded_diags, 0))
        return ckks.decrypt_vec(ct_y, D_in)

    all_sets = bool(preencoded_diags) and all(isinstance(p, ph.diagonal_set) for p in preencoded_diags)
    if all_sets and any(p.shard[1] > 1 for p in preencoded_diags):
        from .sharding import sharded_matvec_batch
        run_batch = lambda cts, sets: sharded_matvec_batch(ckks, cts, sets)      # giant-step shards on every rank
    else:
        def run_batch(cts, sets):
            if len(cts) > 1 and all(c is cts[0] for c in cts) and len({(d.D, d.G, d.B) for d in sets}) == 1:
                return ph.bsgs_hoisted_shared(ckks.ctx, cts[0], sets, ckks.gk)   # one input: the baby steps once [ref: :575-600]
            return ph.bsgs_hoisted_batch(ckks.ctx, cts, sets, ckks.gk)
    if D_out > D_in and all_sets:
        # every chunk pair is an independent mat-vec on the same input: one batched call, then unpack (re, im)
        D, F = D_in, D_out
        pairs = _chunk_pairs(F, D)
        ct_x = ckks.encrypt_replicated(x)
        outs = run_batch([ct_x] * len(pairs), list(preencoded_diags[:len(pairs)]))
        result = np.zeros(F)
        for (c, c2), ct_y in zip(pairs, outs):
            lo1, hi1 = c * D, min((c + 1) * D, F)
            if c2 is None:
                result[lo1:hi1] = ckks.decrypt_vec(ct_y, D)[:hi1 - lo1]
            else:
                lo2, hi2 = c2 * D, min((c2 + 1) * D, F)
                vals = ckks.decrypt_vec_complex(ct_y, D)
                result[lo1:hi1], result[lo2:hi2] = vals.real[:hi1 - lo1], vals.imag[:hi2 - lo2]
        return result
    if D_out < D_in and all_sets:
        # conjugate-packed input chunk pairs: independent mat-vecs whose real parts are summed in plaintext
        D, F = D_out, D_in
        pairs = _chunk_pairs(F, D)
        cts = []
        for c, c2 in pairs:
            x0, x1 = np.zeros(D), np.zeros(D)
            lo, hi = c * D, min((c + 1) * D, F)
            x0[:hi - lo] = x[lo:hi]
            if c2 is None:
                cts.append(ckks.encrypt_replicated(x0))
            else:
                lo1, hi1 = c2 * D, min((c2 + 1) * D, F)
                x1[:hi1 - lo1] = x[lo1:hi1]
                cts.append(ckks.encrypt_replicated_complex(x0, x1))
        outs = run_batch(cts, list(preencoded_diags[:len(pairs)]))
        return sum(ckks.decrypt_vec_complex(ct_y, D).real for ct_y in outs)

    if D_out > D_in:
        D, F = D_in, D_out
        G, B = compute_bsgs_params(D)
        ct_x = ckks.encrypt_replicated(x)
        result = np.zeros(F)
        need_baby = bool(preencoded_diags and not isinstance(preencoded_diags[0], ph.diagonal_set)) or bool(cpu_offloaded_diags)
        ct_baby = _compute_baby_rotations(ckks, ct_x, G) if need_baby else None
        for i, (c, c2) in enumerate(_chunk_pairs(F, D)):
            pe, cpu = pick(preencoded_diags, i), pick(cpu_offloaded_diags, i)
            have = pe is not None or cpu is not None          # chunk matrices are only built when they must be encoded
            lo1, hi1 = c * D, min((c + 1) * D, F)
            M1 = None if have else _key_chunk(W, c, D, F)
            if c2 is not None:
                lo2, hi2 = c2 * D, min((c2 + 1) * D, F)
                M2 = None if have else _key_chunk(W, c2, D, F)
                ct_y = fhe_matmul_bsgs_complex(ckks, ct_x, M1, M2, D, G, B, ct_baby=ct_baby, preencoded=pe,
                                               cpu_offloaded=cpu)
                vals = ckks.decrypt_vec_complex(ct_y, D)
                result[lo1:hi1] = vals.real[:hi1 - lo1]
                result[lo2:hi2] = vals.imag[:hi2 - lo2]
            else:
                ct_y = fhe_matmul_bsgs(ckks, ct_x, M1, D, G, B, ct_baby=ct_baby, preencoded=pe, cpu_offloaded=cpu)
                result[lo1:hi1] = ckks.decrypt_vec(ct_y, D)[:hi1 - lo1]
        return result

    D, F = D_out, D_in
    G, B = compute_bsgs_params(D)
    result = np.zeros(D)
    for i, (c, c2) in enumerate(_chunk_pairs(F, D)):
        pe, cpu = pick(preencoded_diags, i), pick(cpu_offloaded_diags, i)
        have = pe is not None or cpu is not None
        x0 = np.zeros(D)
        lo, hi = c * D, min((c + 1) * D, F)
        x0[:hi - lo] = x[lo:hi]
        M0 = None if have else _val_chunk(W, c, D, F)
        if c2 is not None:
            x1 = np.zeros(D)
            lo1, hi1 = c2 * D, min((c2 + 1) * D, F)
            x1[:hi1 - lo1] = x[lo1:hi1]
            # Enc(x0 + i x1) * (d0 - i d1): real part = M0 x0 + M1 x1  [ref: :630-642]
            M1n = None if have else _val_chunk(W, c2, D, F, -1.0)
            ct_y = fhe_matmul_bsgs_complex(ckks, ckks.encrypt_replicated_complex(x0, x1), M0, M1n, D, G, B,
                                           preencoded=pe, cpu_offloaded=cpu)
            result += ckks.decrypt_vec_complex(ct_y, D).real
        else:
            ct_y = fhe_matmul_bsgs(ckks, ckks.encrypt_replicated(x0), M0, D, G, B, preencoded=pe, cpu_offloaded=cpu)
            result += ckks.decrypt_vec(ct_y, D)
    return result
