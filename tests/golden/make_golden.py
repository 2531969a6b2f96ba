#!/usr/bin/env python3
"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's scripts/bootstrap_generation.py unmodified, with a `pyPhantom`
module injected into sys.modules that is backed by the CPU oracle (the reference's own
PhantomFHE dependency is absent), and records

  bsgs_plan.json       compute_bsgs_params, compute_bsgs_galois_elements (steps handed to
                       get_elts_from_steps), compute_rotation_galois_elements
  bsgs_diagonals.npz   _extract_diagonals + _batch_encode_diags_real: the exact slot vectors the
                       reference hands to the encoder (pre-rotation by +gG, tiling to all slots)
  bsgs_loop.npz        the reference's own Python BSGS loop (fhe_matmul_bsgs fallback,
                       bootstrap_generation.py:464-484) and fhe_projection_bsgs complex /
                       conjugate packing executed over oracle primitives: inputs, keys' seed and
                       the output ciphertext limbs.  tests check orc_bsgs_exact (and, on the GPU,
                       spear_bsgs_multiply_accumulate) against these limbs bit for bit.
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.oracle import Oracle  # noqa: E402

SEED = bytes(range(32))


# ---- a pyPhantom look-alike over the oracle: just enough for the reference's BSGS layer ----------
class Ct:
    def __init__(self, a, scale):
        self.a, self._scale = a, scale

    def chain_index(self):
        return ST.o.L - self.a.shape[1] + 1

    def scale(self):
        return self._scale

    def coeff_modulus_size(self):
        return self.a.shape[1]


class Pt(Ct):
    pass


class ST:   # state shared by the shim
    o = None
    sk = None
    keys = {}
    steps_seen = []
    enc_counter = 0
    captured_vectors = []


def _shim():
    ph = types.ModuleType("pyPhantom")

    class scheme_type:
        ckks = 3
    ph.scheme_type = scheme_type

    class params:
        def __init__(self, s):
            pass

        def set_poly_modulus_degree(self, n):
            self.n = n

        def set_special_modulus_size(self, p):
            self.p = p

        def set_galois_elts(self, e):
            self.elts = list(e)

        def set_coeff_modulus(self, m):
            self.mods = list(m)
    ph.params = params
    ph.create_coeff_modulus = lambda n, bits: [int(x) for x in Oracle.create_coeff_modulus(n, list(bits))]

    def get_elts_from_steps(steps, n):
        ST.steps_seen.append((int(n), [int(s) for s in steps]))
        return [pow(5, int(s), 2 * int(n)) for s in steps]
    ph.get_elts_from_steps = get_elts_from_steps

    class context:
        def __init__(self, p):
            ST.o = Oracle(p.n, np.array(p.mods, dtype=np.uint64), p.p)
            self.elts = p.elts
    ph.context = context

    class secret_key:
        def __init__(self, ctx):
            ST.sk = ST.o.gen_secret(SEED)
            self.ctx = ctx

        def gen_relinkey(self, ctx):
            return None

        def create_galois_keys(self, ctx):
            for e in ctx.elts:
                ST.keys[e] = ST.o.gen_galois_key(SEED, e, ST.sk)
            return ST.keys

        def encrypt_symmetric(self, ctx, pt):
            ST.enc_counter += 1
            return Ct(ST.o.encrypt_symmetric(SEED, ST.enc_counter, ST.sk, pt.a[0]), pt._scale)

        def decrypt(self, ctx, ct):
            return Pt(ST.o.decrypt(ST.sk, ct.a)[None], ct._scale)
    ph.secret_key = secret_key

    class ckks_encoder:
        def __init__(self, ctx):
            pass

        def slot_count(self):
            return ST.o.N // 2

        def encode_double_vector(self, ctx, v, scale, chain_index=1):
            ST.captured_vectors.append(np.asarray(v, dtype=np.float64).copy())
            return Pt(ST.o.encode(np.asarray(v, dtype=np.float64), scale, ST.o.L - chain_index + 1)[None], scale)

        def encode_complex_vector(self, ctx, v, scale, chain_index=1):
            ST.captured_vectors.append(np.asarray(v, dtype=np.complex128).copy())
            return Pt(ST.o.encode(np.asarray(v, dtype=np.complex128), scale, ST.o.L - chain_index + 1)[None], scale)

        def decode_double_vector(self, ctx, pt):
            return list(ST.o.decode(pt.a[0], pt._scale).real)

        def decode_complex_vector(self, ctx, pt):
            return list(ST.o.decode(pt.a[0], pt._scale))
    ph.ckks_encoder = ckks_encoder

    def mod_switch_to(ctx, pt, level):
        l = ST.o.L - level + 1
        return Pt(pt.a[:, :l].copy(), pt._scale)
    ph.mod_switch_to = mod_switch_to
    ph.multiply_plain = lambda ctx, ct, pt: Ct(ST.o.multiply_plain(ct.a, pt.a[0]), ct._scale * pt._scale)
    ph.add = lambda ctx, a, b: Ct(ST.o.add(a.a, b.a), a._scale)
    ph.rotate = lambda ctx, ct, step, gk: Ct(ST.o.rotate(ct.a, step, gk), ct._scale)
    ph.rescale_to_next = lambda ctx, ct: Ct(ST.o.rescale(ct.a), ct._scale / float(ST.o.q[ct.a.shape[1] - 1]))
    # no bsgs_multiply_accumulate / encode_*_batch / offload: the reference falls back to its Python loops
    return ph


def main():
    sys.modules["pyPhantom"] = _shim()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "scripts"))
    import bootstrap_generation as bg   # the unmodified reference

    # ---- plan ---------------------------------------------------------------------------------
    plan = {"source": "reference scripts/bootstrap_generation.py:18-42", "params": [], "galois": []}
    for D in (8, 20, 64, 1024, 2048, 4096, 8192):
        G, B = bg.compute_bsgs_params(D)
        plan["params"].append({"D": D, "G": G, "B": B})
    for N, D, max_dim in ((256, 8, 8), (1024, 20, 32), (32768, 2048, 8192)):
        ST.steps_seen.clear()
        elts = bg.compute_bsgs_galois_elements(N, D)
        (n_seen, steps), = ST.steps_seen
        rot = bg.compute_rotation_galois_elements(N, max_dim)
        n_pow2 = 0
        while (1 << n_pow2) <= max_dim:
            n_pow2 += 1
        plan["galois"].append({"N": N, "D": D, "steps": steps, "bsgs_elts": [int(e) for e in elts],
                               "rot_elts": sorted(int(e) for e in rot), "n_pow2": n_pow2})
    with open(os.path.join(HERE, "bsgs_plan.json"), "w") as f:
        json.dump(plan, f, indent=1)

    # ---- diagonals as the reference hands them to the encoder -------------------------------------
    class FakeCkks:
        pass
    arrays = {}
    rng = np.random.default_rng(2024)
    for D, slots in ((8, 32), (20, 64)):
        W = rng.standard_normal((D, D))
        G, B = bg.compute_bsgs_params(D)
        captured = {}

        class Enc:
            def encode_double_vector_batch(self, ctx, vecs, scale, chain_index=1):
                captured["v"] = np.array(vecs, copy=True)
                return [None] * len(vecs)
        ck = FakeCkks()
        ck.encoder, ck.ctx, ck.diag_scale = Enc(), None, 1.0
        bg._batch_encode_diags_real(ck, bg._extract_diagonals(W, D), D, G, slots, 1)
        arrays[f"W_{D}"] = W
        arrays[f"rolled_{D}"] = captured["v"]
    np.savez_compressed(os.path.join(HERE, "bsgs_diagonals.npz"), **arrays)

    # ---- the reference's own BSGS loop over oracle primitives ---------------------------------------
    N, L0, P, D = 256, 3, 1, 8
    ckks = bg.CKKSBootstrapContext(poly_degree=N, L0=L0, prime_bits=59, special_mod_size=P,
                                   max_rot_dim=D, bsgs_dim=[D], skip_bootstrap=True)
    G, B = bg.compute_bsgs_params(D)
    rng = np.random.default_rng(7)
    W = rng.standard_normal((D, D)) * 0.5
    W2 = rng.standard_normal((D, D)) * 0.5
    x = rng.standard_normal(D)
    out = {"N": N, "L0": L0, "P": P, "D": D, "W": W, "W2": W2, "x": x, "seed": np.frombuffer(SEED, dtype=np.uint8),
           "moduli": ST.o.q.copy()}
    ST.enc_counter = 0
    ct_x = ckks.encrypt_replicated(x)
    out["ct_x"] = ct_x.a
    out["enc_id_x"] = ST.enc_counter
    y = bg.fhe_matmul_bsgs(ckks, ct_x, W, D, G, B)            # real diagonals, Python loop :464-484
    out["ct_y_real"] = y.a
    out["y_real_dec"] = ckks.decrypt_vec(y, D)
    yc = bg.fhe_matmul_bsgs_complex(ckks, ct_x, W, W2, D, G, B)   # complex-packed pair, :488-542
    out["ct_y_complex"] = yc.a
    out["y_complex_dec"] = ckks.decrypt_vec_complex(yc, D)
    assert np.abs(out["y_real_dec"] - W @ x).max() < 1e-9
    assert np.abs(out["y_complex_dec"] - (W @ x + 1j * (W2 @ x))).max() < 1e-9
    # whole projection helper incl. encrypt/decrypt, D->2D (complex packing) and 2D->D (conjugate packing)
    Wk = rng.standard_normal((D, 2 * D)) * 0.5
    Wv = rng.standard_normal((2 * D, D)) * 0.5
    x2 = rng.standard_normal(2 * D)
    out["Wk"], out["Wv"], out["x2"] = Wk, Wv, x2
    out["enc_id_before_proj"] = ST.enc_counter
    out["proj_up"] = bg.fhe_projection_bsgs(ckks, x, Wk, D, 2 * D)
    out["proj_down"] = bg.fhe_projection_bsgs(ckks, x2, Wv, 2 * D, D)
    out["proj_same"] = bg.fhe_projection_bsgs(ckks, x, W, D, D)
    assert np.abs(out["proj_up"] - x @ Wk).max() < 1e-9
    assert np.abs(out["proj_down"] - x2 @ Wv).max() < 1e-9
    assert np.abs(out["proj_same"] - x @ W).max() < 1e-9
    np.savez_compressed(os.path.join(HERE, "bsgs_loop.npz"), **out)

    # ---- the reference's float64 RWKV-7 block (plaintext_block :902-980) on seeded random weights -------------
    from fhe_spear_b200.rwkv_block import RWKVBlockWeights as MyWeights
    Db, Fb, Hb, Sb = 16, 64, 2, 8
    blk = {}
    rngb = np.random.default_rng(99)
    xb = rngb.standard_normal(Db)
    xpa, xpf = rngb.standard_normal(Db) * 0.1, rngb.standard_normal(Db) * 0.1
    st = rngb.standard_normal((Hb, Sb, Sb)) * 0.1
    vf = rngb.standard_normal(Db) * 0.1
    for idx in (0, 1):
        mine = MyWeights.random(Db, Fb, Hb, Sb, block_idx=idx, seed=5 + idx)
        ref_block = object.__new__(bg.RWKVBlockWeights)          # the reference class, filled with the same tensors
        ref_block.__dict__.update(mine.__dict__)
        o = bg.plaintext_block(ref_block, xb, xpa, xpf, st, vf)
        for name, val in zip(("x", "xpa", "xpf", "state", "v_first"), o):
            blk[f"b{idx}_{name}"] = np.asarray(val)
    blk.update(x=xb, x_prev_att=xpa, x_prev_ffn=xpf, state=st, v_first=vf, dims=np.array([Db, Fb, Hb, Sb]))
    np.savez_compressed(os.path.join(HERE, "rwkv_block.npz"), **blk)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
