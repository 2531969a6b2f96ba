#!/usr/bin/env python3
"""Golden fixture for the primitive surface of the reference's fhe_rwkv_inference.py,
generated FROM THE REFERENCE ITSELF.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_inference.py

The reference's fhe_rwkv_inference.py is imported unmodified with a `pyPhantom` module injected into sys.modules
that is backed by the CPU oracle (the PhantomFHE fork is absent).  Its own CKKSContext
(primes [60] + [40]*depth + [60], one special prime, public-key encryption, default Galois keys) and its own
ct_pt_dot / ct_pt_weighted_sum / ct_ct_square / ct_ct_multiply (fhe_rwkv_inference.py:29-108) then run over oracle
primitives; inputs and every resulting ciphertext's limbs go to inference_primitives.npz.  tests/ replay the same call
sequences on the oracle (CPU) and on the CUDA library (GPU) and must reproduce the limbs bit for bit.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle.oracle import Oracle  # noqa: E402

SEED = bytes(range(32))


class Obj:
    def __init__(self, a, scale):
        self.a, self._scale = np.ascontiguousarray(a, dtype=np.uint64), float(scale)


class ST:
    o = None
    sk = None
    enc_counter = 0
    rotations = []


class LazyKeys(dict):
    """Galois keys generated on first use (the library default covers every power-of-two step)."""

    def __missing__(self, elt):
        self[elt] = ST.o.gen_galois_key(SEED, elt, ST.sk)
        return self[elt]


def shim():
    ph = types.ModuleType("pyPhantom")

    class scheme_type:
        ckks = 3
    ph.scheme_type = scheme_type

    class params:
        def __init__(self, s):
            self.p = 1

        def set_poly_modulus_degree(self, n):
            self.n = n

        def set_special_modulus_size(self, p):
            self.p = p

        def set_coeff_modulus(self, m):
            self.mods = list(m)
    ph.params = params
    ph.create_coeff_modulus = lambda n, bits: [int(x) for x in Oracle.create_coeff_modulus(n, list(bits))]

    class context:
        def __init__(self, p):
            ST.o = Oracle(p.n, np.array(p.mods, dtype=np.uint64), p.p)
    ph.context = context

    class public_key:
        def __init__(self, pk):
            self.pk = pk

        def encrypt_asymmetric(self, ctx, pt):
            ST.enc_counter += 1
            return Obj(ST.o.encrypt_asymmetric(ST.o.public_key_seed(SEED), ST.enc_counter, self.pk, pt.a[0]), pt._scale)

    class secret_key:
        def __init__(self, ctx):
            ST.sk = ST.o.gen_secret(SEED)

        def gen_publickey(self, ctx):
            return public_key(ST.o.gen_public_key(SEED, ST.sk))

        def gen_relinkey(self, ctx):
            return ST.o.gen_relin_key(SEED, ST.sk)

        def create_galois_keys(self, ctx):
            return LazyKeys()

        def decrypt(self, ctx, ct):
            return Obj(ST.o.decrypt(ST.sk, ct.a)[None], ct._scale)
    ph.secret_key = secret_key

    class ckks_encoder:
        def __init__(self, ctx):
            pass

        def slot_count(self):
            return ST.o.N // 2

        def encode_double_vector(self, ctx, v, scale, chain_index=1):
            return Obj(ST.o.encode(np.asarray(v, dtype=np.float64), scale, ST.o.L - chain_index + 1)[None], scale)

        def decode_double_vector(self, ctx, pt):
            return list(ST.o.decode(pt.a[0], pt._scale).real)
    ph.ckks_encoder = ckks_encoder

    ph.mod_switch_to = lambda ctx, pt, level: Obj(pt.a[:, :ST.o.L - level + 1].copy(), pt._scale)
    ph.multiply_plain = lambda ctx, ct, pt: Obj(ST.o.multiply_plain(ct.a, pt.a[0]), ct._scale * pt._scale)
    ph.add = lambda ctx, a, b: Obj(ST.o.add(a.a, b.a), a._scale)

    def rotate(ctx, ct, step, gk):
        ST.rotations.append(int(step))
        return Obj(ST.o.rotate(ct.a, step, gk), ct._scale)
    ph.rotate = rotate
    ph.rescale_to_next = lambda ctx, ct: Obj(ST.o.rescale(ct.a), ct._scale / float(ST.o.q[ct.a.shape[1] - 1]))
    ph.multiply = lambda ctx, a, b: Obj(ST.o.multiply(a.a, b.a), a._scale * b._scale)
    ph.relinearize = lambda ctx, ct, rlk: Obj(ST.o.relinearize(ct.a, rlk), ct._scale)
    return ph


def main():
    sys.modules["pyPhantom"] = shim()
    sys.path.insert(0, REF)
    import fhe_rwkv_inference as ri   # the unmodified reference

    N, depth, prime_bits, dim = 2048, 5, 40, 8
    ck = ri.CKKSContext(poly_modulus_degree=N, depth=depth, prime_bits=prime_bits)
    rng = np.random.default_rng(314)
    x = rng.standard_normal(dim)
    w1, w2 = rng.standard_normal(dim), rng.standard_normal(dim)
    mix = np.array([0.5, -1.25])
    ct = ck.encrypt(x)                                              # :46-50 (asymmetric, enc counter 1)
    d1 = ri.ct_pt_dot(ck, ct, w1, dim)                              # :66-76, chain_index 2
    d2 = ri.ct_pt_dot(ck, ct, w2, dim)
    ws = ri.ct_pt_weighted_sum(ck, [d1, d2], list(mix), 2)          # :79-94, chain_index 3
    sq = ri.ct_ct_square(ck, ws)                                    # :97-101, chain_index 4
    pr = ri.ct_ct_multiply(ck, d1, d2)                              # :104-108, chain_index 3
    vals = {"d1": ck.decrypt_slot0(d1), "d2": ck.decrypt_slot0(d2), "ws": ck.decrypt_slot0(ws),
            "sq": ck.decrypt_slot0(sq), "pr": ck.decrypt_slot0(pr)}
    t1, t2 = float(x @ w1), float(x @ w2)
    want = {"d1": t1, "d2": t2, "ws": mix[0] * t1 + mix[1] * t2, "sq": (mix[0] * t1 + mix[1] * t2) ** 2, "pr": t1 * t2}
    for k in want:
        assert abs(vals[k] - want[k]) < 1e-4, (k, vals[k], want[k])   # the reference's own kind of check (printed there)
    np.savez_compressed(os.path.join(HERE, "inference_primitives.npz"),
                        N=N, bits=np.array([60] + [prime_bits] * depth + [60]), P=1, seed=np.frombuffer(SEED, dtype=np.uint8),
                        scale=float(ck.scale), dim=dim, x=x, w1=w1, w2=w2, mix=mix, rotations=np.array(ST.rotations),
                        ct=ct.a, d1=d1.a, d2=d2.a, ws=ws.a, sq=sq.a, pr=pr.a,
                        slot0=np.array([vals[k] for k in ("d1", "d2", "ws", "sq", "pr")]),
                        slot0_float64=np.array([want[k] for k in ("d1", "d2", "ws", "sq", "pr")]))
    print("inference_primitives.npz written; slot-0 values", vals)


if __name__ == "__main__":
    main()
