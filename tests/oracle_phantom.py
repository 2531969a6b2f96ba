"""TEST INFRASTRUCTURE ONLY: the subset of the pyPhantom call surface that the host-level algorithms use
(fhe_spear_b200/bootstrap.py), answered by the CPU oracle.  Lets the same Python orchestration run on the oracle and
on the GPU library, so a whole bootstrap can be compared limb for limb (tests/test_bootstrap_*.py).  Never imported by
the product."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle  # noqa: E402


class _Obj:
    def __init__(self, be, a, scale):
        self.be, self.a, self._scale = be, np.ascontiguousarray(a, dtype=np.uint64), float(scale)

    def chain_index(self):
        return self.be.o.L - self.a.shape[1] + 1

    def scale(self):
        return self._scale

    def set_scale(self, s):
        self._scale = float(s)

    def coeff_modulus_size(self):
        return self.a.shape[1]

    def to_numpy(self):
        return self.a.copy()


class Backend:
    """ph-like namespace over one Oracle instance (ctx arguments are accepted and ignored)."""

    def __init__(self, N, bits, P, seed=bytes(range(32)), elts=()):
        mods = Oracle.create_coeff_modulus(N, list(bits))
        self.o = Oracle(N, np.array(mods, dtype=np.uint64), P)
        self.N, self.P, self.seed = N, P, seed
        self.moduli = [int(m) for m in mods]
        self.sk = self.o.gen_secret(seed)
        self.rlk = self.o.gen_relin_key(seed, self.sk)
        self.keys = {}
        self.add_galois(elts)
        self.enc_id = 0
        self.ctx = self                      # so that callers can pass backend.ctx

    # ---- keys
    def add_galois(self, elts):
        for e in elts:
            e = int(e)
            if e not in self.keys:
                self.keys[e] = self.o.gen_galois_key(self.seed, e, self.sk)

    def get_elt_from_step(self, step, n=None):
        return int(self.o.elt_from_step(int(step)))

    # ---- encoder
    def slot_count(self):
        return self.N // 2

    def encode_complex_vector(self, ctx, v, scale, chain_index=1):
        v = np.asarray(v, dtype=np.complex128)
        buf = np.zeros(self.N // 2, dtype=np.complex128)
        buf[:len(v)] = v
        return _Obj(self, self.o.encode(buf, scale, self.o.L - chain_index + 1)[None], scale)

    def encode_double_vector(self, ctx, v, scale, chain_index=1):
        return self.encode_complex_vector(ctx, np.asarray(v, dtype=np.float64), scale, chain_index)

    def decode_complex_vector(self, ctx, pt):
        return self.o.decode(pt.a[0], pt._scale)

    # ---- encryption
    def encrypt(self, pt, enc_id=None):
        if enc_id is None:
            self.enc_id += 1
            enc_id = self.enc_id
        return _Obj(self, self.o.encrypt_symmetric(self.seed, enc_id, self.sk, pt.a[0]), pt._scale)

    def decrypt(self, ct):
        return _Obj(self, self.o.decrypt(self.sk, ct.a)[None], ct._scale)

    # ---- evaluator (same semantics as fhe_spear_b200/csrc/api.cu)
    def negate(self, ctx, a):
        zero = np.zeros_like(a.a)
        return _Obj(self, self.o.sub(zero, a.a), a._scale)

    def add(self, ctx, a, b):
        assert a.a.shape == b.a.shape
        return _Obj(self, self.o.add(a.a, b.a), a._scale)

    def sub(self, ctx, a, b):
        assert a.a.shape == b.a.shape
        return _Obj(self, self.o.sub(a.a, b.a), a._scale)

    def add_plain(self, ctx, ct, pt):
        l = ct.a.shape[1]
        out = ct.a.copy()
        out[0:1] = self.o.add(ct.a[0:1], pt.a[0:1, :l])
        return _Obj(self, out, ct._scale)

    def sub_plain(self, ctx, ct, pt):
        l = ct.a.shape[1]
        out = ct.a.copy()
        out[0:1] = self.o.sub(ct.a[0:1], pt.a[0:1, :l])
        return _Obj(self, out, ct._scale)

    def multiply_plain(self, ctx, ct, pt):
        l = ct.a.shape[1]
        return _Obj(self, self.o.multiply_plain(ct.a, np.ascontiguousarray(pt.a[0, :l])), ct._scale * pt._scale)

    def multiply(self, ctx, a, b):
        assert a.a.shape == b.a.shape and a.a.shape[0] == 2
        return _Obj(self, self.o.multiply(a.a, b.a), a._scale * b._scale)

    def relinearize(self, ctx, ct, rlk=None):
        return _Obj(self, self.o.relinearize(ct.a, self.rlk), ct._scale) if ct.a.shape[0] == 3 else _Obj(self, ct.a, ct._scale)

    def rescale_to_next(self, ctx, ct):
        l = ct.a.shape[1]
        return _Obj(self, self.o.rescale(ct.a), ct._scale / float(self.moduli[l - 1]))

    def mod_switch_to_next(self, ctx, obj):
        return _Obj(self, obj.a[:, :-1], obj._scale)

    def mod_switch_to(self, ctx, obj, chain_index):
        l = self.o.L - chain_index + 1
        assert l <= obj.a.shape[1]
        return _Obj(self, obj.a[:, :l], obj._scale)

    def mod_raise(self, ctx, ct, chain_index=1):
        return _Obj(self, self.o.mod_raise(ct.a, self.o.L - chain_index + 1), ct._scale)

    def apply_galois(self, ctx, ct, elt, gk=None):
        return _Obj(self, self.o.apply_galois(ct.a, int(elt), self.keys[int(elt)]), ct._scale)

    def rotate(self, ctx, ct, step, gk=None):
        step %= self.N // 2
        if step == 0:
            return _Obj(self, ct.a, ct._scale)
        return self.apply_galois(ctx, ct, self.get_elt_from_step(step))

    def hoisting(self, ctx, ct, gk, steps):
        out = []
        for s in steps:
            e = self.get_elt_from_step(int(s) % (self.N // 2))
            out.append(_Obj(self, self.o.hoisted_rotation(ct.a, e, self.keys[e]), ct._scale))
        return out
