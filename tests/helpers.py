"""Shared fixtures for the parity tests: one small CKKS parameter set, built both in the oracle
(oracle/oracle.py, the checker) and -- on a GPU -- in the product (pyPhantom over libspear_b200.so)."""
import numpy as np

from oracle.oracle import Oracle

SEED = bytes(range(32))


def bsgs_params(D):
    G = int(np.ceil(np.sqrt(D)))
    return G, int(np.ceil(D / G))


def rolled_diagonals(W, D, G, B):
    """diag_k[j] = W[j, (j+k) % D], rows of giant group g rolled right by g*G
    (reference scripts/bootstrap_generation.py:198-203, 361-369)."""
    j = np.arange(D)
    d = np.stack([W[j, (j + k) % D] for k in range(D)])
    for g in range(1, B):
        s, e = g * G, min((g + 1) * G, D)
        d[s:e] = np.roll(d[s:e], g * G, axis=1)
    return d


def tile(v, slots):
    reps, rem = divmod(slots, len(v))
    return np.concatenate([np.tile(v, reps), v[:rem]])


class Setup:
    """Oracle-side parameter set + keys; `gpu()` builds the same thing in the product."""

    def __init__(self, N=2048, bits=(59,) * 7, P=2, seed=SEED):
        self.N, self.P, self.seed = N, P, seed
        self.q = Oracle.create_coeff_modulus(N, list(bits))
        self.L = len(bits) - P
        self.o = Oracle(N, self.q, P)
        self.sk = self.o.gen_secret(seed)
        self.keys = {}
        self.scale = 2.0 ** 40 if min(bits) < 50 else 2.0 ** 59

    def key(self, elt):
        if elt not in self.keys:
            self.keys[elt] = self.o.gen_galois_key(self.seed, elt, self.sk)
        return self.keys[elt]

    def keys_for_steps(self, steps):
        for s in steps:
            self.key(self.o.elt_from_step(s))
        return self.keys

    # ---- product side ------------------------------------------------------------------------
    def gpu(self, steps=()):
        from fhe_spear_b200 import pyPhantom as ph
        parms = ph.params(ph.scheme_type.ckks)
        parms.set_poly_modulus_degree(self.N)
        parms.set_special_modulus_size(self.P)
        parms.set_coeff_modulus([int(x) for x in self.q])
        parms.set_galois_elts([ph.get_elt_from_step(s, self.N) for s in steps] + [2 * self.N - 1])
        ctx = ph.context(parms)
        sk = ph.secret_key(ctx, seed=self.seed)
        return ph, ctx, sk
