"""ctypes front-end of the CPU oracle (oracle/spear_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(fhe_spear_b200/) never imports this module.  Parity unpinned: see the header of
spear_oracle.c and DESIGN.md.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libspear_oracle.so")

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
f64p = C.POINTER(C.c_double)


def build(force=False):
    src = os.path.join(_HERE, "spear_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


def _p(a, t=u64p):
    return a.ctypes.data_as(t)


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a


class Oracle:
    """One CKKS parameter set.  All polynomials are numpy uint64 arrays."""

    def __init__(self, N, moduli, P):
        self.lib = C.CDLL(build())
        L = self.lib
        L.orc_ctx_create.restype = C.c_void_p
        L.orc_ctx_create.argtypes = [C.c_uint64, C.c_int, C.c_int, u64p]
        L.orc_galois_elt_from_step.restype = C.c_uint64
        L.orc_galois_elt_from_step.argtypes = [C.c_uint64, C.c_int]
        L.orc_num_digits.restype = C.c_int
        L.orc_ctx_psi.restype = C.c_uint64
        self.N, self.K, self.P, self.L = int(N), len(moduli), int(P), len(moduli) - int(P)
        self.q = _u64(moduli)
        self.ctx = C.c_void_p(L.orc_ctx_create(self.N, self.K, self.P, _p(self.q)))
        if not self.ctx:
            raise RuntimeError("orc_ctx_create failed")

    def __del__(self):
        try:
            self.lib.orc_ctx_free(self.ctx)
        except Exception:
            pass

    @staticmethod
    def set_threads(n):
        """OpenMP team size of the oracle (explicit: torchrun exports OMP_NUM_THREADS=1); returns the size in effect"""
        lib = C.CDLL(build())
        lib.orc_set_threads.restype = C.c_int
        return int(lib.orc_set_threads(int(n)))

    # ---- parameters -------------------------------------------------
    @staticmethod
    def create_coeff_modulus(N, bits):
        lib = C.CDLL(build())
        out = np.zeros(len(bits), dtype=np.uint64)
        b = (C.c_int * len(bits))(*bits)
        rc = lib.orc_create_coeff_modulus(C.c_uint64(N), b, len(bits), _p(out))
        if rc:
            raise RuntimeError("not enough primes")
        return out

    def elt_from_step(self, step):
        return int(self.lib.orc_galois_elt_from_step(C.c_uint64(self.N), int(step)))

    def num_digits(self, l):
        return int(self.lib.orc_num_digits(self.ctx, int(l)))

    def psi(self, limb):
        return int(self.lib.orc_ctx_psi(self.ctx, int(limb)))

    def rows(self, l, ext=False):
        return l + (self.P if ext else 0)

    # ---- transforms -------------------------------------------------
    def ntt_fwd(self, limb, a):
        a = _u64(a).copy()
        self.lib.orc_ntt_fwd(self.ctx, int(limb), _p(a))
        return a

    def ntt_inv(self, limb, a):
        a = _u64(a).copy()
        self.lib.orc_ntt_inv(self.ctx, int(limb), _p(a))
        return a

    def apply_galois_ntt(self, elt, a):
        a = _u64(a)
        out = np.empty_like(a)
        self.lib.orc_apply_galois_ntt(self.ctx, C.c_uint32(elt), _p(a), _p(out))
        return out

    # ---- PRNG -------------------------------------------------------
    @staticmethod
    def seed_words(seed):
        """32-byte seed -> 8 uint32 ChaCha key words."""
        if isinstance(seed, int):
            seed = seed.to_bytes(32, "little")
        return np.frombuffer(bytes(seed), dtype=np.uint32).copy()

    def prng_words(self, seed, nonce, first, count):
        out = np.empty(count, dtype=np.uint64)
        k = self.seed_words(seed)
        self.lib.orc_prng_words(_p(k, u32p), C.c_uint64(nonce), C.c_uint64(first), C.c_uint64(count), _p(out))
        return out

    # ---- encode / decode -------------------------------------------
    def encode(self, values, scale, l, ext=False, n=None):
        """values: complex/real array of n/2 slots (n = N unless sub-ring)."""
        n = self.N if n is None else int(n)
        v = np.asarray(values)
        assert v.shape[0] == n // 2 and n & (n - 1) == 0 and n <= self.N, "ring size must be a power of two"
        re = np.ascontiguousarray(v.real, dtype=np.float64)
        im = np.ascontiguousarray(v.imag, dtype=np.float64) if np.iscomplexobj(v) else None
        rows = self.rows(l, ext)
        out = np.empty((rows, n), dtype=np.uint64)
        rc = self.lib.orc_encode(self.ctx, C.c_uint64(n), _p(re, f64p), _p(im, f64p) if im is not None else None,
                                 C.c_double(scale), int(l), int(bool(ext)), _p(out))
        if rc:
            raise ValueError("encoded coefficient too large")
        return out

    def decode(self, pt, scale):
        pt = _u64(pt)
        l = pt.shape[0]
        re = np.empty(self.N // 2)
        im = np.empty(self.N // 2)
        self.lib.orc_decode(self.ctx, _p(pt), int(l), C.c_double(scale), _p(re, f64p), _p(im, f64p))
        return re + 1j * im

    # ---- keys -------------------------------------------------------
    def gen_secret(self, seed):
        sk = np.empty((self.K, self.N), dtype=np.uint64)
        self.lib.orc_gen_secret(self.ctx, _p(self.seed_words(seed), u32p), _p(sk))
        return sk

    def key_shape(self):
        return (self.num_digits(self.L), 2, self.K, self.N)

    def gen_galois_key(self, seed, elt, sk):
        key = np.empty(self.key_shape(), dtype=np.uint64)
        self.lib.orc_gen_galois_key(self.ctx, _p(self.seed_words(seed), u32p), C.c_uint32(elt), _p(sk), _p(key))
        return key

    def gen_relin_key(self, seed, sk):
        key = np.empty(self.key_shape(), dtype=np.uint64)
        self.lib.orc_gen_relin_key(self.ctx, _p(self.seed_words(seed), u32p), _p(sk), _p(key))
        return key

    def gen_public_key(self, seed, sk):
        pk = np.empty((2, self.K, self.N), dtype=np.uint64)
        self.lib.orc_gen_public_key(self.ctx, _p(self.seed_words(seed), u32p), _p(sk), _p(pk))
        return pk

    def public_key_seed(self, seed):
        """Seed carried by the public key of secret seed `seed` (the `seed` argument of encrypt_asymmetric)."""
        out = np.empty(8, dtype=np.uint32)
        self.lib.orc_public_key_seed(_p(self.seed_words(seed), u32p), _p(out, u32p))
        return out.tobytes()

    # ---- encryption -------------------------------------------------
    def encrypt_symmetric(self, seed, enc_id, sk, pt):
        pt = _u64(pt)
        l = pt.shape[0]
        ct = np.empty((2, l, self.N), dtype=np.uint64)
        self.lib.orc_encrypt_symmetric(self.ctx, _p(self.seed_words(seed), u32p), C.c_uint64(enc_id), _p(sk),
                                       _p(pt), int(l), _p(ct))
        return ct

    def encrypt_asymmetric(self, seed, enc_id, pk, pt):
        """`seed` is the public key's own seed (public_key_seed of the secret seed), not the secret seed."""
        pt = _u64(pt)
        l = pt.shape[0]
        ct = np.empty((2, l, self.N), dtype=np.uint64)
        self.lib.orc_encrypt_asymmetric(self.ctx, _p(self.seed_words(seed), u32p), C.c_uint64(enc_id), _p(_u64(pk)),
                                        _p(pt), int(l), _p(ct))
        return ct

    def decrypt(self, sk, ct):
        ct = _u64(ct)
        size, l, _ = ct.shape
        pt = np.empty((l, self.N), dtype=np.uint64)
        self.lib.orc_decrypt(self.ctx, _p(sk), _p(ct), int(size), int(l), _p(pt))
        return pt

    # ---- evaluator --------------------------------------------------
    def add(self, a, b):
        a, b = _u64(a), _u64(b)
        o = np.empty_like(a)
        self.lib.orc_add(self.ctx, a.shape[1], a.shape[0], _p(a), _p(b), _p(o))
        return o

    def sub(self, a, b):
        a, b = _u64(a), _u64(b)
        o = np.empty_like(a)
        self.lib.orc_sub(self.ctx, a.shape[1], a.shape[0], _p(a), _p(b), _p(o))
        return o

    def multiply_plain(self, ct, pt):
        ct, pt = _u64(ct), _u64(pt)
        o = np.empty_like(ct)
        self.lib.orc_multiply_plain(self.ctx, ct.shape[1], ct.shape[0], _p(ct), _p(pt), _p(o))
        return o

    def multiply(self, a, b):
        a, b = _u64(a), _u64(b)
        o = np.empty((3,) + a.shape[1:], dtype=np.uint64)
        self.lib.orc_multiply(self.ctx, a.shape[1], _p(a), _p(b), _p(o))
        return o

    def relinearize(self, ct3, rlk):
        ct3 = _u64(ct3)
        o = np.empty((2,) + ct3.shape[1:], dtype=np.uint64)
        self.lib.orc_relinearize(self.ctx, ct3.shape[1], _p(ct3), _p(rlk), _p(o))
        return o

    def rescale(self, ct):
        ct = _u64(ct)
        size, l, _ = ct.shape
        o = np.empty((size, l - 1, self.N), dtype=np.uint64)
        self.lib.orc_rescale(self.ctx, int(l), int(size), _p(ct), _p(o))
        return o

    def mod_raise(self, ct, l):
        ct = _u64(ct)
        size = ct.shape[0]
        assert ct.shape[1] == 1
        o = np.empty((size, l, self.N), dtype=np.uint64)
        self.lib.orc_mod_raise(self.ctx, int(size), _p(ct), int(l), _p(o))
        return o

    def decompose(self, cin):
        cin = _u64(cin)
        l = cin.shape[0]
        E = np.empty((self.num_digits(l), l + self.P, self.N), dtype=np.uint64)
        self.lib.orc_decompose(self.ctx, int(l), _p(cin), _p(E))
        return E

    def moddown(self, x):
        x = _u64(x)
        l = x.shape[0] - self.P
        o = np.empty((l, self.N), dtype=np.uint64)
        self.lib.orc_moddown(self.ctx, int(l), _p(x), _p(o))
        return o

    def keyswitch(self, cin, key):
        cin = _u64(cin)
        l = cin.shape[0]
        o = np.empty((2, l, self.N), dtype=np.uint64)
        self.lib.orc_keyswitch(self.ctx, int(l), _p(cin), _p(key), _p(o[0]), _p(o[1]))
        return o

    def apply_galois(self, ct, elt, key):
        ct = _u64(ct)
        l = ct.shape[1]
        o = np.empty_like(ct)
        self.lib.orc_apply_galois(self.ctx, int(l), _p(ct), C.c_uint32(elt), _p(key), _p(o))
        return o

    def hoisted_rotation(self, ct, elt, key):
        ct = _u64(ct)
        o = np.empty_like(ct)
        self.lib.orc_hoisted_rotation(self.ctx, int(ct.shape[1]), _p(ct), C.c_uint32(elt), _p(key), _p(o))
        return o

    def rotate(self, ct, step, keys):
        elt = self.elt_from_step(step)
        return self.apply_galois(ct, elt, keys[elt])

    # ---- BSGS -------------------------------------------------------
    def bsgs_exact(self, ct_baby, pts, G, B, D, keys):
        """ct_baby: (G,2,l,N); pts: (D,l,N); keys: {elt: key}.  Reference op order."""
        ct_baby, pts = _u64(ct_baby), _u64(pts)
        l = ct_baby.shape[2]
        elts = np.zeros(B, dtype=np.uint32)
        kp = (u64p * B)()
        for g in range(1, B):
            elts[g] = self.elt_from_step(g * G)
            kp[g] = _p(keys[int(elts[g])])
        out = np.empty((2, l - 1, self.N), dtype=np.uint64)
        self.lib.orc_bsgs_exact(self.ctx, int(l), _p(ct_baby), _p(pts), int(G), int(B), int(D),
                                _p(elts, u32p), kp, _p(out))
        return out

    def _hoisted_args(self, G, B, keys):
        belts = np.zeros(G, dtype=np.uint32)
        gelts = np.zeros(B, dtype=np.uint32)
        bk = (u64p * G)()
        gk = (u64p * B)()
        for b in range(1, G):
            belts[b] = self.elt_from_step(b)
            bk[b] = _p(keys[int(belts[b])])
        for g in range(1, B):
            gelts[g] = self.elt_from_step(g * G)
            if int(gelts[g]) in keys:
                gk[g] = _p(keys[int(gelts[g])])
        return belts, bk, gelts, gk

    def bsgs_hoisted(self, ct, diags, G, B, D, keys):
        """ct: (2,l,N); diags: (D, l+P, N >> rshift); keys: {elt: key}."""
        return self.bsgs_finish(self.bsgs_hoisted_partial(ct, diags, G, B, D, keys))

    def bsgs_hoisted_partial(self, ct, diags, G, B, D, keys, g_first=0, g_stride=1):
        """Shard accumulator (2, l+P, N) in basis Q_l*P; diags holds the rows of groups g_first, g_first+g_stride, ..."""
        ct, diags = _u64(ct), _u64(diags)
        l = ct.shape[1]
        dn = diags.shape[2]
        rshift = (self.N // dn).bit_length() - 1
        belts, bk, gelts, gk = self._hoisted_args(G, B, keys)
        R = np.empty((2, l + self.P, self.N), dtype=np.uint64)
        self.lib.orc_bsgs_hoisted_partial(self.ctx, int(l), _p(ct), _p(diags), int(rshift), int(G), int(B), int(D),
                                          int(g_first), int(g_stride), _p(belts, u32p), bk, _p(gelts, u32p), gk, _p(R))
        return R

    def bsgs_finish(self, R):
        R = _u64(R)
        l = R.shape[1] - self.P
        out = np.empty((2, l - 1, self.N), dtype=np.uint64)
        self.lib.orc_bsgs_finish(self.ctx, int(l), _p(R), _p(out))
        return out

    def reduce_rows(self, x, ext):
        x = _u64(x).copy()
        l = x.shape[1] - (self.P if ext else 0)
        self.lib.orc_reduce_rows(self.ctx, int(l), int(bool(ext)), int(x.shape[0]), _p(x))
        return x
