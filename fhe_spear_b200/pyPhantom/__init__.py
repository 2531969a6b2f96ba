"""pyPhantom -- drop-in for the module the reference builds from gpu/phantom_binding.cu.

Same names, argument order and error behaviour as the pybind11 module the reference scripts
import (`import pyPhantom as ph`, scripts/bootstrap_generation.py:14, test_fully_enc_bsgs.py:14,
fhe_rwkv_inference.py:9, fhe_common.py:14), including the fork-only symbols those scripts call
(SURVEY.md section 8b).  Every call goes straight to libspear_b200.so (sm_100a CUDA); there is no
CPU path.  Put the parent directory of this package on sys.path ahead of PHANTOM_PATH.

Extras that the reference does not have (parity hooks, fast path): `secret_key(ctx, seed=...)`,
`ciphertext.to_numpy()/from_numpy()`, `diagonal_set`, `bsgs_hoisted`.
"""
import ctypes as C
import enum
import os

import numpy as np

try:
    from .. import _native as _n
except ImportError:   # imported as top-level `pyPhantom` (sys.path points at fhe_spear_b200/), as the reference does
    import sys as _sys
    _sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from fhe_spear_b200 import _native as _n

_lib = _n.lib
_check = _n.check


def _bootstrapper():
    """fhe_spear_b200.bootstrap.Bootstrapper, whether this module was imported as fhe_spear_b200.pyPhantom or -- the
    way the reference's scripts do -- as top-level `pyPhantom`"""
    try:
        from ..bootstrap import Bootstrapper
    except ImportError:
        from fhe_spear_b200.bootstrap import Bootstrapper
    return Bootstrapper

__version__ = _lib.spear_version().decode()


# ---- enums  [ref: phantom_binding.cu:56-76] -------------------------------------------------------
class scheme_type(enum.IntEnum):
    none = 0
    bgv = 1
    bfv = 2
    ckks = 3


class mul_tech_type(enum.IntEnum):
    none = 0
    behz = 1
    hps = 2
    hps_overq = 3
    hps_overq_leveled = 4


class sec_level_type(enum.IntEnum):
    none = 0
    tc128 = 128
    tc192 = 192
    tc256 = 256


for _e in (scheme_type, mul_tech_type, sec_level_type):   # pybind11 .export_values()
    for _m in _e:
        globals().setdefault(_m.name, _m)


# ---- parameters  [ref: phantom_binding.cu:78-92] ---------------------------------------------------
class modulus:
    def __init__(self, value):
        self._value = int(value)

    def value(self):
        return self._value

    def __int__(self):
        return self._value

    def __repr__(self):
        return f"modulus({self._value})"


def create_coeff_modulus(poly_modulus_degree, bit_sizes):
    bits = [int(b) for b in bit_sizes]
    out = (C.c_uint64 * len(bits))()
    _check(_lib.spear_create_coeff_modulus(int(poly_modulus_degree), (C.c_int * len(bits))(*bits), len(bits), out))
    return [modulus(v) for v in out]


def get_elt_from_step(step, poly_modulus_degree):
    return int(_lib.spear_get_elt_from_step(int(step), int(poly_modulus_degree)))


def get_elts_from_steps(steps, poly_modulus_degree):
    return [get_elt_from_step(s, poly_modulus_degree) for s in steps]


class params:
    def __init__(self, scheme):
        if scheme != scheme_type.ckks:
            raise RuntimeError("only scheme_type.ckks is implemented")
        self.scheme = scheme
        self.poly_modulus_degree = 0
        self.coeff_modulus = []
        self.special_modulus_size = 1
        self.galois_elts = None
        self.mul_tech = mul_tech_type.none

    def set_mul_tech(self, t):
        self.mul_tech = t

    def set_poly_modulus_degree(self, n):
        self.poly_modulus_degree = int(n)

    def set_special_modulus_size(self, p):
        self.special_modulus_size = int(p)

    def set_galois_elts(self, elts):
        self.galois_elts = [int(e) for e in elts]

    def set_coeff_modulus(self, mods):
        self.coeff_modulus = [int(m) for m in mods]

    def set_plain_modulus(self, m):
        raise RuntimeError("plain modulus is a BFV/BGV parameter; only CKKS is implemented")


class cuda_stream:   # [ref: phantom_binding.cu:94] exported by the reference, never passed by any script
    pass


# ---- context  [ref: phantom_binding.cu:97-98] ------------------------------------------------------
class context:
    def __init__(self, parms, device=None):
        n = parms.poly_modulus_degree
        mods = parms.coeff_modulus
        if not mods:
            raise RuntimeError("coeff_modulus is not set")
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", os.environ.get("SPEAR_DEVICE", "0")))
        arr = (C.c_uint64 * len(mods))(*mods)
        h = C.c_void_p()
        _check(_lib.spear_context_create(n, arr, len(mods), parms.special_modulus_size, int(device), C.byref(h)))
        self._h = h
        self.N = n
        self.moduli = list(mods)
        self.P = parms.special_modulus_size
        self.L = len(mods) - self.P
        self.device = int(device)
        if parms.galois_elts is None:   # library default: all power-of-two steps, both directions, and conjugation
            elts = {2 * n - 1}
            s = 1
            while s < n // 2:
                elts.add(get_elt_from_step(s, n))
                elts.add(get_elt_from_step(-s, n))
                s *= 2
            self.galois_elts = sorted(elts)
        else:
            self.galois_elts = list(parms.galois_elts)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.spear_context_destroy(h)

    def synchronize(self):
        _check(_lib.spear_context_sync(self._h))

    def timer_start(self):
        _check(_lib.spear_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        _check(_lib.spear_timer_stop(self._h, C.byref(ms)))
        return ms.value

    PROFILE_CLASSES = ("ks_inner", "pmac", "ntt_fwd_a", "modup", "moddown", "rescale", "ks_baby_fused", "ntt_ks_fused",
                       "ntt_fwd_b", "ntt_inv_a", "ntt_inv_b", "sum_groups", "peer_wait", "peer_reduce")

    def profile(self, on=True):
        """Bracket every launch of each kernel class with a CUDA event pair (bench.py roofline line)."""
        _check(_lib.spear_profile_enable(self._h, int(bool(on))))

    def profile_read(self):
        n = len(self.PROFILE_CLASSES)
        ms, cnt = (C.c_double * n)(), (C.c_uint64 * n)()
        _check(_lib.spear_profile_read(self._h, ms, cnt, n))
        return {k: {"ms": ms[i], "launches": int(cnt[i])} for i, k in enumerate(self.PROFILE_CLASSES)}

    def reserve(self, nbytes):
        """grow the device memory pool to at least `nbytes` now (kept by the pool; see spear_mem_reserve)"""
        _check(_lib.spear_mem_reserve(self._h, int(nbytes)))

    def mem_info(self):
        u, r = C.c_uint64(), C.c_uint64()
        _check(_lib.spear_mem_info(self._h, C.byref(u), C.byref(r)))
        return u.value, r.value


# ---- plaintext / ciphertext  [ref: phantom_binding.cu:158-163 + fork-only accessors] -----------------
class _obj:
    def __init__(self, ctx=None, handle=None):
        self._ctx = ctx
        self._h = handle

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.spear_obj_destroy(h)

    def _info(self):
        if not self._h:
            raise RuntimeError("empty object")
        size, limbs, ext, ring, ci = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        scale = C.c_double()
        _check(_lib.spear_obj_info(self._h, C.byref(size), C.byref(limbs), C.byref(ext), C.byref(ring),
                                   C.byref(scale), C.byref(ci)))
        return size.value, limbs.value, ext.value, ring.value, scale.value, ci.value

    def chain_index(self):
        return self._info()[5]

    def scale(self):
        return self._info()[4]

    def coeff_modulus_size(self):
        return self._info()[1]

    def set_scale(self, s):
        _check(_lib.spear_obj_set_scale(self._h, float(s)))

    def size(self):
        return self._info()[0]

    # parity hooks: raw limbs [size][limbs(+P)][ring_n]
    def to_numpy(self, out=None):
        size, limbs, ext, ring, _, _ = self._info()
        rows = limbs + (self._ctx.P if ext else 0)
        if out is None:
            out = np.empty((size, rows, ring), dtype=np.uint64)
        _check(_lib.spear_obj_export(self._h, out.ctypes.data_as(C.c_void_p), out.size))
        return out

    @classmethod
    def from_numpy(cls, ctx, arr, scale, ext=False):
        arr = np.ascontiguousarray(arr, dtype=np.uint64)
        size, rows, ring = arr.shape
        limbs = rows - (ctx.P if ext else 0)
        h = C.c_void_p()
        _check(_lib.spear_obj_import(ctx._h, arr.ctypes.data_as(C.c_void_p), size, limbs, int(ext), ring,
                                     float(scale), C.byref(h)))
        ctx.synchronize()   # arr may be pageable and freed by the caller
        return cls(ctx, h)


def pinned_empty(shape, dtype=np.uint64):
    """numpy array over page-locked host memory (host legs of the end-to-end path, offloaded plaintexts); the memory is
    released when the last view of the array is garbage-collected."""
    import weakref
    nbytes = max(1, int(np.prod(shape)) * np.dtype(dtype).itemsize)
    p = C.c_void_p()
    _check(_lib.spear_pinned_alloc(nbytes, C.byref(p)))
    buf = (C.c_char * nbytes).from_address(p.value)
    weakref.finalize(buf, _lib.spear_pinned_free, p)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class plaintext(_obj):
    pass


class ciphertext(_obj):
    pass


def _new(cls, ctx, fn, *args):
    h = C.c_void_p()
    _check(fn(ctx._h, *args, C.byref(h)))
    return cls(ctx, h)


# ---- keys  [ref: phantom_binding.cu:100-122] -------------------------------------------------------
class relin_key:
    def __init__(self, ctx=None, handle=None):
        self._ctx, self._h = ctx, handle

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.spear_kswitch_key_destroy(h)

    def to_numpy(self):
        c = self._ctx
        beta = (c.L + c.P - 1) // c.P
        out = np.empty((beta, 2, c.L + c.P, c.N), dtype=np.uint64)
        _check(_lib.spear_kswitch_key_export(self._h, out.ctypes.data_as(C.c_void_p), out.size))
        return out


class galois_key:
    def __init__(self, ctx=None, handle=None):
        self._ctx, self._h = ctx, handle

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.spear_galois_keys_destroy(h)

    def has(self, elt):
        return bool(_lib.spear_galois_keys_has(self._h, int(elt)))

    def to_numpy(self, elt):
        c = self._ctx
        beta = (c.L + c.P - 1) // c.P
        out = np.empty((beta, 2, c.L + c.P, c.N), dtype=np.uint64)
        _check(_lib.spear_galois_key_export(self._h, int(elt), out.ctypes.data_as(C.c_void_p), out.size))
        return out


class public_key:
    def __init__(self, ctx=None, handle=None):
        self._ctx, self._h = ctx, handle
        self._enc = 0

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.spear_public_key_destroy(h)

    def encrypt_asymmetric(self, ctx, pt, enc_id=None):
        if enc_id is None:
            enc_id, self._enc = self._enc, self._enc + 1
        return _new(ciphertext, ctx, _lib.spear_encrypt_asymmetric, self._h, pt._h, int(enc_id))

    def to_numpy(self):
        c = self._ctx
        out = np.empty((2, c.L + c.P, c.N), dtype=np.uint64)
        _check(_lib.spear_public_key_export(self._h, out.ctypes.data_as(C.c_void_p), out.size))
        return out


class secret_key:
    def __init__(self, ctx, seed=None):
        if seed is None:
            # SPEAR_SEED is for reproducible TESTS only: a fixed seed with the per-process counter below restarting at 0
            # reuses (seed, nonce) pairs across runs, which breaks semantic security.  Production keys use os.urandom.
            env = os.environ.get("SPEAR_SEED")
            if env:
                import warnings
                warnings.warn("SPEAR_SEED fixes the secret seed: encryption randomness repeats across runs (tests only)",
                              RuntimeWarning, stacklevel=2)
            seed = bytes.fromhex(env).ljust(32, b"\0")[:32] if env else os.urandom(32)
        if isinstance(seed, int):
            seed = seed.to_bytes(32, "little")
        if len(seed) != 32:
            raise RuntimeError("seed must be 32 bytes")
        self._ctx = ctx
        self.seed = bytes(seed)
        self._enc = 0
        h = C.c_void_p()
        _check(_lib.spear_secret_key_create(ctx._h, self.seed, C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.spear_secret_key_destroy(h)

    def gen_publickey(self, ctx):
        return _new(public_key, ctx, _lib.spear_gen_public_key, self._h)

    def gen_relinkey(self, ctx):
        return _new(relin_key, ctx, _lib.spear_gen_relin_key, self._h)

    def create_galois_keys(self, ctx, elts=None):
        elts = list(ctx.galois_elts if elts is None else elts)
        arr = (C.c_uint32 * len(elts))(*elts)
        h = C.c_void_p()
        _check(_lib.spear_gen_galois_keys(ctx._h, self._h, arr, len(elts), C.byref(h)))
        return galois_key(ctx, h)

    def add_galois_keys(self, ctx, gk, elts):
        elts = list(elts)
        arr = (C.c_uint32 * len(elts))(*elts)
        _check(_lib.spear_galois_keys_add(ctx._h, self._h, gk._h, arr, len(elts)))

    def reserve_enc_ids(self, count=1):
        """`count` fresh encryption ids from the key's ONE monotonic counter (the nonce of the randomness streams of
        encrypt_symmetric).  Callers that pass enc_id explicitly -- e.g. the ranks of a sharded block, which must all
        form the same ciphertext -- draw their ids here, so no id is ever handed out twice under one key."""
        base, self._enc = self._enc, self._enc + int(count)
        return base

    def encrypt_symmetric(self, ctx, pt, enc_id=None):
        if enc_id is None:
            enc_id = self.reserve_enc_ids(1)
        return _new(ciphertext, ctx, _lib.spear_encrypt_symmetric, self._h, pt._h, int(enc_id))

    def decrypt(self, ctx, ct):
        return _new(plaintext, ctx, _lib.spear_decrypt, self._h, ct._h)

    # ---- the client legs of a projection round trip, three launches each (csrc/client.cu) ----------------------
    def encrypt_vector(self, ctx, values, scale, replicate=True, enc_id=None):
        """encode + encrypt_symmetric of `values` (real or complex; tiled over all slots when `replicate`, else
        zero-padded) in one call: the same fresh ciphertext, limb for limb, as
        encrypt_symmetric(encode_complex_vector(replicated values)) with the same enc_id."""
        v = np.ascontiguousarray(np.asarray(values, dtype=np.complex128).ravel())
        if enc_id is None:
            enc_id = self.reserve_enc_ids(1)
        return _new(ciphertext, ctx, _lib.spear_encrypt_vector, self._h, v.view(np.float64).ctypes.data_as(_n.f64p),
                    int(v.size), int(bool(replicate)), float(scale), int(enc_id))

    def decrypt_decode(self, ctx, ct, count=None):
        """first `count` slots of decode(decrypt(ct)) as a complex array, in one call (bit-identical to the two steps)"""
        count = ctx.N // 2 if count is None else int(count)
        out = np.empty(count, dtype=np.complex128)
        _check(_lib.spear_decrypt_decode(ctx._h, self._h, ct._h, out.view(np.float64).ctypes.data_as(_n.f64p), count))
        return out

    def to_numpy(self):
        c = self._ctx
        out = np.empty((c.L + c.P, c.N), dtype=np.uint64)
        _check(_lib.spear_secret_key_export(self._h, out.ctypes.data_as(C.c_void_p), out.size))
        return out


# ---- encoder  [ref: phantom_binding.cu:138-156; fork-only batch forms] --------------------------------
def _as_complex_rows(values, slots):
    v = np.asarray(values)
    if v.ndim == 1:
        v = v[None, :]
    if v.shape[1] > slots:
        raise RuntimeError(f"too many values: {v.shape[1]} > {slots} slots")
    buf = np.zeros((v.shape[0], slots), dtype=np.complex128)
    buf[:, :v.shape[1]] = v
    return buf


class ckks_encoder:
    def __init__(self, ctx):
        self._ctx = ctx

    def slot_count(self):
        return self._ctx.N // 2

    def _encode(self, ctx, values, scale, chain_index, ext=False, ring_n=None):
        ring_n = ctx.N if ring_n is None else ring_n
        buf = _as_complex_rows(values, ring_n // 2)
        count = buf.shape[0]
        outs = (C.c_void_p * count)()
        _check(_lib.spear_encode(ctx._h, buf.view(np.float64).ctypes.data_as(_n.f64p), count, ring_n, float(scale),
                                 int(chain_index), int(ext), outs))
        return [plaintext(ctx, C.c_void_p(h)) for h in outs]

    def encode_double_vector(self, ctx, values, scale, chain_index=1):
        return self._encode(ctx, np.asarray(values, dtype=np.float64), scale, chain_index)[0]

    def encode_complex_vector(self, ctx, values, scale, chain_index=1):
        return self._encode(ctx, np.asarray(values, dtype=np.complex128), scale, chain_index)[0]

    def encode_double_vector_batch(self, ctx, values, scale, chain_index=1):
        return self._encode(ctx, np.asarray(values, dtype=np.float64), scale, chain_index)

    def encode_complex_vector_batch(self, ctx, values, scale, chain_index=1):
        return self._encode(ctx, np.asarray(values, dtype=np.complex128), scale, chain_index)

    def _decode(self, ctx, pt):
        out = np.empty(ctx.N // 2, dtype=np.complex128)
        _check(_lib.spear_decode(ctx._h, pt._h, out.view(np.float64).ctypes.data_as(_n.f64p)))
        return out

    def decode_array(self, ctx, pt):
        """All slots as a numpy complex128 array (extension: the reference's decode_* return Python lists, which costs
        more than the decode itself at 16384 slots; the host mirror uses this when the encoder offers it)."""
        return self._decode(ctx, pt)

    def decode_double_vector(self, ctx, pt):
        return self._decode(ctx, pt).real.tolist()

    def decode_complex_vector(self, ctx, pt):
        return self._decode(ctx, pt).tolist()


class batch_encoder:   # [ref: phantom_binding.cu:128-136] BFV/BGV only
    def __init__(self, ctx):
        raise RuntimeError("batch_encoder is a BFV/BGV feature; only CKKS is implemented")


# ---- evaluator  [ref: phantom_binding.cu:165-205] ----------------------------------------------------
def negate(ctx, ct):
    return _new(ciphertext, ctx, _lib.spear_negate, ct._h)


def add(ctx, a, b):
    return _new(ciphertext, ctx, _lib.spear_add, a._h, b._h)


def sub(ctx, a, b, negate=False):
    if negate:
        a, b = b, a
    return _new(ciphertext, ctx, _lib.spear_sub, a._h, b._h)


def add_plain(ctx, ct, pt):
    return _new(ciphertext, ctx, _lib.spear_add_plain, ct._h, pt._h)


def sub_plain(ctx, ct, pt):
    return _new(ciphertext, ctx, _lib.spear_sub_plain, ct._h, pt._h)


def add_many(ctx, cts, dst=None):
    acc = cts[0]
    for ct in cts[1:]:
        acc = add(ctx, acc, ct)
    return acc


def multiply(ctx, a, b):
    return _new(ciphertext, ctx, _lib.spear_multiply, a._h, b._h)


def multiply_plain(ctx, ct, pt):
    return _new(ciphertext, ctx, _lib.spear_multiply_plain, ct._h, pt._h)


def relinearize(ctx, ct, rlk):
    return _new(ciphertext, ctx, _lib.spear_relinearize, ct._h, rlk._h)


def multiply_and_relin(ctx, a, b, rlk):
    return relinearize(ctx, multiply(ctx, a, b), rlk)


def rescale_to_next(ctx, ct):
    return _new(ciphertext, ctx, _lib.spear_rescale_to_next, ct._h)


def mod_switch_to_next(ctx, obj):
    return _new(type(obj), ctx, _lib.spear_mod_switch_to_next, obj._h)


def mod_raise(ctx, ct, chain_index=1):
    """ModRaise (bootstrapping entry): a one-limb ciphertext re-read modulo the primes of `chain_index`."""
    return _new(ciphertext, ctx, _lib.spear_mod_raise, ct._h, int(chain_index))


def mod_switch_to(ctx, obj, chain_index):
    cur = obj.chain_index()
    if chain_index < cur:
        raise RuntimeError(f"cannot mod-switch from chain_index {cur} back to {chain_index}")
    if chain_index == cur:   # functional style: still hand back a fresh object
        return type(obj).from_numpy(ctx, obj.to_numpy(), obj.scale())
    while obj.chain_index() < chain_index:
        obj = mod_switch_to_next(ctx, obj)
    return obj


def apply_galois(ctx, ct, elt, gk):
    return _new(ciphertext, ctx, _lib.spear_apply_galois, ct._h, int(elt), gk._h)


def _naf(x):
    out, i = [], 0
    while x:
        if x & 1:
            d = 2 - (x & 3)
            out.append(d * (1 << i))
            x -= d
        x >>= 1
        i += 1
    return out


def rotate(ctx, ct, step, gk):
    slots = ctx.N // 2
    step = int(step)
    if step % slots == 0:
        return mod_switch_to(ctx, ct, ct.chain_index())
    elt = get_elt_from_step(step, ctx.N)
    if gk.has(elt):
        return apply_galois(ctx, ct, elt, gk)
    # no key for this step: compose power-of-two rotations (non-adjacent form), as SEAL's rotate_internal
    s = step % slots
    if s > slots // 2:
        s -= slots
    for part in _naf(abs(s)):
        part = part if s > 0 else -part
        e = get_elt_from_step(part, ctx.N)
        if not gk.has(e):
            raise RuntimeError(f"no Galois key for step {step} (element {elt}) nor for its power-of-two parts")
        ct = apply_galois(ctx, ct, e, gk)
    return ct


def hoisting(ctx, ct, gk, steps):
    """Rotations of one ciphertext by several steps sharing one decomposition (hoisted)."""
    elts = [get_elt_from_step(int(s), ctx.N) for s in steps]
    outs = (C.c_void_p * len(elts))()
    _check(_lib.spear_hoisted_rotations(ctx._h, ct._h, (C.c_uint32 * len(elts))(*elts), len(elts), gk._h, outs))
    return [ciphertext(ctx, C.c_void_p(h)) for h in outs]


# ---- fork-only fused ops  [ref: scripts/bootstrap_generation.py:242, 339-357, 449, 459] -----------------
def bsgs_multiply_accumulate(ctx, ct_baby, pts, G, B, D, gk):
    nb, npt = len(ct_baby), len(pts)
    cb = (C.c_void_p * nb)(*[c._h for c in ct_baby])
    pp = (C.c_void_p * npt)(*[p._h for p in pts])
    return _new(ciphertext, ctx, _lib.spear_bsgs_multiply_accumulate, cb, nb, pp, npt, int(G), int(B), int(D), gk._h)


def offload_plaintexts(pts):
    """list[plaintext] -> (data, chain_index, scale, coeff_modulus_size, poly_modulus_degree)
    [ref: fork-only, scripts/bootstrap_generation.py:339].  `data` lives in page-locked host memory, so the way back
    (upload_plaintexts, bsgs_from_cpu) runs at PCIe speed and asynchronously; one synchronisation for the whole batch."""
    size, limbs, ext, ring, scale, ci = pts[0]._info()
    n = len(pts)
    data = pinned_empty((n, limbs, ring))
    _check(_lib.spear_objs_export(pts[0]._ctx._h, (C.c_void_p * n)(*[p._h for p in pts]), n,
                                  data.ctypes.data_as(C.c_void_p), limbs * ring))
    return data, ci, scale, limbs, ring


def upload_plaintexts(data, chain_index, scale, coeff_modulus_size, poly_modulus_degree, ctx=None):
    ctx = ctx or _default_ctx()
    data = np.asarray(data, dtype=np.uint64).reshape(-1, coeff_modulus_size, poly_modulus_degree)
    return [plaintext.from_numpy(ctx, data[k:k + 1], scale) for k in range(data.shape[0])]


def bsgs_from_cpu(ctx, ct_baby, data, ci, sc, cms, pmd, G, B, D, gk):
    """BSGS over offloaded plaintexts [ref: fork-only, scripts/bootstrap_generation.py:449]: the diagonals stream from
    host memory through a two-slot device ring, overlapped with the arithmetic (spear_bsgs_from_host)."""
    data = np.ascontiguousarray(np.asarray(data, dtype=np.uint64).reshape(-1, cms, pmd))
    nb = len(ct_baby)
    cb = (C.c_void_p * nb)(*[c._h for c in ct_baby])
    return _new(ciphertext, ctx, _lib.spear_bsgs_from_host, cb, nb, data.ctypes.data_as(C.c_void_p), int(data.shape[0]),
                int(cms), float(sc), int(G), int(B), int(D), gk._h)


def bsgs_complete_from_cpu(ctx, ct_x, data, ci, sc, cms, pmd, G, B, D, gk):
    ct_baby = [ct_x] + [rotate(ctx, ct_x, b, gk) for b in range(1, G)]
    return bsgs_from_cpu(ctx, ct_baby, data, ci, sc, cms, pmd, G, B, D, gk)


_last_ctx = None


def _default_ctx():
    if _last_ctx is None:
        raise RuntimeError("upload_plaintexts needs a context (none created yet)")
    return _last_ctx


_ctx_init = context.__init__


def _ctx_init_track(self, *a, **k):
    global _last_ctx
    _ctx_init(self, *a, **k)
    import weakref
    _last_ctx = weakref.proxy(self)


context.__init__ = _ctx_init_track


# ---- fast path (not in the reference): pre-encoded diagonal sets + hoisted BSGS -------------------------
class diagonal_set:
    """Pre-rotated period-D diagonals encoded in basis Q_l*P, sub-ring compressed when D is a power of two.

    `diags` is the full (D, D) array of pre-rotated diagonals; with shard=(rank, world) only the giant
    groups g = rank, rank + world, ... are encoded and stored (giant-step sharding over GPUs)."""

    def __init__(self, ctx, diags, G, B, scale, chain_index=1, compress=True, shard=(0, 1)):
        d = np.asarray(diags, dtype=np.complex128)
        D = d.shape[1]
        if d.shape != (D, D):
            raise RuntimeError("diagonal_set expects a (D, D) array of pre-rotated diagonals")
        first, stride = int(shard[0]), int(shard[1])
        rows = [k for g in range(first, int(B), stride) for k in range(g * G, min((g + 1) * G, D))]
        part = np.ascontiguousarray(d[rows]) if rows else np.zeros((0, D), dtype=np.complex128)
        compress = bool(compress) and (D & (D - 1)) == 0 and D >= 2
        h = C.c_void_p()
        _check(_lib.spear_diagset_encode_shard(ctx._h, part.view(np.float64).ctypes.data_as(_n.f64p), len(rows), D,
                                               int(G), int(B), first, stride, float(scale), int(chain_index),
                                               int(compress), C.byref(h)))
        self._ctx, self._h = ctx, h
        self.D, self.G, self.B, self.shard, self.rows = D, int(G), int(B), (first, stride), rows

    @staticmethod
    def _view(M):
        """(array kept alive, pointer, pitch in doubles, transposed) of a 2-D float64 array WITHOUT copying it when one
        of its strides is the item size: row-pitched views (W[lo:hi, :], W[:, lo:hi]) and their transposes."""
        M = np.asarray(M, dtype=np.float64)
        if M.ndim != 2:
            raise RuntimeError("diagonal_set.from_matrix expects a 2-D matrix")
        r, c = M.shape
        s0, s1 = M.strides
        if (s1 == 8 or c <= 1) and s0 % 8 == 0 and s0 >= 8 * c:
            return M, s0 // 8 if r > 1 else max(c, 1), 0
        if (s0 == 8 or r <= 1) and s1 % 8 == 0 and s1 >= 8 * r:
            return M, s1 // 8 if c > 1 else max(r, 1), 1
        M = np.ascontiguousarray(M)
        return M, max(c, 1), 0

    @classmethod
    def from_matrix(cls, ctx, M, G, B, scale, chain_index=1, compress=True, shard=(0, 1), M_imag=None, D=None):
        """Same set from the matrix of y = M x (optionally M + i M_imag for the complex packing): diagonal extraction,
        the +gG pre-rotation and the slot tiling run on the device; only the matrix crosses the bus.  M may be any
        row- or column-pitched VIEW (a chunk W[:, lo:hi].T of a larger weight matrix is taken as it lies in host memory:
        no transposed host copy) and may be smaller than D x D (zero-padded)."""
        M, pitch, tr = cls._view(M)
        D = int(D) if D is not None else max(M.shape)
        if M.shape[0] > D or M.shape[1] > D:
            raise RuntimeError("diagonal_set.from_matrix: matrix larger than D x D")
        Mi = None
        if M_imag is not None:
            Mi, pitch_i, tr_i = cls._view(M_imag)
            if Mi.shape[0] > D or Mi.shape[1] > D:
                raise RuntimeError("diagonal_set.from_matrix: M_imag larger than D x D")
            if Mi.shape != M.shape or (pitch_i, tr_i) != (pitch, tr):   # different shapes / layouts: zero-padded compact copies
                def pad(A):
                    out = np.zeros((D, D))
                    out[:A.shape[0], :A.shape[1]] = A
                    return out
                M, Mi = pad(M), pad(Mi)
                pitch, tr = D, 0
        first, stride = int(shard[0]), int(shard[1])
        compress = bool(compress) and (D & (D - 1)) == 0 and D >= 2
        h = C.c_void_p()
        _check(_lib.spear_diagset_encode_matrix_view(ctx._h, M.ctypes.data_as(_n.f64p),
                                                     Mi.ctypes.data_as(_n.f64p) if Mi is not None else None, D,
                                                     int(M.shape[0]), int(M.shape[1]), int(pitch), int(tr), int(G), int(B),
                                                     first, stride, float(scale), int(chain_index), int(compress), C.byref(h)))
        self = cls.__new__(cls)
        self._ctx, self._h = ctx, h
        self.D, self.G, self.B, self.shard = D, int(G), int(B), (first, stride)
        self.rows = [k for g in range(first, int(B), stride) for k in range(g * int(G), min((g + 1) * int(G), D))]
        return self

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.spear_diagset_destroy(h)

    def info(self):
        D, G, B, limbs, ring = (C.c_int() for _ in range(5))
        scale, nbytes = C.c_double(), C.c_uint64()
        _check(_lib.spear_diagset_info(self._h, C.byref(D), C.byref(G), C.byref(B), C.byref(limbs), C.byref(ring),
                                       C.byref(scale), C.byref(nbytes)))
        return dict(D=D.value, G=G.value, B=B.value, limbs=limbs.value, ring_n=ring.value, scale=scale.value,
                    bytes=nbytes.value)

    @staticmethod
    def share(limbs, P, N, rank, world):
        """(row0, row1, col0, col1): the rows of the limbs + P and the coefficient columns served by `rank` in the first
        phase of a two-phase mat-vec (spear_split_share: rows; beyond four ranks, row groups of two ranks x column halves)"""
        r0, nr, c0, nc = (C.c_int() for _ in range(4))
        _check(_lib.spear_split_share(int(rank), int(world), int(limbs) + int(P), int(N), C.byref(r0), C.byref(nr),
                                      C.byref(c0), C.byref(nc)))
        return r0.value, r0.value + nr.value, c0.value, c0.value + nc.value

    def slice_rows(self, rank, world):
        """The share of the set rank `rank` of `world` holds in a two-phase mat-vec (bsgs_split): every giant group, its
        rows (and column half) only -- 1/world of the set's bytes."""
        if self.shard != (0, 1):
            raise RuntimeError("slice_rows: expected a set holding every giant group")
        i = self.info()
        r0, r1, c0, c1 = self.share(i["limbs"], self._ctx.P, self._ctx.N, rank, world)
        h = C.c_void_p()
        _check(_lib.spear_diagset_slice_share(self._ctx._h, self._h, int(rank), int(world), C.byref(h)))
        out = type(self).__new__(type(self))
        out._ctx, out._h = self._ctx, h
        out.D, out.G, out.B, out.shard, out.rows = self.D, self.G, self.B, self.shard, self.rows
        shift = (self._ctx.N // i["ring_n"]).bit_length() - 1
        out.row_slice, out.col_slice = (r0, r1), (c0 >> shift, c1 >> shift)      # columns in stored values
        return out

    def to_numpy(self):
        i = self.info()
        nrows = i["limbs"] + self._ctx.P
        ncols = i["ring_n"]
        if getattr(self, "row_slice", None):
            nrows = self.row_slice[1] - self.row_slice[0]
            ncols = self.col_slice[1] - self.col_slice[0]
        out = np.empty((len(self.rows), nrows, ncols), dtype=np.uint64)
        if out.size == 0:
            return out
        _check(_lib.spear_diagset_export(self._h, out.ctypes.data_as(C.c_void_p), out.size))
        return out


def bsgs_hoisted_batch(ctx, cts, diag_sets, gk):
    """Independent mat-vecs (e.g. r, k, v of one block) run concurrently on separate streams."""
    n = len(cts)
    outs = (C.c_void_p * n)()
    _check(_lib.spear_bsgs_hoisted_batch(ctx._h, (C.c_void_p * n)(*[c._h for c in cts]),
                                         (C.c_void_p * n)(*[d._h for d in diag_sets]), n, gk._h, outs))
    return [ciphertext(ctx, C.c_void_p(h)) for h in outs]


def bsgs_hoisted_shared(ctx, ct, diag_sets, gk):
    """Several diagonal sets (same D, G, level) times ONE ciphertext -- the chunk pairs of a D -> F projection: the baby
    steps are computed once, as the reference does (scripts/bootstrap_generation.py:575-600).  Same limbs as
    [bsgs_hoisted(ctx, ct, d, gk) for d in diag_sets]."""
    n = len(diag_sets)
    outs = (C.c_void_p * n)()
    _check(_lib.spear_bsgs_hoisted_shared(ctx._h, ct._h, (C.c_void_p * n)(*[d._h for d in diag_sets]), n, gk._h, outs))
    return [ciphertext(ctx, C.c_void_p(h)) for h in outs]


def bsgs_hoisted_batch_host(ctx, host_in, scale, diag_sets, gk, host_out):
    """Serving form of bsgs_hoisted_batch: host_in[i] = (2, l, N) uint64 ciphertext limbs in host memory (pinned_empty for
    asynchronous copies), host_out[i] = (2, l - 1, N) receives result i; uploads, mat-vecs and downloads of the items are
    pipelined over the engine's streams.  Returns the scales of the results."""
    n = len(host_in)
    l = int(host_in[0].shape[1])
    for a, b in zip(host_in, host_out):
        if a.dtype != np.uint64 or b.dtype != np.uint64 or not a.flags.c_contiguous or not b.flags.c_contiguous:
            raise RuntimeError("bsgs_hoisted_batch_host: contiguous uint64 arrays expected")
        if a.shape != (2, l, ctx.N) or b.shape != (2, l - 1, ctx.N):
            raise RuntimeError("bsgs_hoisted_batch_host: in (2, l, N), out (2, l - 1, N) expected")
    scales = (C.c_double * n)()
    _check(_lib.spear_bsgs_hoisted_batch_host(ctx._h, (C.c_void_p * n)(*[a.ctypes.data for a in host_in]), l, float(scale),
                                              (C.c_void_p * n)(*[d._h for d in diag_sets]), n, gk._h,
                                              (C.c_void_p * n)(*[b.ctypes.data for b in host_out]), scales))
    return list(scales)


def bsgs_hoisted_partial(ctx, ct, shard, gk):
    """This shard's accumulator in basis Q_l*P (ciphertext object with the special limbs)."""
    return _new(ciphertext, ctx, _lib.spear_bsgs_hoisted_partial, ct._h, shard._h, gk._h)


def bsgs_hoisted_partial_batch(ctx, cts, shards, gk):
    """Shard accumulators of several independent mat-vecs, computed concurrently on separate streams."""
    n = len(cts)
    outs = (C.c_void_p * n)()
    _check(_lib.spear_bsgs_hoisted_partial_batch(ctx._h, (C.c_void_p * n)(*[c._h for c in cts]),
                                                 (C.c_void_p * n)(*[d._h for d in shards]), n, gk._h, outs))
    return [ciphertext(ctx, C.c_void_p(h)) for h in outs]


def bsgs_split(ctx, ct, rows, gk, window, slot=0):
    """Two-phase mat-vec over the rank group of `window` (include/spear_b200.h): baby steps + diagonal MAC on this rank's
    rows, accumulators scattered to the owners of the giant groups over NVLink, giant steps of this rank's groups.
    Returns this rank's accumulator in basis Q_l*P (sum over the ranks = the full accumulator)."""
    return _new(ciphertext, ctx, _lib.spear_bsgs_split, ct._h, rows._h, gk._h, window._h, int(slot))


def bsgs_split_batch(ctx, cts, rows, gk, window, slot0=0):
    n = len(cts)
    outs = (C.c_void_p * n)()
    _check(_lib.spear_bsgs_split_batch(ctx._h, (C.c_void_p * n)(*[c._h for c in cts]),
                                       (C.c_void_p * n)(*[d._h for d in rows]), n, gk._h, window._h, int(slot0), outs))
    return [ciphertext(ctx, C.c_void_p(h)) for h in outs]


def bsgs_split_shared(ctx, ct, rows, gk, window, slot0=0):
    """bsgs_split for several row-sliced sets multiplying ONE ciphertext: this rank's baby steps once; set i exchanges
    through window slot slot0 + i.  Returns this rank's accumulators."""
    n = len(rows)
    outs = (C.c_void_p * n)()
    _check(_lib.spear_bsgs_split_shared(ctx._h, ct._h, (C.c_void_p * n)(*[d._h for d in rows]), n, gk._h, window._h,
                                        int(slot0), outs))
    return [ciphertext(ctx, C.c_void_p(h)) for h in outs]


def bsgs_split_selftest(ctx, ct, row_sets, gk):
    """One-GPU emulation of bsgs_split over len(row_sets) ranks (test hook): the summed accumulator."""
    n = len(row_sets)
    return _new(ciphertext, ctx, _lib.spear_bsgs_split_selftest, ct._h, (C.c_void_p * n)(*[d._h for d in row_sets]), n, gk._h)


def bsgs_finish(ctx, acc):
    """ModDown + rescale of a (summed) accumulator; the accumulator's contents are consumed."""
    return _new(ciphertext, ctx, _lib.spear_bsgs_finish, acc._h)


def reduce_inplace(ctx, obj):
    _check(_lib.spear_obj_reduce(ctx._h, obj._h))


def device_ptr(obj):
    return int(_lib.spear_obj_device_ptr(obj._h))


class peer_window:
    """Device memory of this rank mapped by the other ranks of its group (CUDA IPC) for the fused NVLink exchange of
    shard accumulators (include/spear_b200.h "peer-memory exchange").  `handle` (64 bytes) goes to every other rank by
    any host channel; connect() takes the handles of all ranks in rank order."""

    def __init__(self, ctx, rank, world, slot_bytes, slots=3):
        self._ctx, self.rank, self.world, self.slots = ctx, rank, world, slots
        buf = C.create_string_buffer(64)
        h = C.c_void_p()
        _check(_lib.spear_peer_window_create(ctx._h, rank, world, int(slot_bytes), slots, buf, C.byref(h)))
        self._h, self.handle = h, bytes(buf.raw)

    def connect(self, handles):
        if len(handles) != self.world or any(len(h) != 64 for h in handles):
            raise RuntimeError("peer_window.connect: one 64-byte handle per rank expected")
        _check(_lib.spear_peer_window_connect(self._ctx._h, self._h, b"".join(handles)))

    def allreduce(self, acc, slot=0):
        """acc <- sum over the ranks of the group, mod q, in place; asynchronous on the context's stream"""
        _check(_lib.spear_peer_allreduce(self._ctx._h, self._h, slot, acc._h))

    def status(self):
        return int(_lib.spear_peer_window_status(self._h))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.spear_peer_window_destroy(h)


def peer_selftest(ctx, accs):
    """One-GPU emulation of the peer exchange over len(accs) ranks (test hook): every acc becomes the sum mod q."""
    arr = (C.c_void_p * len(accs))(*[a._h for a in accs])
    _check(_lib.spear_peer_selftest(ctx._h, arr, len(accs)))


def bsgs_hoisted(ctx, ct, diags, gk):
    return _new(ciphertext, ctx, _lib.spear_bsgs_hoisted, ct._h, diags._h, gk._h)


# ---- bootstrapping  [ref: fork-only ckks_bootstrapper, scripts/bootstrap_generation.py:72-75,110-116] ---
class ckks_bootstrapper:
    """CKKS bootstrapping (fhe_spear_b200/bootstrap.py): ModRaise -> CoeffToSlot -> EvalMod -> SlotToCoeff over the
    evaluator of this module.  The reference's implementation sits in the absent phantom-fhe fork; interface and
    call order follow its call sites, the algorithm is this build's own."""

    DEFAULT_BUDGET = (2, 2)

    def __init__(self, encoder):
        self.encoder = encoder
        self.impl = None

    @staticmethod
    def get_galois_elements(poly_degree, slots, level_budget):
        """Galois elements of every rotation the linear transforms need, plus conjugation (slots = 0: all N/2)."""
        Bootstrapper = _bootstrapper()
        steps = Bootstrapper.rotation_steps(int(poly_degree), tuple(level_budget or ckks_bootstrapper.DEFAULT_BUDGET))
        return sorted(set(get_elts_from_steps(steps, poly_degree)) | {2 * int(poly_degree) - 1})

    @staticmethod
    def get_bootstrap_depth(level_budget, poly_degree=32768):
        """Levels between the raised ciphertext and the bootstrapped one (the reference passes the budget only: the
        default degree is the largest ring of its configurations, smaller rings need at most as many levels)."""
        Bootstrapper = _bootstrapper()
        return Bootstrapper.depth_for(int(poly_degree), tuple(level_budget or ckks_bootstrapper.DEFAULT_BUDGET))

    def setup(self, ctx, level_budget=None):
        Bootstrapper = _bootstrapper()
        import sys
        self.impl = Bootstrapper(sys.modules[__name__], ctx, self.encoder, ctx.N, ctx.moduli, ctx.P,
                                 tuple(level_budget or self.DEFAULT_BUDGET))

    def keygen(self, ctx, sk):
        if self.impl is None:
            raise RuntimeError("ckks_bootstrapper.keygen: call setup(ctx, level_budget) first")
        elts = self.get_galois_elements(ctx.N, 0, self.impl.budget)
        self.impl.gk = sk.create_galois_keys(ctx, elts)
        self.impl.rlk = sk.gen_relinkey(ctx)

    def bootstrap(self, ctx, ct):
        """Returns the refreshed ciphertext with one rescale pending (scale ~ (2^bits)^2), as the reference's call site
        expects: it calls rescale_to_next right after (test_fully_enc_bsgs.py:251-253)."""
        if self.impl is None or self.impl.gk is None:
            raise RuntimeError("ckks_bootstrapper.bootstrap: call setup() and keygen() first")
        return self.impl.bootstrap(ct, rescale_last=False)
