"""CPU tests of the oracle itself (oracle/spear_oracle.c): algebraic properties, decrypt
correctness against float64, and the golden fixtures generated from the reference's own
numpy code (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from helpers import SEED, Setup, bsgs_params, rolled_diagonals, tile

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def S():
    return Setup(N=1024, bits=(59,) * 6, P=2)


def test_ntt_is_negacyclic_convolution(S):
    o, q = S.o, int(S.q[0])
    rng = np.random.default_rng(0)
    a = rng.integers(0, q, S.N, dtype=np.uint64)
    b = np.zeros(S.N, dtype=np.uint64)
    b[[0, 1, S.N - 1]] = [3, 5, 7]
    assert np.array_equal(o.ntt_inv(0, o.ntt_fwd(0, a)), a)
    prod = np.array([(int(x) * int(y)) % q for x, y in zip(o.ntt_fwd(0, a), o.ntt_fwd(0, b))], dtype=np.uint64)
    got = [int(v) for v in o.ntt_inv(0, prod)]
    A = [int(v) for v in a]
    exp = [(3 * A[k] + 5 * (A[k - 1] if k >= 1 else -A[S.N - 1]) + 7 * (-A[k + 1] if k + 1 < S.N else 0)
            + (7 * A[0] if k == S.N - 1 else 0)) % q for k in range(S.N)]
    # x^(N-1) * a: coefficient k gets a[k+1-N]... computed directly instead:
    exp = [0] * S.N
    for i, c in ((0, 3), (1, 5), (S.N - 1, 7)):
        for j in range(S.N):
            k = i + j
            if k < S.N:
                exp[k] = (exp[k] + c * A[j]) % q
            else:
                exp[k - S.N] = (exp[k - S.N] - c * A[j]) % q
    assert got == exp


def test_subring_ntt_is_prefix_of_full_ntt(S):
    """NTT_N(p(X^r))[i] == NTT_{N/r}(p)[i // r] -- the identity behind compressed diagonals."""
    o, q = S.o, int(S.q[1])
    rng = np.random.default_rng(1)
    import ctypes as C
    for r in (2, 8, 64):
        n = S.N // r
        small = rng.integers(0, q, n, dtype=np.uint64)
        full = np.zeros(S.N, dtype=np.uint64)
        full[::r] = small
        sub = small.copy()
        o.lib.orc_ntt_fwd_sub(o.ctx, 1, sub.ctypes.data_as(C.c_void_p), C.c_uint64(n))
        assert np.array_equal(o.ntt_fwd(1, full), np.repeat(sub, r))


def test_galois_permutation_keeps_aligned_blocks(S):
    o = S.o
    idx = np.arange(S.N, dtype=np.uint64)
    for elt in (5, 25, pow(5, 37, 2 * S.N), 2 * S.N - 1):
        src = o.apply_galois_ntt(elt, idx).astype(np.int64)
        assert sorted(src) == list(range(S.N))
        assert np.all(src.reshape(-1, 32) // 32 == (src.reshape(-1, 32) // 32)[:, :1])


def test_encrypt_rotate_multiply_decrypt(S):
    o = S.o
    rng = np.random.default_rng(2)
    z = rng.standard_normal(S.N // 2) + 1j * rng.standard_normal(S.N // 2)
    pt = o.encode(z, S.scale, S.L)
    assert np.abs(o.decode(pt, S.scale) - z).max() < 1e-12
    ct = o.encrypt_symmetric(SEED, 1, S.sk, pt)
    assert np.abs(o.decode(o.decrypt(S.sk, ct), S.scale) - z).max() < 1e-12
    S.keys_for_steps([1, 5, -3])
    for step in (1, 5, -3):
        r = o.rotate(ct, step, S.keys)
        assert np.abs(o.decode(o.decrypt(S.sk, r), S.scale) - np.roll(z, -step)).max() < 1e-9
        h = o.hoisted_rotation(ct, o.elt_from_step(step), S.key(o.elt_from_step(step)))
        assert np.abs(o.decode(o.decrypt(S.sk, h), S.scale) - np.roll(z, -step)).max() < 1e-9
    cj = o.apply_galois(ct, 2 * S.N - 1, S.key(2 * S.N - 1))
    assert np.abs(o.decode(o.decrypt(S.sk, cj), S.scale) - np.conj(z)).max() < 1e-9
    rlk = o.gen_relin_key(SEED, S.sk)
    sq = o.rescale(o.relinearize(o.multiply(ct, ct), rlk))
    assert np.abs(o.decode(o.decrypt(S.sk, sq), S.scale ** 2 / float(S.q[S.L - 1])) - z * z).max() < 1e-9
    pk = o.gen_public_key(SEED, S.sk)
    ca = o.encrypt_asymmetric(o.public_key_seed(SEED), 3, pk, pt)
    assert np.abs(o.decode(o.decrypt(S.sk, ca), S.scale) - z).max() < 1e-9


@pytest.mark.parametrize("D", [16, 20, 64])
def test_bsgs_modes_against_float64(S, D):
    o = S.o
    slots = S.N // 2
    G, B = bsgs_params(D)
    rng = np.random.default_rng(D)
    W, x = rng.standard_normal((D, D)) * 0.1, rng.standard_normal(D)
    rolled = rolled_diagonals(W, D, G, B)
    keys = S.keys_for_steps(list(range(1, G)) + [g * G for g in range(1, B)])
    ct = o.encrypt_symmetric(SEED, 2, S.sk, o.encode(tile(x, slots), S.scale, S.L))
    sc = S.scale ** 2 / float(S.q[S.L - 1])
    pts = np.stack([o.encode(tile(rolled[k], slots), S.scale, S.L) for k in range(D)])
    baby = np.stack([ct] + [o.rotate(ct, b, keys) for b in range(1, G)])
    y = o.decode(o.decrypt(S.sk, o.bsgs_exact(baby, pts, G, B, D, keys)), sc)[:D]
    assert np.abs(y - W @ x).max() < 1e-9                     # tolerance: CKKS noise at scale 2^59
    full = np.stack([o.encode(tile(rolled[k], slots), S.scale, S.L, ext=True) for k in range(D)])
    y = o.decode(o.decrypt(S.sk, o.bsgs_hoisted(ct, full, G, B, D, keys)), sc)[:D]
    assert np.abs(y - W @ x).max() < 1e-9
    if D & (D - 1) == 0:
        comp = np.stack([o.encode(rolled[k], S.scale, S.L, ext=True, n=2 * D) for k in range(D)])
        y = o.decode(o.decrypt(S.sk, o.bsgs_hoisted(ct, comp, G, B, D, keys)), sc)[:D]
        assert np.abs(y - W @ x).max() < 1e-9


def test_bsgs_linearity(S):
    """BSGS(W, x1 + x2) decrypts to BSGS(W, x1) + BSGS(W, x2)  (size-independent property)."""
    o, D = S.o, 16
    G, B = bsgs_params(D)
    rng = np.random.default_rng(9)
    W = rng.standard_normal((D, D)) * 0.1
    comp = np.stack([o.encode(r, S.scale, S.L, ext=True, n=2 * D) for r in rolled_diagonals(W, D, G, B)])
    keys = S.keys_for_steps(list(range(1, G)) + [g * G for g in range(1, B)])
    xs = rng.standard_normal((2, D))
    cts = [o.encrypt_symmetric(SEED, 10 + i, S.sk, o.encode(tile(x, S.N // 2), S.scale, S.L)) for i, x in enumerate(xs)]
    sc = S.scale ** 2 / float(S.q[S.L - 1])
    ys = [o.decode(o.decrypt(S.sk, o.bsgs_hoisted(c, comp, G, B, D, keys)), sc)[:D].real for c in cts]
    ysum = o.decode(o.decrypt(S.sk, o.bsgs_hoisted(o.add(cts[0], cts[1]), comp, G, B, D, keys)), sc)[:D].real
    assert np.abs(ysum - (ys[0] + ys[1])).max() < 1e-9
    assert np.abs(ysum - W @ (xs[0] + xs[1])).max() < 1e-9


def test_golden_fixtures_from_reference_numpy_code(S):
    """tests/golden/bsgs_plan.json was produced by importing the reference's
    scripts/bootstrap_generation.py (make_golden.py): G/B split, Galois elements, diagonal
    extraction + pre-rotation + slot tiling.  The oracle-side helpers must agree."""
    with open(os.path.join(GOLD, "bsgs_plan.json")) as f:
        gold = json.load(f)
    for case in gold["params"]:
        assert list(bsgs_params(case["D"])) == [case["G"], case["B"]]
    for case in gold["galois"]:
        N = case["N"]
        from oracle.oracle import Oracle
        o = S.o if N == S.N else None
        elts = [pow(5, s, 2 * N) for s in case["steps"]]
        assert elts == case["bsgs_elts"]
        if o is not None:
            assert [o.elt_from_step(s) for s in case["steps"]] == case["bsgs_elts"]
        assert sorted(case["rot_elts"]) == sorted({2 * N - 1} | {pow(5, 1 << i, 2 * N) for i in range(case["n_pow2"])})
    npz = np.load(os.path.join(GOLD, "bsgs_diagonals.npz"))
    for D in (8, 20):
        W = npz[f"W_{D}"]
        G, B = bsgs_params(D)
        assert np.array_equal(rolled_diagonals(W, D, G, B), npz[f"rolled_{D}"][:, :D])
        slots = npz[f"rolled_{D}"].shape[1]
        assert np.array_equal(np.stack([tile(r, slots) for r in rolled_diagonals(W, D, G, B)]), npz[f"rolled_{D}"])


def _golden_loop():
    g = np.load(os.path.join(GOLD, "bsgs_loop.npz"))
    from oracle.oracle import Oracle
    o = Oracle(int(g["N"]), g["moduli"], int(g["P"]))
    return g, o


def test_oracle_bsgs_equals_reference_python_loop():
    """bsgs_loop.npz holds the output limbs of the reference's own fhe_matmul_bsgs /
    fhe_matmul_bsgs_complex Python loops (scripts/bootstrap_generation.py:464-484, :521-542) run
    over oracle primitives.  The fused oracle restatement must give the same limbs."""
    g, o = _golden_loop()
    seed = bytes(g["seed"])
    D, L0 = int(g["D"]), int(g["L0"])
    G, B = bsgs_params(D)
    slots = o.N // 2
    scale = 2.0 ** 59
    sk = o.gen_secret(seed)
    ct_x = o.encrypt_symmetric(seed, int(g["enc_id_x"]), sk, o.encode(tile(g["x"], slots), scale, L0))
    assert np.array_equal(ct_x, g["ct_x"])
    keys = {}
    for s in list(range(1, G)) + [k * G for k in range(1, B)]:
        e = o.elt_from_step(s)
        keys[e] = o.gen_galois_key(seed, e, sk)
    baby = np.stack([ct_x] + [o.rotate(ct_x, b, keys) for b in range(1, G)])
    rolled = rolled_diagonals(g["W"], D, G, B)
    pts = np.stack([o.encode(tile(r, slots), scale, L0) for r in rolled])
    assert np.array_equal(o.bsgs_exact(baby, pts, G, B, D, keys), g["ct_y_real"])
    rolled_c = rolled + 1j * rolled_diagonals(g["W2"], D, G, B)
    pts_c = np.stack([o.encode(tile(r, slots), scale, L0) for r in rolled_c])
    assert np.array_equal(o.bsgs_exact(baby, pts_c, G, B, D, keys), g["ct_y_complex"])
    sc = scale * scale / float(o.q[L0 - 1])
    assert np.array_equal(o.decode(o.decrypt(sk, g["ct_y_real"]), sc)[:D].real, g["y_real_dec"])
    assert np.abs(g["y_real_dec"] - g["W"] @ g["x"]).max() < 1e-9


def test_inference_primitives_golden_replay():
    """tests/golden/inference_primitives.npz was produced by the reference's own fhe_rwkv_inference.py functions
    (CKKSContext, ct_pt_dot, ct_pt_weighted_sum, ct_ct_square, ct_ct_multiply, :29-108) running over oracle primitives
    (tests/golden/make_golden_inference.py).  Replaying those call sequences directly on the oracle must give the
    same limbs: pins the oracle's public-key encryption, plaintext multiply, rescale, rotate, mod-switch, tensor
    product and relinearisation at the reference's [60] + [40]*d + [60], P = 1 parameter style."""
    g = np.load(os.path.join(GOLD, "inference_primitives.npz"))
    S = Setup(N=int(g["N"]), bits=tuple(int(b) for b in g["bits"]), P=int(g["P"]), seed=bytes(g["seed"]))
    o, scale, dim, slots = S.o, float(g["scale"]), int(g["dim"]), int(g["N"]) // 2
    pk, rlk = o.gen_public_key(S.seed, S.sk), o.gen_relin_key(S.seed, S.sk)
    pad = lambda v: np.concatenate([np.asarray(v, dtype=np.float64), np.zeros(slots - len(v))])
    ct = o.encrypt_asymmetric(o.public_key_seed(S.seed), 1, pk, o.encode(pad(g["x"]), scale, S.L))
    assert np.array_equal(ct, g["ct"])

    def dot(w):                                   # fhe_rwkv_inference.py:66-76
        prod = o.rescale(o.multiply_plain(ct, o.encode(pad(w), scale, S.L)))
        step = 1
        while step < dim:
            elt = o.elt_from_step(step)
            prod = o.add(prod, o.apply_galois(prod, elt, S.key(elt)))
            step *= 2
        return prod
    d1, d2 = dot(g["w1"]), dot(g["w2"])
    assert np.array_equal(d1, g["d1"]) and np.array_equal(d2, g["d2"])
    terms = [o.rescale(o.multiply_plain(d, o.encode(np.full(slots, w), scale, S.L)[:d.shape[1]]))   # :79-94, level 2
             for d, w in zip((d1, d2), g["mix"])]
    ws = o.add(terms[0], terms[1])
    assert np.array_equal(ws, g["ws"])
    sq = o.rescale(o.relinearize(o.multiply(ws, ws), rlk))       # :97-101
    pr = o.rescale(o.relinearize(o.multiply(d1, d2), rlk))       # :104-108
    assert np.array_equal(sq, g["sq"]) and np.array_equal(pr, g["pr"])
    for name, arr, sc in (("d1", d1, scale ** 2 / float(S.q[S.L - 1])), ):
        val = o.decode(o.decrypt(S.sk, arr), sc)[0].real
        assert abs(val - float(g["slot0_float64"][0])) < 1e-6
