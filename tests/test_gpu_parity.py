"""GPU parity: every C-ABI entry point of the hot path against the CPU oracle, bit for bit.

All calls go through fhe_spear_b200.pyPhantom -> libspear_b200.so (ctypes); the oracle
(oracle/spear_oracle.c) is only the checker.  Integer outputs must be identical; decoded floats
must match W.x within the tolerance written next to each check.
"""
import numpy as np
import pytest

from helpers import SEED, Setup, bsgs_params, rolled_diagonals, tile

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    return Setup(N=2048, bits=(59,) * 7, P=2)   # L=5, P=2: beta=3 with a ragged last digit


@pytest.fixture(scope="module")
def G_(S):
    steps = list(range(1, 8)) + [8 * g for g in range(1, 8)] + [5 * g for g in range(1, 5)] + [-1, 16]
    ph, ctx, sk = S.gpu(steps)
    gk = sk.create_galois_keys(ctx)
    return dict(ph=ph, ctx=ctx, sk=sk, gk=gk, enc=ph.ckks_encoder(ctx))


def test_ntt_matches_oracle(S, G_):
    import ctypes as C
    from fhe_spear_b200 import _native as n
    rng = np.random.default_rng(1)
    for ring in (S.N, 512, 64, 8):
        limbs = [0, 3, S.L, S.L + S.P - 1]
        a = np.stack([rng.integers(0, int(S.q[t]), ring, dtype=np.uint64) for t in limbs])
        got = a.copy()
        ids = (C.c_int * len(limbs))(*limbs)
        n.check(n.lib.spear_ntt_host(G_["ctx"]._h, got.ctypes.data_as(C.c_void_p), len(limbs), ids, ring, 0))
        for r, t in enumerate(limbs):
            exp = a[r].copy()
            S.o.lib.orc_ntt_fwd_sub(S.o.ctx, t, exp.ctypes.data_as(C.c_void_p), C.c_uint64(ring))
            assert np.array_equal(got[r], exp), (ring, t)
        back = got.copy()
        n.check(n.lib.spear_ntt_host(G_["ctx"]._h, back.ctypes.data_as(C.c_void_p), len(limbs), ids, ring, 1))
        assert np.array_equal(back, a)


@pytest.mark.parametrize("N", [4096, 8192, 16384, 32768, 65536])
def test_ntt_large_sizes(N):
    import ctypes as C
    from fhe_spear_b200 import _native as n
    s = Setup(N=N, bits=(59, 59, 60), P=1)
    ph, ctx, _ = s.gpu()
    rng = np.random.default_rng(N)
    limbs = [0, 1, 2]
    a = np.stack([rng.integers(0, int(s.q[t]), N, dtype=np.uint64) for t in limbs])
    got = a.copy()
    ids = (C.c_int * 3)(*limbs)
    n.check(n.lib.spear_ntt_host(ctx._h, got.ctypes.data_as(C.c_void_p), 3, ids, N, 0))
    for r, t in enumerate(limbs):
        assert np.array_equal(got[r], s.o.ntt_fwd(t, a[r]))
    n.check(n.lib.spear_ntt_host(ctx._h, got.ctypes.data_as(C.c_void_p), 3, ids, N, 1))
    assert np.array_equal(got, a)


def test_primes_and_galois_elements(S, G_):
    ph = G_["ph"]
    for bits in ([59] * 27, [60] + [40] * 9 + [60], [54] * 4):
        for N in (8192, 32768):
            from oracle.oracle import Oracle
            assert [int(m) for m in ph.create_coeff_modulus(N, bits)] == [int(x) for x in Oracle.create_coeff_modulus(N, bits)]
    for step in (1, 2, 45, 46, 2024, -1, -7, 0):
        assert ph.get_elt_from_step(step, S.N) == S.o.elt_from_step(step)


def test_keys_match_oracle(S, G_):
    ph, ctx, sk, gk = G_["ph"], G_["ctx"], G_["sk"], G_["gk"]
    assert np.array_equal(sk.to_numpy(), S.sk)
    for step in (1, 3, 8):
        elt = ph.get_elt_from_step(step, S.N)
        assert np.array_equal(gk.to_numpy(elt), S.key(elt)), step
    assert np.array_equal(gk.to_numpy(2 * S.N - 1), S.key(2 * S.N - 1))
    assert np.array_equal(sk.gen_relinkey(ctx).to_numpy(), S.o.gen_relin_key(SEED, S.sk))
    assert np.array_equal(sk.gen_publickey(ctx).to_numpy(), S.o.gen_public_key(SEED, S.sk))


def test_encode_decode_match_oracle(S, G_):
    ph, ctx, enc = G_["ph"], G_["ctx"], G_["enc"]
    rng = np.random.default_rng(2)
    slots = S.N // 2
    z = rng.standard_normal(slots) + 1j * rng.standard_normal(slots)
    for ci in (1, 3):
        l = S.L - ci + 1
        pt = enc.encode_complex_vector(ctx, z, S.scale, chain_index=ci)
        assert pt.chain_index() == ci and pt.coeff_modulus_size() == l
        assert np.array_equal(pt.to_numpy()[0], S.o.encode(z, S.scale, l))
    x = rng.standard_normal(100)
    pt = enc.encode_double_vector(ctx, list(x), S.scale)
    xp = np.zeros(slots)
    xp[:100] = x
    exp = S.o.encode(xp, S.scale, S.L)
    assert np.array_equal(pt.to_numpy()[0], exp)
    got = np.array(enc.decode_complex_vector(ctx, pt))
    assert np.array_equal(got, S.o.decode(exp, S.scale))          # bit-identical doubles
    assert np.abs(got.real - xp).max() < 1e-12
    # batch + large magnitudes (coefficients beyond 2^64)
    big = rng.standard_normal((3, slots)) * 2.0 ** 30
    pts = enc.encode_double_vector_batch(ctx, big, S.scale, chain_index=2)
    for k in range(3):
        assert np.array_equal(pts[k].to_numpy()[0], S.o.encode(big[k], S.scale, S.L - 1))
    with pytest.raises(RuntimeError):
        enc.encode_double_vector(ctx, [2.0 ** 80], S.scale)


def test_encrypt_decrypt_match_oracle(S, G_):
    ph, ctx, sk, enc = G_["ph"], G_["ctx"], G_["sk"], G_["enc"]
    rng = np.random.default_rng(3)
    z = rng.standard_normal(S.N // 2)
    pt = enc.encode_double_vector(ctx, z, S.scale)
    pt_o = S.o.encode(z, S.scale, S.L)
    ct = sk.encrypt_symmetric(ctx, pt, enc_id=11)
    ct_o = S.o.encrypt_symmetric(SEED, 11, S.sk, pt_o)
    assert np.array_equal(ct.to_numpy(), ct_o)
    dec = sk.decrypt(ctx, ct)
    assert np.array_equal(dec.to_numpy()[0], S.o.decrypt(S.sk, ct_o))
    assert np.abs(np.array(enc.decode_double_vector(ctx, dec)) - z).max() < 1e-10
    # asymmetric
    pk = sk.gen_publickey(ctx)
    ca = pk.encrypt_asymmetric(ctx, pt, enc_id=5)
    assert np.array_equal(ca.to_numpy(), S.o.encrypt_asymmetric(S.o.public_key_seed(SEED), 5, S.o.gen_public_key(SEED, S.sk), pt_o))
    assert np.abs(np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, ca))) - z).max() < 1e-9


def test_evaluator_matches_oracle(S, G_):
    ph, ctx, sk, gk, enc = (G_[k] for k in ("ph", "ctx", "sk", "gk", "enc"))
    o = S.o
    rng = np.random.default_rng(4)
    slots = S.N // 2
    za, zb = rng.standard_normal(slots), rng.standard_normal(slots)
    pa, pb = (enc.encode_double_vector(ctx, z, S.scale) for z in (za, zb))
    a, b = sk.encrypt_symmetric(ctx, pa, enc_id=1), sk.encrypt_symmetric(ctx, pb, enc_id=2)
    ao, bo = a.to_numpy(), b.to_numpy()
    pbo = pb.to_numpy()[0]
    assert np.array_equal(ph.add(ctx, a, b).to_numpy(), o.add(ao, bo))
    assert np.array_equal(ph.sub(ctx, a, b).to_numpy(), o.sub(ao, bo))
    assert np.array_equal(ph.negate(ctx, a).to_numpy(), o.sub(np.zeros_like(ao), ao))
    mp = ph.multiply_plain(ctx, a, pb)
    assert np.array_equal(mp.to_numpy(), o.multiply_plain(ao, pbo))
    assert mp.scale() == S.scale * S.scale
    ap = ph.add_plain(ctx, a, pb).to_numpy()
    assert np.array_equal(ap[0], o.add(ao[:1], pbo[None])[0]) and np.array_equal(ap[1], ao[1])
    # rescale / mod switch
    rs = ph.rescale_to_next(ctx, mp)
    assert np.array_equal(rs.to_numpy(), o.rescale(mp.to_numpy()))
    assert rs.chain_index() == 2 and rs.scale() == S.scale * S.scale / float(S.q[S.L - 1])
    ms = ph.mod_switch_to_next(ctx, a)
    assert np.array_equal(ms.to_numpy(), ao[:, :-1])
    assert ph.mod_switch_to(ctx, pb, 3).coeff_modulus_size() == S.L - 2
    # ct x ct, relinearize
    rlk = sk.gen_relinkey(ctx)
    t3 = ph.multiply(ctx, a, b)
    t3o = o.multiply(ao, bo)
    assert np.array_equal(t3.to_numpy(), t3o)
    rl = ph.relinearize(ctx, t3, rlk)
    assert np.array_equal(rl.to_numpy(), o.relinearize(t3o, o.gen_relin_key(SEED, S.sk)))
    prod = np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, ph.rescale_to_next(ctx, rl))))
    assert np.abs(prod - za * zb).max() < 1e-8
    # rotations at two levels (ragged digit at l=4: digits of 2,2 ; l=5: 2,2,1 ; l=3: 2,1)
    for ct, cto in ((a, ao), (ms, ao[:, :-1]), (ph.mod_switch_to_next(ctx, ms), ao[:, :-2])):
        for step in (1, 7, 16, -1):
            r = ph.rotate(ctx, ct, step, gk)
            elt = o.elt_from_step(step)
            assert np.array_equal(r.to_numpy(), o.apply_galois(cto, elt, S.key(elt))), (step, cto.shape)
    r = ph.rotate(ctx, a, 7, gk)
    assert np.abs(np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, r))) - np.roll(za, -7)).max() < 1e-9
    cj = ph.apply_galois(ctx, a, 2 * S.N - 1, gk)
    assert np.array_equal(cj.to_numpy(), o.apply_galois(ao, 2 * S.N - 1, S.key(2 * S.N - 1)))
    # step without its own key -> composed from power-of-two keys (NAF): 3 = 4 - 1
    steps_pow2 = [1, 2, 4, -1]
    gk2 = sk.create_galois_keys(ctx, [ph.get_elt_from_step(s, S.N) for s in steps_pow2])
    r3 = ph.rotate(ctx, a, 3, gk2)
    assert np.abs(np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, r3))) - np.roll(za, -3)).max() < 1e-9
    with pytest.raises(RuntimeError):
        ph.rotate(ctx, a, 24, gk2)


def test_hoisted_rotations(S, G_):
    ph, ctx, sk, gk, enc = (G_[k] for k in ("ph", "ctx", "sk", "gk", "enc"))
    rng = np.random.default_rng(5)
    z = rng.standard_normal(S.N // 2)
    ct = sk.encrypt_symmetric(ctx, enc.encode_double_vector(ctx, z, S.scale), enc_id=9)
    outs = ph.hoisting(ctx, ct, gk, [1, 2, 5])
    cto = ct.to_numpy()
    for r, step in zip(outs, (1, 2, 5)):
        elt = S.o.elt_from_step(step)
        assert np.array_equal(r.to_numpy(), S.o.hoisted_rotation(cto, elt, S.key(elt)))
        got = np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, r)))
        assert np.abs(got - np.roll(z, -step)).max() < 1e-9


@pytest.mark.parametrize("D,complex_", [(64, False), (64, True), (20, False), (1024, False)])
def test_bsgs_exact_and_hoisted_match_oracle(S, G_, D, complex_):
    ph, ctx, sk, enc = (G_[k] for k in ("ph", "ctx", "sk", "enc"))
    o = S.o
    slots = S.N // 2
    G, B = bsgs_params(D)
    rng = np.random.default_rng(D)
    W = rng.standard_normal((D, D)) * 0.1
    x = rng.standard_normal(D)
    rolled = rolled_diagonals(W, D, G, B).astype(np.complex128)
    expect = W @ x
    if complex_:
        W2 = rng.standard_normal((D, D)) * 0.1
        rolled = rolled + 1j * rolled_diagonals(W2, D, G, B)
        expect = W @ x + 1j * (W2 @ x)
    steps = list(range(1, G)) + [g * G for g in range(1, B)]
    gk = sk.create_galois_keys(ctx, [ph.get_elt_from_step(s, S.N) for s in steps])
    keys = S.keys_for_steps(steps)

    ct = sk.encrypt_symmetric(ctx, enc.encode_double_vector(ctx, tile(x, slots), S.scale), enc_id=100 + D)
    cto = ct.to_numpy()
    out_scale = S.scale * S.scale / float(S.q[S.L - 1])

    # exact mode: reference op order over pre-encoded plaintexts
    if D <= 64:
        vecs = np.stack([tile(rolled[k], slots) for k in range(D)])
        pts = (enc.encode_complex_vector_batch if complex_ else enc.encode_double_vector_batch)(
            ctx, vecs if complex_ else vecs.real, S.scale, chain_index=1)
        baby = [ct] + [ph.rotate(ctx, ct, b, gk) for b in range(1, G)]
        y = ph.bsgs_multiply_accumulate(ctx, baby, pts, G, B, D, gk)
        pts_o = np.stack([p.to_numpy()[0] for p in pts])
        baby_o = np.stack([c.to_numpy() for c in baby])
        assert np.array_equal(y.to_numpy(), o.bsgs_exact(baby_o, pts_o, G, B, D, keys))
        assert y.chain_index() == 2 and y.scale() == out_scale
        dec = np.array(enc.decode_complex_vector(ctx, sk.decrypt(ctx, y)))[:D]
        assert np.abs(dec - expect).max() < 1e-9, "exact mode vs float64 W.x"
        # python fallback loop of the reference (multiply_plain / add / rotate / rescale) gives the same limbs
        acc = None
        for g in range(B):
            inner = None
            for b in range(G):
                k = g * G + b
                if k >= D:
                    continue
                term = ph.multiply_plain(ctx, baby[b], pts[k])
                inner = term if inner is None else ph.add(ctx, inner, term)
            if inner is None:
                continue
            if g:
                inner = ph.rotate(ctx, inner, g * G, gk)
            acc = inner if acc is None else ph.add(ctx, acc, inner)
        assert np.array_equal(ph.rescale_to_next(ctx, acc).to_numpy(), y.to_numpy())

    # hoisted mode: compressed (sub-ring) when D is a power of two, full-ring otherwise
    for compress in ([True, False] if (D & (D - 1)) == 0 and D <= 64 else [(D & (D - 1)) == 0]):
        ds = ph.diagonal_set(ctx, rolled, G, B, S.scale, chain_index=1, compress=compress)
        info = ds.info()
        ring = 2 * D if compress else S.N
        assert info["ring_n"] == ring and info["limbs"] == S.L
        if compress:
            diag_o = np.stack([o.encode(rolled[k], S.scale, S.L, ext=True, n=ring) for k in range(D)])
        else:
            diag_o = np.stack([o.encode(tile(rolled[k], slots), S.scale, S.L, ext=True) for k in range(D)])
        assert np.array_equal(ds.to_numpy(), diag_o)
        y = ph.bsgs_hoisted(ctx, ct, ds, gk)
        assert np.array_equal(y.to_numpy(), o.bsgs_hoisted(cto, diag_o, G, B, D, keys)), (D, compress)
        assert y.chain_index() == 2 and y.scale() == out_scale
        dec = np.array(enc.decode_complex_vector(ctx, sk.decrypt(ctx, y)))[:D]
        assert np.abs(dec - expect).max() < 1e-9, "hoisted mode vs float64 W.x"


def test_bsgs_at_lower_level_and_errors(S, G_):
    ph, ctx, sk, enc = (G_[k] for k in ("ph", "ctx", "sk", "enc"))
    D = 16
    G, B = bsgs_params(D)
    rng = np.random.default_rng(77)
    W, x = rng.standard_normal((D, D)) * 0.1, rng.standard_normal(D)
    rolled = rolled_diagonals(W, D, G, B)
    steps = list(range(1, G)) + [g * G for g in range(1, B)]
    gk = sk.create_galois_keys(ctx, [ph.get_elt_from_step(s, S.N) for s in steps])
    keys = S.keys_for_steps(steps)
    ct = sk.encrypt_symmetric(ctx, enc.encode_double_vector(ctx, tile(x, S.N // 2), S.scale), enc_id=7)
    ct3 = ph.mod_switch_to(ctx, ct, 3)          # 3 limbs: digits of sizes 2 and 1
    ds3 = ph.diagonal_set(ctx, rolled, G, B, S.scale, chain_index=3)
    y = ph.bsgs_hoisted(ctx, ct3, ds3, gk)
    diag_o = np.stack([S.o.encode(rolled[k].astype(complex), S.scale, 3, ext=True, n=2 * D) for k in range(D)])
    assert np.array_equal(y.to_numpy(), S.o.bsgs_hoisted(ct3.to_numpy(), diag_o, G, B, D, keys))
    dec = np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, y)))[:D]
    assert np.abs(dec - W @ x).max() < 1e-9
    with pytest.raises(RuntimeError):           # level mismatch between diagonals and ciphertext
        ph.bsgs_hoisted(ctx, ct, ds3, gk)
    gk_small = sk.create_galois_keys(ctx, [ph.get_elt_from_step(1, S.N)])
    with pytest.raises(RuntimeError):           # missing rotation key
        ph.bsgs_hoisted(ctx, ct3, ds3, gk_small)


def test_mixed_prime_sizes_public_key_path():
    """Parameter style of fhe_rwkv_inference.py:29-35: [60] + [40]*d + [60], P=1, asymmetric encryption."""
    s = Setup(N=1024, bits=(60, 40, 40, 40, 60), P=1)
    ph, ctx, sk = s.gpu([1, 2, 4])
    enc = ph.ckks_encoder(ctx)
    pk, rlk, gk = sk.gen_publickey(ctx), sk.gen_relinkey(ctx), sk.create_galois_keys(ctx)
    rng = np.random.default_rng(8)
    z = rng.standard_normal(512)
    w = rng.standard_normal(512)
    ct = pk.encrypt_asymmetric(ctx, enc.encode_double_vector(ctx, z, 2.0 ** 40))
    prod = ph.rescale_to_next(ctx, ph.multiply_plain(ctx, ct, enc.encode_double_vector(ctx, w, 2.0 ** 40)))
    assert np.array_equal(prod.to_numpy(), s.o.rescale(s.o.multiply_plain(ct.to_numpy(), s.o.encode(w, 2.0 ** 40, s.L))))
    step = 1
    acc = prod
    while step < 8:                              # ct_pt_dot's rotate-and-sum, fhe_rwkv_inference.py:66-76
        acc = ph.add(ctx, acc, ph.rotate(ctx, acc, step, gk))
        step *= 2
    got = enc.decode_double_vector(ctx, sk.decrypt(ctx, acc))[0]
    assert abs(got - np.dot(z[:8], w[:8])) < 1e-5
    sq = ph.rescale_to_next(ctx, ph.relinearize(ctx, ph.multiply(ctx, prod, prod), rlk))
    got = np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, sq)))
    assert np.abs(got - (z * w) ** 2).max() < 1e-4


def test_wide_primes_take_the_exact_fused_path():
    """60-bit primes (fhe_rwkv_inference.py:29-35 style) at N = 2048: the fused transform + key product kernel runs its
    non-lazy butterflies, folds every 8 digits and reduces with the full Barrett step.  Rotation and relinearisation
    must equal the oracle limb for limb."""
    s = Setup(N=2048, bits=(60, 59, 60, 40, 40, 60), P=1)   # the rescale drops a 40-bit prime: scale stays 2^40
    ph, ctx, sk = s.gpu([1, 5])
    enc = ph.ckks_encoder(ctx)
    rlk, gk = sk.gen_relinkey(ctx), sk.create_galois_keys(ctx)
    rng = np.random.default_rng(18)
    z = rng.standard_normal(s.N // 2)
    ct = sk.encrypt_symmetric(ctx, enc.encode_double_vector(ctx, z, 2.0 ** 40), enc_id=9)
    cto = ct.to_numpy()
    for step in (1, 5):
        elt = s.o.elt_from_step(step)
        assert np.array_equal(ph.rotate(ctx, ct, step, gk).to_numpy(), s.o.apply_galois(cto, elt, s.key(elt)))
    sq = ph.relinearize(ctx, ph.multiply(ctx, ct, ct), rlk)
    assert np.array_equal(sq.to_numpy(), s.o.relinearize(s.o.multiply(cto, cto), s.o.gen_relin_key(s.seed, s.sk)))
    got = np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, ph.rescale_to_next(ctx, sq))))
    assert np.abs(got - z * z).max() < 1e-4


@pytest.mark.parametrize("N", [8192, 65536])
def test_rotation_at_large_ring_sizes(N):
    """A full rotate (permute, decompose, transform fused with the key product, ModDown) at N = 8192 and at the
    largest ring the engine supports (N = 65536: pass A of eight stages in front of the fused pass) vs the oracle."""
    s = Setup(N=N, bits=(59,) * 5, P=2)
    ph, ctx, sk = s.gpu([7])
    enc = ph.ckks_encoder(ctx)
    gk = sk.create_galois_keys(ctx)
    rng = np.random.default_rng(N)
    z = rng.standard_normal(64)
    ct = sk.encrypt_symmetric(ctx, enc.encode_double_vector(ctx, z, s.scale), enc_id=4)
    elt = s.o.elt_from_step(7)
    rot = ph.rotate(ctx, ct, 7, gk)
    assert np.array_equal(rot.to_numpy(), s.o.apply_galois(ct.to_numpy(), elt, s.key(elt)))
    got = np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, rot)))[:57]
    assert np.abs(got - z[7:]).max() < 1e-9


def test_reference_inference_primitives_golden_limbs():
    """The call sequences of the reference's fhe_rwkv_inference.py (CKKSContext.encrypt :46-50, ct_pt_dot :66-76,
    ct_pt_weighted_sum :79-94, ct_ct_square :97-101, ct_ct_multiply :104-108) through pyPhantom on the GPU must
    reproduce the limbs recorded when the reference's own functions ran over the oracle
    (tests/golden/inference_primitives.npz, generator tests/golden/make_golden_inference.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inference_primitives.npz"))
    N, dim = int(g["N"]), int(g["dim"])
    s = Setup(N=N, bits=tuple(int(b) for b in g["bits"]), P=int(g["P"]), seed=bytes(g["seed"]))
    ph, ctx, sk = s.gpu([1, 2, 4])
    enc = ph.ckks_encoder(ctx)
    pk, rlk, gk = sk.gen_publickey(ctx), sk.gen_relinkey(ctx), sk.create_galois_keys(ctx)
    scale, slots = float(g["scale"]), enc.slot_count()
    pad = lambda v: list(v) + [0.0] * (slots - len(v))
    ct = pk.encrypt_asymmetric(ctx, enc.encode_double_vector(ctx, pad(g["x"]), scale), enc_id=1)
    assert np.array_equal(ct.to_numpy(), g["ct"])

    def ct_pt_dot(w):
        prod = ph.rescale_to_next(ctx, ph.multiply_plain(ctx, ct, enc.encode_double_vector(ctx, pad(w), scale)))
        step = 1
        while step < dim:
            prod = ph.add(ctx, prod, ph.rotate(ctx, prod, step, gk))
            step *= 2
        return prod
    d1, d2 = ct_pt_dot(g["w1"]), ct_pt_dot(g["w2"])
    assert np.array_equal(d1.to_numpy(), g["d1"]) and np.array_equal(d2.to_numpy(), g["d2"])
    assert d1.chain_index() == 2
    terms = []
    for d, w in zip((d1, d2), g["mix"]):
        w_pt = ph.mod_switch_to(ctx, enc.encode_double_vector(ctx, [float(w)] * slots, scale), 2)
        terms.append(ph.rescale_to_next(ctx, ph.multiply_plain(ctx, d, w_pt)))
    ws = ph.add(ctx, terms[0], terms[1])
    assert np.array_equal(ws.to_numpy(), g["ws"])
    sq = ph.rescale_to_next(ctx, ph.relinearize(ctx, ph.multiply(ctx, ws, ws), rlk))
    pr = ph.rescale_to_next(ctx, ph.relinearize(ctx, ph.multiply(ctx, d1, d2), rlk))
    assert np.array_equal(sq.to_numpy(), g["sq"]) and np.array_equal(pr.to_numpy(), g["pr"])
    got = [enc.decode_double_vector(ctx, sk.decrypt(ctx, c))[0] for c in (d1, d2, ws, sq, pr)]
    assert np.abs(np.array(got) - g["slot0"]).max() < 1e-9 and np.abs(np.array(got) - g["slot0_float64"]).max() < 1e-4


@pytest.mark.parametrize("N,bits,P", [(2048, (59,) * 4, 1), (4096, (60, 40, 40, 40, 60), 1), (8192, (59,) * 6, 2),
                                      (32768, (59,) * 5, 2), (65536, (59,) * 4, 1)])
def test_fused_client_legs_match_the_staged_path_bit_for_bit(N, bits, P):
    """csrc/client.cu: encode + encrypt and decrypt + decode in three launches each give the
    same limbs / the same doubles as encode -> encrypt_symmetric and decrypt -> decode (which the tests above pin to
    the oracle), for every pass-A split (N = 2^11 .. 2^16), real and complex, replicated and zero-padded inputs,
    mixed prime sizes, and size-3 ciphertexts."""
    S = Setup(N=N, bits=bits, P=P)
    ph, ctx, sk = S.gpu([1])
    enc = ph.ckks_encoder(ctx)
    slots = N // 2
    rng = np.random.default_rng(N)
    for D, cplx, rep in [(16, False, True), (20, True, True), (slots, True, True), (33, False, False), (1, False, True)]:
        v = rng.standard_normal(D) + (1j * rng.standard_normal(D) if cplx else 0)
        full = np.zeros(slots, dtype=np.complex128)
        if rep:
            full[:] = np.concatenate([np.tile(v, slots // D), v[:slots % D]])
        else:
            full[:D] = v
        two_step = sk.encrypt_symmetric(ctx, enc.encode_complex_vector(ctx, full, S.scale), enc_id=77)
        fused = sk.encrypt_vector(ctx, v, S.scale, replicate=rep, enc_id=77)
        assert fused.chain_index() == 1 and fused.scale() == two_step.scale()
        assert np.array_equal(fused.to_numpy(), two_step.to_numpy()), (N, D, cplx, rep)
        ref = np.array(enc.decode_complex_vector(ctx, sk.decrypt(ctx, two_step)))
        got = sk.decrypt_decode(ctx, fused)
        assert np.array_equal(got, ref) and np.abs(got - full).max() < 1e-6
        assert np.array_equal(sk.decrypt_decode(ctx, fused, 7), ref[:7])
    # deeper levels and a size-3 ciphertext (the decoder reads the first min(l, 3) limbs)
    a = sk.encrypt_vector(ctx, rng.standard_normal(8), S.scale, enc_id=5)
    sq = ph.multiply(ctx, a, a)
    assert np.array_equal(sk.decrypt_decode(ctx, sq), np.array(enc.decode_complex_vector(ctx, sk.decrypt(ctx, sq))))
    low = ph.rescale_to_next(ctx, ph.relinearize(ctx, sq, sk.gen_relinkey(ctx)))
    while low.coeff_modulus_size() > 1:
        assert np.array_equal(sk.decrypt_decode(ctx, low), np.array(enc.decode_complex_vector(ctx, sk.decrypt(ctx, low))))
        low = ph.mod_switch_to_next(ctx, low)
    assert np.array_equal(sk.decrypt_decode(ctx, low), np.array(enc.decode_complex_vector(ctx, sk.decrypt(ctx, low))))
    # the auto counter: two encryptions of the same vector never share randomness
    c1, c2 = sk.encrypt_vector(ctx, [1.0], S.scale), sk.encrypt_vector(ctx, [1.0], S.scale)
    assert not np.array_equal(c1.to_numpy()[1], c2.to_numpy()[1])
