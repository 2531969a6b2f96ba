"""CKKS bootstrapping (fhe_spear_b200/bootstrap.py).

The reference's bootstrapper lives in the absent phantom-fhe fork: parity against it is unpinned.  What is pinned:
  * the algorithm's own invariants on the CPU oracle (ModRaise decrypts to m + q0*I, the special-FFT factorisation
    equals the encoder's slot map, the refreshed ciphertext decrypts to the message within a stated bound);
  * on a GPU, the product runs the same orchestration and every limb of the bootstrapped ciphertext equals the
    oracle's (same seed, same encryption randomness).
"""
import numpy as np
import pytest

from helpers import SEED
from oracle_phantom import Backend


def _setup(N, L0, P):
    from fhe_spear_b200.bootstrap import Bootstrapper
    be = Backend(N, [59] * (L0 + P), P, seed=SEED)
    steps = Bootstrapper.rotation_steps(N, (2, 2))
    be.add_galois([be.get_elt_from_step(s) for s in steps] + [2 * N - 1])
    bt = Bootstrapper(be, be.ctx, be, N, be.moduli, P, (2, 2))
    return be, bt, steps


def test_special_fft_stages_equal_the_slot_map():
    from fhe_spear_b200 import bootstrap as B
    N = 64
    n, M = N // 2, 2 * N
    zeta = np.exp(2j * np.pi / M)
    U0 = np.array([[zeta ** ((pow(5, j, M) * k) % M) for k in range(n)] for j in range(n)])
    bits = n.bit_length() - 1
    R = np.zeros((n, n))
    for i in range(n):
        R[i, int(format(i, f"0{bits}b")[::-1], 2)] = 1

    def dense(dm):
        A = np.zeros((n, n), complex)
        for k, d in dm.items():
            A[np.arange(n), (np.arange(n) + k) % n] += d
        return A
    F = np.eye(n, dtype=complex)
    for st in B._s2c_stages(n, N):
        F = dense(st) @ F
    Finv = np.eye(n, dtype=complex)
    for st in B._c2s_stages(n, N):
        Finv = dense(st) @ Finv
    assert np.abs(F @ R - U0).max() < 1e-12                     # slots = U0 (c_lo + i c_hi)
    assert np.abs(Finv @ F - np.eye(n)).max() < 1e-12
    merged = B._merge(B._s2c_stages(n, N), 2, n, scalar=3.0)
    assert np.abs(dense(merged[1]) @ dense(merged[0]) - 3.0 * F).max() < 1e-11
    q, r = B._cheb_divide(np.arange(1.0, 12.0), 8)             # c = q*T_8 + r in the Chebyshev basis
    x = np.linspace(-1, 1, 7)
    T = np.polynomial.chebyshev.chebval
    assert np.abs(T(x, np.arange(1.0, 12.0)) - (T(x, q) * T(x, [0] * 8 + [1]) + T(x, r))).max() < 1e-12


def test_mod_raise_decrypts_to_message_plus_multiple_of_q0():
    be, bt, _ = _setup(256, 4, 1)
    rng = np.random.default_rng(0)
    msg = rng.standard_normal(128) * 0.01
    ct = be.encrypt(be.encode_complex_vector(None, msg, 2.0 ** 40, be.o.L))     # one limb
    raised = be.mod_raise(None, ct, 1)
    assert raised.coeff_modulus_size() == be.o.L
    q0 = be.moduli[0]
    # limb 0 is untouched; every limb holds the same small integers I*q0 + m
    assert np.array_equal(raised.a[:, 0], ct.a[:, 0])
    pt = be.o.decrypt(be.sk, raised.a)
    L = be.o.L
    coef = [be.o.ntt_inv(i, pt[i].copy()) for i in range(L)]
    Q = 1
    for q in be.moduli[:L]:
        Q *= q
    crt = [(Q // q) * pow(Q // q, -1, q) for q in be.moduli[:L]]
    small = be.o.ntt_inv(0, be.o.decrypt(be.sk, ct.a)[0].copy())            # m + e modulo q0 before the raise
    worst = 0
    for k in range(be.N):
        v = sum(int(coef[i][k]) * crt[i] for i in range(L)) % Q
        v = v - Q if v > Q // 2 else v                                          # the integer the raised ciphertext holds
        assert (v - int(small[k])) % q0 == 0                                    # = m + q0 * I
        worst = max(worst, abs(v) // q0)
    assert 1 <= worst <= bt.K                                                   # |I| within the EvalMod range


def test_bootstrap_refreshes_levels_and_keeps_the_message():
    from fhe_spear_b200.bootstrap import Bootstrapper
    N, L0, P = 512, 19, 2
    be, bt, steps = _setup(N, L0, P)
    n = N // 2
    rng = np.random.default_rng(1)
    msg = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    ct = be.encrypt(be.encode_complex_vector(None, msg, 2.0 ** 59, L0 - 1))      # two limbs left
    out = bt.bootstrap(ct)
    depth = Bootstrapper.depth_for(N, (2, 2))
    assert out.coeff_modulus_size() == L0 - depth and depth <= 17
    dec = be.decode_complex_vector(None, be.decrypt(out))
    assert np.abs(dec - msg).max() < 2e-3                                         # |msg| up to ~4, scale 2^59
    # the refreshed ciphertext is usable: one more plaintext multiplication and rescale
    half = be.encode_complex_vector(None, np.full(n, 0.5), float(be.moduli[out.coeff_modulus_size() - 1]), out.chain_index())
    again = be.rescale_to_next(None, be.multiply_plain(None, out, half))
    assert np.abs(be.decode_complex_vector(None, be.decrypt(again)) - 0.5 * msg).max() < 2e-3
    # galois elements published for the context cover every rotation the transforms make
    assert set(steps) == set(Bootstrapper.rotation_steps(N, (2, 2)))


@pytest.mark.gpu
def test_gpu_mod_raise_and_bootstrap_match_the_oracle_limb_for_limb():
    from fhe_spear_b200 import pyPhantom as ph
    N, L0, P = 1024, 19, 2
    be, bt_ref, steps = _setup(N, L0, P)
    parms = ph.params(ph.scheme_type.ckks)
    parms.set_poly_modulus_degree(N)
    parms.set_special_modulus_size(P)
    parms.set_coeff_modulus(be.moduli)
    parms.set_galois_elts(ph.ckks_bootstrapper.get_galois_elements(N, 0, [2, 2]))
    ctx = ph.context(parms)
    sk = ph.secret_key(ctx, seed=SEED)
    enc = ph.ckks_encoder(ctx)
    bt = ph.ckks_bootstrapper(enc)
    bt.setup(ctx, [2, 2])
    bt.keygen(ctx, sk)
    n = N // 2
    rng = np.random.default_rng(2)
    msg = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    ct_g = sk.encrypt_symmetric(ctx, enc.encode_complex_vector(ctx, msg, 2.0 ** 59, L0 - 1), enc_id=7)
    ct_o = be.encrypt(be.encode_complex_vector(None, msg, 2.0 ** 59, L0 - 1), enc_id=7)
    assert np.array_equal(ct_g.to_numpy(), ct_o.a)
    one_g = ph.mod_switch_to_next(ctx, ct_g)
    raised = ph.mod_raise(ctx, one_g, 1)
    assert np.array_equal(raised.to_numpy(), be.mod_raise(None, be.mod_switch_to_next(None, ct_o), 1).a)
    pending = bt.bootstrap(ctx, ct_g)                  # reference convention: one rescale left to the caller
    assert abs(pending.scale() / (2.0 ** 59) ** 2 - 1) < 1e-6
    out_g = ph.rescale_to_next(ctx, pending)
    out_o = bt_ref.bootstrap(ct_o)
    assert out_g.coeff_modulus_size() == out_o.coeff_modulus_size() == L0 - ph.ckks_bootstrapper.get_bootstrap_depth([2, 2], N)
    assert np.array_equal(out_g.to_numpy(), out_o.a)                             # bit-exact, ~1500 primitive calls deep
    dec = np.array(enc.decode_complex_vector(ctx, sk.decrypt(ctx, out_g)))
    assert np.abs(dec - msg).max() < 3e-3
