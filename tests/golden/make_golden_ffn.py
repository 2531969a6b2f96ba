#!/usr/bin/env python3
"""Golden fixture for the fully encrypted FFN block, generated FROM THE REFERENCE ITSELF.
Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_ffn.py

The reference's test_fully_enc_bsgs.py is imported unmodified (with scripts/bootstrap_generation.py behind it) over the
oracle-backed `pyPhantom` look-alike of make_golden.py, extended by the calls this flow adds (multiply, relinearize,
mod_switch_to_next, set_scale, a real relinearisation key).  Its own fully_encrypted_ffn_block (:26-118) -- shared baby
rotations, the Python BSGS loop, CT-CT square, per-chunk value mat-vecs, level alignment, set_scale, residual add --
runs on a small parameter set; input, weights and the output ciphertext's limbs go to ffn_block.npz.  The mirror
fhe_spear_b200/ffn_block.py in its reference-order mode must reproduce the limbs on the CUDA library (GPU test).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (shim + oracle state)

REF = mg.REF


def main():
    ph = mg._shim()
    ST, Ct = mg.ST, mg.Ct

    # ---- what test_fully_enc_bsgs.py uses beyond the BSGS layer ------------------------------------------------------
    def gen_relinkey(self, ctx):
        return ST.o.gen_relin_key(mg.SEED, ST.sk)
    ph.secret_key.gen_relinkey = gen_relinkey
    Ct.set_scale = lambda self, s: setattr(self, "_scale", float(s))
    ph.multiply = lambda ctx, a, b: Ct(ST.o.multiply(a.a, b.a), a._scale * b._scale)
    ph.relinearize = lambda ctx, ct, rlk: Ct(ST.o.relinearize(ct.a, rlk), ct._scale)
    ph.mod_switch_to_next = lambda ctx, ct: Ct(np.ascontiguousarray(ct.a[:, :-1]), ct._scale)
    sys.modules["pyPhantom"] = ph
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "scripts"))
    import test_fully_enc_bsgs as tf   # the unmodified reference (imports bootstrap_generation itself)

    N, L0, P, D, F = 256, 6, 2, 8, 16
    ckks = tf.CKKSBootstrapContext(poly_degree=N, L0=L0, prime_bits=59, special_mod_size=P, max_rot_dim=D, bsgs_dim=[D],
                                   skip_bootstrap=True)
    rng = np.random.default_rng(77)
    W_key, W_val = rng.standard_normal((D, F)) * 0.3, rng.standard_normal((F, D)) * 0.3
    x = rng.standard_normal(D) * 0.5
    ST.enc_counter = 0
    ct_x = ckks.encrypt_replicated(x)
    out, used = tf.fully_encrypted_ffn_block(ckks, ct_x, W_key, W_val, D, F)
    dec = ckks.decrypt_vec(out, D)
    want = tf.plaintext_ffn_block(x, W_key, W_val)
    assert np.abs(dec - want).max() < 1e-6 and used == 3, (np.abs(dec - want).max(), used)
    np.savez_compressed(os.path.join(HERE, "ffn_block.npz"), N=N, L0=L0, P=P, D=D, F=F, W_key=W_key, W_val=W_val, x=x,
                        seed=np.frombuffer(mg.SEED, dtype=np.uint8), enc_id_x=1, ct_x=ct_x.a, ct_out=out.a,
                        out_scale=out._scale, levels_used=used, dec=dec, plaintext=want)
    print("ffn_block.npz written: levels used", used, "max |err| vs float64", float(np.abs(dec - want).max()))


if __name__ == "__main__":
    main()
