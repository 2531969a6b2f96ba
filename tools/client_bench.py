#!/usr/bin/env python3
"""Client legs of one projection round trip at config C3: encode + encrypt of a replicated
d=2048 vector and decrypt + decode of the result, one-call forms (csrc/client.cu, three launches each) against the
two-step forms of the reference's CKKSBootstrapContext (scripts/bootstrap_generation.py:119-147).  Host wall clock per
call (what a token loop pays), synchronised."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=32768)
    ap.add_argument("--L0", type=int, default=24)
    ap.add_argument("--D", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=200)
    a = ap.parse_args()
    from fhe_spear_b200 import bsgs as hb
    from fhe_spear_b200 import pyPhantom as ph
    from fhe_spear_b200 import _native
    ckks = hb.CKKSBootstrapContext(poly_degree=a.N, L0=a.L0, prime_bits=59, special_mod_size=3, max_rot_dim=1,
                                   bsgs_dim=0, skip_bootstrap=True, seed=bytes(range(32)), verbose=False)
    ctx, sk, enc = ckks.ctx, ckks.sk, ckks.encoder
    x = np.random.default_rng(0).standard_normal(a.D)
    rep = hb._replicate_to_slots(x, ckks.slots)

    def two_step_enc():
        return sk.encrypt_symmetric(ctx, enc.encode_double_vector(ctx, rep, ckks.scale))

    def one_call_enc():
        return sk.encrypt_vector(ctx, x, ckks.scale, replicate=True)
    ct = ph.rescale_to_next(ctx, one_call_enc())          # a result ciphertext has l - 1 limbs

    def two_step_dec():
        return enc.decode_array(ctx, sk.decrypt(ctx, ct))[:a.D]

    def one_call_dec():
        return sk.decrypt_decode(ctx, ct, a.D)
    out = {}
    for name, fn in (("encode_encrypt_two_step", two_step_enc), ("encode_encrypt_one_call", one_call_enc),
                     ("decrypt_decode_two_step", two_step_dec), ("decrypt_decode_one_call", one_call_dec)):
        for _ in range(10):
            fn()
        ctx.synchronize()
        l0 = _native.launch_count()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            fn()
        ctx.synchronize()
        out[name] = {"ms": (time.perf_counter() - t0) / a.reps * 1e3, "launches": (_native.launch_count() - l0) / a.reps}
    assert np.array_equal(two_step_dec(), one_call_dec())
    print(json.dumps({"what": "client legs per call, host clock, synchronised at the end of the batch", "N": a.N, "L0": a.L0,
                      "D": a.D, **out}))


if __name__ == "__main__":
    main()
