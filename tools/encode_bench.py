#!/usr/bin/env python3
"""Diagonal supply chain: time of one diagonal set from the D x D matrix
(ph.diagonal_set.from_matrix: H2D of the matrix, extraction + pre-rotation on the device, encoding of D diagonals on
l + P limbs) at the sizes of BASELINE config 5 (N=16384, L0=36) and config 3 (N=32768, L0=24).
SPEAR_FUSED_ENCODE=0 selects the staged encoder (one launch per FFT / NTT stage) for the A/B."""
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
    ap.add_argument("--N", type=int, default=16384)
    ap.add_argument("--L0", type=int, default=36)
    ap.add_argument("--P", type=int, default=3)
    ap.add_argument("--D", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from fhe_spear_b200 import bsgs as hb
    from fhe_spear_b200 import pyPhantom as ph
    ckks = hb.CKKSBootstrapContext(poly_degree=a.N, L0=a.L0, prime_bits=59, special_mod_size=a.P, max_rot_dim=1,
                                   bsgs_dim=[a.D], skip_bootstrap=True, seed=bytes(range(32)), verbose=False)
    D = a.D
    G, B = hb.compute_bsgs_params(D)
    rng = np.random.default_rng(0)
    Wk = rng.standard_normal((D, 2 * D)) * 0.02
    t_prep, t_enc = [], []
    ref = None
    for _ in range(a.reps + 1):
        t0 = time.perf_counter()
        M = np.zeros((D, D))
        M[:, :] = Wk[:, :D].T                      # the host-side chunk extraction of the FFN blocks
        t1 = time.perf_counter()
        ds = ph.diagonal_set.from_matrix(ckks.ctx, M, G, B, ckks.diag_scale, chain_index=1)
        ckks.ctx.synchronize()
        t2 = time.perf_counter()
        t_prep.append(t1 - t0), t_enc.append(t2 - t1)
        del ds
    ds = ph.diagonal_set.from_matrix(ckks.ctx, M, G, B, ckks.diag_scale, chain_index=1)
    import hashlib
    digest = hashlib.sha256(ds.to_numpy()[:64].tobytes()).hexdigest()
    print(json.dumps({"what": "one diagonal set from the matrix", "N": a.N, "L0": a.L0, "D": D, "bytes": ds.info()["bytes"],
                      "fused_encoder": os.environ.get("SPEAR_FUSED_ENCODE", "1") != "0",
                      "host_matrix_prep_ms": float(np.median(t_prep[1:]) * 1e3),
                      "from_matrix_ms": float(np.median(t_enc[1:]) * 1e3), "first_64_rows_sha256": digest}))


if __name__ == "__main__":
    main()
