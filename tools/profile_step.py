#!/usr/bin/env python3
"""Set up config C3 (or another) exactly like bench.py and run a few hoisted BSGS mat-vecs.
Used under ncu:  S=$(python tools/profile_step.py --count-only) && ncu -s $S -c ... python tools/profile_step.py
--count-only prints the number of kernel launches that precede the profiled steps (setup + warm-up)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--count-only", action="store_true")
    ap.add_argument("--mode", default="hoisted", choices=["hoisted", "exact"])
    ap.add_argument("--weight", type=float, default=1.0, help="split G = ceil(sqrt(weight * D)); 1 = the reference's")
    ap.add_argument("--classes", action="store_true", help="print per-kernel-class times (event pair around each launch)")
    a = ap.parse_args()
    from fhe_spear_b200 import _native
    from fhe_spear_b200 import bsgs as hb
    from fhe_spear_b200 import pyPhantom as ph
    N, L0, P, D = bench.CONFIGS[a.config]
    G, B = hb.compute_bsgs_params(D, a.weight)
    ckks = hb.CKKSBootstrapContext(poly_degree=N, L0=L0, prime_bits=59, special_mod_size=P, max_rot_dim=1,
                                   bsgs_dim=[D], skip_bootstrap=True, seed=bytes(range(32)), verbose=False,
                                   baby_weights=(a.weight,))
    rng = np.random.default_rng(1000)
    W, x = rng.standard_normal((D, D)) * 0.02, rng.standard_normal(D) * 0.1
    diags = hb.pre_encode_real_diags(ckks, W, D, G, B, level=1)
    ct = ckks.encrypt_replicated(x)
    for _ in range(a.warmup):
        ph.bsgs_hoisted(ckks.ctx, ct, diags, ckks.gk)
    ckks.ctx.synchronize()
    if a.count_only:
        print(_native.launch_count())
        return
    if a.classes:
        ckks.ctx.timer_start()
        for _ in range(a.steps):
            ph.bsgs_hoisted(ckks.ctx, ct, diags, ckks.gk)
        print(f"G={G} B={B}: {ckks.ctx.timer_stop() / a.steps:.3f} ms per mat-vec")
        ckks.ctx.profile(True)
    for _ in range(a.steps):
        y = ph.bsgs_hoisted(ckks.ctx, ct, diags, ckks.gk)
    ckks.ctx.synchronize()
    if a.classes:
        for k, v in ckks.ctx.profile_read().items():
            print(f"  {k:14s} {v['ms'] / a.steps:8.3f} ms  {v['launches'] // a.steps:4d} launches")
        ckks.ctx.profile(False)
    err = float(np.abs(ckks.decrypt_vec(y, D) - W @ x).max())
    print(f"steps={a.steps} max_abs_err={err:.3e} launches={_native.launch_count()}")


if __name__ == "__main__":
    main()
