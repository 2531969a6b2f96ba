#!/usr/bin/env python3
"""Driver (ours) for the reference's OWN functions on the CUDA drop-in: imports the unmodified, staged
scripts/bootstrap_generation.py (tests/_ref/, see tools/stage_reference.py) with `pyPhantom` resolving to
fhe_spear_b200/pyPhantom, and runs

  fallback   the reference's Python BSGS loop (fhe_matmul_bsgs :464-484, fhe_matmul_bsgs_complex :521-542) and all three
             shapes of fhe_projection_bsgs with the fused fork-only entry points REMOVED from the module, so that
             rotate / multiply_plain / add / rescale_to_next are driven one call at a time exactly as the reference
             drives PhantomFHE without its fork;
  fused      the same with the fork-only entry points present (bsgs_multiply_accumulate, encode_*_vector_batch);
  block      the reference's client_aided_block against its own plaintext_block on random RWKV-7 weights built through
             the reference's RWKVBlockWeights (pre-encoded and CPU-offloaded plaintexts included).

Prints one line per check, ending in `match` or `MISMATCH`; exit code 1 on any mismatch."""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["fallback", "fused", "block"])
    ap.add_argument("--D", type=int, default=16)
    ap.add_argument("--N", type=int, default=4096)
    a = ap.parse_args()
    sys.path.insert(0, os.path.join(ROOT, "fhe_spear_b200"))     # `import pyPhantom` -> the CUDA drop-in
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "scripts"))
    import pyPhantom as ph
    assert "fhe_spear_b200" in ph.__file__, ph.__file__
    if a.mode == "fallback":
        for name in ("bsgs_multiply_accumulate", "bsgs_from_cpu", "bsgs_complete_from_cpu"):
            if hasattr(ph, name):
                delattr(ph, name)
        for name in ("encode_double_vector_batch", "encode_complex_vector_batch"):
            if hasattr(ph.ckks_encoder, name):
                delattr(ph.ckks_encoder, name)
    import bootstrap_generation as bg
    assert bg.__file__.startswith(REF), bg.__file__
    D, F = a.D, 2 * a.D
    ok = True

    def check(label, got, want, tol=1e-7):
        nonlocal ok
        err = float(np.abs(np.asarray(got) - np.asarray(want)).max())
        good = err < tol
        ok &= good
        print(f"{a.mode}: {label}: max_err={err:.3e} ({'match' if good else 'MISMATCH'})", flush=True)

    rng = np.random.default_rng(7)
    if a.mode in ("fallback", "fused"):
        ckks = bg.CKKSBootstrapContext(poly_degree=a.N, L0=4, prime_bits=59, special_mod_size=2, max_rot_dim=F,
                                       bsgs_dim=[D, F], skip_bootstrap=True)
        W, x = rng.standard_normal((D, D)) * 0.1, rng.standard_normal(D)
        G, B = bg.compute_bsgs_params(D)
        ct = ckks.encrypt_replicated(x)
        check("fhe_matmul_bsgs", ckks.decrypt_vec(bg.fhe_matmul_bsgs(ckks, ct, W, D, G, B), D), W @ x)
        W2 = rng.standard_normal((D, D)) * 0.1
        yc = ckks.decrypt_vec_complex(bg.fhe_matmul_bsgs_complex(ckks, ct, W, W2, D, G, B), D)
        check("fhe_matmul_bsgs_complex", yc, W @ x + 1j * (W2 @ x))
        check("fhe_projection_bsgs D->D", bg.fhe_projection_bsgs(ckks, x, W, D, D), x @ W)
        Wk = rng.standard_normal((D, F)) * 0.1
        check("fhe_projection_bsgs D->F", bg.fhe_projection_bsgs(ckks, x, Wk, D, F), x @ Wk)
        Wv, xf = rng.standard_normal((F, D)) * 0.1, rng.standard_normal(F)
        check("fhe_projection_bsgs F->D", bg.fhe_projection_bsgs(ckks, xf, Wv, F, D), xf @ Wv)
    else:
        import torch
        H, S = 2, D // 2
        F = 4 * D
        w = {}
        for b in range(2):
            p = f"blocks.{b}."
            t = lambda *shape, s=1.0: torch.from_numpy(rng.standard_normal(shape) * s).float()
            for k in ("ln1", "ln2", "att.ln_x"):
                w[p + k + ".weight"], w[p + k + ".bias"] = torch.ones(D), torch.zeros(D)
            for k in ("att.x_r", "att.x_k", "att.x_v", "att.x_g", "att.x_w", "att.x_a", "ffn.x_k", "att.k_k", "att.k_a"):
                w[p + k] = torch.from_numpy(rng.uniform(0.2, 1.0, D)).float()
            for k, r in (("w", 8), ("a", 8), ("v", 8)):
                w[p + f"att.{k}0"], w[p + f"att.{k}1"], w[p + f"att.{k}2"] = t(D, s=0.1), t(D, r, s=0.01), t(r, D, s=0.01)
            w[p + "att.g1"], w[p + "att.g2"], w[p + "att.r_k"] = t(D, 8, s=0.01), t(8, D, s=0.01), t(H, S, s=0.01)
            for k in ("receptance", "key", "value", "output"):
                w[p + f"att.{k}.weight"] = t(D, D, s=0.1)
            w[p + "ffn.key.weight"], w[p + "ffn.value.weight"] = t(D, F, s=0.1), t(F, D, s=0.1)
        ckks = bg.CKKSBootstrapContext(poly_degree=a.N, L0=4, prime_bits=59, special_mod_size=2, max_rot_dim=1,
                                       bsgs_dim=[D, F], skip_bootstrap=True)
        x = rng.standard_normal(D)
        xf, xp = x.copy(), x.copy()
        st_f = st_p = np.zeros((H, S, S))
        pa_f = pa_p = pf_f = pf_p = np.zeros(D)
        vf_f = vf_p = None
        for b in range(2):
            blk = bg.RWKVBlockWeights(w, b, D, F, H, S)
            pe = bg.pre_encode_block(ckks, blk, D, F) if b == 0 else None
            cpu = bg.offload_block_plaintexts(pe) if pe is not None and hasattr(ph, "offload_plaintexts") else None
            out = bg.client_aided_block(ckks, blk, xf, pa_f, pf_f, st_f, vf_f, use_bsgs=True,
                                        preencoded_block=None if cpu else pe, cpu_offloaded_block=cpu)
            xf, pa_f, pf_f, st_f, vf_f = out[:5]
            xp, pa_p, pf_p, st_p, vf_p = bg.plaintext_block(blk, xp, pa_p, pf_p, st_p, vf_p)[:5]
            check(f"client_aided_block {b} vs plaintext_block", xf, xp, tol=1e-6)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
