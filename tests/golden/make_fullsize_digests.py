#!/usr/bin/env python3
"""SHA-256 digests of the ORACLE's ciphertext limbs at the full BASELINE sizes (C3, C2, the first level of C5) and on
the hoisting-aware splits the token benchmark runs on.  Run in the build container (CPU only, ~15 min on 8 cores):

    python tests/golden/make_fullsize_digests.py [case ...]

The oracle needs minutes per case at these sizes (keys for up to 142 rotations, 2048 diagonal encodings, the hoisted
BSGS itself), which is GPU-box time nobody should pay at test time: the digests are committed as
tests/golden/fullsize_digests.json and tests/test_gpu_fullsize_parity.py compares the SHA-256 of the CUDA library's
limbs with them -- same seeds, same inputs, same algorithm, bit for bit.  Inputs are fully determined by the seeds
below (numpy default_rng) and the 32-byte key seed; nothing here reads /root/reference.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import Setup, rolled_diagonals, tile  # noqa: E402

OUT = os.path.join(HERE, "fullsize_digests.json")

# name: N, L0 (data limbs), P, D, [(G, B, what)], complex diagonals?
CASES = {
    # BASELINE config 3 / 4: the reference split, the hoisting-aware splits of tools/token_bench.py (chunked diagonal
    # MAC for G > 64), one shard of a world-8 giant-step split, complex-packed diagonals (FFN key shape)
    "c3": dict(N=32768, L0=24, P=3, D=2048, splits=[(46, 45), (128, 16), (91, 23)], shards=[(46, 45, 0, 8), (46, 45, 5, 8)],
               complex_split=(46, 45), primitives=True),
    # BASELINE config 2
    "c2": dict(N=16384, L0=24, P=3, D=1024, splits=[(32, 32)], shards=[], complex_split=None, primitives=False),
    # BASELINE config 5, first level: l = 36 limbs, beta = 12 digits
    "c5": dict(N=16384, L0=36, P=3, D=2048, splits=[(46, 45)], shards=[], complex_split=None, primitives=False),
}
DIAG_ROWS = (0, 1, 45, 46, 2047)   # rows of the diagonal set whose encodings are digested (clipped to D)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def inputs(D, seed=0):
    rng = np.random.default_rng(seed)
    W = rng.standard_normal((D, D)) * 0.02
    W2 = rng.standard_normal((D, D)) * 0.02
    x = rng.standard_normal(D) * 0.1
    return W, W2, x


def run_case(name, cfg):
    N, L0, P, D = cfg["N"], cfg["L0"], cfg["P"], cfg["D"]
    t0 = time.time()
    S = Setup(N=N, bits=(59,) * (L0 + P), P=P)
    o = S.o
    W, W2, x = inputs(D)
    pt = o.encode(tile(x, N // 2).astype(complex), S.scale, S.L)
    ct = o.encrypt_symmetric(S.seed, 1, S.sk, pt)
    res = {"params": {"N": N, "L0": L0, "P": P, "D": D, "scale_log2": 59, "enc_id": 1, "input_seed": 0},
           "ct_x": sha(ct), "splits": {}, "shards": {}, "reference_split": "%dx%d" % cfg["splits"][0]}
    print(f"[{name}] context + input {time.time() - t0:.1f}s", flush=True)

    def encode_set(rolled, G, B, rows=None):
        rows = range(D) if rows is None else rows
        return np.stack([o.encode(rolled[k].astype(complex), S.scale, S.L, ext=True, n=2 * D) for k in rows])

    def keys_for(G, B):
        steps = list(range(1, G)) + [g * G for g in range(1, B)]
        t = time.time()
        have = len(S.keys)
        keys = S.keys_for_steps(steps)
        print(f"[{name}] keys for G={G} B={B}: +{len(S.keys) - have} in {time.time() - t:.1f}s", flush=True)
        return keys

    for G, B in cfg["splits"]:
        keys = keys_for(G, B)
        t = time.time()
        rolled = rolled_diagonals(W, D, G, B)
        diag = encode_set(rolled, G, B)
        print(f"[{name}] {D} diagonals encoded in {time.time() - t:.1f}s", flush=True)
        t = time.time()
        y = o.bsgs_hoisted(ct, diag, G, B, D, keys)
        dec = o.decode(o.decrypt(S.sk, y), S.scale * S.scale / float(S.q[S.L - 1]))[:D].real
        err = float(np.abs(dec - W @ x).max())
        assert err < 1e-9, err
        res["splits"][f"{G}x{B}"] = {"y": sha(y), "max_abs_err": err,
                                     "diag_rows": {str(k): sha(diag[k]) for k in DIAG_ROWS if k < D}}
        print(f"[{name}] hoisted {G}x{B} in {time.time() - t:.1f}s, err {err:.2e}", flush=True)
        for (g, b, rank, world) in cfg["shards"]:
            if (g, b) != (G, B):
                continue
            t = time.time()
            rows = [k for gg in range(rank, B, world) for k in range(gg * G, min((gg + 1) * G, D))]
            R = o.bsgs_hoisted_partial(ct, diag[rows], G, B, D, keys, g_first=rank, g_stride=world)
            res["shards"][f"{G}x{B}:{rank}/{world}"] = sha(R)
            print(f"[{name}] shard {rank}/{world} in {time.time() - t:.1f}s", flush=True)
        if cfg["complex_split"] == (G, B):
            t = time.time()
            rolled_c = rolled_diagonals(W, D, G, B) + 1j * rolled_diagonals(W2, D, G, B)
            diag_c = np.stack([o.encode(rolled_c[k], S.scale, S.L, ext=True, n=2 * D) for k in range(D)])
            yc = o.bsgs_hoisted(ct, diag_c, G, B, D, keys)
            decc = o.decode(o.decrypt(S.sk, yc), S.scale * S.scale / float(S.q[S.L - 1]))[:D]
            errc = float(np.abs(decc - (W @ x + 1j * (W2 @ x))).max())
            assert errc < 1e-9, errc
            res["complex"] = {"split": f"{G}x{B}", "y": sha(yc), "max_abs_err": errc}
            print(f"[{name}] complex-packed {G}x{B} in {time.time() - t:.1f}s, err {errc:.2e}", flush=True)
            del diag_c
        if cfg["primitives"] and (G, B) == cfg["splits"][0]:
            # the un-hoisted primitives of the reference's own loop at full size (bootstrap_generation.py:215-220, 464-484)
            r3 = o.rotate(ct, 3, keys)
            pt_d = o.encode(tile(rolled[1], N // 2).astype(complex), S.scale, S.L)
            prod = o.multiply_plain(ct, pt_d)
            res["primitives"] = {"rotate_3": sha(r3), "multiply_plain_diag1": sha(prod), "rescale": sha(o.rescale(prod)),
                                 "hoisted_rotation_3": sha(o.hoisted_rotation(ct, o.elt_from_step(3), keys[o.elt_from_step(3)]))}
        del diag
    res["seconds"] = round(time.time() - t0, 1)
    return res


def main():
    want = sys.argv[1:] or list(CASES)
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in want:
        out[name] = run_case(name, CASES[name])
        json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    print("written", OUT)


if __name__ == "__main__":
    main()
