"""CPU tests of the host-side mirror (fhe_spear_b200/bsgs.py) against the golden fixtures produced by the
reference's own numpy code (tests/golden/make_golden.py): planning, diagonal extraction, pre-rotation and
slot tiling must be identical; chunk packing for D->F / F->D must match the reference matrices."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_planning_matches_reference():
    from fhe_spear_b200 import bsgs as hb
    gold = json.load(open(os.path.join(GOLD, "bsgs_plan.json")))
    for case in gold["params"]:
        assert list(hb.compute_bsgs_params(case["D"])) == [case["G"], case["B"]]
    for case in gold["galois"]:
        assert hb.bsgs_steps(case["D"]) == case["steps"]
        assert hb.compute_bsgs_galois_elements(case["N"], case["D"]) == case["bsgs_elts"]
        got = hb.compute_rotation_galois_elements(case["N"], (1 << (case["n_pow2"] - 1)))
        assert sorted(got) == case["rot_elts"]
    assert hb.compute_bsgs_params(2048) == (46, 45)      # 89 rotations, reference README.md:17


def test_diagonals_match_reference():
    from fhe_spear_b200 import bsgs as hb
    npz = np.load(os.path.join(GOLD, "bsgs_diagonals.npz"))
    for D in (8, 20):
        W = npz[f"W_{D}"]
        G, _ = hb.compute_bsgs_params(D)
        gold = npz[f"rolled_{D}"]
        slots = gold.shape[1]
        d = hb._extract_diagonals(W, D)
        assert np.array_equal(d, np.stack(hb.compute_diagonals(W, D)))
        assert np.array_equal(hb._tile_rows(hb._pre_rotate(d, D, G), slots), gold)
        assert np.array_equal(hb._replicate_to_slots(W[0], slots), np.array(hb.replicate_vector(W[0], slots)))


def test_chunk_packing_reproduces_projection():
    """D->F pairs output chunks as (re, im); F->D pairs input chunks with a negated second matrix."""
    from fhe_spear_b200 import bsgs as hb
    rng = np.random.default_rng(0)
    D, F = 8, 20                                     # ragged: 3 chunks, the last one short
    Wk, Wv = rng.standard_normal((D, F)), rng.standard_normal((F, D))
    x, xf = rng.standard_normal(D), rng.standard_normal(F)
    up = np.zeros(F)
    pad = lambda M: hb._padded(M, D)                 # chunk helpers return views; the device zero-pads them to D x D
    for c, c2 in hb._chunk_pairs(F, D):
        lo, hi = c * D, min((c + 1) * D, F)
        up[lo:hi] = (pad(hb._key_chunk(Wk, c, D, F)) @ x)[:hi - lo]
        if c2 is not None:
            lo2, hi2 = c2 * D, min((c2 + 1) * D, F)
            up[lo2:hi2] = (pad(hb._key_chunk(Wk, c2, D, F)) @ x)[:hi2 - lo2]
    assert np.allclose(up, x @ Wk)
    down = np.zeros(D)
    for c, c2 in hb._chunk_pairs(F, D):
        x0 = np.zeros(D)
        lo, hi = c * D, min((c + 1) * D, F)
        x0[:hi - lo] = xf[lo:hi]
        M0 = pad(hb._val_chunk(Wv, c, D, F))
        if c2 is None:
            down += M0 @ x0
        else:
            x1 = np.zeros(D)
            lo1, hi1 = c2 * D, min((c2 + 1) * D, F)
            x1[:hi1 - lo1] = xf[lo1:hi1]
            M1n = pad(hb._val_chunk(Wv, c2, D, F, -1.0))
            down += ((M0 + 1j * M1n) @ (x0 + 1j * x1)).real     # diagonals d0 + i*(-d1): Enc(x0 + i x1) * (d0 - i d1)
    assert np.allclose(down, xf @ Wv)


def test_plaintext_block_matches_reference():
    """tests/golden/rwkv_block.npz: outputs of the reference's plaintext_block (bootstrap_generation.py:902-980)
    on the seeded random weights of RWKVBlockWeights.random; the mirror must agree to rounding."""
    from fhe_spear_b200.rwkv_block import RWKVBlockWeights, plaintext_block
    g = np.load(os.path.join(GOLD, "rwkv_block.npz"))
    D, F, H, S = (int(v) for v in g["dims"])
    for idx in (0, 1):
        blk = RWKVBlockWeights.random(D, F, H, S, block_idx=idx, seed=5 + idx)
        out = plaintext_block(blk, g["x"], g["x_prev_att"], g["x_prev_ffn"], g["state"], g["v_first"])
        for name, val in zip(("x", "xpa", "xpf", "state", "v_first"), out):
            assert np.abs(np.asarray(val) - g[f"b{idx}_{name}"]).max() < 1e-12, (idx, name)
