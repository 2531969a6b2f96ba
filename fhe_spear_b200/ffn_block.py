"""Host mirror of the reference's fully encrypted FFN block (test_fully_enc_bsgs.py:26-125): x + ((x W_key)^2) W_val with
the squared activation computed under encryption -- BSGS mat-vecs for the key chunks, CT-CT square + relinearize +
rescale, BSGS mat-vecs for the value chunks, level alignment, residual add.  Three levels per block.

The mat-vecs go through the hoisted path of this build (`fhe_matmul_bsgs` without baby ciphertexts: diagonals are
encoded at the ciphertext's level and consumed by `spear_bsgs_hoisted`), so the separately rotated baby ciphertexts
of the reference loop are not materialised.  `reference_order=True` runs the reference's own op order instead
(separately rotated baby ciphertexts, plaintext diagonals, `bsgs_multiply_accumulate`): its output ciphertext equals
the one the reference's function produces limb for limb (tests/golden/ffn_block.npz)."""
import time

import numpy as np

from . import bsgs as hb
from . import pyPhantom as ph


PHASES = None   # set to a dict to collect host-clock seconds per phase (synchronising; tools/fully_enc_bench.py --phases)


def plaintext_ffn_block(x, W_key, W_val):
    """[ref: :121-125]"""
    return x + ((x @ W_key) ** 2) @ W_val


def _align(ctx, a, b):
    """bring two ciphertexts to the deeper of their levels  [ref: :83-91, :97-107]"""
    while a.chain_index() < b.chain_index():
        a = ph.mod_switch_to_next(ctx, a)
    while b.chain_index() < a.chain_index():
        b = ph.mod_switch_to_next(ctx, b)
    return a, b


def _matvecs(ckks, cts, mats, D, G, B, shard, reference_order=False):
    """The independent mat-vecs of one phase: diagonals encoded at the ciphertexts' level, then one batched call
    (giant-step sharded over the ranks when shard = (rank, world) with world > 1)."""
    if reference_order:   # [ref: :36-49, :67-79] baby rotations per distinct input, one reference-order mat-vec per chunk
        outs, baby, last = [], None, None
        for ct, M in zip(cts, mats):
            if ct is not last:
                baby, last = hb._compute_baby_rotations(ckks, ct, G), ct
            outs.append(hb.fhe_matmul_bsgs(ckks, ct, M, D, G, B, baby))
        return outs
    level = cts[0].chain_index()
    t0 = time.perf_counter()
    sets = [hb.pre_encode_real_diags(ckks, M, D, G, B, level, shard=shard) for M in mats]
    if PHASES is not None:
        ckks.ctx.synchronize()
        PHASES["encode_diagonals"] = PHASES.get("encode_diagonals", 0.0) + time.perf_counter() - t0
        t0 = time.perf_counter()
    if shard[1] > 1:
        from .sharding import sharded_matvec_batch
        outs = sharded_matvec_batch(ckks, cts, sets)
    else:
        outs = ph.bsgs_hoisted_batch(ckks.ctx, cts, sets, ckks.gk)
    if PHASES is not None:
        ckks.ctx.synchronize()
        PHASES["matvecs"] = PHASES.get("matvecs", 0.0) + time.perf_counter() - t0
    return outs


def fully_encrypted_ffn_block(ckks, ct_x_rep, W_key, W_val, D, F, block_idx=0, split=None, shard=(0, 1), verbose=False,
                              reference_order=False):
    """Enc(x replicated) -> (Enc(x + ((x W_key)^2) W_val), levels used)  [ref: :26-118]"""
    t0 = time.perf_counter()
    G, B = split if split else hb.compute_bsgs_params(D)
    n_chunks = int(np.ceil(F / D))
    start_level = ct_x_rep.chain_index()
    mats = []
    for c in range(n_chunks):                                   # FFN key: one mat-vec per chunk of D outputs
        lo, hi = c * D, min((c + 1) * D, F)
        mats.append(W_key[:, lo:hi].T if not reference_order else hb._padded(W_key[:, lo:hi].T, D))   # a view: no host transpose
    ct_sq = [ph.rescale_to_next(ckks.ctx, ph.relinearize(ckks.ctx, ph.multiply(ckks.ctx, fk, fk), ckks.rlk))
             for fk in _matvecs(ckks, [ct_x_rep] * n_chunks, mats, D, G, B, shard, reference_order)]
    mats = []
    for c in range(n_chunks):                                   # FFN value: chunk partials summed homomorphically
        lo, hi = c * D, min((c + 1) * D, F)
        mats.append(W_val[lo:hi, :].T if not reference_order else hb._padded(W_val[lo:hi, :].T, D))
    acc = None
    for part in _matvecs(ckks, ct_sq, mats, D, G, B, shard, reference_order):
        if acc is None:
            acc = part
        else:
            acc, part = _align(ckks.ctx, acc, part)
            acc = ph.add(ckks.ctx, acc, part)
    x_al, acc = _align(ckks.ctx, ct_x_rep, acc)
    acc.set_scale(x_al.scale())
    out = ph.add(ckks.ctx, x_al, acc)
    if verbose:
        print(f"  Block {block_idx}: levels {start_level} -> {out.chain_index()}, {time.perf_counter() - t0:.3f}s")
    return out, out.chain_index() - start_level
