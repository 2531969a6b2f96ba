"""GPU tests of the callers either side of the hot path: the reference's own loop output (golden limbs),
giant-step sharding emulated on one GPU, and the host-side mirror of fhe_projection_bsgs."""
import os

import numpy as np
import pytest

from helpers import SEED, Setup, bsgs_params, rolled_diagonals, tile

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_reference_python_loop_golden_limbs():
    """tests/golden/bsgs_loop.npz: ciphertext limbs produced by the reference's fhe_matmul_bsgs /
    fhe_matmul_bsgs_complex Python loops (scripts/bootstrap_generation.py:464-484, :521-542).
    The CUDA path must reproduce them from the same seed, inputs and randomness."""
    g = np.load(os.path.join(GOLD, "bsgs_loop.npz"))
    from fhe_spear_b200 import pyPhantom as ph
    N, L0, P, D = int(g["N"]), int(g["L0"]), int(g["P"]), int(g["D"])
    G, B = bsgs_params(D)
    steps = list(range(1, G)) + [k * G for k in range(1, B)]
    parms = ph.params(ph.scheme_type.ckks)
    parms.set_poly_modulus_degree(N)
    parms.set_special_modulus_size(P)
    parms.set_coeff_modulus([int(q) for q in g["moduli"]])
    parms.set_galois_elts(ph.get_elts_from_steps(steps, N))
    ctx = ph.context(parms)
    sk = ph.secret_key(ctx, seed=bytes(g["seed"]))
    gk = sk.create_galois_keys(ctx)
    enc = ph.ckks_encoder(ctx)
    scale = 2.0 ** 59
    slots = N // 2
    ct = sk.encrypt_symmetric(ctx, enc.encode_double_vector(ctx, tile(g["x"], slots), scale), enc_id=int(g["enc_id_x"]))
    assert np.array_equal(ct.to_numpy(), g["ct_x"])
    baby = [ct] + [ph.rotate(ctx, ct, b, gk) for b in range(1, G)]
    rolled = rolled_diagonals(g["W"], D, G, B)
    pts = enc.encode_double_vector_batch(ctx, np.stack([tile(r, slots) for r in rolled]), scale, chain_index=1)
    y = ph.bsgs_multiply_accumulate(ctx, baby, pts, G, B, D, gk)
    assert np.array_equal(y.to_numpy(), g["ct_y_real"])
    rolled_c = rolled + 1j * rolled_diagonals(g["W2"], D, G, B)
    pts_c = enc.encode_complex_vector_batch(ctx, np.stack([tile(r, slots) for r in rolled_c]), scale, chain_index=1)
    yc = ph.bsgs_multiply_accumulate(ctx, baby, pts_c, G, B, D, gk)
    assert np.array_equal(yc.to_numpy(), g["ct_y_complex"])
    dec = np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, y)))[:D]
    assert np.array_equal(dec, g["y_real_dec"])                      # decode is bit-identical too
    # offload / upload round trip and the *_from_cpu entry points (reference :336-358, :449)
    data, ci, sc, cms, pmd = ph.offload_plaintexts(pts)
    assert (ci, cms, pmd) == (1, L0, N) and data.shape == (D, L0, N)
    y2 = ph.bsgs_from_cpu(ctx, baby, data, ci, sc, cms, pmd, G, B, D, gk)
    assert np.array_equal(y2.to_numpy(), g["ct_y_real"])
    y3 = ph.bsgs_complete_from_cpu(ctx, ct, data, ci, sc, cms, pmd, G, B, D, gk)
    assert np.array_equal(y3.to_numpy(), g["ct_y_real"])
    pageable = np.array(data)                                        # the same stream from ordinary host memory
    assert np.array_equal(ph.bsgs_from_cpu(ctx, baby, pageable, ci, sc, cms, pmd, G, B, D, gk).to_numpy(), g["ct_y_real"])
    back = ph.upload_plaintexts(data, ci, sc, cms, pmd)
    assert len(back) == D and all(np.array_equal(a.to_numpy(), b.to_numpy()) for a, b in zip(back[:3], pts[:3]))
    del data, pageable


@pytest.mark.parametrize("D,world", [(64, 2), (64, 3), (20, 2)])
def test_giant_sharding_on_one_gpu(D, world):
    """Each shard's accumulator equals the oracle's; their sum (mod-add or lazy integer sum + one
    reduction) finished once equals the unsharded result."""
    from fhe_spear_b200 import sharding as sh
    S = Setup(N=2048, bits=(59,) * 6, P=2)
    G, B = bsgs_params(D)
    steps = list(range(1, G)) + [g * G for g in range(1, B)]
    ph, ctx, sk = S.gpu(steps)
    gk = sk.create_galois_keys(ctx)
    enc = ph.ckks_encoder(ctx)
    keys = S.keys_for_steps(steps)
    rng = np.random.default_rng(D + world)
    W, x = rng.standard_normal((D, D)) * 0.1, rng.standard_normal(D)
    rolled = rolled_diagonals(W, D, G, B)
    ct = sk.encrypt_symmetric(ctx, enc.encode_double_vector(ctx, tile(x, S.N // 2), S.scale), enc_id=3)
    full = ph.diagonal_set(ctx, rolled, G, B, S.scale)
    ref = ph.bsgs_hoisted(ctx, ct, full, gk)
    accs = []
    for r in range(world):
        shard = ph.diagonal_set(ctx, rolled, G, B, S.scale, shard=(r, world))
        assert shard.rows == sh.shard_rows(D, G, B, r, world)
        acc = ph.bsgs_hoisted_partial(ctx, ct, shard, gk)
        exp = S.o.bsgs_hoisted_partial(ct.to_numpy(), shard.to_numpy(), G, B, D, keys, g_first=r, g_stride=world)
        assert np.array_equal(acc.to_numpy(), exp), r
        accs.append(acc)
    shards = [ph.diagonal_set(ctx, rolled, G, B, S.scale, shard=(r, world)) for r in range(world)]
    batch = ph.bsgs_hoisted_partial_batch(ctx, [ct] * world, shards, gk)   # same accumulators, concurrent streams
    for a, b in zip(accs, batch):
        assert np.array_equal(a.to_numpy(), b.to_numpy())
    lazy = sum(a.to_numpy().astype(object) for a in accs)            # what an integer all-reduce would hold
    assert int(lazy.max()) < 1 << 63
    total = accs[0]
    for a in accs[1:]:
        total = ph.add(ctx, total, a)                                # modular-add path
    lazy_obj = ph.ciphertext.from_numpy(ctx, lazy.astype(np.uint64), total.scale(), ext=True)
    ph.reduce_inplace(ctx, lazy_obj)                                 # lazy-sum path
    assert np.array_equal(lazy_obj.to_numpy(), total.to_numpy())
    y = ph.bsgs_finish(ctx, total)
    assert np.array_equal(y.to_numpy(), ref.to_numpy())
    dec = np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, y)))[:D]
    assert np.abs(dec - W @ x).max() < 1e-9


@pytest.mark.parametrize("D,world,weight", [(64, 2, 1.0), (64, 3, 1.0), (20, 2, 1.0), (16, 4, 1.0), (64, 8, 1.0), (64, 6, 1.0), (20, 8, 1.0), (512, 2, 32.0), (512, 8, 32.0),
                                             (512, 3, 32.0)])
def test_two_phase_split_on_one_gpu(D, world, weight):
    """Two-phase mat-vec (rows split for the baby steps and the diagonal MAC, giant groups split for the giant steps;
    include/spear_b200.h) emulated on one GPU: the summed accumulator equals the unsharded accumulator limb for limb,
    hence (the unsharded path is pinned to the oracle) the oracle's.  Covers the TMA-staged MAC (rshift 1..5), its
    fall-back (D = 16: rshift 6; D = 20: full ring) and sets walked in two baby-step chunks (G = 128)."""
    S = Setup(N=2048, bits=(59,) * 6, P=2)
    G = int(np.ceil(np.sqrt(weight * D)))
    B = -(-D // G)
    steps = list(range(1, G)) + [g * G for g in range(1, B)]
    ph, ctx, sk = S.gpu(steps)
    gk = sk.create_galois_keys(ctx)
    enc = ph.ckks_encoder(ctx)
    rng = np.random.default_rng(7 * D + world)
    W, x = rng.standard_normal((D, D)) * 0.1, rng.standard_normal(D)
    rolled = rolled_diagonals(W, D, G, B)
    ct = sk.encrypt_symmetric(ctx, enc.encode_double_vector(ctx, tile(x, S.N // 2), S.scale), enc_id=5)
    full = ph.diagonal_set(ctx, rolled, G, B, S.scale)
    ref_acc = ph.bsgs_hoisted_partial(ctx, ct, full, gk)
    slices = [full.slice_rows(r, world) for r in range(world)]
    rows_total = full.info()["limbs"] + ctx.P
    whole = full.to_numpy()
    cover = np.zeros(whole.shape[1:], dtype=np.int64)
    shift = (S.N // full.info()["ring_n"]).bit_length() - 1
    for r, sl in enumerate(slices):                       # the shares partition the (rows x columns) of the set
        r0, r1, c0, c1 = ph.diagonal_set.share(full.info()["limbs"], ctx.P, S.N, r, world)
        assert sl.row_slice == (r0, r1) and sl.col_slice == (c0 >> shift, c1 >> shift)
        assert np.array_equal(sl.to_numpy(), whole[:, r0:r1, c0 >> shift:c1 >> shift])
        cover[r0:r1, c0 >> shift:c1 >> shift] += 1
    assert cover.shape[0] == rows_total and (cover == 1).all()
    acc = ph.bsgs_split_selftest(ctx, ct, slices, gk)
    assert np.array_equal(acc.to_numpy(), ref_acc.to_numpy())
    y = ph.bsgs_finish(ctx, acc)
    assert np.array_equal(y.to_numpy(), ph.bsgs_hoisted(ctx, ct, full, gk).to_numpy())
    dec = np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, y)))[:D]
    assert np.abs(dec - W @ x).max() < 1e-9
    with pytest.raises(RuntimeError):                     # a row slice is not a giant-group shard
        ph.bsgs_hoisted(ctx, ct, slices[0], gk)


@pytest.mark.parametrize("D,weight", [(64, 1.0), (512, 32.0)])
def test_shared_baby_steps_match_separate_calls(D, weight):
    """Several diagonal sets times one ciphertext (the chunk pairs of a D -> F projection, for which the reference computes
    the baby rotations once, scripts/bootstrap_generation.py:575-600): spear_bsgs_hoisted_shared and -- through a one-rank
    window -- spear_bsgs_split_shared return the limbs of separate bsgs_hoisted calls."""
    S = Setup(N=2048, bits=(59,) * 6, P=2)
    G = int(np.ceil(np.sqrt(weight * D)))
    B = -(-D // G)
    steps = list(range(1, G)) + [g * G for g in range(1, B)]
    ph, ctx, sk = S.gpu(steps)
    gk = sk.create_galois_keys(ctx)
    enc = ph.ckks_encoder(ctx)
    rng = np.random.default_rng(3 * D)
    Ws = [rng.standard_normal((D, D)) * 0.1 for _ in range(3)]
    x = rng.standard_normal(D)
    ct = sk.encrypt_symmetric(ctx, enc.encode_double_vector(ctx, tile(x, S.N // 2), S.scale), enc_id=9)
    sets = [ph.diagonal_set(ctx, rolled_diagonals(W, D, G, B), G, B, S.scale) for W in Ws]
    refs = [ph.bsgs_hoisted(ctx, ct, d, gk).to_numpy() for d in sets]
    for _ in range(2):
        outs = ph.bsgs_hoisted_shared(ctx, ct, sets, gk)
        assert all(np.array_equal(y.to_numpy(), r) for y, r in zip(outs, refs))
    dec = np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, outs[2])))[:D]
    assert np.abs(dec - Ws[2] @ x).max() < 1e-9
    # the two-phase form over a group of one rank: same code path as on several GPUs, minus the waits
    rows = [d.slice_rows(0, 1) for d in sets]
    win = ph.peer_window(ctx, 0, 1, B * 2 * (S.L + S.P) * S.N * 8, slots=3)
    for _ in range(2):
        accs = ph.bsgs_split_shared(ctx, ct, rows, gk, win, 0)
        assert all(np.array_equal(ph.bsgs_finish(ctx, a).to_numpy(), r) for a, r in zip(accs, refs))
    accs = ph.bsgs_split_batch(ctx, [ct] * 3, rows, gk, win, 0)
    assert all(np.array_equal(ph.bsgs_finish(ctx, a).to_numpy(), r) for a, r in zip(accs, refs))
    with pytest.raises(RuntimeError):                     # sets of different splits cannot share baby steps
        other = ph.diagonal_set(ctx, rolled_diagonals(Ws[0], D, 2 * G, -(-D // (2 * G))), 2 * G, -(-D // (2 * G)), S.scale)
        ph.bsgs_hoisted_shared(ctx, ct, [sets[0], other], gk)


def test_host_buffer_serving_batch_matches_device_batch():
    """spear_bsgs_hoisted_batch_host (ciphertexts in and out of page-locked host memory, transfers pipelined with the
    mat-vecs over the engine's streams) returns exactly the limbs of the device-resident batch call, for more items than
    streams"""
    D = 64
    S = Setup(N=2048, bits=(59,) * 6, P=2)
    G, B = bsgs_params(D)
    steps = list(range(1, G)) + [g * G for g in range(1, B)]
    ph, ctx, sk = S.gpu(steps)
    gk = sk.create_galois_keys(ctx)
    enc = ph.ckks_encoder(ctx)
    rng = np.random.default_rng(21)
    Ws = [rng.standard_normal((D, D)) * 0.1 for _ in range(5)]
    xs = [rng.standard_normal(D) for _ in range(5)]
    cts = [sk.encrypt_symmetric(ctx, enc.encode_double_vector(ctx, tile(x, S.N // 2), S.scale), enc_id=40 + i)
           for i, x in enumerate(xs)]
    sets = [ph.diagonal_set(ctx, rolled_diagonals(W, D, G, B), G, B, S.scale) for W in Ws]
    ref = ph.bsgs_hoisted_batch(ctx, cts, sets, gk)
    l = cts[0].coeff_modulus_size()
    h_in = [ph.pinned_empty((2, l, S.N)) for _ in cts]
    h_out = [ph.pinned_empty((2, l - 1, S.N)) for _ in cts]
    for ct, h in zip(cts, h_in):
        ct.to_numpy(out=h)
    for _ in range(2):                                     # buffers and streams are reused
        scales = ph.bsgs_hoisted_batch_host(ctx, h_in, cts[0].scale(), sets, gk, h_out)
        for y, h, sc in zip(ref, h_out, scales):
            assert np.array_equal(y.to_numpy(), h) and sc == y.scale()
    with pytest.raises(RuntimeError):
        ph.bsgs_hoisted_batch_host(ctx, h_in, cts[0].scale(), sets, gk, h_in)   # wrong output shape


def test_host_mirror_projections():
    """fhe_projection_bsgs for D->D, D->F (complex-packed, ragged last chunk) and F->D (conjugate-packed),
    hoisted path and reference-order path, against float64 x @ W (tolerance 1e-8 at scale 2^59)."""
    from fhe_spear_b200 import bsgs as hb
    from fhe_spear_b200 import pyPhantom as ph
    D, F = 16, 40
    ckks = hb.CKKSBootstrapContext(poly_degree=2048, L0=4, prime_bits=59, special_mod_size=2, max_rot_dim=D,
                                   bsgs_dim=[D], skip_bootstrap=True, seed=SEED, verbose=False)
    rng = np.random.default_rng(11)
    W = rng.standard_normal((D, D)) * 0.1
    Wk, Wv = rng.standard_normal((D, F)) * 0.1, rng.standard_normal((F, D)) * 0.1
    x, xf = rng.standard_normal(D), rng.standard_normal(F)
    assert np.abs(hb.fhe_projection_bsgs(ckks, x, W, D, D) - x @ W).max() < 1e-8
    assert np.abs(hb.fhe_projection_bsgs(ckks, x, Wk, D, F) - x @ Wk).max() < 1e-8
    assert np.abs(hb.fhe_projection_bsgs(ckks, xf, Wv, F, D) - xf @ Wv).max() < 1e-8
    # pre-encoded forms: diagonal sets (fast path) and plaintext lists (reference op order) agree with float64
    G, B = hb.compute_bsgs_params(D)
    ct = ckks.encrypt_replicated(x)
    level = ct.chain_index()
    ds = hb.pre_encode_real_diags(ckks, W.T, D, G, B, level)
    pts = hb.pre_encode_real_diags(ckks, W.T, D, G, B, level, as_plaintexts=True)
    y_fast = ckks.decrypt_vec(hb.fhe_matmul_bsgs(ckks, ct, None, D, G, B, preencoded=ds), D)
    y_ref = ckks.decrypt_vec(hb.fhe_matmul_bsgs(ckks, ct, None, D, G, B, preencoded=pts), D)
    assert np.abs(y_fast - x @ W).max() < 1e-8 and np.abs(y_ref - x @ W).max() < 1e-8
    cpu = ph.offload_plaintexts(pts)
    y_cpu = ckks.decrypt_vec(hb.fhe_matmul_bsgs(ckks, ct, None, D, G, B, cpu_offloaded=cpu), D)
    assert np.array_equal(y_cpu, y_ref)
    # a second level: chain_index is preserved through the call surface (reference test_fully_enc_bsgs.py:32)
    y_ct = hb.fhe_matmul_bsgs(ckks, ct, W.T, D)
    assert y_ct.chain_index() == level + 1 and y_ct.coeff_modulus_size() == ckks.L0 - 1
    with pytest.raises(RuntimeError):                      # [ref: :149-151] no bootstrapper without skip_bootstrap=False
        ckks.bootstrap(ct)


def test_full_size_c3_properties():
    """BASELINE config C3 at full size (N=32768, L0=24 x 59 bit, P=3, D=2048, G=46, B=45: 89 rotations).
    The oracle would need minutes here, so this checks size-independent properties of the CUDA path:
    decrypt error against float64 W.x (bound 1e-9; the reference's acceptance is corr > 0.999,
    test_fully_enc_bsgs.py:298), linearity in the ciphertext, agreement of full-ring and sub-ring diagonal
    sets, and that one hoisted rotation decrypts to the rotated vector."""
    from fhe_spear_b200 import bsgs as hb
    from fhe_spear_b200 import pyPhantom as ph
    N, L0, P, D = 32768, 24, 3, 2048
    ckks = hb.CKKSBootstrapContext(poly_degree=N, L0=L0, prime_bits=59, special_mod_size=P, max_rot_dim=1,
                                   bsgs_dim=[D], skip_bootstrap=True, seed=SEED, verbose=False)
    G, B = hb.compute_bsgs_params(D)
    assert (G, B) == (46, 45) and len(hb.bsgs_steps(D)) == 89
    rng = np.random.default_rng(0)
    W = rng.standard_normal((D, D)) * 0.02
    x1, x2 = rng.standard_normal(D) * 0.1, rng.standard_normal(D) * 0.1
    ds = hb.pre_encode_real_diags(ckks, W, D, G, B, level=1)
    assert ds.info()["ring_n"] == 2 * D and ds.info()["bytes"] == D * (L0 + P) * 2 * D * 8
    c1, c2 = ckks.encrypt_replicated(x1), ckks.encrypt_replicated(x2)
    y1 = ph.bsgs_hoisted(ckks.ctx, c1, ds, ckks.gk)
    y2 = ph.bsgs_hoisted(ckks.ctx, c2, ds, ckks.gk)
    assert y1.chain_index() == 2 and y1.coeff_modulus_size() == L0 - 1
    d1, d2 = ckks.decrypt_vec(y1, D), ckks.decrypt_vec(y2, D)
    assert np.abs(d1 - W @ x1).max() < 1e-9 and np.abs(d2 - W @ x2).max() < 1e-9
    assert np.corrcoef(d1, W @ x1)[0, 1] > 0.999999
    y12 = ph.bsgs_hoisted(ckks.ctx, ph.add(ckks.ctx, c1, c2), ds, ckks.gk)
    assert np.abs(ckks.decrypt_vec(y12, D) - (d1 + d2)).max() < 1e-9
    # every slot block carries the same result (the replicated layout the next layer relies on)
    full = np.array(ckks.encoder.decode_double_vector(ckks.ctx, ckks.sk.decrypt(ckks.ctx, y1)))
    assert np.abs(full.reshape(-1, D) - d1).max() < 1e-9
    # exact mode on a thin slice of the same matrix: first giant group only (46 diagonals)
    r3 = ph.hoisting(ckks.ctx, c1, ckks.gk, [3])[0]
    assert np.abs(ckks.decrypt_vec(r3, D) - np.roll(x1, -3)).max() < 1e-9
    r3e = ph.rotate(ckks.ctx, c1, 3, ckks.gk)
    assert np.abs(ckks.decrypt_vec(r3e, D) - np.roll(x1, -3)).max() < 1e-9
    # D -> 2D complex-packed and 2D -> D conjugate-packed projections (config C3's FFN key / value shapes, scaled to one call each)
    Wk = rng.standard_normal((D, 2 * D)) * 0.02
    assert np.abs(hb.fhe_projection_bsgs(ckks, x1, Wk, D, 2 * D) - x1 @ Wk).max() < 1e-9


def test_batched_matvecs_match_single_calls():
    """bsgs_hoisted_batch (three independent projections on separate streams) == three single calls."""
    from fhe_spear_b200 import bsgs as hb
    from fhe_spear_b200 import pyPhantom as ph
    D = 64
    ckks = hb.CKKSBootstrapContext(poly_degree=4096, L0=6, prime_bits=59, special_mod_size=3, max_rot_dim=1,
                                   bsgs_dim=[D], skip_bootstrap=True, seed=SEED, verbose=False)
    G, B = hb.compute_bsgs_params(D)
    rng = np.random.default_rng(3)
    Ws = [rng.standard_normal((D, D)) * 0.1 for _ in range(4)]
    xs = [rng.standard_normal(D) for _ in range(4)]
    sets = [hb.pre_encode_real_diags(ckks, W, D, G, B, level=1) for W in Ws]
    cts = [ckks.encrypt_replicated(x) for x in xs]
    for rep in range(3):                               # repeated: exercises stream reuse and the pool
        singles = [ph.bsgs_hoisted(ckks.ctx, c, s, ckks.gk) for c, s in zip(cts, sets)]
        batch = ph.bsgs_hoisted_batch(ckks.ctx, cts, sets, ckks.gk)
        for a, b, W, x in zip(singles, batch, Ws, xs):
            assert np.array_equal(a.to_numpy(), b.to_numpy())
            assert np.abs(ckks.decrypt_vec(b, D) - W @ x).max() < 1e-9


def test_fully_encrypted_ffn_block_flow():
    """The call sequence of the reference's fully_encrypted_ffn_block (test_fully_enc_bsgs.py:26-118) at small
    size: shared baby rotations, per-chunk BSGS for the FFN key, CT-CT square + relinearize + rescale, BSGS
    for the FFN value, level alignment with mod_switch_to_next, set_scale and residual add -- 3 levels per block,
    two blocks back to back, checked against the float64 block (corr > 0.999 is the reference's own bar, :298)."""
    from fhe_spear_b200 import bsgs as hb
    from fhe_spear_b200 import pyPhantom as ph
    D, F, L0 = 16, 32, 8
    ckks = hb.CKKSBootstrapContext(poly_degree=4096, L0=L0, prime_bits=59, special_mod_size=2, max_rot_dim=D,
                                   bsgs_dim=[D], skip_bootstrap=True, seed=SEED, verbose=False)
    G, B = hb.compute_bsgs_params(D)
    rng = np.random.default_rng(42)
    x = rng.standard_normal(D) * 0.1
    ct = ckks.encrypt_replicated(x)
    ref = x.copy()
    for blk in range(2):
        Wk, Wv = rng.standard_normal((D, F)) * 0.2, rng.standard_normal((F, D)) * 0.2
        start = ct.chain_index()
        n_chunks = F // D
        ct_baby = hb._compute_baby_rotations(ckks, ct, G)
        sq = []
        for c in range(n_chunks):
            M = Wk[:, c * D:(c + 1) * D].T.copy()
            fk = hb.fhe_matmul_bsgs(ckks, ct, M, D, G, B, ct_baby)          # reference-order path (plaintext list)
            s = ph.rescale_to_next(ckks.ctx, ph.relinearize(ckks.ctx, ph.multiply(ckks.ctx, fk, fk), ckks.rlk))
            sq.append(s)
        acc = None
        for c, s in enumerate(sq):
            M = Wv[c * D:(c + 1) * D, :].T.copy()
            part = hb.fhe_matmul_bsgs(ckks, s, M, D)                        # hoisted path at a lower level
            acc = part if acc is None else ph.add(ckks.ctx, acc, part)
        xa = ct
        while xa.chain_index() < acc.chain_index():
            xa = ph.mod_switch_to_next(ckks.ctx, xa)
        acc.set_scale(xa.scale())
        ct = ph.add(ckks.ctx, xa, acc)
        assert ct.chain_index() == start + 3
        ref = ref + ((ref @ Wk) ** 2) @ Wv
        got = ckks.decrypt_vec(ct, D)
        assert np.corrcoef(got, ref)[0, 1] > 0.999999
        assert np.abs(got - ref).max() < 1e-6      # set_scale fudges the scale by q_i/2^59 - 1 ~ 1e-13 relative


def test_client_aided_rwkv_block_matches_plaintext_block():
    """Two chained RWKV-7 blocks (reference client_aided_block, scripts/bootstrap_generation.py:756-899) with the
    eight projections per block on the GPU, pre-encoded and on-the-fly, against the float64 plaintext_block."""
    from fhe_spear_b200 import bsgs as hb
    from fhe_spear_b200.rwkv_block import RWKVBlockWeights, client_aided_block, plaintext_block
    D, F, H, S = 16, 64, 2, 8
    ckks = hb.CKKSBootstrapContext(poly_degree=2048, L0=3, prime_bits=59, special_mod_size=1, max_rot_dim=1,
                                   bsgs_dim=[D], skip_bootstrap=True, seed=SEED, verbose=False)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(D)
    xf, xp = x.copy(), x.copy()
    st_f = st_p = np.zeros((H, S, S))
    pa_f = pa_p = pf_f = pf_p = np.zeros(D)
    vf_f = vf_p = None
    for idx in range(2):
        blk = RWKVBlockWeights.random(D, F, H, S, block_idx=idx, seed=20 + idx)
        pe = hb.pre_encode_block(ckks, blk, D, F) if idx == 0 else None     # block 1 encodes its diagonals on the fly
        xf, pa_f, pf_f, st_f, vf_f, tm = client_aided_block(ckks, blk, xf, pa_f, pf_f, st_f, vf_f, use_bsgs=True,
                                                            preencoded_block=pe)
        xp, pa_p, pf_p, st_p, vf_p = plaintext_block(blk, xp, pa_p, pf_p, st_p, vf_p)
        assert set(tm) >= {"server_rkv", "server_wo", "server_ffn_key", "server_ffn_val"}
        assert np.abs(xf - xp).max() < 1e-7 and np.abs(st_f - st_p).max() < 1e-7
        assert np.corrcoef(xf, xp)[0, 1] > 0.999999


def test_hybrid_block_single_rank_and_encryption_ids():
    """sharding.HybridBlock (the multi-GPU server of a block, here with one rank): same block output as
    plaintext_block through client_aided_block, and -- ADVICE round 1 -- encryption ids come from the secret key's
    one monotonic counter: two HybridBlocks on one context never reuse a (seed, nonce) pair, so no two of their
    ciphertexts share the c1 polynomial (equal c1 would leak m_i - m_j)."""
    from fhe_spear_b200 import bsgs as hb
    from fhe_spear_b200.rwkv_block import RWKVBlockWeights, client_aided_block, plaintext_block
    from fhe_spear_b200.sharding import HybridBlock, PhasePlan
    D, F, H, S = 16, 64, 2, 8
    ckks = hb.CKKSBootstrapContext(poly_degree=2048, L0=3, prime_bits=59, special_mod_size=1, max_rot_dim=1,
                                   bsgs_dim=[D], skip_bootstrap=True, seed=SEED, verbose=False,
                                   baby_weights=HybridBlock.required_weights(1, D, F))
    blocks = [RWKVBlockWeights.random(D, F, H, S, block_idx=i, seed=30 + i) for i in range(2)]
    servers = [HybridBlock(ckks, b, D, F) for b in blocks]
    rng = np.random.default_rng(2)
    x = rng.standard_normal(D)
    xf, xp = x.copy(), x.copy()
    st_f = st_p = np.zeros((H, S, S))
    pa_f = pa_p = pf_f = pf_p = np.zeros(D)
    vf_f = vf_p = None
    for blk, srv in zip(blocks, servers):
        xf, pa_f, pf_f, st_f, vf_f, tm = client_aided_block(ckks, blk, xf, pa_f, pf_f, st_f, vf_f, preencoded_block=srv)
        xp, pa_p, pf_p, st_p, vf_p = plaintext_block(blk, xp, pa_p, pf_p, st_p, vf_p)
        assert np.abs(xf - xp).max() < 1e-7 and np.abs(st_f - st_p).max() < 1e-7
    # the same phase of two blocks (two layers of a model) and a plain encrypt call in between: all c1 differ
    ins = [rng.standard_normal(D) for _ in range(3)]
    c1s = []
    for srv in servers:
        base = ckks.sk.reserve_enc_ids(3)
        c1s += [ct.to_numpy()[1] for ct in srv._encrypt_inputs(ins, [0, 1, 2], base)]
        c1s.append(ckks.encrypt_replicated(ins[0]).to_numpy()[1])
    for i in range(len(c1s)):
        for j in range(i + 1, len(c1s)):
            assert not np.array_equal(c1s[i], c1s[j]), (i, j)
    # plans: every mat-vec of a phase has exactly one leader, every rank carries work
    for world in (2, 4, 8):
        for k in (1, 2, 3):
            plan = PhasePlan(k, world)
            assert sorted(j for j, _ in plan.assign) == list(range(k))
            assert all(plan.mine(r) for r in range(world))


def test_diagonal_sets_from_matrix_views():
    """diagonal_set.from_matrix reads row-pitched, transposed and short views in place (chunks of a larger weight
    matrix, reference fhe_projection_bsgs :575-591 / :630-642): same limbs as from the zero-padded compact copy, and
    the same as the host-side diagonal extraction (both encoders: one-kernel sub-ring and staged)."""
    from fhe_spear_b200 import bsgs as hb
    from fhe_spear_b200 import pyPhantom as ph
    D, F = 32, 80                                                    # ragged: the last chunk has 16 columns
    ckks = hb.CKKSBootstrapContext(poly_degree=2048, L0=4, prime_bits=59, special_mod_size=2, max_rot_dim=1,
                                   bsgs_dim=[D], skip_bootstrap=True, seed=SEED, verbose=False)
    G, B = hb.compute_bsgs_params(D)
    rng = np.random.default_rng(5)
    Wk, Wv = rng.standard_normal((D, F)), rng.standard_normal((F, D))
    views = [Wk[:, 0:32].T, Wk[:, 64:80].T, Wv[32:64, :].T, Wv[64:80, :].T, Wk[:, 32:64], Wv[10:42, :], Wk[:20, 5:30]]
    for V in views:
        compact = hb._padded(np.ascontiguousarray(V), D)
        a = ph.diagonal_set.from_matrix(ckks.ctx, V, G, B, ckks.diag_scale, D=D).to_numpy()
        b = ph.diagonal_set.from_matrix(ckks.ctx, compact, G, B, ckks.diag_scale).to_numpy()
        c = ph.diagonal_set(ckks.ctx, hb._pre_rotate(hb._extract_diagonals(compact, D), D, G), G, B, ckks.diag_scale).to_numpy()
        assert np.array_equal(a, b) and np.array_equal(a, c)
    # complex pair with different shapes (last chunk pair of a ragged F) and a full-ring (uncompressed) set
    a = ph.diagonal_set.from_matrix(ckks.ctx, views[0], G, B, ckks.diag_scale, M_imag=views[1], D=D).to_numpy()
    b = ph.diagonal_set.from_matrix(ckks.ctx, hb._padded(views[0], D), G, B, ckks.diag_scale, M_imag=hb._padded(views[1], D)).to_numpy()
    assert np.array_equal(a, b)
    full = ph.diagonal_set.from_matrix(ckks.ctx, views[2], G, B, ckks.diag_scale, compress=False, D=D)
    assert full.info()["ring_n"] == 2048
    x = rng.standard_normal(D)
    y = ph.bsgs_hoisted(ckks.ctx, ckks.encrypt_replicated(x), full, ckks.gk)
    assert np.abs(ckks.decrypt_vec(y, D) - hb._padded(views[2], D) @ x).max() < 1e-8
    # the whole D -> F and F -> D projections over views (ragged F) still match float64
    assert np.abs(hb.fhe_projection_bsgs(ckks, x, Wk, D, F) - x @ Wk).max() < 1e-8
    xf = rng.standard_normal(F)
    assert np.abs(hb.fhe_projection_bsgs(ckks, xf, Wv, F, D) - xf @ Wv).max() < 1e-8


def test_retrieval_wrapper_flow():
    """The call sequence of the reference's PhantomFHE retrieval wrapper (fhe_common.py:83-194; BASELINE config 1
    uses the same primitives): N=8192, primes [60, 40, 40, 60], P=1, asymmetric encryption, complex-packed
    Lorentz vectors, batched CT-PT and CT-CT dot products.  Scores must match the plaintext Lorentz inner
    product (tolerance 1e-5 at scale 2^40)."""
    from fhe_spear_b200 import pyPhantom as ph
    parms = ph.params(ph.scheme_type.ckks)
    parms.set_poly_modulus_degree(8192)
    parms.set_coeff_modulus(ph.create_coeff_modulus(8192, [60, 40, 40, 60]))
    parms.set_special_modulus_size(1)                       # after set_coeff_modulus, as the reference does
    ctx = ph.context(parms)
    sk = ph.secret_key(ctx, seed=SEED)
    pk, rlk, enc = sk.gen_publickey(ctx), sk.gen_relinkey(ctx), ph.ckks_encoder(ctx)
    scale, slots = 2.0 ** 40, enc.slot_count()
    rng = np.random.default_rng(12)

    def lorentz(v):
        return np.concatenate([np.sqrt(1 + (v ** 2).sum(-1, keepdims=True)), v], axis=-1)

    def pack(v, conj=False):
        v = np.concatenate([v, [0.0]]) if len(v) % 2 else v
        return v[0::2] + (-1j if conj else 1j) * v[1::2]

    docs = rng.standard_normal((300, 64))
    docs /= np.linalg.norm(docs, axis=1, keepdims=True)
    q = rng.standard_normal(64)
    q /= np.linalg.norm(q)
    ql, dl = lorentz(q), lorentz(docs)
    sign = np.concatenate([[-1.0], np.ones(64)])
    truth = dl @ (ql * sign)
    qp = pack(ql * sign, conj=True)                          # conj-packed query: Re(q_conj * d) sums both halves
    dp = [pack(d) for d in dl]
    spd = len(qp)
    per = slots // spd
    scores_pt, scores_ct = [], []
    for s0 in range(0, len(dp), per):
        batch = dp[s0:s0 + per]
        qs = np.zeros(slots, dtype=complex)
        ds = np.zeros(slots, dtype=complex)
        for i, d in enumerate(batch):
            qs[i * spd:(i + 1) * spd] = qp
            ds[i * spd:(i + 1) * spd] = d
        enc_q = pk.encrypt_asymmetric(ctx, enc.encode_complex_vector(ctx, list(qs), scale))
        d_pt = enc.encode_complex_vector(ctx, list(ds), scale)
        r1 = ph.rescale_to_next(ctx, ph.multiply_plain(ctx, enc_q, d_pt))
        v1 = enc.decode_complex_vector(ctx, sk.decrypt(ctx, r1))
        enc_d = pk.encrypt_asymmetric(ctx, d_pt)
        r2 = ph.rescale_to_next(ctx, ph.relinearize(ctx, ph.multiply(ctx, enc_q, enc_d), rlk))
        v2 = enc.decode_complex_vector(ctx, sk.decrypt(ctx, r2))
        for i in range(len(batch)):
            scores_pt.append(sum(c.real for c in v1[i * spd:(i + 1) * spd]))
            scores_ct.append(sum(c.real for c in v2[i * spd:(i + 1) * spd]))
    assert np.abs(np.array(scores_pt) - truth).max() < 1e-5
    assert np.abs(np.array(scores_ct) - truth).max() < 1e-5
    assert int(np.argmax(scores_pt)) == int(np.argmax(truth)) == int(np.argmax(scores_ct))


@pytest.mark.parametrize("L0,P,N", [(12, 1, 1024), (20, 2, 1024), (9, 4, 1024), (12, 1, 2048), (9, 4, 2048), (36, 3, 2048)])
def test_many_digit_parameter_sets_match_oracle(L0, P, N):
    """More than 8 key-switch digits (config C5 has beta = 12) takes the generic, non-unrolled kernel paths;
    P = 4 exercises a wider digit.  N = 2048 runs the key product fused into the forward transform's last pass
    (k_ntt_b_ks) with more digits than warps, N = 1024 the stand-alone kernels.  Rotation, exact BSGS and hoisted BSGS
    must stay bit-exact with the oracle."""
    S = Setup(N=N, bits=(59,) * (L0 + P), P=P)
    D = 16
    G, B = bsgs_params(D)
    steps = list(range(1, G)) + [g * G for g in range(1, B)]
    ph, ctx, sk = S.gpu(steps)
    gk = sk.create_galois_keys(ctx)
    enc = ph.ckks_encoder(ctx)
    keys = S.keys_for_steps(steps)
    rng = np.random.default_rng(L0)
    W, x = rng.standard_normal((D, D)) * 0.1, rng.standard_normal(D)
    rolled = rolled_diagonals(W, D, G, B)
    ct = sk.encrypt_symmetric(ctx, enc.encode_double_vector(ctx, tile(x, S.N // 2), S.scale), enc_id=5)
    cto = ct.to_numpy()
    r = ph.rotate(ctx, ct, 3, gk)
    assert np.array_equal(r.to_numpy(), S.o.apply_galois(cto, S.o.elt_from_step(3), S.key(S.o.elt_from_step(3))))
    ds = ph.diagonal_set(ctx, rolled, G, B, S.scale)
    diag_o = np.stack([S.o.encode(v.astype(complex), S.scale, S.L, ext=True, n=2 * D) for v in rolled])
    y = ph.bsgs_hoisted(ctx, ct, ds, gk)
    assert np.array_equal(y.to_numpy(), S.o.bsgs_hoisted(cto, diag_o, G, B, D, keys))
    assert np.abs(np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, y)))[:D] - W @ x).max() < 1e-9
    pts = enc.encode_double_vector_batch(ctx, np.stack([tile(v, S.N // 2) for v in rolled]), S.scale, chain_index=1)
    baby = [ct] + [ph.rotate(ctx, ct, b, gk) for b in range(1, G)]
    ye = ph.bsgs_multiply_accumulate(ctx, baby, pts, G, B, D, gk)
    assert np.array_equal(ye.to_numpy(), S.o.bsgs_exact(np.stack([c.to_numpy() for c in baby]),
                                                         np.stack([p.to_numpy()[0] for p in pts]), G, B, D, keys))


def test_reference_ffn_block_golden_limbs():
    """tests/golden/ffn_block.npz holds the output ciphertext of the reference's own fully_encrypted_ffn_block
    (test_fully_enc_bsgs.py:26-118) run over oracle primitives (tests/golden/make_golden_ffn.py).  The mirror in its
    reference-order mode on the CUDA library must reproduce it limb for limb; the default hoisted mode must decrypt to
    the same values."""
    import os
    from fhe_spear_b200 import bsgs as hb
    from fhe_spear_b200 import ffn_block as fb
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ffn_block.npz"))
    N, L0, P, D, F = (int(g[k]) for k in ("N", "L0", "P", "D", "F"))
    ckks = hb.CKKSBootstrapContext(poly_degree=N, L0=L0, prime_bits=59, special_mod_size=P, max_rot_dim=D, bsgs_dim=[D],
                                   skip_bootstrap=True, seed=bytes(g["seed"]), verbose=False)
    rep = hb._replicate_to_slots(np.asarray(g["x"], dtype=np.float64), ckks.slots)
    ct_x = ckks.sk.encrypt_symmetric(ckks.ctx, ckks.encoder.encode_double_vector(ckks.ctx, rep, ckks.scale),
                                     enc_id=int(g["enc_id_x"]))
    assert np.array_equal(ct_x.to_numpy(), g["ct_x"])
    out, used = fb.fully_encrypted_ffn_block(ckks, ct_x, g["W_key"], g["W_val"], D, F, reference_order=True)
    assert used == int(g["levels_used"]) == 3
    assert np.array_equal(out.to_numpy(), g["ct_out"])
    assert out.scale() == float(g["out_scale"])
    assert np.array_equal(ckks.decrypt_vec(out, D), g["dec"])
    fast, used2 = fb.fully_encrypted_ffn_block(ckks, ct_x, g["W_key"], g["W_val"], D, F)      # hoisted mat-vecs
    assert used2 == 3 and np.abs(ckks.decrypt_vec(fast, D) - g["plaintext"]).max() < 1e-6
