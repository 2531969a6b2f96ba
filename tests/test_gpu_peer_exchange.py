"""The fused NVLink peer-memory exchange of shard accumulators (csrc/peer.cu).

  * one GPU: `ph.peer_selftest` runs the reduce kernels of all ranks one after the other over local windows (no epoch
    waits -- kernels that wait on one another must not share a GPU as separate launches, B200_PROFILING.md) and the
    result is compared with plain integer arithmetic for every group size 2..8, and with the unsharded mat-vec;
  * two or more GPUs (`gpurun --gpus 2`): the real thing, driven the way fhe_spear_b200.sharding drives it -- one
    process per GPU, windows mapped through CUDA IPC, handles swapped over torch.distributed (gloo: host plumbing
    only), epoch flags across NVLink."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import Setup, bsgs_params, rolled_diagonals, tile

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, D, out):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SPEAR_DEVICE=str(rank))   # one GPU per rank
        os.environ.pop("LOCAL_RANK", None)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from fhe_spear_b200 import sharding as sh

        class K:          # the two attributes sharded_matvec needs of a CKKSBootstrapContext
            pass
        S = Setup(N=2048, bits=(59,) * 6, P=2)
        G, B = bsgs_params(D)
        steps = list(range(1, G)) + [g * G for g in range(1, B)]
        ph, ctx, sk = S.gpu(steps)
        ckks = K()
        ckks.ctx, ckks.gk = ctx, sk.create_galois_keys(ctx)
        enc = ph.ckks_encoder(ctx)
        rng = np.random.default_rng(D)
        W, x = rng.standard_normal((D, D)) * 0.1, rng.standard_normal(D)
        rolled = rolled_diagonals(W, D, G, B)
        ct = sk.encrypt_symmetric(ctx, enc.encode_double_vector(ctx, tile(x, S.N // 2), S.scale), enc_id=3)
        ref = ph.bsgs_hoisted(ctx, ct, ph.diagonal_set(ctx, rolled, G, B, S.scale), ckks.gk).to_numpy()
        shard = ph.diagonal_set(ctx, rolled, G, B, S.scale, shard=(rank, world))
        ex = sh.PeerExchange.get(ctx)
        assert ex is not None, "peer windows could not be mapped"
        same = True
        for it in range(5):                               # epochs advance, the window is reused
            y = sh.sharded_matvec(ckks, ct, shard)
            same &= bool(np.array_equal(y.to_numpy(), ref))
        ys = sh.sharded_matvec_batch(ckks, [ct, ct, ct, ct], [shard] * 4)   # slots 0, 1, 2, 0
        same &= all(np.array_equal(y.to_numpy(), ref) for y in ys)
        # the exchange itself against plain integer arithmetic: sum of both ranks' accumulators mod q
        acc = ph.bsgs_hoisted_partial(ctx, ct, shard, ckks.gk)
        mine = acc.to_numpy()
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        ex.allreduce(acc, 1)
        q = np.array([int(v) for v in S.q], dtype=object)    # top level: all L data limbs, then the P special ones
        exp = sum(p.astype(object) for p in parts) % q[None, :, None]
        same &= bool(np.array_equal(acc.to_numpy().astype(object), exp))
        dec = np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, ys[-1])))[:D]
        out.put((rank, same, float(np.abs(dec - W @ x).max()), ex.window.status()))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:   # noqa: BLE001 -- reported to the parent, which fails the test
        out.put((rank, False, repr(e), -1))
        raise


def _split_worker(rank, world, port, D, weight, out):
    """two-phase mat-vecs through the real windows: slots reused over several epochs, batches on the auxiliary streams"""
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SPEAR_DEVICE=str(rank))
        os.environ.pop("LOCAL_RANK", None)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from fhe_spear_b200 import sharding as sh

        class K:
            pass
        S = Setup(N=2048, bits=(59,) * 6, P=2)
        G = int(np.ceil(np.sqrt(weight * D)))
        B = -(-D // G)
        steps = list(range(1, G)) + [g * G for g in range(1, B)]
        ph, ctx, sk = S.gpu(steps)
        ckks = K()
        ckks.ctx, ckks.gk = ctx, sk.create_galois_keys(ctx)
        enc = ph.ckks_encoder(ctx)
        rng = np.random.default_rng(D)
        Ws = [rng.standard_normal((D, D)) * 0.1 for _ in range(2)]
        x = rng.standard_normal(D)
        ct = sk.encrypt_symmetric(ctx, enc.encode_double_vector(ctx, tile(x, S.N // 2), S.scale), enc_id=3)
        fulls = [ph.diagonal_set(ctx, rolled_diagonals(W, D, G, B), G, B, S.scale) for W in Ws]
        refs = [ph.bsgs_hoisted(ctx, ct, f, ckks.gk).to_numpy() for f in fulls]
        rows = [f.slice_rows(rank, world) for f in fulls]
        same = True
        for it in range(4):                               # epochs advance, slots are reused
            ys = sh.split_matvec_batch(ckks, [ct, ct], rows)
            same &= all(np.array_equal(y.to_numpy(), r) for y, r in zip(ys, refs))
        ys = sh.split_matvec_batch(ckks, [ct] * 5, [rows[0], rows[1], rows[0], rows[1], rows[0]])   # two rounds of slots
        same &= all(np.array_equal(y.to_numpy(), refs[i % 2]) for i, y in enumerate(ys))
        dec = np.array(enc.decode_double_vector(ctx, sk.decrypt(ctx, ys[1])))[:D]
        status = max(sh.PeerExchange.get(ctx).window.status(), sh.PeerExchange.get(ctx, tag="split").window.status())
        out.put((rank, same, float(np.abs(dec - Ws[1] @ x).max()), status))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:   # noqa: BLE001 -- reported to the parent, which fails the test
        out.put((rank, False, repr(e), -1))
        raise


def _block_worker(rank, world, port, out):
    """a whole client-aided RWKV-7 block served by sharding.HybridBlock in two-phase mode over real windows"""
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SPEAR_DEVICE=str(rank))
        os.environ.pop("LOCAL_RANK", None)
        os.environ.pop("SPEAR_TWO_PHASE", None)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from helpers import SEED
        from fhe_spear_b200 import bsgs as hb
        from fhe_spear_b200.rwkv_block import RWKVBlockWeights, client_aided_block, plaintext_block
        from fhe_spear_b200.sharding import HybridBlock
        D, F, H, S = 32, 128, 2, 16
        ckks = hb.CKKSBootstrapContext(poly_degree=2048, L0=3, prime_bits=59, special_mod_size=1, max_rot_dim=1,
                                       bsgs_dim=[D], skip_bootstrap=True, seed=SEED, verbose=False, device=rank,
                                       baby_weights=HybridBlock.required_weights(world, D, F, two_phase=True))
        blocks = [RWKVBlockWeights.random(D, F, H, S, block_idx=i, seed=30 + i) for i in range(2)]
        servers = [HybridBlock(ckks, b, D, F, rank, world) for b in blocks]
        rng = np.random.default_rng(2)
        x = rng.standard_normal(D)
        xf, xp = x.copy(), x.copy()
        st_f = st_p = np.zeros((H, S, S))
        pa_f = pa_p = pf_f = pf_p = np.zeros(D)
        vf_f = vf_p = None
        err = 0.0
        for blk, srv in zip(blocks, servers):
            xf, pa_f, pf_f, st_f, vf_f, tm = client_aided_block(ckks, blk, xf, pa_f, pf_f, st_f, vf_f, preencoded_block=srv)
            xp, pa_p, pf_p, st_p, vf_p = plaintext_block(blk, xp, pa_p, pf_p, st_p, vf_p)
            err = max(err, float(np.abs(xf - xp).max()), float(np.abs(st_f - st_p).max()))
        outs = [None] * world
        dist.all_gather_object(outs, xf)                       # every rank ends with the same activations
        same = all(np.array_equal(o, outs[0]) for o in outs)
        out.put((rank, same and all(s.two_phase for s in servers), err, 0))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:   # noqa: BLE001 -- reported to the parent, which fails the test
        out.put((rank, False, repr(e), -1))
        raise


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 3, 4, 5, 6, 7, 8])
def test_peer_reduce_kernels_emulated_on_one_gpu(world):
    """every rank's slice of the fused reduce-scatter + Barrett + all-gather, for every group size"""
    S = Setup(N=2048, bits=(59,) * 5, P=2)
    ph, ctx, sk = S.gpu([1])
    rng = np.random.default_rng(world)
    q = np.array([int(v) for v in S.q], dtype=object)
    raws = [np.stack([np.stack([rng.integers(0, int(S.q[i]), S.N, dtype=np.uint64) for i in range(S.L + S.P)])
                      for _ in range(2)]) for _ in range(world)]
    accs = [ph.ciphertext.from_numpy(ctx, r, 1.0, ext=True) for r in raws]
    ph.peer_selftest(ctx, accs)
    exp = (sum(r.astype(object) for r in raws) % q[None, :, None]).astype(np.uint64)
    for a in accs:
        assert np.array_equal(a.to_numpy(), exp)


@pytest.mark.parametrize("D,world", [(64, 2), (20, 3), (64, 8)])
def test_emulated_exchange_of_shard_accumulators_matches_unsharded(D, world):
    """shard accumulators of all ranks (computed one after the other on this GPU) -> emulated exchange -> finish
    == the unsharded hoisted mat-vec, bit for bit"""
    S = Setup(N=2048, bits=(59,) * 6, P=2)
    G, B = bsgs_params(D)
    steps = list(range(1, G)) + [g * G for g in range(1, B)]
    ph, ctx, sk = S.gpu(steps)
    gk = sk.create_galois_keys(ctx)
    enc = ph.ckks_encoder(ctx)
    rng = np.random.default_rng(D)
    W, x = rng.standard_normal((D, D)) * 0.1, rng.standard_normal(D)
    rolled = rolled_diagonals(W, D, G, B)
    ct = sk.encrypt_symmetric(ctx, enc.encode_double_vector(ctx, tile(x, S.N // 2), S.scale), enc_id=3)
    ref = ph.bsgs_hoisted(ctx, ct, ph.diagonal_set(ctx, rolled, G, B, S.scale), gk).to_numpy()
    accs = [ph.bsgs_hoisted_partial(ctx, ct, ph.diagonal_set(ctx, rolled, G, B, S.scale, shard=(r, world)), gk)
            for r in range(world)]
    ph.peer_selftest(ctx, accs)
    for a in accs:
        assert np.array_equal(ph.bsgs_finish(ctx, a).to_numpy(), ref)


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs (kernels that wait on one another must not share one)")
@pytest.mark.parametrize("D", [64, 20])
def test_two_process_peer_exchange_matches_unsharded(D):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mpc = mp.get_context("spawn")
    out = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, D, out)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = [out.get(timeout=240) for _ in procs]
        for p in procs:
            p.join(timeout=60)
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    for rank, same, err, status in res:
        assert status == 0, f"rank {rank}: window status {status} ({err})"
        assert same, f"rank {rank}: peer-exchanged result differs from the unsharded one"
        assert err < 1e-9
    assert all(p.exitcode == 0 for p in procs)


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs (kernels that wait on one another must not share one)")
@pytest.mark.parametrize("D,weight", [(64, 1.0), (512, 32.0)])
def test_two_process_two_phase_matvec_matches_unsharded(D, weight):
    """the two-phase mat-vec over real peer windows (rows | giant groups, scatter fused into the MAC's stores)"""
    world = min(_gpu_count(), 4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mpc = mp.get_context("spawn")
    out = mpc.Queue()
    procs = [mpc.Process(target=_split_worker, args=(r, world, port, D, weight, out)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [out.get(timeout=240) for _ in procs]
        for p in procs:
            p.join(timeout=60)
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    for rank, same, err, status in res:
        assert status == 0, f"rank {rank}: window status {status} ({err})"
        assert same, f"rank {rank}: two-phase result differs from the unsharded one ({err})"
        assert err < 1e-9
    assert all(p.exitcode == 0 for p in procs)


@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs (kernels that wait on one another must not share one)")
def test_two_process_hybrid_block_two_phase_matches_plaintext_block():
    """sharding.HybridBlock over 2 GPUs (two-phase mat-vecs, real windows) drives client_aided_block
    (reference scripts/bootstrap_generation.py:756-899) to the float64 block within 1e-7, identically on every rank"""
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mpc = mp.get_context("spawn")
    out = mpc.Queue()
    procs = [mpc.Process(target=_block_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [out.get(timeout=240) for _ in procs]
        for p in procs:
            p.join(timeout=60)
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    for rank, same, err, status in res:
        assert status == 0, f"rank {rank}: {err}"
        assert same, f"rank {rank}: ranks disagree or the two-phase path was not taken"
        assert err < 1e-7
    assert all(p.exitcode == 0 for p in procs)


def test_single_rank_window_is_a_plain_reduction():
    S = Setup(N=2048, bits=(59,) * 5, P=2)
    ph, ctx, sk = S.gpu([1])
    w = ph.peer_window(ctx, 0, 1, 2 * (S.L + S.P) * S.N * 8)
    rng = np.random.default_rng(1)
    raw = rng.integers(0, 1 << 62, size=(2, S.L + S.P, S.N), dtype=np.uint64)
    obj = ph.ciphertext.from_numpy(ctx, raw, 1.0, ext=True)
    w.allreduce(obj)
    q = np.array(list(S.q[:S.L]) + list(S.q[S.L:]), dtype=np.uint64)
    assert np.array_equal(obj.to_numpy(), raw % q[None, :, None])
    assert w.status() == 0
    with pytest.raises(RuntimeError):
        ph.peer_window(ctx, 3, 2, 1024)
