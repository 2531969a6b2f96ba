"""The fused NVLink peer-memory exchange of shard accumulators (csrc/peer.cu), driven the way
fhe_spear_b200.sharding drives it: one process per rank, windows mapped through CUDA IPC, handles swapped over
torch.distributed (gloo here: host plumbing only).  On a one-GPU box both ranks share cuda:0 -- IPC mapping, the
epoch flags and the reduce kernel are the same code that runs across NVLink; `tools/sharded_latency.py` is the
multi-GPU run of it."""
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
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SPEAR_DEVICE="0")
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
