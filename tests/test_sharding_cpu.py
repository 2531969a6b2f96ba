"""World-size-2 gloo tests of the giant-step sharding logic (host side of SURVEY.md section 8e).

The arithmetic of each rank is done by the CPU oracle (this is a CPU test); what is under test is the
shard plan of fhe_spear_b200.sharding, the lazy integer all-reduce of the Q_l*P accumulators over
torch.distributed, and that reduce + one ModDown/rescale reproduces the unsharded ciphertext bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import SEED, Setup, bsgs_params, rolled_diagonals, tile


def test_shard_plan_partitions_the_matrix():
    from fhe_spear_b200 import sharding as sh
    for D in (16, 20, 2048):
        G, B = bsgs_params(D)
        for world in (1, 2, 3, 8):
            rows = [sh.shard_rows(D, G, B, r, world) for r in range(world)]
            assert sorted(k for part in rows for k in part) == list(range(D))
            groups = [sh.giant_groups(B, r, world) for r in range(world)]
            assert sorted(g for part in groups for g in part) == list(range(B))
            for r in range(world):
                steps = sh.shard_steps(D, G, B, r, world)
                assert set(range(1, G)) <= set(steps)
                assert {g * G for g in groups[r] if g} == set(steps) - set(range(1, G))
    assert [sh.projection_owner(i, 3) for i in range(8)] == [0, 1, 2, 0, 1, 2, 0, 1]
    q59 = [(1 << 59) - 1] * 4
    assert sh.lazy_sum_is_safe(q59, 8) and not sh.lazy_sum_is_safe([(1 << 60) - 1], 16)


def test_two_phase_plan_partitions_rows_and_groups():
    """host side of the two-phase mat-vec: the row ranges of the ranks partition the l + P rows of the RNS basis (some may be
    empty when there are more ranks than rows), the giant groups are dealt round-robin, a window slot holds a rank's share"""
    from fhe_spear_b200 import pyPhantom as ph
    from fhe_spear_b200 import sharding as sh
    for limbs, P in ((24, 3), (36, 3), (4, 2), (1, 1)):
        for world in (1, 2, 3, 4, 5, 6, 8):
            N = 2048
            cover = np.zeros((limbs + P, N // 128), dtype=np.int64)
            area = []
            for r in range(world):
                r0, r1, c0, c1 = ph.diagonal_set.share(limbs, P, N, r, world)
                assert 0 <= r0 <= r1 <= limbs + P and 0 <= c0 < c1 <= N and c0 % 128 == 0 and c1 % 128 == 0
                cover[r0:r1, c0 // 128:c1 // 128] += 1
                area.append((r1 - r0) * (c1 - c0))
            assert (cover == 1).all()                       # the shares partition rows x columns
            assert max(area) - min(area) <= N               # at most one row (or one half row) apart
            if world == 8 and limbs + P == 27:
                assert max(area) == 7 * N // 2              # 3.5 row-equivalents instead of 4

    class Ctx:
        L, P, N = 24, 3, 32768
    assert sh.split_slot_bytes(Ctx, 45, 8) == 6 * 2 * 27 * 32768 * 8      # ceil(45 / 8) groups of [2][L+P][N] words
    assert sh.split_slot_bytes(Ctx, 16, 8) == 2 * 2 * 27 * 32768 * 8
    assert sh.HybridBlock.required_weights(8, 2048, 8192, two_phase=True) == (8.0,)
    assert not sh.HybridBlock.two_phase_default(1)


def test_two_phase_entry_points_refuse_a_single_rank_loudly():
    """no silent fallback: the two-phase batch needs a rank group (and peer windows); alone it raises before touching a device"""
    from fhe_spear_b200 import sharding as sh

    class K:
        ctx = gk = None
    with pytest.raises(RuntimeError, match="rank group of two or more"):
        sh.split_matvec_batch(K, [object()], [object()])
    assert sh.two_phase_ready(None, 45, 1) is False


def _worker(rank, world, port, D, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fhe_spear_b200 import sharding as sh
    S = Setup(N=512, bits=(59,) * 5, P=2)
    o = S.o
    G, B = bsgs_params(D)
    rng = np.random.default_rng(5)
    W, x = rng.standard_normal((D, D)) * 0.1, rng.standard_normal(D)
    rolled = rolled_diagonals(W, D, G, B)
    keys = S.keys_for_steps(sh.shard_steps(D, G, B, rank, world))
    ct = o.encrypt_symmetric(SEED, 1, S.sk, o.encode(tile(x, S.N // 2), S.scale, S.L))
    rows = sh.shard_rows(D, G, B, rank, world)
    pow2 = D & (D - 1) == 0

    def enc(v):    # sub-ring form when D is a power of two, else the full ring
        if pow2:
            return o.encode(v.astype(complex), S.scale, S.L, ext=True, n=2 * D)
        return o.encode(tile(v, S.N // 2), S.scale, S.L, ext=True)
    diag = np.stack([enc(rolled[k]) for k in rows])
    acc = o.bsgs_hoisted_partial(ct, diag, G, B, D, keys, g_first=rank, g_stride=world)
    assert sh.lazy_sum_is_safe(S.q, world)
    t = torch.from_numpy(acc.view(np.int64))
    sh.allreduce_residues(t)                       # plain integer sum over ranks
    y = o.bsgs_finish(o.reduce_rows(acc, ext=True))
    if rank == 0:
        all_keys = S.keys_for_steps(list(range(1, G)) + [g * G for g in range(1, B)])
        full = np.stack([enc(r) for r in rolled])
        ref = o.bsgs_hoisted(ct, full, G, B, D, all_keys)
        dec = o.decode(o.decrypt(S.sk, y), S.scale ** 2 / float(S.q[S.L - 1]))[:D].real
        out.put((bool(np.array_equal(y, ref)), float(np.abs(dec - W @ x).max())))
    dist.barrier()
    dist.destroy_process_group()




def test_phase_plans_cover_every_matvec_once_and_balance_the_ranks():
    from fhe_spear_b200 import sharding as sh
    for world in (1, 2, 3, 4, 8):
        for k in (1, 2, 3, 5):
            plan = sh.PhasePlan(k, world)
            assert sorted(j for j, _ in plan.assign) == list(range(k))
            load = np.zeros(world)
            for j, ranks in plan.assign:
                assert len(set(ranks)) == len(ranks) and all(0 <= r < world for r in ranks)
                load[list(ranks)] += 1.0 / len(ranks)
                assert plan.leader(j) == ranks[0]
            if world <= k or k == 1:
                assert np.ptp(load) < 1e-12 or k % world                # equal shares when the phase fills the ranks
            for rank in range(world):
                mine = plan.mine(rank)
                assert all(rank in g for g, _ in mine)
                assert sum(len(js) for _, js in mine) == sum(1 for _, g in plan.assign if rank in g)
    p = sh.PhasePlan(3, 2)      # r -> rank 0, k -> rank 1, v sharded over both: 1.5 mat-vecs each
    assert p.assign == [(0, (0,)), (1, (1,)), (2, (0, 1))]
    assert sh.HybridBlock.required_weights(8, 2048, 8192, two_phase=False) == (1.0, 2.0, 8.0 / 3.0, 4.0)


@pytest.mark.parametrize("D", [16, 20])
def test_two_rank_giant_sharding_matches_unsharded(D):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, D, out)) for r in range(2)]
    for p in procs:
        p.start()
    same, err = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same, "sharded ciphertext differs from the unsharded one"
    assert err < 1e-9
