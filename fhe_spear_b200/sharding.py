"""Multi-GPU split of the BSGS mat-vec path: one process per GPU, torch.distributed.

Two levels, no parameter exchange after setup:

  * across mat-vecs -- the projections of one RWKV-7 block (r, k, v | o | 2 x ffn_key | 2 x ffn_val,
    reference scripts/bootstrap_generation.py:784-893) are independent: `projection_owner` deals them
    round-robin over the ranks; no collective on the data path (weak scaling, what bench.py --gpus N times).
  * within a mat-vec -- giant groups g = rank, rank + world, ... (`giant_groups`).  Every rank runs the
    hoisted baby steps, the diagonal MAC and the giant key switches of its groups on its own shard of the
    diagonals (`pyPhantom.diagonal_set(..., shard=(rank, world))`) and ends with an accumulator in basis
    Q_l*P.  The accumulators are summed with ONE integer all-reduce: residues are < 2^60, so the plain
    uint64 sum of <= 8 shards cannot wrap (`lazy_sum_is_safe`), and a single Barrett pass
    (`spear_obj_reduce`) turns it into the mod-q sum -- NCCL's own sum is not mod q, but it does not have
    to be.  One ModDown + rescale then finishes the ciphertext on every rank.
"""
import numpy as np


def giant_groups(B, rank, world):
    """Giant groups served by `rank`."""
    return list(range(rank, B, world))


def shard_rows(D, G, B, rank, world):
    """Indices of the (pre-rotated) diagonals stored by `rank`, in storage order."""
    return [k for g in giant_groups(B, rank, world) for k in range(g * G, min((g + 1) * G, D))]


def shard_steps(D, G, B, rank, world):
    """Rotation steps whose Galois keys `rank` needs: all baby steps, its own giant steps."""
    return list(range(1, min(G, D))) + [g * G for g in giant_groups(B, rank, world) if g > 0 and g * G < D]


def projection_owner(index, world):
    """Rank serving projection `index` of a block (level-1 sharding)."""
    return index % world


def lazy_sum_is_safe(moduli, world):
    """True when the plain 64-bit sum of `world` residues cannot wrap (then one reduction suffices)."""
    return world * (max(int(q) for q in moduli) - 1) < (1 << 63)


def allreduce_residues(tensor, group=None):
    """In-place SUM all-reduce of a torch int64 tensor holding residues (NCCL on GPU, gloo on CPU)."""
    import torch.distributed as dist
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


class _DevView:
    """torch-importable view (CUDA array interface) of an engine object's limbs, as int64."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i8", "data": (ptr, False), "version": 3,
                                         "strides": None}


def sharded_matvec(ckks, ct, shard_set, group=None):
    """One giant-step-sharded mat-vec: this rank's accumulator, integer all-reduce, Barrett pass, ModDown + rescale.
    Every rank ends with the same ciphertext (bit-identical to the unsharded result)."""
    import torch
    import torch.distributed as dist
    from . import pyPhantom as ph
    ctx = ckks.ctx
    acc = ph.bsgs_hoisted_partial(ctx, ct, shard_set, ckks.gk)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        size, limbs, ext, ring, _, _ = acc._info()
        count = size * (limbs + ctx.P) * ring
        ctx.synchronize()                                   # engine stream -> torch stream hand-off
        t = torch.as_tensor(_DevView(ph.device_ptr(acc), count), device=f"cuda:{ctx.device}")
        allreduce_residues(t, group)
        torch.cuda.synchronize(ctx.device)
        ph.reduce_inplace(ctx, acc)
    return ph.bsgs_finish(ctx, acc)


def sharded_matvec_batch(ckks, cts, shard_sets, group=None):
    """The independent mat-vecs of one block phase (r, k, v | the ffn chunk pairs), giant-step sharded: all shard
    accumulators are computed concurrently on the engine's streams, then one host hand-off, the all-reduces issued
    back to back, one Barrett pass and the ModDown + rescale of each.  Same results as sharded_matvec per item."""
    import torch
    import torch.distributed as dist
    from . import pyPhantom as ph
    ctx = ckks.ctx
    accs = ph.bsgs_hoisted_partial_batch(ctx, list(cts), list(shard_sets), ckks.gk)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        ctx.synchronize()
        works = []
        for acc in accs:
            size, limbs, ext, ring, _, _ = acc._info()
            t = torch.as_tensor(_DevView(ph.device_ptr(acc), size * (limbs + ctx.P) * ring), device=f"cuda:{ctx.device}")
            works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True))
        for w in works:
            w.wait()
        torch.cuda.synchronize(ctx.device)
        for acc in accs:
            ph.reduce_inplace(ctx, acc)
    return [ph.bsgs_finish(ctx, acc) for acc in accs]


class ShardedMatvec:
    """Giant-step-sharded hoisted BSGS mat-vec  Enc(x) -> Enc(W @ x)  over the ranks of `group`.

    `ckks` is a fhe_spear_b200.bsgs.CKKSBootstrapContext built identically (same seed) on every rank."""

    def __init__(self, ckks, W, D, level=1, group=None, compress=True, baby_weight=1.0):
        import torch.distributed as dist
        from . import bsgs as hb
        from . import pyPhantom as ph
        self.ph, self.ckks, self.group = ph, ckks, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if not lazy_sum_is_safe(ckks.ctx.moduli, self.world):
            raise RuntimeError("too many shards for a lazy 64-bit sum of residues")
        G, B = hb.compute_bsgs_params(D, baby_weight)   # the context must hold the keys of this split
        rolled = hb._pre_rotate(hb._extract_diagonals(np.asarray(W, dtype=np.float64), D), D, G)
        self.D, self.G, self.B = D, G, B
        self.shard = ph.diagonal_set(ckks.ctx, rolled, G, B, ckks.diag_scale, chain_index=level, compress=compress,
                                     shard=(self.rank, self.world))

    def __call__(self, ct):
        return sharded_matvec(self.ckks, ct, self.shard, self.group)
