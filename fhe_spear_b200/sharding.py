"""Multi-GPU split of the BSGS mat-vec path: one process per GPU, torch.distributed.

Two levels, no parameter exchange after setup:

  * across mat-vecs -- the projections of one RWKV-7 block (r, k, v | o | 2 x ffn_key | 2 x ffn_val,
    reference scripts/bootstrap_generation.py:784-893) are independent: `projection_owner` deals them
    round-robin over the ranks; no collective on the data path (weak scaling, what bench.py --gpus N times).
    `PhasePlan` / `HybridBlock` combine the two levels for token latency: the k independent mat-vecs of a phase
    are dealt to k rank groups, each group shards its mat-vec by giant step; the decrypted vectors (a few KB) are
    exchanged with one small all-reduce per phase.
  * within a mat-vec -- giant groups g = rank, rank + world, ... (`giant_groups`).  Every rank runs the
    hoisted baby steps, the diagonal MAC and the giant key switches of its groups on its own shard of the
    diagonals (`pyPhantom.diagonal_set(..., shard=(rank, world))`) and ends with an accumulator in basis
    Q_l*P.  The accumulators are summed mod q by the engine itself over NVLink peer memory (`PeerExchange`,
    csrc/peer.cu): every rank maps the others' exchange windows through CUDA IPC and one kernel per rank does
    reduce-scatter + Barrett + all-gather on the engine's stream, without a host hand-off.  Residues are < 2^60,
    so the plain uint64 sum of <= 8 shards cannot wrap (`lazy_sum_is_safe`) and one Barrett step per element
    gives the mod-q sum.  Where windows cannot be mapped (gloo CPU tests, GPUs hidden from each other) the same
    sum is ONE integer all-reduce followed by a Barrett pass (`spear_obj_reduce`) -- NCCL's own sum is not mod q,
    but it does not have to be.  One ModDown + rescale then finishes the ciphertext on every rank.
"""
import os
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


class PeerExchange:
    """The exchange windows of one rank group (pyPhantom.peer_window): created once per (context, group), handles
    swapped with one all_gather_object over the group -- host plumbing only; the data path is csrc/peer.cu."""

    _cache = {}
    _retired = []      # replaced windows stay mapped: peers may hold their IPC mappings for the life of the process

    def __init__(self, ctx, group=None, slots=3, slot_bytes=None):
        """Every rank runs the SAME collective sequence whether or not a local step fails (handle all-gather, outcome
        all-gather after mapping), so a failure on one rank degrades the whole group to the NCCL path instead of
        leaving the others in a mismatched collective.  `self.ok` is the agreed outcome."""
        import torch.distributed as dist
        from . import pyPhantom as ph
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.slots, self.window, err = slots, None, None
        if slot_bytes is None:
            slot_bytes = 2 * (ctx.L + ctx.P) * ctx.N * 8        # an accumulator at the top level
        self.slot_bytes = int(slot_bytes)
        try:
            self.window = ph.peer_window(ctx, self.rank, self.world, slot_bytes, slots)
        except RuntimeError as e:
            err = e
        handles = [None] * self.world
        dist.all_gather_object(handles, self.window.handle if self.window is not None else None, group=group)
        if err is None and all(h is not None for h in handles):
            try:
                self.window.connect(handles)
            except RuntimeError as e:
                err = e
        elif err is None:
            err = RuntimeError("a peer could not create its window")
        flags = [None] * self.world
        dist.all_gather_object(flags, err is None, group=group)   # doubles as the barrier: every window is mapped before anyone posts
        self.ok = all(flags)
        if not self.ok:
            print(f"[spear] peer windows unavailable on rank {dist.get_rank()}: {err or 'a peer failed'}; using the NCCL all-reduce")
            self.window = None

    @classmethod
    def get(cls, ctx, group=None, tag="acc", slot_bytes=None):
        """The exchange `tag` of (ctx, group) -- "acc": accumulator all-reduces, "split": the two-phase mat-vec's scatter
        windows (slot_bytes = what a slot must hold; a cached window that is too small is replaced, on every rank alike)
        -- or None when peer windows are switched off (SPEAR_PEER=0), not on CUDA ranks, or cannot be mapped (then the
        callers use the integer all-reduce)."""
        import torch.distributed as dist
        if os.environ.get("SPEAR_PEER", "1") == "0" or not dist.is_initialized() or dist.get_world_size(group) < 2:
            return None
        import weakref
        key = (id(ctx), id(group) if group is not None else 0, tag)
        hit = cls._cache.get(key)
        stale = hit is not None and hit[1] is not None and slot_bytes is not None and hit[1].slot_bytes < slot_bytes
        if hit is None or hit[0]() is not ctx or stale:   # ids are recycled: the entry must belong to this very context
            if stale:
                ctx.synchronize()
                dist.barrier(group)                        # nobody is still writing into the window that is replaced
                cls._retired.append(hit[1])
            ex = cls(ctx, group, slot_bytes=slot_bytes)
            hit = cls._cache[key] = (weakref.ref(ctx), ex if ex.ok else None)
        return hit[1]

    def allreduce(self, acc, slot=0):
        self.window.allreduce(acc, slot % self.slots)

    def check(self, ctx):
        """Raise if a peer never arrived.  The exchange is asynchronous, so the status word is only meaningful after a
        synchronisation: called by sharded_matvec* before a result leaves (the decrypt that follows synchronises anyway)."""
        ctx.synchronize()
        st = self.window.status()
        if st != 0:
            raise RuntimeError(f"peer exchange failed: rank {st - 1} of the group never arrived (CUDA peer window timed out); "
                               "the accumulator was poisoned")


def sharded_matvec(ckks, ct, shard_set, group=None):
    """One giant-step-sharded mat-vec: this rank's accumulator, the mod-q sum over the group (fused peer-memory
    exchange; integer all-reduce + Barrett pass as the fallback), ModDown + rescale.
    Every rank ends with the same ciphertext (bit-identical to the unsharded result)."""
    import torch
    import torch.distributed as dist
    from . import pyPhantom as ph
    ctx = ckks.ctx
    acc = ph.bsgs_hoisted_partial(ctx, ct, shard_set, ckks.gk)
    ex = PeerExchange.get(ctx, group)
    if ex is not None:
        ex.allreduce(acc)
    elif dist.is_initialized() and dist.get_world_size(group) > 1:
        size, limbs, ext, ring, _, _ = acc._info()
        count = size * (limbs + ctx.P) * ring
        ctx.synchronize()                                   # engine stream -> torch stream hand-off
        t = torch.as_tensor(_DevView(ph.device_ptr(acc), count), device=f"cuda:{ctx.device}")
        allreduce_residues(t, group)
        torch.cuda.synchronize(ctx.device)
        ph.reduce_inplace(ctx, acc)
    out = ph.bsgs_finish(ctx, acc)
    if ex is not None:
        ex.check(ctx)
    return out


def sharded_matvec_batch(ckks, cts, shard_sets, group=None):
    """The independent mat-vecs of one block phase (r, k, v | the ffn chunk pairs), giant-step sharded: all shard
    accumulators are computed concurrently on the engine's streams, then one host hand-off, the all-reduces issued
    back to back, one Barrett pass and the ModDown + rescale of each.  Same results as sharded_matvec per item."""
    import torch
    import torch.distributed as dist
    from . import pyPhantom as ph
    ctx = ckks.ctx
    accs = ph.bsgs_hoisted_partial_batch(ctx, list(cts), list(shard_sets), ckks.gk)
    ex = PeerExchange.get(ctx, group)
    if ex is not None:
        for i, acc in enumerate(accs):
            ex.allreduce(acc, i)
    elif dist.is_initialized() and dist.get_world_size(group) > 1:
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
    outs = [ph.bsgs_finish(ctx, acc) for acc in accs]
    if ex is not None:
        ex.check(ctx)
    return outs


def two_phase_ready(ctx, groups, world, group=None):
    """Bring up the two windows of the two-phase mat-vec (accumulator all-reduce + scatter) for mat-vecs of up to `groups`
    giant groups; False -- on every rank alike -- when they cannot be mapped (no CUDA IPC / peer access between the GPUs,
    SPEAR_PEER=0): the callers then use the giant-step sharding of round 1, whose exchange has an NCCL stand-in."""
    if world < 2:
        return False
    acc = PeerExchange.get(ctx, group)
    sc = PeerExchange.get(ctx, group, tag="split", slot_bytes=split_slot_bytes(ctx, groups, world)) if acc is not None else None
    return acc is not None and sc is not None


def split_slot_bytes(ctx, groups, world):
    """bytes a scatter-window slot must hold: the accumulators of ceil(groups / world) giant groups at the top level"""
    return -(-int(groups) // world) * 2 * (ctx.L + ctx.P) * ctx.N * 8


def split_matvec_batch(ckks, cts, row_sets, group=None):
    """Independent mat-vecs served by ALL ranks of `group` in two phases each (include/spear_b200.h, "two-phase mat-vec"):
    the hoisted baby steps and the diagonal MAC split by rows of the RNS basis, the MAC's epilogue scattering the giant
    groups' accumulators to their owners over NVLink peer memory, the giant steps split by group, the ranks' accumulators
    summed by the fused peer all-reduce, one ModDown + rescale.  row_sets[i] = diagonal_set.slice_rows(rank, world) of
    mat-vec i.  Every rank ends with the same ciphertexts, bit-identical to the unsharded ones.  The data path needs
    peer-mapped windows: there is no NCCL stand-in for a scatter fused into a kernel's stores."""
    import torch.distributed as dist
    from . import pyPhantom as ph
    ctx = ckks.ctx
    if not dist.is_initialized() or dist.get_world_size(group) < 2:
        raise RuntimeError("split_matvec_batch needs a rank group of two or more (use bsgs_hoisted_batch on one GPU)")
    world = dist.get_world_size(group)
    exr = PeerExchange.get(ctx, group)
    outs = []
    for i0 in range(0, len(cts), 3):                      # three window slots, three auxiliary streams
        part = list(range(i0, min(i0 + 3, len(cts))))
        need = max(split_slot_bytes(ctx, -(-row_sets[i].D // row_sets[i].G), world) for i in part)
        exa = PeerExchange.get(ctx, group, tag="split", slot_bytes=need)
        if exa is None or exr is None:
            raise RuntimeError("two-phase mat-vec: the ranks' peer windows could not be mapped (CUDA IPC / peer access)")
        if len(part) > 1 and all(cts[i] is cts[part[0]] for i in part) and len({(row_sets[i].D, row_sets[i].G) for i in part}) == 1:
            # one input for all of them (the chunk pairs of a D -> F projection): this rank's baby steps once
            accs = ph.bsgs_split_shared(ctx, cts[part[0]], [row_sets[i] for i in part], ckks.gk, exa.window, 0)
        else:
            accs = ph.bsgs_split_batch(ctx, [cts[i] for i in part], [row_sets[i] for i in part], ckks.gk, exa.window, 0)
        for k, acc in enumerate(accs):
            exr.allreduce(acc, k)
        outs += [ph.bsgs_finish(ctx, acc) for acc in accs]
    exr.check(ctx)
    PeerExchange.get(ctx, group, tag="split").check(ctx)
    return outs


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


class PhasePlan:
    """Deal the k independent mat-vecs of a block phase to rank groups: min(k, world) contiguous groups take one
    mat-vec each per round; when fewer mat-vecs than groups remain (k = 3 on 2 GPUs) the rest is giant-step sharded
    over all ranks, so every rank carries the same load."""

    def __init__(self, k, world):
        self.k, self.world = k, world
        ng = min(k, world)
        parts = [tuple(int(r) for r in part) for part in np.array_split(np.arange(world), ng)]
        full = (k // ng) * ng
        self.assign = [(j, parts[j % ng]) for j in range(full)] + [(j, tuple(range(world))) for j in range(full, k)]
        self.groups = sorted({g for _, g in self.assign}, key=lambda g: (len(g), g))

    def mine(self, rank):
        """[(ranks of the group, [mat-vecs it serves])] for the groups `rank` belongs to, smallest group first"""
        return [(g, [j for j, gj in self.assign if gj == g]) for g in self.groups if rank in g]

    def leader(self, j):
        return dict(self.assign)[j][0]


class HybridBlock:
    """The eight projections of one client-aided RWKV-7 block (reference scripts/bootstrap_generation.py:756-899)
    served by all ranks: phases r,k,v | o | ffn_key pairs | ffn_val pairs, each phase planned by PhasePlan.
    Every rank runs the same client code and ends every phase with the same plaintext vectors."""

    PHASES = ("rkv", "o", "ffn_key", "ffn_val")

    @staticmethod
    def plans(world, D, F):
        from . import bsgs as hb
        npairs = len(hb._chunk_pairs(F, D))
        return {"rkv": PhasePlan(3, world), "o": PhasePlan(1, world), "ffn_key": PhasePlan(npairs, world),
                "ffn_val": PhasePlan(npairs, world)}

    @staticmethod
    def two_phase_default(world):
        """two-phase mat-vecs (rows | giant groups, split_matvec_batch) unless SPEAR_TWO_PHASE=0 asks for the round-1 plan
        (mat-vecs dealt to rank groups, giant steps sharded inside a group, baby steps replicated)"""
        return world > 1 and os.environ.get("SPEAR_TWO_PHASE", "1") != "0"

    @staticmethod
    def required_weights(world, D, F, two_phase=None):
        """baby weights (BSGS splits) whose rotation keys the context must hold"""
        from . import bsgs as hb
        if two_phase is None:
            two_phase = HybridBlock.two_phase_default(world)
        if two_phase:      # both phases are divided by the ranks: the single-GPU optimum holds for every world size
            return (hb.hoisting_weight(1),)
        return tuple(sorted({hb.hoisting_weight(len(g)) for p in HybridBlock.plans(world, D, F).values() for g in p.groups}))

    def __init__(self, ckks, block, D, F, rank=0, world=1, two_phase=None):
        import torch.distributed as dist
        from . import bsgs as hb
        self.ckks, self.D, self.F, self.rank, self.world = ckks, D, F, rank, world
        self.two_phase = self.two_phase_default(world) if two_phase is None else bool(two_phase)
        if self.two_phase:   # the windows come up here, outside any timed region; without them: the round-1 plan
            Gs, Bs = hb.compute_bsgs_params(D, hb.hoisting_weight(1))
            self.two_phase = two_phase_ready(ckks.ctx, Bs, world)
            if not self.two_phase and rank == 0:
                print("[spear] peer windows unavailable: two-phase mat-vecs off, giant-step sharding with the NCCL exchange "
                      "(the context must hold the rotation keys of HybridBlock.required_weights(..., two_phase=False))")
        self.plan = self.plans(world, D, F)
        self.pairs = hb._chunk_pairs(F, D)
        level = ckks.encrypt_replicated(np.zeros(1)).chain_index()
        # one process group per distinct rank group (every rank creates all of them, in the same order)
        self.pg = {}
        if world > 1:
            for ranks in sorted({g for p in self.plan.values() for g in p.groups}):
                self.pg[ranks] = dist.group.WORLD if len(ranks) == world else dist.new_group(list(ranks))
        mats = {"rkv": [("real", block.W_r.T), ("real", block.W_k.T), ("real", block.W_v.T)], "o": [("real", block.W_o.T)],
                "ffn_key": [], "ffn_val": []}
        for c, c2 in self.pairs:
            M1 = hb._key_chunk(block.W_key_ffn, c, D, F)
            mats["ffn_key"].append(("real", M1) if c2 is None else ("complex", M1, hb._key_chunk(block.W_key_ffn, c2, D, F)))
            M0 = hb._val_chunk(block.W_val_ffn, c, D, F)
            mats["ffn_val"].append(("real", M0) if c2 is None else ("complex", M0, hb._val_chunk(block.W_val_ffn, c2, D, F, -1.0)))
        self.sets = {}
        if self.two_phase:
            G, B = hb.compute_bsgs_params(D, hb.hoisting_weight(1))
            self.split_GB = (G, B)
            for phase in self.PHASES:
                for j, m in enumerate(mats[phase]):
                    enc = hb.pre_encode_real_diags if m[0] == "real" else hb.pre_encode_complex_diags
                    full = enc(ckks, *m[1:], D, G, B, level)
                    self.sets[(phase, j)] = full.slice_rows(rank, world)
                    del full
            return
        for phase in self.PHASES:
            for ranks, js in self.plan[phase].mine(rank):
                G, B = hb.compute_bsgs_params(D, hb.hoisting_weight(len(ranks)))
                for j in js:
                    m = mats[phase][j]
                    enc = hb.pre_encode_real_diags if m[0] == "real" else hb.pre_encode_complex_diags
                    self.sets[(phase, j)] = enc(ckks, *m[1:], D, G, B, level, shard=(ranks.index(rank), len(ranks)))

    def _encrypt_inputs(self, inputs, js, base):
        """Enc(inputs[j] replicated) with encryption id base + j for j in js: identical on every rank of the group"""
        ckks, D = self.ckks, self.D
        cts = []
        for j in js:
            cts.append(ckks.sk.encrypt_vector(ckks.ctx, np.asarray(inputs[j], dtype=np.complex128), ckks.scale,
                                              replicate=True, enc_id=base + j))
        return cts

    def _serve(self, phase, inputs):
        """inputs: k complex (or real) vectors of length D -> k complex result vectors, identical on every rank"""
        import torch
        import torch.distributed as dist
        ckks, D, plan = self.ckks, self.D, self.plan[phase]
        if self.two_phase:
            # every rank runs the same client code (identical ciphertexts: same key seed, same encryption ids), serves its
            # rows and its giant groups of EVERY mat-vec of the phase, and ends with every result ciphertext
            js = list(range(plan.k))
            if plan.k > 1 and all(v is inputs[0] for v in inputs):     # one vector for every mat-vec of the phase (ffn_key):
                cts = self._encrypt_inputs(inputs, [0], ckks.sk.reserve_enc_ids(1)) * plan.k   # one ciphertext, shared baby steps
            else:
                cts = self._encrypt_inputs(inputs, js, ckks.sk.reserve_enc_ids(plan.k))
            outs = split_matvec_batch(ckks, cts, [self.sets[(phase, j)] for j in js])
            return np.stack([ckks.decrypt_vec_complex(ct_y, D) for ct_y in outs])
        # Encryption ids come from the secret key's one monotonic counter: every rank runs the same client code in the
        # same order, so all ranks reserve the same range (the ranks of a group form identical ciphertexts) and no
        # (seed, nonce) pair is ever used twice -- neither across blocks, tokens, nor against plain encrypt calls.
        base = ckks.sk.reserve_enc_ids(plan.k)
        res = np.zeros((plan.k, D), dtype=np.complex128)
        for ranks, js in plan.mine(self.rank):
            cts = self._encrypt_inputs(inputs, js, base)
            outs = sharded_matvec_batch(ckks, cts, [self.sets[(phase, j)] for j in js], group=self.pg.get(ranks))
            for j, ct_y in zip(js, outs):
                res[j] = ckks.decrypt_vec_complex(ct_y, D)
        if self.world > 1 and any(len(g) < self.world for g in plan.groups):
            keep = np.array([plan.leader(j) == self.rank for j in range(plan.k)])
            res[~keep] = 0                                      # every result has exactly one contributing rank
            t = torch.from_numpy(res.view(np.float64)).to(f"cuda:{ckks.ctx.device}")
            dist.all_reduce(t)
            res = t.cpu().numpy().view(np.complex128)
        return res

    def rkv(self, mixed):
        out = self._serve("rkv", [mixed[n] for n in "rkv"])
        return out[0].real, out[1].real, out[2].real

    def o(self, gated):
        return self._serve("o", [gated])[0].real

    def ffn_key(self, x):
        out = self._serve("ffn_key", [x] * len(self.pairs))
        D, F, result = self.D, self.F, np.zeros(self.F)
        for (c, c2), vals in zip(self.pairs, out):
            lo1, hi1 = c * D, min((c + 1) * D, F)
            result[lo1:hi1] = vals.real[:hi1 - lo1]
            if c2 is not None:
                lo2, hi2 = c2 * D, min((c2 + 1) * D, F)
                result[lo2:hi2] = vals.imag[:hi2 - lo2]
        return result

    def ffn_val(self, x):
        D, F, ins = self.D, self.F, []
        for c, c2 in self.pairs:
            v = np.zeros(D, dtype=np.complex128)
            lo, hi = c * D, min((c + 1) * D, F)
            v[:hi - lo] = x[lo:hi]
            if c2 is not None:
                lo1, hi1 = c2 * D, min((c2 + 1) * D, F)
                v[:hi1 - lo1] += 1j * x[lo1:hi1]
            ins.append(v)
        return sum(vals.real for vals in self._serve("ffn_val", ins))
