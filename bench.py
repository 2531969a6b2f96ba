#!/usr/bin/env python3
"""bench.py -- d=2048 BSGS CKKS mat-vecs/s and server ms per RWKV-7 token on B200 (BASELINE.json metric, config C3/C4).

  python bench.py --gpus N --steps K --warmup W            this build (CUDA, hoisted BSGS)
  python bench.py --impl reference --gpus N ...            CPU arm: the oracle port of the reference's un-hoisted
                                                           op order on the host cores (PhantomFHE itself is not
                                                           available: SURVEY.md section 8c)

A step is one pass over the r, k, v projections of one RWKV-7 block: THREE 2048x2048 encrypted mat-vecs (CKKS N=32768,
L0=24 x 59-bit, P=3, G=46, B=45: 89 rotations each) over pre-encoded diagonals, the reference's
scripts/bootstrap_generation.py:784-792.  Rotation keys (10.1 GB) and diagonals are far larger than L2, so no explicit
flush is needed between iterations.

  N = 1   the three mat-vecs run on three streams of one GPU (spear_bsgs_hoisted_batch).
  N > 1   STRONG scaling: the same three mat-vecs are served by all N GPUs -- dealt to rank groups, giant steps sharded
          inside a group (fhe_spear_b200.sharding.PhasePlan), shard accumulators combined by the fused peer-memory
          exchange (spear_peer_allreduce over NVLink, csrc/peer.cu) inside the timed region.

`value` is timed with CUDA events on the engine's stream with the input ciphertexts already in HBM, max over ranks;
`e2e` goes through the pyPhantom call surface with pinned host buffers (H2D of the input ciphertexts and D2H of the
results inside the timed region).  `token` is a MEASURED 24-block RWKV-7 token loop (client-aided blocks, 192 mat-vecs,
reference :983-1011) on the same N GPUs.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "d=2048 BSGS CKKS matvecs/s"
UNIT = "matvecs/s"
CONFIGS = {
    # name: (N, L0, P, D)
    "c3": (32768, 24, 3, 2048),
    "c2": (16384, 24, 3, 1024),
    "small": (4096, 6, 3, 64),
}
# Issue cost of the multiplier instructions on sm_100a, cycles per warp instruction and SM sub-partition, measured by
# tools/ubench/imad.cu and mac2.cu on this pool's B200s (profiles/r1_ubench_imad.log, r1_ubench_mac.log)
CYC_BUTTERFLY = 4 * 2.05 + 5.4 + 4 * 2.0     # lazy Harvey butterfly: 4 IMAD.WIDE, 1 accumulating IMAD.WIDE, 4 IMAD
CYC_MAC_TERM = 15.5                          # one split-30 Karatsuba term: 3 accumulating IMAD.WIDE (measured chain)
SMSP_PER_GPU = 148 * 4


def split_of(D):
    G = int(np.ceil(np.sqrt(D)))
    return G, int(np.ceil(D / G))


def workload_name(cfg):
    N, L0, P, D = CONFIGS[cfg]
    G, B = split_of(D)
    return (f"{cfg.upper()}: r,k,v projections of one RWKV-7 block = 3 x ({D}x{D} BSGS mat-vec, CKKS N={N}, L0={L0}x59-bit, "
            f"P={P}, G={G} B={B}, {G + B - 2} rotations), pre-encoded diagonals")


# ---- clocks -----------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled in the background.  It is started well before the timed region (its start-up takes the
    driver lock for a while, which would stall kernel launches) and samples are filtered to the timed window."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[1]), float(parts[2]),
                             {n for n, v in zip(names, parts[4:8]) if v.lower().startswith("active")}))
            except ValueError:
                continue
        os.unlink(self.f.name)
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.05 <= r[0] <= (self.t1 or 1e18) + 0.05]
        use = inside or rows
        if use:
            out.update(sm_mhz=float(np.median([r[1] for r in use])), sm_max_mhz=float(max(r[2] for r in use)),
                       reasons=sorted(set().union(*[r[3] for r in use])), samples=len(use),
                       window="timed region" if inside else "whole run (no sample fell inside the timed region)")
        return out


# ---- CPU arm: oracle port of the reference's op order -----------------------------------------------------
class CpuMatvec:
    """The reference's BSGS loop (un-hoisted baby rotations scripts/bootstrap_generation.py:215-220, the Python loop
    :464-484: multiply_plain + add per diagonal, one rotate per giant group, one rescale) over the oracle's primitives
    on the host cores, at full size.  Timing does not depend on the values, so random residues stand in for the
    ciphertext, the rotation keys (four distinct 113 MB buffers used in turn, so none is cache-resident) and the
    plaintext diagonals (G distinct ones).  The OpenMP team is set explicitly to every core: launchers such as
    torch.distributed.run export OMP_NUM_THREADS=1."""

    def __init__(self, cfg, seed=0):
        from oracle.oracle import Oracle
        self.N, self.L0, self.P, self.D = CONFIGS[cfg]
        self.G, self.B = split_of(self.D)
        self.cores = os.cpu_count() or 1
        self.threads = Oracle.set_threads(self.cores)
        N, L0, P = self.N, self.L0, self.P
        q = Oracle.create_coeff_modulus(N, [59] * (L0 + P))
        self.o = o = Oracle(N, q, P)
        rng = np.random.default_rng(seed)
        K = L0 + P

        def residues(shape_front, limbs):
            out = np.empty(tuple(shape_front) + (len(limbs), N), dtype=np.uint64)
            for pos, i in enumerate(limbs):
                out[..., pos, :] = rng.integers(0, int(q[i]), tuple(shape_front) + (N,), dtype=np.uint64)
            return out
        self.ct = residues((2,), range(L0))
        beta = o.num_digits(L0)
        self.keys = [residues((beta, 2), range(K)) for _ in range(4)]
        self.pts = [residues((), range(L0)) for _ in range(self.G)]
        self.belt = [0] + [o.elt_from_step(b) for b in range(1, self.G)]
        self.gelt = [0] + [o.elt_from_step(g * self.G) for g in range(1, self.B)]
        self.baby = None

    def whole(self):
        """one whole mat-vec, op for op; seconds"""
        o, G, B, D = self.o, self.G, self.B, self.D
        t0 = time.perf_counter()
        baby = [self.ct] + [o.apply_galois(self.ct, self.belt[b], self.keys[b % 4]) for b in range(1, G)]
        res = None
        for g in range(B):
            nb = min(G, D - g * G)
            inner = None
            for b in range(nb):
                prod = o.multiply_plain(baby[b], self.pts[b])
                inner = prod if inner is None else o.add(inner, prod)
            if g > 0:
                inner = o.apply_galois(inner, self.gelt[g], self.keys[g % 4])
            res = inner if res is None else o.add(res, inner)
        o.rescale(res)
        dt = time.perf_counter() - t0
        self.baby = baby
        return dt

    def slice(self, g=1):
        """one giant group's share of a mat-vec: one baby rotation, G plaintext MACs, one giant rotation, one
        accumulation = 2/(G+B-2) of the rotations and G/D of the MACs (both 1/44.5 at C3); seconds"""
        o, G = self.o, self.G
        if self.baby is None:
            self.baby = [self.ct] * G
        t0 = time.perf_counter()
        self.baby[1 + g % (G - 1)] = o.apply_galois(self.ct, self.belt[1 + g % (G - 1)], self.keys[g % 4])
        inner = None
        for b in range(G):
            prod = o.multiply_plain(self.baby[b], self.pts[b])
            inner = prod if inner is None else o.add(inner, prod)
        inner = o.apply_galois(inner, self.gelt[1 + g % (self.B - 1)], self.keys[(g + 1) % 4])
        o.add(self.ct, inner)
        return time.perf_counter() - t0

    @property
    def slice_fraction(self):
        return 2.0 / (self.G + self.B - 2)


def cpu_baseline(cfg):
    """cpu_baseline of our own arm: ONE whole un-hoisted mat-vec on the host cores (about 10 s at C3)."""
    m = CpuMatvec(cfg)
    s = m.whole()
    return {"value": 1.0 / s, "unit": UNIT, "cores": m.threads, "kind": "port",
            "sample": (f"one whole un-hoisted {cfg.upper()} mat-vec, op for op ({m.G - 1} baby + {m.B - 1} giant full rotations, {m.D} "
                       f"multiply_plain + add, 1 rescale) = {s:.2f} s on {m.threads} OpenMP threads of {m.cores} host cores; "
                       "oracle C port of the reference's loop (scripts/bootstrap_generation.py:215-220, 464-484)")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_wall = time.perf_counter()
    m = CpuMatvec(args.config)
    whole_s = m.whole()                    # one whole mat-vec first: the anchor of the sampled steps (and their warm-up)
    for w in range(args.warmup):
        m.slice(w)
    t0 = time.perf_counter()
    for k in range(args.steps):
        m.slice(k)
    dt = time.perf_counter() - t0
    frac = m.slice_fraction
    value = args.steps * frac / dt
    wall = time.perf_counter() - t_wall
    sample = (f"a step = one giant group's share of a mat-vec (1 baby rotation + {m.G} multiply_plain/add + 1 giant rotation + 1 add "
              f"= {frac:.5f} mat-vec) at full {args.config.upper()} size, {dt / args.steps * 1e3:.0f} ms each; one WHOLE un-hoisted "
              f"mat-vec run first in the same process: {whole_s:.2f} s = {1.0 / whole_s:.4f} mat-vecs/s; {m.threads} OpenMP threads "
              f"of {m.cores} host cores; oracle C port of the reference's loop")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": workload_name(args.config), "mode": "reference op order on CPU (un-hoisted)",
                   "matvecs_per_step": frac,
                   "note": "PhantomFHE (the reference's GPU library) is absent and unpinned; this arm is the oracle port"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": m.threads, "kind": "port", "sample": sample},
        "whole_matvec": {"seconds": whole_s, "value": 1.0 / whole_s, "unit": UNIT},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line))


# ---- op counts of one mat-vec (integer roofline) -------------------------------------------------------------------
def integer_work(N, l, P, D, G, B, n_giant=None, n_baby=None, row_share=None):
    """Multiplier-pipe work of one hoisted mat-vec: NTT butterflies and modular multiply-accumulate terms, from the
    algorithm, in SM-sub-partition cycles at the measured issue cost of the instructions that
    carry them.  n_giant / n_baby: rotations done by this rank (sharded runs)."""
    n_giant = B - 1 if n_giant is None else n_giant
    n_baby = G - 1 if n_baby is None else n_baby
    rows, beta, logn = l + P, -(-l // P), N.bit_length() - 1
    dig = beta * rows - l                                   # ModUp'd rows transformed per decomposition
    row_ntts = (l + dig) + n_giant * (P + l + l + dig) + (2 * (P + l) + 2 + 2 * (l - 1))
    butterflies = row_ntts * (N // 2) * logn
    d_share = (n_giant + 1) / B                             # share of the diagonals this rank multiplies
    if row_share is not None:                               # two-phase mat-vec: this rank's rows of every baby step and diagonal
        d_share, n_baby = row_share, n_baby * row_share
    mac_terms = ((n_baby + n_giant) * rows * N * beta * 2   # key inner products
                 + D * d_share * 2 * rows * N               # diagonal MAC
                 + (1 + n_giant) * beta * (rows - P) * N * P   # ModUp
                 + (n_giant + 2) * l * N * P)               # ModDown
    cycles = butterflies / 32 * CYC_BUTTERFLY + mac_terms / 32 * CYC_MAC_TERM
    return {"row_ntts": row_ntts, "butterflies": butterflies, "mac_terms": mac_terms, "multiplier_warp_cycles": cycles}


KERNEL_OF = {
    "ks_baby_fused": "k_ks_baby_fused (all hoisted baby-step rotation-key inner products, TMA-staged key stream)",
    "ntt_ks_fused": "k_ntt_b_ks_all / k_ntt_b_ks (forward pass B of the ModUp'd digits fused with the giant-step key inner product)",
    "pmac": "k_pmac_tma (plaintext-diagonal multiply-accumulate, TMA-staged diagonals)",
    "modup": "k_intt_modup_fwd_a (inverse pass A + n^-1*hatinv + ModUp + forward pass A in one kernel) / k_modup",
    "ntt_fwd_a": "ntt_fwd_a2 (forward NTT pass A)", "ntt_fwd_b": "ntt_fwd_b2 (forward NTT pass B)",
    "ntt_inv_a": "ntt_inv_a2 (inverse NTT pass A)", "ntt_inv_b": "ntt_inv_b2 (inverse NTT pass B)",
    "moddown": "k_moddown_conv + k_moddown_final", "rescale": "k_rescale_*", "ks_inner": "k_ks_inner_tma",
    "sum_groups": "k_sum_groups (sum of the giant groups' partial results)",
    "peer_wait": "k_peer_sync (epoch flags over NVLink: post + wait for the peers -- the time this rank waits)",
    "peer_reduce": "k_peer_reduce + window copies (fused reduce-scatter + Barrett + all-gather over peer memory)",
}


# ---- this build ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from fhe_spear_b200 import _native
    from fhe_spear_b200 import bsgs as hb
    from fhe_spear_b200 import pyPhantom as ph
    from fhe_spear_b200 import sharding as sh
    from fhe_spear_b200 import rwkv_block as rb

    N, L0, P, D = CONFIGS[args.config]
    G, B = hb.compute_bsgs_params(D)
    F = 4 * D
    do_token = args.config == "c3" and not args.no_token
    do_tuned = not args.no_tuned and world == 1
    t_setup = time.perf_counter()
    weights = {1.0}
    if do_tuned:
        weights.add(args.tuned_weight)
    if do_token:   # (both multi-GPU plans' splits: the two-phase path falls back to the round-1 plan without peer windows)
        weights |= (set(sh.HybridBlock.required_weights(world, D, F, two_phase=True)) |
                    set(sh.HybridBlock.required_weights(world, D, F, two_phase=False))) if world > 1 else {hb.hoisting_weight(1)}
    ckks = hb.CKKSBootstrapContext(poly_degree=N, L0=L0, prime_bits=59, special_mod_size=P, max_rot_dim=1,
                                   bsgs_dim=[D], skip_bootstrap=True, seed=bytes(range(32)), device=local,
                                   verbose=(rank == 0 and args.verbose), baby_weights=tuple(sorted(weights)))
    ctx = ckks.ctx
    nb = args.batch
    rng = np.random.default_rng(1000)                     # every rank draws the same matrices and inputs
    Ws = [rng.standard_normal((D, D)) * 0.02 for _ in range(nb)]
    xs = [rng.standard_normal(D) * 0.1 for _ in range(nb)]
    cts = [ckks.encrypt_replicated(x) for x in xs]        # identical on every rank (same key seed, same counter)
    compress = not args.full_diagonals

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()   # the barrier itself (and NCCL's lazy set-up) must be over before timing

    if world == 1:
        dsets = [hb.pre_encode_real_diags(ckks, W, D, G, B, level=1, compress=compress) for W in Ws]
        info = dsets[0].info()
        mine = [((0,), list(range(nb)))]
        pg = {}
        two_phase = False

        def step():
            return ph.bsgs_hoisted_batch(ctx, cts, dsets, ckks.gk)
        parallelism = "1 GPU: three mat-vecs on three streams"
        n_giant_rank = B - 1
    elif sh.HybridBlock.two_phase_default(world) and sh.two_phase_ready(ctx, max(B, hb.compute_bsgs_params(D, hb.hoisting_weight(1))[1]), world):
        # strong scaling, two-phase mat-vecs: EVERY mat-vec is served by all N ranks -- baby steps and diagonal MAC split
        # by rows of the RNS basis, the MAC's epilogue scattering the giant groups' accumulators to their owners over NVLink
        # peer memory, giant steps split by group, accumulators summed by the fused peer all-reduce
        pg = {}
        two_phase = True
        all_ranks = tuple(range(world))
        mine = [(all_ranks, list(range(nb)))]
        dsets = {}
        for j in range(nb):
            full = hb.pre_encode_real_diags(ckks, Ws[j], D, G, B, level=1, compress=compress)
            if j == 0:
                info = full.info()
            dsets[j] = full.slice_rows(rank, world)
            del full

        def step():
            return dict(enumerate(sh.split_matvec_batch(ckks, cts, [dsets[j] for j in range(nb)])))
        r0_, r1_, c0_, c1_ = ph.diagonal_set.share(L0, P, N, rank, world)
        parallelism = (f"{world} GPUs, strong scaling, two-phase mat-vecs: every mat-vec on all ranks -- hoisted baby steps + diagonal "
                       f"MAC split by rows of the {L0 + P}-limb basis{' and column halves' if world >= 5 else ''} (k_pmac_tma's epilogue stores scatter the giant groups' "
                       f"accumulators to their owners over NVLink peer memory), giant steps split by group, accumulators summed by "
                       f"spear_peer_allreduce; all inside the timed region")
        n_giant_rank = max(len(sh.giant_groups(B, r, world)) - (1 if r == 0 else 0) for r in range(world))
    else:
        # strong scaling (round-1 plan, SPEAR_TWO_PHASE=0): the same nb mat-vecs dealt to rank groups, giant steps sharded
        # inside a group; every rank creates every process group in the same order
        two_phase = False
        plan = sh.PhasePlan(nb, world)
        pg = {ranks: (dist.group.WORLD if len(ranks) == world else dist.new_group(list(ranks))) for ranks in plan.groups}
        mine = plan.mine(rank)
        dsets = {}
        for ranks, js in mine:
            for j in js:
                dsets[j] = hb.pre_encode_real_diags(ckks, Ws[j], D, G, B, level=1, compress=compress,
                                                    shard=(ranks.index(rank), len(ranks)))
        info = next(iter(dsets.values())).info()
        for ranks in plan.groups:                          # peer windows of every group, outside the timed regions
            if rank in ranks and len(ranks) > 1:
                sh.PeerExchange.get(ctx, pg[ranks])

        def step():
            outs = {}
            for ranks, js in mine:
                ys_ = sh.sharded_matvec_batch(ckks, [cts[j] for j in js], [dsets[j] for j in js], group=pg[ranks])
                outs.update(zip(js, ys_))
            return outs
        parallelism = (f"{world} GPUs, strong scaling: mat-vecs dealt to rank groups {plan.groups}, giant steps sharded inside a group, "
                       f"shard accumulators combined by spear_peer_allreduce (fused reduce-scatter + Barrett + all-gather over NVLink "
                       f"peer memory) inside the timed region")
        n_giant_rank = max(len(sh.giant_groups(B, ranks.index(rank), len(ranks))) for ranks, _ in mine)
    ctx.synchronize()
    t_setup = time.perf_counter() - t_setup

    # correctness of exactly what is timed: decrypt error vs float64 W.x, and sharded == unsharded limb for limb
    outs = step()
    outs = dict(enumerate(outs)) if world == 1 else outs
    err = max(float(np.abs(ckks.decrypt_vec(yy, D) - Ws[j] @ xs[j]).max()) for j, yy in outs.items())
    if not err < 1e-6:
        raise SystemExit(f"bench: decrypted result is wrong (max abs err {err})")
    j0 = sorted(outs)[0]
    full0 = dsets[j0] if world == 1 else hb.pre_encode_real_diags(ckks, Ws[j0], D, G, B, level=1, compress=compress)
    y_single = ph.bsgs_hoisted(ctx, cts[j0], full0, ckks.gk)
    ref_limbs = y_single.to_numpy()
    if not np.array_equal(ref_limbs, outs[j0].to_numpy()):
        raise SystemExit("bench: batched / sharded result differs from the single-call result")
    del outs

    if world > 1:                      # bring the communicator up outside every timed region
        warm = torch.zeros(1, device="cuda")
        dist.all_reduce(warm)
        barrier()

    clocks = ClockSampler(local)
    time.sleep(1.0)                       # let nvidia-smi finish starting up before anything is timed
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    if args.count_only:
        clocks.stop()
        print(_native.launch_count())
        return
    # three back-to-back timed regions of exactly K steps each (barrier + synchronize on both sides, CUDA events on
    # the engine stream, max over ranks); the median region is reported, all three are listed
    region_ms, launches = [], 0
    clocks.mark_start()
    for _rep in range(3):
        barrier()
        launches0 = _native.launch_count()
        ctx.timer_start()
        for _ in range(args.steps):
            step()
        ms = ctx.timer_stop()
        launches = _native.launch_count() - launches0
        barrier()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        region_ms.append(float(t.item()))
    clocks.mark_end()
    clk = clocks.stop()
    ms_max = float(np.median(region_ms))
    value = args.steps * nb / (ms_max * 1e-3)             # the whole job: nb mat-vecs per step, whatever N
    if world > 1:
        tl = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tl)
        launches = int(tl.item())

    # latency and per-kernel times of ONE mat-vec alone on the engine stream (rank-local: the unsharded mat-vec at N = 1,
    # this rank's shard accumulator + finish at N > 1), event pair around each launch, for the roofline line
    if world == 1:
        def single():
            return ph.bsgs_hoisted(ctx, cts[j0], full0, ckks.gk)
    elif two_phase:
        def single():          # collective: every rank runs it the same number of times
            return sh.split_matvec_batch(ckks, [cts[j0]], [dsets[j0]])[0]
    else:
        def single():
            return ph.bsgs_finish(ctx, ph.bsgs_hoisted_partial(ctx, cts[j0], dsets[j0], ckks.gk))
    for _ in range(2):
        single()
    ctx.timer_start()
    for _ in range(args.steps):
        single()
    single_ms = ctx.timer_stop() / args.steps
    ctx.profile(True)
    for _ in range(args.steps):
        single()
    prof = ctx.profile_read()
    ctx.profile(False)

    # End to end through the public call surface, from the client's float vector to the client's float result: encode +
    # encrypt (ckks.sk.encrypt_vector: H2D of the vector, three launches), the ciphertext crosses the client/server
    # boundary through pinned host memory (D2H + H2D), the mat-vecs, the result ciphertext crosses back (D2H + H2D),
    # decrypt + decode (three launches, D2H of the result vector).  The reference's fhe_projection_bsgs does the same
    # round trip in one process (scripts/bootstrap_generation.py:545-556).
    l = cts[0].coeff_modulus_size()
    my_js = sorted({j for _, js in mine for j in js})
    h_x = {j: ph.pinned_empty((D,), dtype=np.complex128) for j in my_js}
    h_in = {j: ph.pinned_empty((2, l, N)) for j in my_js}
    h_out = {j: ph.pinned_empty((2, l - 1, N)) for j in my_js}
    for j in my_js:
        h_x[j][:] = xs[j]
    scale = cts[0].scale()
    e2e_base = ckks.sk.reserve_enc_ids(nb * (args.steps + 4))      # the same ids on every rank of a group
    e2e_calls = [0]
    y_host = {}

    def e2e_step():
        base = e2e_base + nb * e2e_calls[0]
        e2e_calls[0] += 1
        for j in my_js:                                                                  # client: float vector -> ciphertext
            ckks.sk.encrypt_vector(ctx, h_x[j], ckks.scale, replicate=True, enc_id=base + j).to_numpy(out=h_in[j])   # -> wire (D2H)
        if world == 1:
            # server, host buffers in and out: upload, mat-vec and download of each item pipelined over the engine's streams
            sc = ph.bsgs_hoisted_batch_host(ctx, [h_in[j] for j in range(nb)], scale, dsets, ckks.gk, [h_out[j] for j in range(nb)])
            for j in my_js:                                                              # wire -> client (H2D), decrypt + decode (D2H)
                back = ph.ciphertext.from_numpy(ctx, h_out[j], sc[j])
                y_host[j] = ckks.sk.decrypt_decode(ctx, back, D)
            return
        ins = {j: ph.ciphertext.from_numpy(ctx, h_in[j], scale) for j in my_js}          # wire -> server (H2D)
        if two_phase:
            res = dict(enumerate(sh.split_matvec_batch(ckks, [ins[j] for j in range(nb)], [dsets[j] for j in range(nb)])))
        else:
            res = {}
            for ranks, js in mine:
                res.update(zip(js, sh.sharded_matvec_batch(ckks, [ins[j] for j in js], [dsets[j] for j in js],
                                                           group=pg[ranks])))
        for j in my_js:
            res[j].to_numpy(out=h_out[j])                                                # server -> wire (D2H)
        for j in my_js:                                                                  # wire -> client (H2D), decrypt + decode (D2H)
            back = ph.ciphertext.from_numpy(ctx, h_out[j], res[j].scale())
            y_host[j] = ckks.sk.decrypt_decode(ctx, back, D)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    per_step_h2d = sum(h_x[j].nbytes + h_in[j].nbytes + h_out[j].nbytes for j in my_js)
    per_step_d2h = sum(h_in[j].nbytes + h_out[j].nbytes + D * 16 for j in my_js)
    te = torch.tensor([e2e_s, float(per_step_h2d), float(per_step_d2h)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = te.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(te)
        te[0] = tmax[0]
    e2e_value = args.steps * nb / float(te[0].item())
    h2d_bytes, d2h_bytes = int(te[1].item()), int(te[2].item())
    e2e_err = max(float(np.abs(y_host[j].real - Ws[j] @ xs[j]).max()) for j in my_js)
    if not e2e_err < 1e-6:
        raise SystemExit(f"bench: end-to-end result is wrong (max abs err {e2e_err})")

    # secondary measurement (not the headline): the same mat-vecs with a hoisting-aware split G = ceil(sqrt(w D))
    tuned = None
    if do_tuned:
        G2, B2 = hb.compute_bsgs_params(D, args.tuned_weight)
        dsets2 = [hb.pre_encode_real_diags(ckks, Wm, D, G2, B2, level=1) for Wm in Ws]
        ys2 = ph.bsgs_hoisted_batch(ctx, cts, dsets2, ckks.gk)
        err2 = max(float(np.abs(ckks.decrypt_vec(yy, D) - WW @ xx).max()) for yy, WW, xx in zip(ys2, Ws, xs))
        for _ in range(3):
            ph.bsgs_hoisted_batch(ctx, cts, dsets2, ckks.gk)
        barrier()
        ctx.timer_start()
        for _ in range(args.steps):
            ph.bsgs_hoisted_batch(ctx, cts, dsets2, ckks.gk)
        t2 = ctx.timer_stop()
        for _ in range(2):                                   # the single-stream path grows its own workspace once
            ph.bsgs_hoisted(ctx, cts[0], dsets2[0], ckks.gk)
        ctx.timer_start()
        for _ in range(args.steps):
            ph.bsgs_hoisted(ctx, cts[0], dsets2[0], ckks.gk)
        lat2 = ctx.timer_stop() / args.steps
        tuned = {"split": f"G={G2} B={B2} ({G2 + B2 - 2} rotations)", "value": args.steps * nb / (t2 * 1e-3),
                 "unit": UNIT, "latency_ms_single_matvec": lat2, "max_abs_err_vs_float64": err2,
                 "note": "same matrices and ciphertexts; not the BASELINE config (which names G=46 B=45)"}
        del dsets2, ys2

    # the second half of the metric, MEASURED: one RWKV-7 token = 24 client-aided blocks x 8 projections on these N GPUs
    token = None
    if do_token:
        token = measure_token(ckks, hb, rb, sh, D, F, rank, world, args.tokens)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the DOMINANT kernel of the un-overlapped mat-vec, picked live ------------------------------
    beta = (l + P - 1) // P
    key_bytes = beta * 2 * (l + P) * N * 8
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    n_baby = G - 1
    row_share = 1.0
    if world > 1 and two_phase:      # this rank's rows of the baby-step keys and of the diagonals
        row_share = (r1_ - r0_) * (c1_ - c0_) / float((l + P) * N)
    alg_bytes = {   # algorithmic bytes per MAT-VEC of each kernel: SURVEY.md section 8(d) -- keys and diagonals read once
        "ks_baby_fused": n_baby * key_bytes * row_share, "ntt_ks_fused": n_giant_rank * key_bytes,
        "ks_inner": n_giant_rank * key_bytes, "pmac": info["bytes"] * row_share,
        # the decomposition front end streams no keys or diagonals: its compulsory traffic is the polynomial it reads (both
        # forms) and the digits it writes, once each, per decomposition (one per giant step + the baby steps' own)
        "modup": (n_giant_rank + 1) * (2 * l + beta * (l + P)) * N * 8,
    }
    alg_kind = {"modup": "compulsory intermediates: inputs read once + digits written once (this kernel streams no keys or diagonals)"}
    traffic_tab = {}
    tpath = os.path.join(ROOT, "profiles", "r2_kernel_traffic.json")
    if args.config == "c3" and world == 1 and os.path.exists(tpath):
        traffic_tab = json.load(open(tpath))     # ncu dram__bytes_read.sum + dram__bytes_write.sum per launch (static: ncu cannot run inside a timed bench)
    kernels = {}
    for k, v in prof.items():
        if v["launches"] == 0:
            continue
        per_matvec_ms = v["ms"] / args.steps
        lpm = v["launches"] / args.steps
        ab = alg_bytes.get(k, 0)
        kernels[k] = {"kernel": KERNEL_OF.get(k, k), "ms_per_matvec": per_matvec_ms, "launches_per_matvec": lpm,
                      "avg_launch_ms": v["ms"] / v["launches"], "share_of_step": per_matvec_ms / single_ms,
                      "algorithmic_bytes_per_launch": ab / lpm,
                      "algorithmic_bytes_kind": alg_kind.get(k, "rotation keys / diagonals, each read once (SURVEY.md section 8d)" if ab else "none"),
                      "achieved": ab / (per_matvec_ms * 1e-3) / 1e9 if per_matvec_ms > 0 else 0.0}
        kernels[k]["frac"] = kernels[k]["achieved"] / peak
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_matvec"])
    dk = kernels[dom]
    iw = integer_work(N, l, P, D, G, B, n_giant=n_giant_rank, row_share=row_share if (world > 1 and two_phase) else None)
    avail = single_ms * 1e-3 * (clk["sm_mhz"] or 1965.0) * 1e6 * SMSP_PER_GPU
    mv_bytes = (n_baby * row_share + n_giant_rank) * key_bytes + info["bytes"] * row_share + (4 * l - 2) * N * 8
    roofline = {
        "bound": "hbm", "kernel": dk["kernel"], "class": dom, "achieved": dk["achieved"], "peak": peak, "unit": "GB/s",
        "frac": dk["frac"], "traffic": (traffic_tab.get(dom) or {}).get("dram_bytes_per_launch"),
        "traffic_source": (traffic_tab.get(dom) or {}).get("source"),
        "algorithmic_bytes_per_launch": dk["algorithmic_bytes_per_launch"], "algorithmic_bytes_kind": dk["algorithmic_bytes_kind"],
        "avg_launch_ms": dk["avg_launch_ms"],
        "launches_per_matvec": dk["launches_per_matvec"], "share_of_step": dk["share_of_step"],
        "how": ("dominant kernel = the kernel class with the largest CUDA-event time in an un-overlapped mat-vec, picked live; achieved = "
                "algorithmic bytes (rotation keys / diagonals, each read once; for a kernel that streams neither, the intermediates it must "
                "read and write once) / its time"),
        "peak_source": peak_src,
        "kernels": kernels,
        "matvec": {"algorithmic_bytes": mv_bytes, "achieved_gbs": mv_bytes / (single_ms * 1e-3) / 1e9,
                   "frac": mv_bytes / (single_ms * 1e-3) / 1e9 / peak, "diagonal_bytes": info["bytes"],
                   "key_bytes": (n_baby * row_share + n_giant_rank) * key_bytes},
        "integer": {"bound": "integer multiplier pipe", "unit": "SM-sub-partition cycles per mat-vec",
                    "achieved": iw["multiplier_warp_cycles"], "peak": avail, "frac": iw["multiplier_warp_cycles"] / avail,
                    "butterflies": iw["butterflies"], "mac_terms": iw["mac_terms"], "row_transforms": iw["row_ntts"],
                    "how": (f"work the algorithm needs on the multiplier pipe: butterflies x {CYC_BUTTERFLY:.1f} + multiply-accumulate terms x "
                            f"{CYC_MAC_TERM} cycles per warp (issue costs measured by tools/ubench on this pool's B200s), against the cycles "
                            f"{SMSP_PER_GPU} sub-partitions offer in the measured {single_ms:.3f} ms at the sampled SM clock")},
    }

    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_baseline(args.config)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_max / args.steps, "region_ms": region_ms, "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(args.config), "mode": "hoisted BSGS (spear_bsgs_hoisted)",
                   "matvecs_per_step": nb,
                   "diagonals": f"pre-encoded, basis Q_l*P, ring {info['ring_n']} ({'sub-ring compressed' if info['ring_n'] < N else 'full ring'})",
                   "l2": "inputs larger than L2 (rotation keys + diagonals >> 126 MB); no flush",
                   "latency_ms_single_matvec": single_ms,
                   "parallelism": parallelism,
                   "max_abs_err_vs_float64": err},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "path": "float vector -> encrypt_vector -> ciphertext over pinned host memory -> mat-vecs (N = 1: "
                        "spear_bsgs_hoisted_batch_host, uploads / mat-vecs / downloads pipelined over three streams) -> ciphertext over "
                        "pinned host memory -> decrypt_decode -> float vector", "max_abs_err_vs_float64": e2e_err},
        "gpu_launches": int(launches),
        "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"]},
        "roofline": roofline,
        "cpu_baseline": cpu,
        "token": token,
        "tuned_split": tuned,
        "setup_s": t_setup,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measure_token(ckks, hb, rb, sh, D, F, rank, world, n_tokens):
    """Server time of one RWKV-7 token through the client-aided block loop (reference scripts/bootstrap_generation.py:983-1011,
    756-899): 24 blocks x (r,k,v | o | 2 ffn_key pairs | 2 ffn_val pairs) = 192 encrypted mat-vecs with their client legs.
    Random-init weights of the named shapes; one block's eight diagonal sets are shared by all blocks (a 24-block
    model holds 348 GB of diagonals: 8 GPUs x 43 GB).  Timed as the reference times it (host clock around every server
    round, which includes encode+encrypt and decrypt+decode) and, beside it, with CUDA events over the whole token."""
    import copy
    import torch
    import torch.distributed as dist
    H, S = max(1, D // 64), min(64, D)
    blocks_n = 24
    base = rb.RWKVBlockWeights.random(D, F, H, S, block_idx=0, seed=0)
    if world > 1:
        pe = sh.HybridBlock(ckks, base, D, F, rank, world)
        if pe.two_phase:
            split = (f"G={pe.split_GB[0]} B={pe.split_GB[1]} (hoisting-aware), every mat-vec two-phase over all {world} ranks: rows | giant groups")
        else:
            split = "per rank group: G = ceil(sqrt(8 / size * D))"
    else:
        w = hb.hoisting_weight(1)
        Gt, Bt = hb.compute_bsgs_params(D, w)
        pe = hb.pre_encode_block(ckks, base, D, F, G=Gt, B=Bt)
        split = f"G={Gt} B={Bt} (hoisting-aware; limb-exact vs the oracle in tests/test_gpu_fullsize_parity.py)"
    blocks = []
    for i in range(blocks_n):
        b = copy.copy(base)
        b.block_idx = i
        blocks.append(b)
    rng = np.random.default_rng(3)
    vocab = 512
    emb, head = rng.standard_normal((vocab, D)) * 0.1, rng.standard_normal((D, vocab)) * 0.02
    ones, zeros = np.ones(D), np.zeros(D)
    xa = [zeros.copy() for _ in blocks]
    xf = [zeros.copy() for _ in blocks]
    st = [np.zeros((H, S, S)) for _ in blocks]
    pxa, pxf, pst = list(xa), list(xf), list(st)
    tok, rows = 3, []
    for _step in range(n_tokens):
        ckks.ctx.synchronize()
        if world > 1:
            dist.barrier()
        ckks.ctx.timer_start()
        t0 = time.perf_counter()
        logits, xa, xf, st, tms = rb.generate_token_fhe(ckks, blocks, emb, head, ones, zeros, ones, zeros, tok, xa, xf, st, D,
                                                        use_bsgs=True, preencoded_blocks=[pe] * len(blocks))
        wall = time.perf_counter() - t0
        dev_ms = ckks.ctx.timer_stop()
        ref, pxa, pxf, pst = rb.generate_token_plaintext(blocks, emb, head, ones, zeros, ones, zeros, tok, pxa, pxf, pst, D)
        server = sum(v for tm in tms for k, v in tm.items() if k.startswith("server_"))
        rows.append({"server_ms": server * 1e3, "device_ms": dev_ms, "wall_ms": wall * 1e3,
                     "max_abs_logit_err": float(np.abs(logits - ref).max())})
        tok = int(np.argmax(ref))
    best = min(rows, key=lambda r: r["server_ms"])
    t = torch.tensor([best["server_ms"], best["device_ms"], best["wall_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del pe
    return {"metric": "server ms per RWKV-7 token (24 client-aided blocks, d=2048, d_ffn=8192, 192 mat-vecs, measured)",
            "server_ms_per_token": float(t[0].item()), "device_event_ms_per_token": float(t[1].item()),
            "wall_ms_per_token": float(t[2].item()), "n_gpus": world, "tokens_run": n_tokens, "split": split,
            "max_abs_logit_err_vs_float64": max(r["max_abs_logit_err"] for r in rows),
            "note": "server_ms sums the reference's server_* timers (host clock; they include client encode+encrypt and decrypt+decode "
                    "of every projection), max over ranks; best of the tokens run"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--full-diagonals", action="store_true", help="store diagonals on the full ring (12.9+ GB at C3)")
    ap.add_argument("--batch", type=int, default=3, help="independent projections per step (3 = r, k, v of one block)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tuned", action="store_true", help="skip the secondary hoisting-aware-split measurement")
    ap.add_argument("--no-token", action="store_true", help="skip the measured 24-block token loop")
    ap.add_argument("--tokens", type=int, default=2, help="tokens run by the token loop (best one reported)")
    ap.add_argument("--tuned-weight", type=float, default=8.0, help="secondary split: G = ceil(sqrt(weight * D))")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--count-only", action="store_true",
                    help="print the number of kernel launches that precede the first timed region and exit "
                         "(ncu: -s that number captures the timed steps of the same command)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
