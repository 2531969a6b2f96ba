#!/usr/bin/env python3
"""Giant-step-sharded single mat-vec over the ranks of one node (second level):
torchrun --nproc-per-node N tools/sharded_latency.py.  Checks the result against the unsharded mat-vec
on every rank and prints the latency (CUDA events, max over ranks)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from fhe_spear_b200 import bsgs as hb
    from fhe_spear_b200 import pyPhantom as ph
    from fhe_spear_b200.sharding import ShardedMatvec
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("config", nargs="?", default="c3")
    ap.add_argument("--weight", type=float, default=0.0, help="split G = ceil(sqrt(weight * D)); 0 = 8 / world")
    a = ap.parse_args()
    cfg = a.config
    weight = a.weight if a.weight > 0 else max(1.0, 8.0 / world)
    N, L0, P, D = bench.CONFIGS[cfg]
    G, B = hb.compute_bsgs_params(D, weight)
    ckks = hb.CKKSBootstrapContext(poly_degree=N, L0=L0, prime_bits=59, special_mod_size=P, max_rot_dim=1, bsgs_dim=[D],
                                   skip_bootstrap=True, seed=bytes(range(32)), device=local, verbose=False,
                                   baby_weights=(weight,))
    rng = np.random.default_rng(7)                      # same matrix and input on every rank
    W, x = rng.standard_normal((D, D)) * 0.02, rng.standard_normal(D) * 0.1
    ct = ckks.encrypt_replicated(x)                     # same seed, same enc counter -> identical ciphertext
    mv = ShardedMatvec(ckks, W, D, baby_weight=weight)
    y = mv(ct)
    full = hb.pre_encode_real_diags(ckks, W, D, G, B, level=1)
    ref = ph.bsgs_hoisted(ckks.ctx, ct, full, ckks.gk)
    same = bool(np.array_equal(y.to_numpy(), ref.to_numpy()))
    err = float(np.abs(ckks.decrypt_vec(y, D) - W @ x).max())
    steps = 10
    import time
    from fhe_spear_b200.sharding import PeerExchange

    def latency():
        for _ in range(3):
            mv(ct)
        ckks.ctx.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            mv(ct)
        ckks.ctx.synchronize()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps * 1e3
    peer_on = PeerExchange.get(ckks.ctx) is not None
    dt = latency()                                      # fused peer-memory exchange (csrc/peer.cu) when available
    os.environ["SPEAR_PEER"] = "0"
    dt_nccl = latency()                                 # int64 NCCL all-reduce + Barrett pass
    same_nccl = bool(np.array_equal(mv(ct).to_numpy(), ref.to_numpy()))
    os.environ["SPEAR_PEER"] = "1"
    # the exchange alone, on the engine stream (CUDA events), accumulators already computed
    ex_ms = float("nan")
    if peer_on:
        acc = ph.bsgs_hoisted_partial(ckks.ctx, ct, mv.shard, ckks.gk)
        ex = PeerExchange.get(ckks.ctx)
        for _ in range(3):
            ex.allreduce(acc)
        ckks.ctx.synchronize()
        dist.barrier()
        ckks.ctx.timer_start()
        for _ in range(steps):
            ex.allreduce(acc)
        ex_ms = ckks.ctx.timer_stop() / steps
    # phase breakdown (host clock, synchronised after every phase)
    from fhe_spear_b200.sharding import _DevView, allreduce_residues
    ctx = ckks.ctx
    ph_ms = {"partial": 0.0, "allreduce": 0.0, "reduce_finish": 0.0}
    for _ in range(steps):
        dist.barrier()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        acc = ph.bsgs_hoisted_partial(ctx, ct, mv.shard, ckks.gk)
        ctx.synchronize()
        t2 = time.perf_counter()
        size, limbs, ext, ring, _, _ = acc._info()
        tt = torch.as_tensor(_DevView(ph.device_ptr(acc), size * (limbs + ctx.P) * ring), device=f"cuda:{ctx.device}")
        allreduce_residues(tt)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        ph.reduce_inplace(ctx, acc)
        ph.bsgs_finish(ctx, acc)
        ctx.synchronize()
        t4 = time.perf_counter()
        ph_ms["partial"] += (t2 - t1) * 1e3 / steps
        ph_ms["allreduce"] += (t3 - t2) * 1e3 / steps
        ph_ms["reduce_finish"] += (t4 - t3) * 1e3 / steps
    pt = torch.tensor([ph_ms["partial"], ph_ms["allreduce"], ph_ms["reduce_finish"]], dtype=torch.float64, device="cuda")
    dist.all_reduce(pt, op=dist.ReduceOp.MAX)
    t = torch.tensor([dt, dt_nccl, ex_ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = torch.tensor([1.0 if same and same_nccl else 0.0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"what": "giant-step-sharded single mat-vec latency", "config": cfg, "n_gpus": world,
                          "split": f"G={G} B={B}", "ms_per_matvec": float(t[0].item()),
                          "exchange": "fused peer-memory kernel (csrc/peer.cu)" if peer_on else "NCCL int64 all-reduce",
                          "ms_per_matvec_nccl_allreduce_path": float(t[1].item()),
                          "peer_exchange_ms_alone": float(t[2].item()),
                          "phase_ms_max_over_ranks": dict(zip(("partial", "allreduce", "reduce_finish"), [float(v) for v in pt.tolist()])), "bit_identical_to_unsharded_on_all_ranks": bool(ok.item() > 0.5),
                          "max_abs_err_vs_float64": err, "shard_diag_bytes": mv.shard.info()["bytes"]}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
