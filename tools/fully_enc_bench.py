#!/usr/bin/env python3
"""BASELINE config C5: the reference's test_fully_enc_bsgs.py flow (fully encrypted FFN blocks chained without
decryption, bootstrapping when fewer than 4 levels remain) on this build: random seeded weights with the reference's
magnitude calibration, verification against the float64 block after every block.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--D", type=int, default=2048)
    ap.add_argument("--F", type=int, default=4096)
    ap.add_argument("--num_blocks", type=int, default=24)
    ap.add_argument("--L0", type=int, default=36)
    ap.add_argument("--P", type=int, default=3)
    ap.add_argument("--N", type=int, default=16384)
    ap.add_argument("--no-bootstrap", action="store_true")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--weight", type=float, default=1.0, help="BSGS split G = ceil(sqrt(weight * D)); 1 = the reference's")
    ap.add_argument("--pageable-weights", action="store_true", help="keep the weight matrices in ordinary (pageable) host memory")
    ap.add_argument("--reserve-gb", type=int, default=40, help="device memory pool grown to this size in the warm-up")
    ap.add_argument("--no-warmup", action="store_true", help="time the very first block too (includes one-off allocations)")
    ap.add_argument("--phases", action="store_true", help="host-clock seconds per phase (adds synchronisations)")
    a = ap.parse_args()
    rank, world, local = 0, 1, int(os.environ.get("LOCAL_RANK", "0"))
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:      # torchrun: every mat-vec giant-step sharded over the ranks
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        rank, world = dist.get_rank(), dist.get_world_size()
    from fhe_spear_b200 import bsgs as hb
    from fhe_spear_b200 import pyPhantom as ph
    from fhe_spear_b200.ffn_block import fully_encrypted_ffn_block, plaintext_ffn_block
    D, F, L0 = a.D, a.F, a.L0
    if a.phases:
        from fhe_spear_b200 import ffn_block as fb
        fb.PHASES = {}
    np.random.seed(a.seed)                                          # [ref: test_fully_enc_bsgs.py:153, 171-201]
    W_keys = [np.random.randn(D, F) * 0.02 for _ in range(a.num_blocks)]
    W_raw = [np.random.randn(F, D) * 0.02 for _ in range(a.num_blocks)]
    x0 = np.random.randn(D) * 0.1
    W_vals, ref = [], [x0.copy()]
    x = x0.copy()
    for b in range(a.num_blocks):                                   # magnitude control folded into W_val
        fv = ((x @ W_keys[b]) ** 2) @ W_raw[b]
        ms = 1.0 / (np.max(np.abs(fv)) + 1e-12)
        W_vals.append(W_raw[b] * ms)
        x = plaintext_ffn_block(x, W_keys[b], W_vals[b])
        ref.append(x.copy())
    if not a.pageable_weights:
        # a server keeps its weights in page-locked memory: diagonal sets are then encoded from views of them by direct DMA
        def pinned(Wm):
            buf = ph.pinned_empty(Wm.shape, dtype=np.float64)
            buf[...] = Wm
            return buf
        W_keys, W_vals = [pinned(Wm) for Wm in W_keys], [pinned(Wm) for Wm in W_vals]
    t0 = time.perf_counter()
    ckks = hb.CKKSBootstrapContext(poly_degree=a.N, L0=L0, prime_bits=59, special_mod_size=a.P,
                                   level_budget=None if a.no_bootstrap else [2, 2], max_rot_dim=1, bsgs_dim=[D],
                                   skip_bootstrap=a.no_bootstrap, seed=bytes(range(32)), verbose=False,
                                   baby_weights=(a.weight,), device=local)
    split = hb.compute_bsgs_params(D, a.weight)
    setup_s = time.perf_counter() - t0
    ct = ckks.encrypt_replicated(x0)
    if not a.no_warmup:
        # untimed warm-up at the top level (the largest shapes of the run): grows the stream workspaces, the memory pool
        # and the page-locked staging buffer once, as any timed GPU measurement does before its first step
        t_w = time.perf_counter()
        ckks.ctx.reserve(a.reserve_gb << 30)
        fully_encrypted_ffn_block(ckks, ct, W_keys[0], W_vals[0], D, F, block_idx=0, split=split, shard=(rank, world))
        if not a.no_bootstrap:
            ckks.bootstrap(ct)
        ckks.ctx.synchronize()
        warmup_s = time.perf_counter() - t_w
    else:
        warmup_s = 0.0
    rows, boots, done = [], [], 0
    t_all = time.perf_counter()
    for b in range(a.num_blocks):
        remaining = (L0 - 1) - ct.chain_index()
        if remaining < 4:                                           # [ref: :238-266]
            if a.no_bootstrap:
                break
            ckks.ctx.synchronize()
            tb = time.perf_counter()
            ct = ph.rescale_to_next(ckks.ctx, ckks.bootstrap(ct))
            ckks.ctx.synchronize()
            tb = time.perf_counter() - tb
            err = float(np.abs(ckks.decrypt_vec(ct, D) - ref[b]).max())
            boots.append({"before_block": b, "seconds": tb, "chain_index_after": ct.chain_index(), "max_abs_err": err})
        ckks.ctx.synchronize()
        ph0 = dict(fb.PHASES) if a.phases else None
        tb = time.perf_counter()
        ct, used = fully_encrypted_ffn_block(ckks, ct, W_keys[b], W_vals[b], D, F, block_idx=b, split=split,
                                             shard=(rank, world))
        ckks.ctx.synchronize()
        tb = time.perf_counter() - tb
        got = ckks.decrypt_vec(ct, D)
        rows.append({"block": b, "seconds": tb, "levels_used": used, "chain_index": ct.chain_index(),
                     "phases": ({k: round(v - ph0.get(k, 0.0), 4) for k, v in fb.PHASES.items()} if a.phases else None),
                     "corr": float(np.corrcoef(got, ref[b + 1])[0, 1]), "max_abs_err": float(np.abs(got - ref[b + 1]).max())})
        done = b + 1
    total = time.perf_counter() - t_all
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total = float(t.item())
        dist.destroy_process_group()
        if rank != 0:
            return
    print(json.dumps({"what": "fully encrypted FFN blocks (reference test_fully_enc_bsgs.py flow)",
                      "config": {"D": D, "F": F, "N": a.N, "L0": L0, "P": a.P, "num_blocks": a.num_blocks,
                                 "bootstrap": not a.no_bootstrap, "split": f"G={split[0]} B={split[1]}", "n_gpus": world},
                      "blocks_completed": done, "bootstraps": len(boots), "total_s": total,
                      "s_per_block": float(np.mean([r["seconds"] for r in rows])) if rows else None,
                      "s_per_bootstrap": float(np.mean([r["seconds"] for r in boots])) if boots else None,
                      "final_corr": rows[-1]["corr"] if rows else None, "final_max_abs_err": rows[-1]["max_abs_err"] if rows else None,
                      "match": bool(rows and rows[-1]["corr"] > 0.999), "setup_s": setup_s, "warmup_s": warmup_s,
                      "phase_seconds_total": (fb.PHASES if a.phases else None), "blocks": rows, "boot": boots}))


if __name__ == "__main__":
    main()
