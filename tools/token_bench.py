#!/usr/bin/env python3
"""Server time per RWKV-7 token through the client-aided block (BASELINE config C4, second half of the metric):
24 blocks x 8 BSGS projections (r, k, v | o | 2 complex-packed ffn_key | 2 conjugate-packed ffn_val), d=2048,
d_ffn=8192, CKKS N=32768 L0=24 P=3, random-init weights of the named shapes.  One block's eight diagonal sets
(14.5 GB) are pre-encoded and shared by all blocks (a 24-block model is 348 GB of diagonals: 8 GPUs x 43 GB).
Prints one JSON line: server / client milliseconds per token and the error against the float64 plaintext token."""
import argparse
import copy
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--embed_dim", type=int, default=2048)
    ap.add_argument("--ffn_dim", type=int, default=8192)
    ap.add_argument("--num_blocks", type=int, default=24)
    ap.add_argument("--num_tokens", type=int, default=2)
    ap.add_argument("--N", type=int, default=32768)
    ap.add_argument("--L0", type=int, default=24)
    ap.add_argument("--P", type=int, default=3)
    ap.add_argument("--plan", default="hybrid", choices=["hybrid", "giant"],
                    help="hybrid: mat-vecs of a phase dealt to rank groups, giant steps sharded inside a group; "
                         "giant: every mat-vec sharded by giant step over all ranks")
    ap.add_argument("--weight", type=float, default=0.0,
                    help="BSGS split G = ceil(sqrt(weight * D)); 1 = the reference's, 0 = hoisting-aware (8 / world)")
    a = ap.parse_args()
    from fhe_spear_b200 import bsgs as hb
    from fhe_spear_b200 import rwkv_block as rb
    rank, world, local = 0, 1, int(os.environ.get("LOCAL_RANK", "0"))
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
        # torchrun: every mat-vec is split by giant step over all ranks (each rank holds 1/world of every diagonal set);
        # all ranks run the same deterministic client code, so their input ciphertexts are identical
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        rank, world = dist.get_rank(), dist.get_world_size()
    D, F = a.embed_dim, a.ffn_dim
    H, S = max(1, D // 64), min(64, D)
    t0 = time.perf_counter()
    from fhe_spear_b200.sharding import HybridBlock
    hybrid = a.plan == "hybrid" and world > 1
    weight = a.weight if a.weight > 0 else hb.hoisting_weight(world)
    G, B = hb.compute_bsgs_params(D, weight)
    weights = HybridBlock.required_weights(world, D, F) if hybrid else (weight,)
    ckks = hb.CKKSBootstrapContext(poly_degree=a.N, L0=a.L0, prime_bits=59, special_mod_size=a.P, max_rot_dim=1,
                                   bsgs_dim=[D], skip_bootstrap=True, seed=bytes(range(32)), verbose=False, device=local,
                                   baby_weights=weights)
    base = rb.RWKVBlockWeights.random(D, F, H, S, block_idx=0, seed=0)
    if hybrid:
        pe = HybridBlock(ckks, base, D, F, rank, world)
    else:
        pe = hb.pre_encode_block(ckks, base, D, F, G=G, B=B, shard=(rank, world))
    ckks.ctx.synchronize()
    setup_s = time.perf_counter() - t0
    blocks = []
    for i in range(a.num_blocks):
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
    token, rows = 3, []
    for step in range(a.num_tokens):
        t0 = time.perf_counter()
        logits, xa, xf, st, tms = rb.generate_token_fhe(ckks, blocks, emb, head, ones, zeros, ones, zeros, token, xa, xf,
                                                        st, D, use_bsgs=True, preencoded_blocks=[pe] * len(blocks))
        wall = time.perf_counter() - t0
        ref, pxa, pxf, pst = rb.generate_token_plaintext(blocks, emb, head, ones, zeros, ones, zeros, token, pxa, pxf, pst, D)
        server = sum(v for tm in tms for k, v in tm.items() if k.startswith("server_"))
        client = sum(v for tm in tms for k, v in tm.items() if k.startswith("client_"))
        rows.append({"token": step, "server_ms": server * 1e3, "client_numpy_ms": client * 1e3, "wall_ms": wall * 1e3,
                     "max_abs_logit_err": float(np.abs(logits - ref).max()),
                     "logit_corr": float(np.corrcoef(logits, ref)[0, 1]),
                     "same_argmax": bool(int(np.argmax(logits)) == int(np.argmax(ref)))})
        token = int(np.argmax(ref))
    best = min(r["server_ms"] for r in rows)
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([best], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t.item())
        if rank != 0:
            dist.destroy_process_group()
            return
    print(json.dumps({"metric": "server ms per RWKV-7 token (client-aided, BSGS, pre-encoded diagonals)", "n_gpus": world,
                      "parallelism": ("1 GPU" if world == 1 else
                                      "hybrid: the mat-vecs of a phase dealt to rank groups " + str({k: v.groups for k, v in pe.plan.items()}) + ", giant steps sharded inside a group" if hybrid else
                                      "giant steps of every mat-vec sharded over all ranks; the mat-vecs of a block phase run together, one int64 all-reduce each"),
                      "config": {"embed_dim": D, "ffn_dim": F, "num_blocks": a.num_blocks, "N": a.N, "L0": a.L0, "P": a.P,
                                 "matvecs_per_token": 8 * a.num_blocks, "split": "per group size: G = ceil(sqrt(8 / size * D))" if hybrid else f"G={G} B={B} ({G + B - 2} rotations)",
                                 "note": "server_* timings as in the reference: they include client encode+encrypt and decrypt+decode of every projection"},
                      "server_ms_per_token": best, "tokens": rows, "setup_s": setup_s}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
