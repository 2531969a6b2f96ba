"""Client-aided RWKV-7 block over the BSGS mat-vec path: host-side mirror of the caller layer
(reference scripts/bootstrap_generation.py:662-1032: RWKVBlockWeights, client_aided_block, plaintext_block,
generate_token_fhe, generate_token_plaintext).

The server only ever runs the eight encrypted projections of a block (r, k, v | o | ffn_key x2 | ffn_val x2 when
F = 4D); everything non-linear happens on the client in float64 NumPy, exactly as in the reference.  Names,
arguments, return values and the `timings` keys are the reference's, so its drivers can import from here.
Additions: `RWKVBlockWeights.random` (checkpoints are not downloadable offline) and a batched server round for
r/k/v (`ph.bsgs_hoisted_batch`) when the block's diagonal sets are pre-encoded.
"""
import time

import numpy as np

from . import bsgs as hb
from . import pyPhantom as ph


class RWKVBlockWeights:
    """Per-block tensors with the reference's attribute names [ref: :662-716]; x @ W convention ([in, out])."""

    VEC = ("ln1_w", "ln1_b", "ln2_w", "ln2_b", "ln_x_w", "ln_x_b", "x_r", "x_k", "x_v", "x_g", "x_w", "x_a",
           "x_k_ffn", "k_k", "k_a", "w0", "a0", "v0")

    def __init__(self, tensors, block_idx, D, F, n_head, head_size):
        self.D, self.F, self.n_head, self.head_size, self.block_idx = D, F, n_head, head_size, block_idx
        for k, v in tensors.items():
            setattr(self, k, np.asarray(v, dtype=np.float64))

    @classmethod
    def random(cls, D, F, n_head, head_size, block_idx=0, seed=0, lora=(96, 96, 64, 128)):
        """Random-init weights of the named shapes: projections N(0, 0.02^2),
        LayerNorm weight 1 / bias 0, token-shift mixes U(0,1), low-rank adapters N(0, 0.01^2)."""
        rng = np.random.default_rng(seed)
        t = {k: np.zeros(D) for k in cls.VEC}
        for k in ("ln1_w", "ln2_w", "ln_x_w"):
            t[k] = np.ones(D)
        for k in ("x_r", "x_k", "x_v", "x_g", "x_w", "x_a", "x_k_ffn"):
            t[k] = rng.uniform(0.0, 1.0, D)
        t["k_k"], t["k_a"] = rng.uniform(0.5, 1.0, D), rng.uniform(0.5, 1.0, D)
        rw, ra, rv, rg = lora
        for name, rank in (("w", rw), ("a", ra), ("v", rv)):
            t[name + "1"] = rng.standard_normal((D, rank)) * 0.01
            t[name + "2"] = rng.standard_normal((rank, D)) * 0.01
        t["g1"], t["g2"] = rng.standard_normal((D, rg)) * 0.01, rng.standard_normal((rg, D)) * 0.01
        t["r_k"] = rng.standard_normal((n_head, head_size)) * 0.01
        for name in ("W_r", "W_k", "W_v", "W_o"):
            t[name] = rng.standard_normal((D, D)) * 0.02
        t["W_key_ffn"] = rng.standard_normal((D, F)) * 0.02
        t["W_val_ffn"] = rng.standard_normal((F, D)) * 0.02
        return cls(t, block_idx, D, F, n_head, head_size)


# ---- client-side float64 pieces  [ref: :719-740] --------------------------------------------------------
def layer_norm(x, weight, bias, eps=1e-5):
    return (x - x.mean()) / np.sqrt(x.var() + eps) * weight + bias


def group_norm(x, n_groups, weight, bias, eps=64e-5):
    g = x.reshape(n_groups, -1)
    g = (g - g.mean(axis=1, keepdims=True)) / np.sqrt(g.var(axis=1, keepdims=True) + eps)
    return g.reshape(-1) * weight + bias


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-np.clip(x, -500, 500)))


def _time_mix_inputs(block, x, x_prev_att):
    x_ln = layer_norm(x, block.ln1_w, block.ln1_b)
    delta = x_prev_att - x_ln
    mixed = {name: x_ln + delta * getattr(block, "x_" + name) for name in ("r", "k", "v", "g", "w", "a")}
    return x_ln, mixed


def _wkv_and_gate(block, mixed, r, k, v, state, v_first):
    """WKV-7 state update, GroupNorm, bonus term and output gate  [ref: :800-846] -- client side."""
    H, S = block.n_head, block.head_size
    heads = lambda t: t.reshape(H, S)
    w = sigmoid(block.w0 + np.tanh(mixed["w"] @ block.w1) @ block.w2)
    decay = np.exp(-np.exp(-0.5) * heads(w))
    a = heads(sigmoid(block.a0 + (mixed["a"] @ block.a1) @ block.a2))
    kk = heads(k) * heads(block.k_k)
    kk = kk / (np.linalg.norm(kk, axis=1, keepdims=True) + 1e-12)
    k_h = heads(k) * (1.0 + (a - 1.0) * heads(block.k_a))
    if block.block_idx == 0:
        v_first_out = v.copy()
    else:
        v = v + (v_first - v) * sigmoid(block.v0 + (mixed["v"] @ block.v1) @ block.v2)
        v_first_out = v_first
    v_h, r_h = heads(v), heads(r)
    # S <- S * decay (per column) + (S @ -kk) (kk*a)^T + v k^T ;  out = S @ r      (all heads at once)
    sa = np.einsum("hij,hj->hi", state, -kk)
    new_state = state * decay[:, None, :] + sa[:, :, None] * (kk * a)[:, None, :] + v_h[:, :, None] * k_h[:, None, :]
    wkv = np.einsum("hij,hj->hi", new_state, r_h).reshape(-1)
    wkv = group_norm(wkv, H, block.ln_x_w, block.ln_x_b)
    wkv = wkv + ((r_h * k_h * block.r_k).sum(axis=1, keepdims=True) * v_h).reshape(-1)
    gate = sigmoid(mixed["g"] @ block.g1) @ block.g2
    return wkv * gate, new_state, v_first_out


def _ffn_input(block, x, x_prev_ffn):
    x_ln = layer_norm(x, block.ln2_w, block.ln2_b)
    return x_ln, x_ln + (x_prev_ffn - x_ln) * block.x_k_ffn


def client_aided_block(ckks, block, x, x_prev_att, x_prev_ffn, state, v_first, use_bsgs=True,
                       preencoded_block=None, cpu_offloaded_block=None):
    """One RWKV-7 block with the 3 + 1 + 2 + 2 projections on the server  [ref: :756-899]."""
    if not use_bsgs:
        raise RuntimeError("only the BSGS projection path is built (the per-column path is ~30x slower, reference README.md:60)")
    D, F = block.D, block.F
    pe, cpu = preencoded_block, cpu_offloaded_block
    tm = {}
    if hasattr(pe, "rkv"):       # sharding.HybridBlock: the phases are served by rank groups
        return _client_aided_block_hybrid(block, pe, x, x_prev_att, x_prev_ffn, state, v_first)
    t0 = time.perf_counter()
    x_ln, mixed = _time_mix_inputs(block, x, x_prev_att)
    tm["client_mix"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    if pe and all(isinstance(pe[n], ph.diagonal_set) for n in "rkv"):
        # server round 1 as one batched call: three independent mat-vecs sharing the keys
        cts = [ckks.encrypt_replicated(mixed[n]) for n in "rkv"]
        if any(pe[n].shard[1] > 1 for n in "rkv"):
            from .sharding import sharded_matvec_batch
            outs = sharded_matvec_batch(ckks, cts, [pe[n] for n in "rkv"])
        else:
            outs = ph.bsgs_hoisted_batch(ckks.ctx, cts, [pe[n] for n in "rkv"], ckks.gk)
        r, k, v = (ckks.decrypt_vec(o, D) for o in outs)
    else:
        r, k, v = (hb.fhe_projection_bsgs(ckks, mixed[n], getattr(block, "W_" + n), D, D, n,
                                          preencoded_diags=[pe[n]] if pe else None,
                                          cpu_offloaded_diags=[cpu[n]] if cpu else None) for n in "rkv")
    tm["server_rkv"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    gated, new_state, v_first_out = _wkv_and_gate(block, mixed, r, k, v, state, v_first)
    tm["client_wkv_gate"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    att_out = hb.fhe_projection_bsgs(ckks, gated, block.W_o, D, D, "o", preencoded_diags=[pe["o"]] if pe else None,
                                     cpu_offloaded_diags=[cpu["o"]] if cpu else None)
    tm["server_wo"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    x = x + att_out
    x_ffn_ln, x_k_ffn = _ffn_input(block, x, x_prev_ffn)
    tm["client_ffn_prep"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    fk = hb.fhe_projection_bsgs(ckks, x_k_ffn, block.W_key_ffn, D, F, "ffn_key",
                                preencoded_diags=pe.get("ffn_key") if pe else None,
                                cpu_offloaded_diags=cpu.get("ffn_key") if cpu else None)
    tm["server_ffn_key"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    fk_sq = np.maximum(fk, 0.0) ** 2
    tm["client_relu_sq"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    v_ffn = hb.fhe_projection_bsgs(ckks, fk_sq, block.W_val_ffn, F, D, "ffn_val",
                                   preencoded_diags=pe.get("ffn_val") if pe else None,
                                   cpu_offloaded_diags=cpu.get("ffn_val") if cpu else None)
    tm["server_ffn_val"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    x = x + v_ffn
    tm["client_residual"] = time.perf_counter() - t0
    return x, x_ln, x_ffn_ln, new_state, v_first_out, tm


def _client_aided_block_hybrid(block, server, x, x_prev_att, x_prev_ffn, state, v_first):
    """Same flow as client_aided_block with the four server rounds answered by a sharding.HybridBlock."""
    tm = {}
    t0 = time.perf_counter()
    x_ln, mixed = _time_mix_inputs(block, x, x_prev_att)
    tm["client_mix"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    r, k, v = server.rkv(mixed)
    tm["server_rkv"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    gated, new_state, v_first_out = _wkv_and_gate(block, mixed, r, k, v, state, v_first)
    tm["client_wkv_gate"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    att_out = server.o(gated)
    tm["server_wo"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    x = x + att_out
    x_ffn_ln, x_k_ffn = _ffn_input(block, x, x_prev_ffn)
    tm["client_ffn_prep"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    fk = server.ffn_key(x_k_ffn)
    tm["server_ffn_key"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    fk_sq = np.maximum(fk, 0.0) ** 2
    tm["client_relu_sq"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    v_ffn = server.ffn_val(fk_sq)
    tm["server_ffn_val"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    x = x + v_ffn
    tm["client_residual"] = time.perf_counter() - t0
    return x, x_ln, x_ffn_ln, new_state, v_first_out, tm


def plaintext_block(block, x, x_prev_att, x_prev_ffn, state, v_first):
    """float64 oracle of the same block  [ref: :902-980]"""
    x_ln, mixed = _time_mix_inputs(block, x, x_prev_att)
    r, k, v = mixed["r"] @ block.W_r, mixed["k"] @ block.W_k, mixed["v"] @ block.W_v
    gated, new_state, v_first_out = _wkv_and_gate(block, mixed, r, k, v, state, v_first)
    x = x + gated @ block.W_o
    x_ffn_ln, x_k_ffn = _ffn_input(block, x, x_prev_ffn)
    x = x + (np.maximum(x_k_ffn @ block.W_key_ffn, 0.0) ** 2) @ block.W_val_ffn
    return x, x_ln, x_ffn_ln, new_state, v_first_out


def generate_token_fhe(ckks, blocks, emb, head_w, ln_out_w, ln_out_b, ln0_w, ln0_b, token_id, x_prevs_att,
                       x_prevs_ffn, states, D, use_bsgs=True, preencoded_blocks=None, cpu_offloaded_blocks=None):
    """[ref: :983-1011]"""
    x = layer_norm(emb[token_id].copy(), ln0_w, ln0_b)
    xpa, xpf, sts, tms = [], [], [], []
    v_first = None
    for i, block in enumerate(blocks):
        x, a, f, st, v_first, tm = client_aided_block(
            ckks, block, x, x_prevs_att[i], x_prevs_ffn[i], states[i], v_first, use_bsgs=use_bsgs,
            preencoded_block=preencoded_blocks[i] if preencoded_blocks else None,
            cpu_offloaded_block=cpu_offloaded_blocks[i] if cpu_offloaded_blocks else None)
        xpa.append(a), xpf.append(f), sts.append(st), tms.append(tm)
    return layer_norm(x, ln_out_w, ln_out_b) @ head_w, xpa, xpf, sts, tms


def generate_token_plaintext(blocks, emb, head_w, ln_out_w, ln_out_b, ln0_w, ln0_b, token_id, x_prevs_att,
                             x_prevs_ffn, states, D):
    """[ref: :1014-1032]"""
    x = layer_norm(emb[token_id].copy(), ln0_w, ln0_b)
    xpa, xpf, sts = [], [], []
    v_first = None
    for i, block in enumerate(blocks):
        x, a, f, st, v_first = plaintext_block(block, x, x_prevs_att[i], x_prevs_ffn[i], states[i], v_first)
        xpa.append(a), xpf.append(f), sts.append(st)
    return layer_norm(x, ln_out_w, ln_out_b) @ head_w, xpa, xpf, sts
