"""CKKS bootstrapping over the evaluator primitives of this package.

The reference reaches bootstrapping through `ph.ckks_bootstrapper(encoder)` with `setup(ctx, level_budget)`,
`keygen(ctx, sk)`, `bootstrap(ctx, ct)` and the static `get_galois_elements(N, slots, level_budget)`,
`get_bootstrap_depth(level_budget)` (reference scripts/bootstrap_generation.py:72-75, 110-116, 149-154).  Its
implementation lives in the absent phantom-fhe fork, so this is an independent design ("parity unpinned"): the
standard pipeline

    ModRaise -> CoeffToSlot -> EvalMod (Chebyshev cosine + double angle) -> SlotToCoeff

written only in terms of primitives that are each bit-exact against the oracle (hoisted rotations, plaintext
multiply, ciphertext multiply + relinearize, rescale, conjugation, the new `mod_raise`).  The same orchestration
runs on any object offering that call surface, which is how tests compare a whole bootstrap on the GPU library and
on the CPU oracle limb for limb.

Conventions.  Slot j of the encoder is the evaluation at zeta^(5^j); with n = N/2 and w = c_lo + i c_hi (the two
halves of the coefficient vector) the slot vector is z = U0 w, U0[j][k] = zeta^(5^j k).  U0 = F R with R the
bit-reversal permutation and F a product of log2(n) butterfly stages (three generalized diagonals each), so
CoeffToSlot applies F^-1 and SlotToCoeff applies F; the bit reversal cancels because EvalMod acts slot-wise.
Stages are merged into `level_budget` groups per direction and applied with a baby-step/giant-step over their
diagonals; every scalar factor of the pipeline is folded into the diagonals.
"""
import math

import numpy as np


# ---- linear maps as generalized diagonals: (M x)_i = sum_k d_k[i] * x[(i + k) mod n] -------------------------
def _s2c_stages(n, N):
    """Butterfly stages of F (applied in list order: len = 2, 4, ..., n)."""
    M = 2 * N
    rot = [pow(5, j, M) for j in range(max(1, n // 2))]
    out, ln = [], 2
    while ln <= n:
        lenh, lenq = ln // 2, ln * 4
        j = np.arange(lenh)
        ksi = np.exp(2j * np.pi * np.array([(rot[t] % lenq) * (M // lenq) for t in j]) / M)
        first = np.zeros(n, dtype=bool)
        first.reshape(-1, ln)[:, :lenh] = True
        ksi_t = np.tile(np.concatenate([ksi, ksi]), n // ln)
        d0 = np.where(first, 1.0, -ksi_t)
        dp = np.where(first, ksi_t, 0.0)
        dm = np.where(first, 0.0, 1.0)
        st = {}
        for k, d in ((0, d0), (lenh % n, dp), ((-lenh) % n, dm)):
            st[k] = st.get(k, 0) + d
        out.append(st)
        ln *= 2
    return out


def _c2s_stages(n, N):
    """Inverse butterflies, applied in list order: len = n, n/2, ..., 2 (their product is F^-1)."""
    M = 2 * N
    rot = [pow(5, j, M) for j in range(max(1, n // 2))]
    out, ln = [], n
    while ln >= 2:
        lenh, lenq = ln // 2, ln * 4
        j = np.arange(lenh)
        ksi = np.exp(2j * np.pi * np.array([(rot[t] % lenq) * (M // lenq) for t in j]) / M)
        first = np.zeros(n, dtype=bool)
        first.reshape(-1, ln)[:, :lenh] = True
        inv_t = np.tile(np.concatenate([1 / ksi, 1 / ksi]), n // ln)
        d0 = np.where(first, 0.5, -0.5 * inv_t)          # u = (a + b)/2 ; v = (a - b)/(2 ksi)
        dp = np.where(first, 0.5, 0.0)
        dm = np.where(first, 0.0, 0.5 * inv_t)
        st = {}
        for k, d in ((0, d0), (lenh % n, dp), ((-lenh) % n, dm)):
            st[k] = st.get(k, 0) + d
        out.append(st)
        ln //= 2
    return out


def _compose(A, B, n):
    """A after B."""
    out = {}
    for a, da in A.items():
        for b, db in B.items():
            k = (a + b) % n
            out[k] = out.get(k, 0) + da * np.roll(db, -a)
    return out


def _merge(stages, groups, n, scalar=1.0):
    """Merge consecutive stages into `groups` maps (applied in list order); `scalar` is spread over the groups."""
    groups = max(1, min(groups, len(stages)))
    sizes = [len(stages) // groups + (1 if i < len(stages) % groups else 0) for i in range(groups)]
    out, pos = [], 0
    f = abs(scalar) ** (1.0 / groups)
    for gi, sz in enumerate(sizes):
        m = stages[pos]
        for s in stages[pos + 1:pos + sz]:
            m = _compose(s, m, n)
        pos += sz
        fac = f * (np.sign(scalar) if gi == 0 and np.isreal(scalar) else 1.0)
        out.append({k: d * fac for k, d in m.items() if np.abs(d).max() > 1e-14})
    return out


def _bsgs_split(offsets, n):
    """Baby count N1 (a power of two times the offset stride) minimising #babies + #giants; returns (N1, babies, giants)."""
    offs = sorted(int(k) % n for k in offsets)
    best = None
    n1 = 1
    while n1 <= n:
        babies = sorted({k % n1 for k in offs})
        giants = sorted({k - k % n1 for k in offs})
        cost = len([b for b in babies if b]) + len([g for g in giants if g])
        if best is None or cost < best[0]:
            best = (cost, n1, babies, giants)
        n1 *= 2
    return best[1], best[2], best[3]


class LinearTransform:
    """One merged group of butterfly stages, applied as  sum_g rot_g( sum_j roll(d_{g+j}, g) * rot_j(x) )."""

    def __init__(self, diags, n):
        self.n, self.diags = n, diags
        self.n1, self.babies, self.giants = _bsgs_split(diags.keys(), n)
        self._pts = {}

    def rotation_steps(self):
        return sorted({b for b in self.babies if b} | {g for g in self.giants if g})

    def plaintexts(self, bt, l):
        """Pre-rotated diagonals encoded for a ciphertext with l limbs at the scale that makes the rescaled product land
        exactly on the ladder scale of level l-1."""
        if l not in self._pts:
            sc, ci = bt.S[l - 1] * float(bt.moduli[l - 1]) / bt.S[l], bt.L - l + 1
            self._pts[l] = {(g, j): bt.encoder.encode_complex_vector(bt.ctx, np.roll(self.diags[(g + j) % self.n], g), sc, ci)
                            for g in self.giants for j in self.babies if (g + j) % self.n in self.diags}
        return self._pts[l]

    def apply(self, bt, ct, rescale=True):
        ph, ctx = bt.ph, bt.ctx
        pts = self.plaintexts(bt, ct.coeff_modulus_size())
        steps = [b for b in self.babies if b]
        rotated = dict(zip(steps, ph.hoisting(ctx, ct, bt.gk, steps))) if steps else {}
        rotated[0] = ct
        acc = None
        for g in self.giants:
            inner = None
            for j in self.babies:
                if (g, j) not in pts:
                    continue
                t = ph.multiply_plain(ctx, rotated[j], pts[(g, j)])
                inner = t if inner is None else ph.add(ctx, inner, t)
            if inner is None:
                continue
            if g:
                inner = ph.rotate(ctx, inner, g, bt.gk)
            acc = inner if acc is None else ph.add(ctx, acc, inner)
        l = ct.coeff_modulus_size()
        if not rescale:
            acc.set_scale(bt.S[l - 1] * float(bt.moduli[l - 1]))   # the caller rescales
            return acc
        out = ph.rescale_to_next(ctx, acc)
        out.set_scale(bt.S[l - 1])
        return out


# ---- Chebyshev series on [-1, 1] --------------------------------------------------------------------------------
def _cheb_coeffs(f, tol=1e-13, max_deg=255):
    deg = 15
    while True:
        c = np.polynomial.chebyshev.chebinterpolate(f, deg)
        tail = np.abs(c[-4:]).max()
        if tail < tol or deg >= max_deg:
            break
        deg = deg * 2 + 1
    keep = len(c)
    while keep > 1 and abs(c[keep - 1]) < tol:
        keep -= 1
    return c[:keep]


def _cheb_divide(c, t):
    """c (Chebyshev coefficients, degree < 2t) = q * T_t + r with deg q, deg r < t."""
    c = np.asarray(c, dtype=float)
    r = np.zeros(t)
    q = np.zeros(max(1, len(c) - t))
    r[:min(t, len(c))] = c[:t]
    for k in range(t, len(c)):
        if k == t:
            q[0] += c[k]
        else:                       # T_k = 2 T_{k-t} T_t - T_{2t-k}
            q[k - t] += 2 * c[k]
            r[2 * t - k] -= c[k]
    return q, r


class _DryCt:
    def __init__(self, be, l, scale):
        self.be, self.l, self._s = be, l, scale

    def coeff_modulus_size(self):
        return self.l

    def chain_index(self):
        return self.be.L - self.l + 1

    def scale(self):
        return self._s

    def set_scale(self, s):
        self._s = s


class _DryRun:
    """Stand-in evaluator that only tracks levels and scales (Bootstrapper.depth_for)."""

    def __init__(self, N, L=64, P=1):
        self.N, self.L, self.P = N, L, P
        self.moduli = [(1 << 59) - 1] * (L + P)

    def encode_complex_vector(self, ctx, v, scale, chain_index=1):
        return _DryCt(self, self.L - chain_index + 1, scale)

    def multiply_plain(self, ctx, a, p):
        return _DryCt(self, a.l, a._s * p._s)

    def multiply(self, ctx, a, b):
        assert a.l == b.l
        return _DryCt(self, a.l, a._s * b._s)

    def relinearize(self, ctx, a, rlk):
        return a

    def add(self, ctx, a, b):
        assert a.l == b.l
        return _DryCt(self, a.l, a._s)

    sub = add

    def add_plain(self, ctx, a, p):
        return _DryCt(self, a.l, a._s)

    def rescale_to_next(self, ctx, a):
        return _DryCt(self, a.l - 1, a._s / self.moduli[a.l - 1])

    def mod_switch_to_next(self, ctx, a):
        return _DryCt(self, a.l - 1, a._s)

    def mod_switch_to(self, ctx, a, chain_index):
        return _DryCt(self, self.L - chain_index + 1, a._s)

    def mod_raise(self, ctx, a, chain_index=1):
        return _DryCt(self, self.L - chain_index + 1, a._s)

    def apply_galois(self, ctx, a, elt, gk):
        return a

    def rotate(self, ctx, a, step, gk):
        return a

    def hoisting(self, ctx, a, gk, steps):
        return [a for _ in steps]


class Bootstrapper:
    """Backend-agnostic CKKS bootstrapper.  `ph` supplies the evaluator functions (fhe_spear_b200.pyPhantom or a
    look-alike), `encoder` the CKKS encoder, `moduli` the data primes q_0..q_{L-1} followed by the special ones."""

    RATIO_BITS = 12     # the message is scaled down by 2^RATIO_BITS before ModRaise so that sin(x) ~ x holds
    CONST_CACHE_BYTES = 3 << 30

    def __init__(self, ph, ctx, encoder, N, moduli, special, level_budget=(2, 2), K=None, doublings=None,
                 ratio_bits=None):
        self.ph, self.ctx, self.encoder = ph, ctx, encoder
        self.N, self.n = int(N), int(N) // 2
        self.moduli = [int(q) for q in moduli]
        self.L = len(self.moduli) - int(special)
        self.budget = tuple(int(b) for b in level_budget)
        # |I| bound for a uniform ternary secret: the coefficients of c1*s/q0 have sigma = sqrt(N/18)
        self.K = int(K) if K else int(2 ** math.ceil(math.log2(6.5 * math.sqrt(self.N / 18.0) + 1)))
        # every double-angle step multiplies the error by up to 4: prefer a higher Chebyshev degree (one more level per
        # doubling of the degree) over doublings; 2 pi K / 2^r ~ 100 keeps the degree near 144
        self.r = int(doublings) if doublings is not None else max(0, int(round(math.log2(self.K))) - 4)
        a, phi = 2 * math.pi * self.K / 2 ** self.r, math.pi / 2 ** (self.r + 1)
        self.cheb = _cheb_coeffs(lambda u: np.cos(a * u - phi))
        q0 = float(self.moduli[0])
        self.c2s = [LinearTransform(d, self.n) for d in
                    _merge(_c2s_stages(self.n, self.N), self.budget[0], self.n, scalar=0.5 / self.K)]
        self.ratio_bits = int(ratio_bits) if ratio_bits is not None else self.RATIO_BITS
        self.post = 2.0 ** self.ratio_bits / (2 * math.pi)      # sin(2 pi y)/(2 pi) * q0/Delta'
        self.s2c = [LinearTransform(d, self.n) for d in
                    _merge(_s2c_stages(self.n, self.N), self.budget[1], self.n, scalar=self.post)]
        self.q0 = q0
        # scale ladder: a ciphertext with l limbs always has scale S[l]; S[l-1] = S[l]^2 / q_{l-1}, so the product of two
        # level-l ciphertexts lands exactly on the scale of level l-1 and every addition is between equal scales.  (The
        # primes differ from one another by ~1e-11 relative; forcing one nominal scale instead loses ~1e-9 per
        # bootstrap before the double-angle steps amplify it.)
        self.S = [0.0] * (self.L + 1)
        self.S[self.L] = q0
        for l in range(self.L, 1, -1):
            self.S[l - 1] = self.S[l] * self.S[l] / float(self.moduli[l - 1])
        self.gk = self.rlk = None
        self._consts, self._const_bytes = {}, 0

    # ---- static helpers of the reference interface
    @staticmethod
    def rotation_steps(N, level_budget=(2, 2)):
        n = N // 2
        steps = set()
        for stages, groups in ((_c2s_stages(n, N), level_budget[0]), (_s2c_stages(n, N), level_budget[1])):
            for d in _merge(stages, groups, n):
                steps.update(LinearTransform(d, n).rotation_steps())
        return sorted(steps)

    @staticmethod
    def depth_for(N, level_budget=(2, 2), **kw):
        """Levels a bootstrap spends after ModRaise, found by running the pipeline on level-counting stand-ins."""
        dry = _DryRun(N)
        bt = Bootstrapper(dry, dry, dry, N, dry.moduli, dry.P, level_budget, **kw)
        out = bt.bootstrap(_DryCt(dry, 2, 2.0 ** 59))
        return dry.L - out.coeff_modulus_size()

    @staticmethod
    def _baby_count(d):
        return 1 << max(1, int(math.ceil(math.log2(max(2, d + 1)) / 2)))

    # ---- level / scale discipline: a ciphertext with l limbs has exactly the scale self.S[l]
    def _drop(self, a, l):
        """Bring `a` down to l limbs and onto that level's ladder scale: plain limb drop to l+1, then one multiplication
        by the constant 1 encoded at S[l] q_l / S[la] and a rescale."""
        ph, ctx = self.ph, self.ctx
        la = a.coeff_modulus_size()
        if la == l:
            return a
        if la > l + 1:
            a = ph.mod_switch_to(ctx, a, self.L - l)
        one = self._const(1.0, l + 1, self.S[l] * float(self.moduli[l]) / self.S[la])
        out = ph.rescale_to_next(ctx, ph.multiply_plain(ctx, a, one))
        out.set_scale(self.S[l])
        return out

    def _align(self, a, b):
        l = min(a.coeff_modulus_size(), b.coeff_modulus_size())
        return self._drop(a, l), self._drop(b, l)

    def _mul(self, a, b):
        ph, ctx = self.ph, self.ctx
        a, b = self._align(a, b)
        out = ph.rescale_to_next(ctx, ph.relinearize(ctx, ph.multiply(ctx, a, b), self.rlk))
        out.set_scale(self.S[out.coeff_modulus_size()])          # S[l]^2 / q_{l-1} by construction of the ladder
        return out

    def _add(self, a, b, sub=False):
        a, b = self._align(a, b)
        return self.ph.sub(self.ctx, a, b) if sub else self.ph.add(self.ctx, a, b)

    def _const(self, value, l, scale):
        """Constant slot vector as a plaintext with l limbs.  The ladder makes (value, l, scale) repeat exactly from one
        bootstrap to the next, so the encodings are kept (bounded by CONST_CACHE_BYTES)."""
        key = (complex(value), int(l), float(scale))
        pt = self._consts.get(key)
        if pt is None:
            pt = self.encoder.encode_complex_vector(self.ctx, np.full(self.n, value, dtype=np.complex128), float(scale),
                                                    self.L - l + 1)
            self._const_bytes += 8 * l * self.N
            if self._const_bytes <= self.CONST_CACHE_BYTES:
                self._consts[key] = pt
        return pt

    def _add_const(self, a, value):
        return self.ph.add_plain(self.ctx, a, self._const(value, a.coeff_modulus_size(), a.scale()))

    def _double_minus_one(self, sq):
        """2*sq - 1"""
        return self._add_const(self.ph.add(self.ctx, sq, sq), -1.0)

    def _linear_combination(self, terms, const):
        """sum_k c_k * ct_k + const, one level below the deepest term.  Terms from higher levels are limb-dropped (scale
        unchanged) and their constants encoded at S[l-1] q_{l-1} / scale(ct_k), so every product has the same scale."""
        ph, ctx = self.ph, self.ctx
        terms = [(c, t) for c, t in terms if abs(c) > 1e-300]
        if not terms:
            return None
        lmin = min(t.coeff_modulus_size() for _, t in terms)
        target = self.S[lmin - 1] * float(self.moduli[lmin - 1])
        acc = None
        for c, t in terms:
            st = t.scale()
            if t.coeff_modulus_size() > lmin:
                t = ph.mod_switch_to(ctx, t, self.L - lmin + 1)
            p = ph.multiply_plain(ctx, t, self._const(c, lmin, target / st))
            acc = p if acc is None else ph.add(ctx, acc, p)
        acc.set_scale(target)
        if const:
            acc = ph.add_plain(ctx, acc, self._const(const, lmin, target))
        out = ph.rescale_to_next(ctx, acc)
        out.set_scale(self.S[lmin - 1])
        return out

    # ---- Chebyshev series by baby-step / giant-step (Paterson-Stockmeyer in the Chebyshev basis)
    def _eval_chebyshev(self, x, coeffs):
        d = len(coeffs) - 1
        m = self._baby_count(d)
        T = {1: x}
        for k in range(2, m + 1):                      # T_k from the two halves: depth ceil(log2 k)
            a, b = (k + 1) // 2, k // 2
            prod = self._mul(T[a], T[b])
            dbl = self.ph.add(self.ctx, prod, prod)
            T[k] = self._add_const(dbl, -1.0) if a == b else self._add(dbl, T[a - b], sub=True)
        t = m
        while 2 * t <= d:                               # giants T_2m, T_4m, ...
            T[2 * t] = self._double_minus_one(self._mul(T[t], T[t]))
            t *= 2

        def rec(c, t):
            """series with coefficients c of degree < 2t (t a giant index or m)"""
            c = np.asarray(c, dtype=float)
            if len(c) <= m:                             # leaf: degree < m
                out = self._linear_combination([(c[k], T[k]) for k in range(1, len(c))], c[0])
                return out, (c[0] if out is None else 0.0)
            while t >= len(c):
                t //= 2
            q, r = _cheb_divide(c, t)
            qc, q0 = rec(q, t // 2 if t > m else m)
            rc, r0 = rec(r, t // 2 if t > m else m)
            if qc is None:                              # quotient is a constant
                prod = self._linear_combination([(q0, T[t])], 0.0)
            else:
                if q0:
                    qc = self._add_const(qc, q0)
                prod = self._mul(qc, T[t])
            if rc is None:
                return (self._add_const(prod, r0) if r0 else prod), 0.0
            if r0:
                rc = self._add_const(rc, r0)
            return self._add(prod, rc), 0.0

        out, c0 = rec(coeffs, t)
        return self._add_const(out, c0) if c0 else out

    def _eval_mod(self, u):
        """u = y / K  ->  sin(2 pi y)"""
        c = self._eval_chebyshev(u, self.cheb)
        for _ in range(self.r):
            c = self._double_minus_one(self._mul(c, c))
        return c

    # ---- the pipeline
    def bootstrap(self, ct, rescale_last=True):
        """ct: at most ... two limbs.  rescale_last=False leaves the rescale of the last SlotToCoeff group to the caller
        (the reference's convention: its call site rescales the bootstrapped ciphertext, test_fully_enc_bsgs.py:251-253)."""
        ph, ctx = self.ph, self.ctx
        while ct.coeff_modulus_size() > 2:
            ct = ph.mod_switch_to_next(ctx, ct)
        scale_in = ct.scale()
        if ct.coeff_modulus_size() == 2:               # scale the message down by 2^-RATIO_BITS on the way to one limb
            q1 = float(self.moduli[1])
            ct = ph.rescale_to_next(ctx, ph.multiply_plain(ctx, ct, self._const(2.0 ** -self.ratio_bits, 2, q1)))
        else:
            raise RuntimeError("bootstrap: the ciphertext must still have two limbs (one is spent scaling the message down)")
        raised = ph.mod_raise(ctx, ct, 1)
        raised.set_scale(self.S[self.L])               # relabel (S[L] = q0): slots are now (m + q0 I)/q0
        x = raised
        for lt in self.c2s:                            # slots: (c_lo + i c_hi) / (2 K q0), bit-reversed order
            x = lt.apply(self, x)
        conj = ph.apply_galois(ctx, x, 2 * self.N - 1, self.gk)
        re = ph.add(ctx, x, conj)                      # c_lo / (K q0)
        im = ph.multiply_plain(ctx, ph.sub(ctx, x, conj), self._const(-1j, x.coeff_modulus_size(), 1.0))
        im.set_scale(x.scale())
        re, im = self._eval_mod(re), self._eval_mod(im)
        re, im = self._align(re, im)
        y = ph.add(ctx, re, ph.multiply_plain(ctx, im, self._const(1j, im.coeff_modulus_size(), 1.0)))
        y.set_scale(re.scale())
        for i, lt in enumerate(self.s2c):
            y = lt.apply(self, y, rescale=rescale_last or i + 1 < len(self.s2c))
        # relabel: value * label = msg * scale_in * label / q0 (times the pending prime when the rescale is left out)
        y.set_scale(scale_in * y.scale() / self.q0)
        return y
