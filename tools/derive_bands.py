"""Error-band derivation for the fp32 fast path of K1 (fwd DCT+quant) and K2 (dequant+IDCT).

The fast kernels compute the 8x8 transform with a scaled 30-op butterfly in fp32 and round.
The reference (src/dct.c, src/quantization.c) computes in fp64.  The integer results are
identical whenever the fp32 value is farther from a .5 rounding boundary than the worst-case
fp32 error; everything inside that band is replayed in fp64 (K3).  This script derives that
worst-case error RIGOROUSLY by pushing (linear functional, error bound) pairs through the very
same flowgraph the CUDA code executes (dct_b200/csrc/butterfly.cuh mirrors `fdct8` / `idct8`
below op for op), and writes dct_b200/csrc/band_tables.h.

    python tools/derive_bands.py            # regenerate the header
    python tools/derive_bands.py --check    # empirical check of the bound with an fp32 emulation

(The byte loader of fast_core.cuh feeds the first two forward stages with a 2^15 bias per sample; every value there is
an integer below 2^19, so those operations are exact with or without it and the model below, which sees the centred
samples, describes the same values.)

Model of one fp32 operation (round-to-nearest, u = 2^-24):
    fl(a+b)   = (a+b)(1+d), |d| <= u ; exact when both are error-free integers and |a+b| < 2^24
    fl(a*c+b) = (a*c+b)(1+d)          (fused, single rounding; c is a float32 constant)
"""
import argparse
import math
import os

import numpy as np

U = 2.0 ** -24
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# butterfly constants (exact reals; the CUDA code holds their float32 roundings)
C4 = math.cos(math.pi / 4)             # 0.70710678
C382 = math.cos(3 * math.pi / 8)       # 0.38268343
C541 = math.sqrt(2) * math.cos(3 * math.pi / 8)   # 0.54119610
C1306 = math.sqrt(2) * math.cos(math.pi / 8)      # 1.30656296
SQRT2 = math.sqrt(2.0)
C1847 = 2 * math.cos(math.pi / 8)      # 1.84775907
C1082 = 2 * (math.cos(math.pi / 8) - math.cos(3 * math.pi / 8))   # 1.08239220
C2613 = 2 * (math.cos(math.pi / 8) + math.cos(3 * math.pi / 8))   # 2.61312593
AAN = np.array([1.0] + [math.cos(k * math.pi / 16) * math.sqrt(2) for k in range(1, 8)])


def fdct8(o, x):
    """Scaled forward 8-point DCT, 30 operations.  out[k] = AAN[k] * sqrt(8) * (orthonormal DCT)[k]."""
    s07, d07 = o.add(x[0], x[7]), o.sub(x[0], x[7])
    s16, d16 = o.add(x[1], x[6]), o.sub(x[1], x[6])
    s25, d25 = o.add(x[2], x[5]), o.sub(x[2], x[5])
    s34, d34 = o.add(x[3], x[4]), o.sub(x[3], x[4])
    e0, e3 = o.add(s07, s34), o.sub(s07, s34)
    e1, e2 = o.add(s16, s25), o.sub(s16, s25)
    y0, y4 = o.add(e0, e1), o.sub(e0, e1)
    t = o.add(e2, e3)
    y2, y6 = o.fma(t, C4, e3), o.fma(t, -C4, e3)
    a, b, c = o.add(d34, d25), o.add(d25, d16), o.add(d16, d07)
    z5 = o.mul(o.sub(a, c), C382)
    z2, z4 = o.fma(a, C541, z5), o.fma(c, C1306, z5)
    z11, z13 = o.fma(b, C4, d07), o.fma(b, -C4, d07)
    y5, y3 = o.add(z13, z2), o.sub(z13, z2)
    y1, y7 = o.add(z11, z4), o.sub(z11, z4)
    return [y0, y1, y2, y3, y4, y5, y6, y7]


def idct8(o, v):
    """Scaled inverse 8-point DCT, 30 operations; inputs pre-multiplied by AAN[k]/sqrt(8)."""
    t10, t11 = o.add(v[0], v[4]), o.sub(v[0], v[4])
    t13 = o.add(v[2], v[6])
    t12n = o.fma(o.sub(v[2], v[6]), -SQRT2, t13)        # t13 - d*sqrt2  (= -t12; no operand negation needed)
    e0, e3 = o.add(t10, t13), o.sub(t10, t13)
    e1, e2 = o.sub(t11, t12n), o.add(t11, t12n)
    z13, z10 = o.add(v[5], v[3]), o.sub(v[5], v[3])
    z11, z12 = o.add(v[1], v[7]), o.sub(v[1], v[7])
    t7 = o.add(z11, z13)
    zd = o.sub(z11, z13)
    z5 = o.mul(o.add(z10, z12), C1847)
    t10o = o.fma(z12, -C1082, z5)
    t12o = o.fma(z10, -C2613, z5)
    t6 = o.sub(t12o, t7)
    t5n = o.fma(zd, -SQRT2, t6)                          # t6 - zd*sqrt2  (= -t5)
    t4 = o.add(t10o, t5n)
    return [o.add(e0, t7), o.add(e1, t6), o.sub(e2, t5n), o.add(e3, t4),
            o.sub(e3, t4), o.add(e2, t5n), o.sub(e1, t6), o.sub(e0, t7)]


# ---------------------------------------------------------------- numeric back-ends
class F64Ops:
    def add(self, a, b): return a + b
    def sub(self, a, b): return a - b
    def mul(self, a, c): return a * c
    def fma(self, a, c, b): return a * c + b
    def fms(self, a, c, b): return a * c - b


class F32Ops:
    """numpy float32 emulation; fma via float64 (product exact, one extra rounding, harmless here)."""
    def add(self, a, b): return np.float32(a) + np.float32(b)
    def sub(self, a, b): return np.float32(a) - np.float32(b)
    def mul(self, a, c): return np.float32(a) * np.float32(c)
    def fma(self, a, c, b):
        return (np.float64(a) * np.float64(np.float32(c)) + np.float64(b)).astype(np.float32)
    def fms(self, a, c, b):
        return (np.float64(a) * np.float64(np.float32(c)) - np.float64(b)).astype(np.float32)


class Node:
    """One value of the flowgraph: its exact linear functional over the inputs, a bound on its accumulated error,
    the operands it was computed from (with the float32 multipliers the CUDA code really uses) and the bound on the
    error INJECTED by this operation alone."""
    __slots__ = ("L", "E", "isint", "parents", "inj", "idx")


class BoundOps:
    """Error analysis of the flowgraph for inputs bounded by `mag` (per-input |x| bounds).

    Every operation computes fl(f(operands)) = f(operands) (1 + d), |d| <= u, on the COMPUTED operands.  The flowgraph is
    linear, so the error of an output is exactly
        sum over operations n of   T(out, n) * injected_n      +   sum over inputs k of T(out, k) * input_error_k
    where T(out, n) is the transfer coefficient from node n to the output (the product of the float32 multipliers along
    the paths, WITH their signs: cancellations between paths are real) and
        |injected_n| <= u * (|exact value| + |accumulated error|)  [+ |c32 - c| * |operand|  for a rounded constant].
    `E` carries the coarse bound of the accumulated error (absolute values path by path, no cancellation): it is only
    used inside the parentheses above, i.e. for the second-order term, so its looseness costs nothing.
    `transfer_error(out)` evaluates the sum with |T| and the bounds: first-order exact, second-order covered.
    (Round 1 used `E` itself as the bound; it over-estimates low-frequency gains up to 22-fold because it cannot see
    that the errors an input or an early rounding sends down two paths cancel exactly as the signal does.)"""

    def __init__(self, mag):
        self.mag = np.asarray(mag, dtype=np.float64)
        self.nodes = []

    def M(self, L):
        return float(np.sum(np.abs(L) * self.mag))

    def _new(self, L, E, isint, parents, inj):
        n = Node()
        n.L, n.E, n.isint, n.parents, n.inj, n.idx = L, E, isint, parents, inj, len(self.nodes)
        self.nodes.append(n)
        return n

    def inp(self, L, E, isint):
        return self._new(L, E, isint, [], E)

    def _addsub(self, a, b, sign):
        L = a.L + sign * b.L
        if a.isint and b.isint and a.E == 0.0 and b.E == 0.0 and self.M(L) < 2.0 ** 24:
            return self._new(L, 0.0, True, [(a, 1.0), (b, sign)], 0.0)
        e = a.E + b.E
        inj = U * (self.M(L) + e)
        return self._new(L, e + inj, False, [(a, 1.0), (b, sign)], inj)

    def add(self, a, b): return self._addsub(a, b, 1.0)
    def sub(self, a, b): return self._addsub(a, b, -1.0)

    def mul(self, a, c):
        c32 = float(np.float32(c))
        L = a.L * c
        e = abs(c32) * a.E + abs(c32 - c) * self.M(a.L)
        inj = abs(c32 - c) * self.M(a.L) + U * (self.M(L) + e)
        return self._new(L, e + U * (self.M(L) + e), False, [(a, c32)], inj)

    def _fma(self, a, c, b, sign):
        c32 = float(np.float32(c))
        L = a.L * c + sign * b.L
        e = abs(c32) * a.E + abs(c32 - c) * self.M(a.L) + b.E
        inj = abs(c32 - c) * self.M(a.L) + U * (self.M(L) + e)
        return self._new(L, e + U * (self.M(L) + e), False, [(a, c32), (b, sign)], inj)

    def fma(self, a, c, b): return self._fma(a, c, b, 1.0)
    def fms(self, a, c, b): return self._fma(a, c, b, -1.0)

    def transfer_error(self, out):
        """Bound on |computed - exact| of node `out`: sum_n |T(out, n)| * inj_n (inputs carry their own error as inj)."""
        adj = np.zeros(len(self.nodes))
        adj[out.idx] = 1.0
        total = 0.0
        for n in reversed(self.nodes[:out.idx + 1]):
            t = adj[n.idx]
            if t == 0.0:
                continue
            total += abs(t) * n.inj
            for p, c in n.parents:
                adj[p.idx] += t * c
        return total


def two_d(o, f, X, rows_first):
    """X: 8x8 list-of-lists of numbers.  rows_first: transform each row, then each column."""
    if rows_first:
        T = [f(o, X[i]) for i in range(8)]
        cols = [f(o, [T[i][j] for i in range(8)]) for j in range(8)]
        return [[cols[j][i] for j in range(8)] for i in range(8)]
    cols = [f(o, [X[i][j] for i in range(8)]) for j in range(8)]
    T = [[cols[j][i] for j in range(8)] for i in range(8)]
    return [f(o, T[i]) for i in range(8)]


def fwd_tables(float_pixels=False):
    """beta[k]: bound on |fp32 coefficient - exact coefficient| + |exact coefficient| * u,
    in orthonormal-coefficient units (the K1 band is beta[k] / Q[k], see band_tables.h).
    float_pixels: the inputs are fl32(p - 128) of arbitrary floats p in [0, 255] -- not integers,
    each already carrying a rounding error <= 128 u -- instead of exact integers."""
    o = BoundOps(np.full(64, 128.0))
    if float_pixels:
        X = [[o.inp(np.eye(64)[8 * i + j], 128.0 * U, False) for j in range(8)] for i in range(8)]
    else:
        X = [[o.inp(np.eye(64)[8 * i + j], 0.0, True) for j in range(8)] for i in range(8)]
    Y = two_d(o, fdct8, X, rows_first=True)
    S = 8.0 * np.outer(AAN, AAN)
    beta, cmax = np.zeros(64), np.zeros(64)
    for u in range(8):
        for v in range(8):
            k = 8 * u + v
            cmax[k] = o.M(Y[u][v].L) / S[u, v]
            beta[k] = o.transfer_error(Y[u][v]) / S[u, v] + cmax[k] * U
    return beta, cmax, S.ravel()


def inv_tables():
    """G[k]: error gain.  |fp32 pixel - exact pixel| <= u * sum_k G[k] * |v_k| where v_k are the
    prescaled dequantized inputs of the butterfly (each assumed to carry a relative error <= 8u:
    int16 -> float exact, adaptive scale 4u, two roundings and the table entry's own rounding)."""
    G = np.zeros(64)
    for k in range(64):
        mag = np.zeros(64)
        mag[k] = 1.0
        o = BoundOps(mag)
        X = [[o.inp(np.eye(64)[8 * i + j], 8 * U * mag[8 * i + j], False) for j in range(8)] for i in range(8)]
        Y = two_d(o, idct8, X, rows_first=False)
        G[k] = max(o.transfer_error(Y[i][j]) for i in range(8) for j in range(8)) / U
    P = np.outer(AAN, AAN).ravel() / 8.0
    return G, P


def emit(path):
    beta, cmax, S = fwd_tables()
    beta_f, _, _ = fwd_tables(float_pixels=True)
    G, P = inv_tables()

    def arr(name, a, fmt="%.9ef"):
        body = ",\n    ".join(", ".join(fmt % v for v in a[i:i + 8]) for i in range(0, 64, 8))
        return f"static const float {name}[64] = {{\n    {body}}};\n"

    with open(path, "w") as f:
        f.write("// GENERATED by tools/derive_bands.py -- do not edit.\n"
                "// Worst-case fp32 error tables of the scaled butterflies in butterfly.cuh.\n"
                "//   kFwdBeta[k] : |fp32 coef - exact coef| + |coef|max * 2^-24, orthonormal units\n"
                "//   kFwdBetaF32[k]: the same for float pixel tiles (inputs fl32(p - 128), p in [0, 255])\n"
                "//   kFwdCmax[k] : max |orthonormal coefficient k| over u8 blocks\n"
                "//   kFwdScale[k]: 8*a_u*a_v, the butterfly's output scale (double precision below)\n"
                "//   kInvGain[k] : |fp32 pixel - exact pixel| <= 2^-24 * sum_k kInvGain[k]*|v_k|\n"
                "//   kInvPrescale[k]: a_u*a_v/8, folded into the dequantisation multiplier\n"
                "#pragma once\n")
        f.write(arr("kFwdBeta", beta))
        f.write(arr("kFwdBetaF32", beta_f))
        f.write(arr("kFwdCmax", cmax))
        f.write(arr("kInvGain", G))
        d = lambda name, a: (f"static const double {name}[64] = {{\n    " + ",\n    ".join(
            ", ".join("%.17g" % v for v in a[i:i + 8]) for i in range(0, 64, 8)) + "};\n")
        f.write(d("kFwdScale", S))
        f.write(d("kInvPrescale", P))
    print("wrote", path)
    print("fwd beta max %.3e  (cmax max %.1f)   inv gain max %.2f" % (beta.max(), cmax.max(), G.max()))


def check(nblocks=20000):
    """Empirical: the fp32 emulation's error never exceeds the bound (and shows the slack)."""
    D = np.array([[(1 / math.sqrt(8) if i == 0 else 0.5) * math.cos(math.pi * (2 * j + 1) * i / 16)
                   for j in range(8)] for i in range(8)])
    beta, cmax, S = fwd_tables()
    rng = np.random.default_rng(0)
    X = rng.integers(0, 256, size=(nblocks, 8, 8)).astype(np.float64) - 128.0
    # adversarial: sign patterns of every basis function at full amplitude
    basis = np.einsum("ui,vj->uvij", D, D).reshape(64, 8, 8)
    adv = np.where(basis >= 0, 127.0, -128.0)
    X = np.concatenate([X, adv, -adv - 1.0])
    exact = np.einsum("ui,nij,vj->nuv", D, X, D)
    o32 = F32Ops()
    Xl = [[X[:, i, j].astype(np.float32) for j in range(8)] for i in range(8)]
    Y = two_d(o32, fdct8, Xl, rows_first=True)
    Y = np.stack([np.stack(r, -1) for r in Y], -2).astype(np.float64)   # n,u,v
    err = np.abs(Y / S.reshape(8, 8) - exact).reshape(-1, 64).max(0)
    print("fwd: max fp32 error %.3e, bound min/max %.3e/%.3e, worst ratio err/bound %.3f"
          % (err.max(), beta.min(), beta.max(), (err / beta).max()))
    assert np.all(err <= beta)
    G, P = inv_tables()
    Cq = rng.integers(-300, 300, size=(nblocks, 8, 8)).astype(np.float64) * rng.random((nblocks, 8, 8))
    V = (Cq * P.reshape(8, 8)).astype(np.float32)
    exact = np.einsum("ui,nuv,vj->nij", D, V.astype(np.float64) / P.reshape(8, 8), D)
    Vl = [[V[:, i, j] for j in range(8)] for i in range(8)]
    Xr = two_d(o32, idct8, Vl, rows_first=False)
    Xr = np.stack([np.stack(r, -1) for r in Xr], -2).astype(np.float64)
    bound = U * (np.abs(V.astype(np.float64)).reshape(-1, 64) @ G)
    err = np.abs(Xr - exact).reshape(-1, 64).max(1)
    print("inv: max fp32 error %.3e, worst ratio err/bound %.3f" % (err.max(), (err / bound).max()))
    assert np.all(err <= bound)


def check_inverse_sparse(trials=2000, seed=5):
    """The inverse gains one at a time: a single coefficient (then three) at random magnitudes.  Dense random inputs
    average the gains out; these inputs stress each kInvGain[k] on its own.  Returns the worst err / bound."""
    D = np.array([[(1 / math.sqrt(8) if i == 0 else 0.5) * math.cos(math.pi * (2 * j + 1) * i / 16)
                   for j in range(8)] for i in range(8)])
    G, P = inv_tables()
    rng = np.random.default_rng(seed)
    o32 = F32Ops()

    def ratio(V):
        Vl = [[V[:, 8 * i + j] for j in range(8)] for i in range(8)]
        Xr = two_d(o32, idct8, Vl, rows_first=False)
        Xr = np.stack([np.stack(r, -1) for r in Xr], -2).astype(np.float64)
        exact = np.einsum("ui,nuv,vj->nij", D, (V.astype(np.float64) / P).reshape(-1, 8, 8), D)
        bound = U * (np.abs(V.astype(np.float64)) @ G)
        return float((np.abs(Xr - exact).reshape(-1, 64).max(1) / np.maximum(bound, 1e-300)).max())

    worst = 0.0
    for k in range(64):
        V = np.zeros((trials, 64), np.float32)
        V[:, k] = (rng.random(trials) * 2000 - 1000).astype(np.float32)
        worst = max(worst, ratio(V))
    V = np.zeros((20 * trials, 64), np.float32)
    for _ in range(3):
        idx = rng.integers(0, 64, V.shape[0])
        V[np.arange(V.shape[0]), idx] = (rng.random(V.shape[0]) * 2000 - 1000).astype(np.float32)
    worst = max(worst, ratio(V))
    print("inv, single-coefficient and sparse inputs: worst ratio err/bound %.3f" % worst)
    assert worst <= 1.0
    return worst


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    if args.check:
        check()
        check_inverse_sparse()
    else:
        emit(os.path.join(ROOT, "dct_b200", "csrc", "band_tables.h"))
