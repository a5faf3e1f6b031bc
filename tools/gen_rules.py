#!/usr/bin/env python
"""Generate the nested 1-D quadrature tables used by the Smolyak builder.

Genz-Keister (1996) nested extensions of Gauss-Hermite for the standard-normal
weight phi(x) = exp(-x^2/2)/sqrt(2 pi): 1 -> 3 -> 9 -> 19 -> 35 points with
polynomial exactness 1, 5, 15, 29, 51.  The reference reaches these rules through
`SparseQuadratureGrids.GenzKeister` (reference test/runtests.jl:42); that package is
not vendored, so the tables are re-derived here from the defining property
(Kronrod/Patterson-style extension):

  given the n nodes of the previous rule, P_n(x) = prod (x - x_i), add m new nodes =
  roots of the monic polynomial E_m with   int phi(x) P_n(x) E_m(x) x^k dx = 0, k<m,
  then take interpolatory weights on all n+m nodes (moment matching).

Everything is done in 120-digit mpmath arithmetic, nodes are refined by Newton, the
exactness degree of every rule is verified, and the output is written as C hex-float
literals so the oracle and the CUDA library hold bit-identical tables.

Also emits the nested Gauss-Kronrod-Patterson rule for the uniform weight on [-1,1]
(1 -> 3 -> 7 -> 15 -> 31 -> 63, `SparseQuadratureGrids.KronrodPatterson`,
reference test/runtests.jl:42) by the same construction.

usage: python tools/gen_rules.py <out.h> [<out2.h> ...]
"""
import sys
import mpmath as mp

mp.mp.dps = 120


def gauss_moment(k):
    # int x^k phi(x) dx = (k-1)!! for even k
    if k % 2:
        return mp.mpf(0)
    r = mp.mpf(1)
    for j in range(k - 1, 0, -2):
        r *= j
    return r


def unif_moment(k):
    # (1/2) int_{-1}^{1} x^k dx  (probability-normalised so the weights sum to 1)
    if k % 2:
        return mp.mpf(0)
    return mp.mpf(1) / (k + 1)


def polymul(a, b):
    r = [mp.mpf(0)] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            r[i + j] += x * y
    return r


def polyval(c, x):  # ascending coefficients
    r = mp.mpf(0)
    for a in reversed(c):
        r = r * x + a
    return r


def polyder(c):
    return [i * c[i] for i in range(1, len(c))]


def extend(nodes, m, moment):
    """Return the m new nodes extending the symmetric node set `nodes`."""
    P = [mp.mpf(1)]
    for x in nodes:
        P = polymul(P, [-x, mp.mpf(1)])
    # E(x) = x^m + sum_{j<m} e_j x^j ; orthogonality against x^k, k<m
    A = mp.matrix(m, m)
    b = mp.matrix(m, 1)
    for k in range(m):
        for j in range(m):
            A[k, j] = sum(P[i] * moment(i + j + k) for i in range(len(P)))
        b[k] = -sum(P[i] * moment(i + m + k) for i in range(len(P)))
    e = mp.lu_solve(A, b)
    E = [e[j] for j in range(m)] + [mp.mpf(1)]
    roots = mp.polyroots(list(reversed(E)), maxsteps=2000, extraprec=2000)
    out = []
    dE = polyder(E)
    for r in roots:
        if abs(mp.im(r)) > mp.mpf(10) ** (-60):
            raise RuntimeError("extension has complex roots: %s" % r)
        x = mp.re(r)
        for _ in range(8):  # Newton polish
            x = x - polyval(E, x) / polyval(dE, x)
        out.append(x)
    return out


def weights(nodes, moment):
    n = len(nodes)
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for k in range(n):
        for j, x in enumerate(nodes):
            A[k, j] = x ** k
        b[k] = moment(k)
    w = mp.lu_solve(A, b)
    return [w[j] for j in range(n)]


def exactness(nodes, w, moment, upto):
    deg = -1
    for k in range(upto + 1):
        q = sum(wi * x ** k for wi, x in zip(w, nodes))
        if abs(q - moment(k)) > mp.mpf(10) ** (-80) * (1 + abs(moment(k))):
            break
        deg = k
    return deg


def symmetrise(new):
    """Force exact +-pairs and an exact 0."""
    pos = sorted(x for x in new if x > mp.mpf(10) ** (-50))
    neg = sorted((-x for x in new if x < -mp.mpf(10) ** (-50)))
    assert len(pos) == len(neg)
    out = []
    for a, b in zip(pos, neg):
        assert abs(a - b) < mp.mpf(10) ** (-70), (a, b)
        out.append((a + b) / 2)
    return out  # positive generators only


def build_family(adds, moment, expect_deg):
    """adds: list of numbers of new nodes per level (level 1 is the single node 0)."""
    gens = []  # positive generators in order of first appearance
    levels = []  # per level: list of (global_index) sorted by node value, weights
    order = [mp.mpf(0)]  # master list: index 0 = centre, then +g1, -g1, +g2, -g2 ...
    cur = [mp.mpf(0)]
    rules = []
    w = weights(cur, moment)
    rules.append((list(cur), w))
    for lvl, m in enumerate(adds):
        new = extend(cur, m, moment)
        pos = symmetrise(new)
        assert 2 * len(pos) == m
        for g in pos:
            order.append(g)
            order.append(-g)
        cur = list(order)
        w = weights(cur, moment)
        rules.append((list(cur), w))
    for (nodes, w), ed in zip(rules, expect_deg):
        deg = exactness(nodes, w, moment, ed + 2)
        assert deg == ed + 0 or deg == ed, ("exactness", len(nodes), deg, ed)
        assert abs(sum(w) - 1) < mp.mpf(10) ** (-90)
    return order, rules


def hexf(x):
    return float(x).hex()


def emit(fh, prefix, order, rules, comment):
    n = len(order)
    L = len(rules)
    fh.write("/* %s */\n" % comment)
    fh.write("#define %s_LEVELS %d\n" % (prefix, L))
    fh.write("#define %s_NMAX %d\n" % (prefix, n))
    fh.write("static const int %s_npts[%d] = {%s};\n" % (prefix.lower(), L, ", ".join(str(len(r[0])) for r in rules)))
    fh.write("/* master node list: index 0 = centre, then (+g,-g) pairs in order of first appearance;\n"
             "   the level-l rule uses master indices 0 .. npts[l]-1 (nested) */\n")
    fh.write("static const double %s_nodes[%d] = {\n" % (prefix.lower(), n))
    for x in order:
        fh.write("  %s, /* %s */\n" % (hexf(x), mp.nstr(x, 20)))
    fh.write("};\n")
    fh.write("/* weights[l][j]: weight of master node j in the level-(l+1) rule (0 beyond npts[l]) */\n")
    fh.write("static const double %s_weights[%d][%d] = {\n" % (prefix.lower(), L, n))
    for nodes, w in rules:
        row = [hexf(x) for x in w] + ["0.0"] * (n - len(w))
        fh.write("  { %s },\n" % ", ".join(row))
    fh.write("};\n\n")


def main():
    gk_order, gk_rules = build_family([2, 6, 10, 16], gauss_moment, [1, 5, 15, 29, 51])
    kp_order, kp_rules = build_family([2, 4, 8, 16, 32], unif_moment, [1, 5, 11, 23, 47, 95])
    for path in sys.argv[1:]:
        with open(path, "w") as fh:
            fh.write("/* GENERATED by tools/gen_rules.py -- do not edit.\n"
                     " * Nested 1-D quadrature tables (hex-float literals: bit-identical on CPU and GPU). */\n"
                     "#pragma once\n\n")
            emit(fh, "JP_GK", gk_order, gk_rules,
                 "Genz-Keister nested rules for the weight exp(-x^2/2)/sqrt(2 pi): 1,3,9,19,35 points; exactness 1,5,15,29,51")
            emit(fh, "JP_KP", kp_order, kp_rules,
                 "Gauss-Kronrod-Patterson nested rules for the weight 1/2 on [-1,1]: 1,3,7,15,31,63 points; exactness 1,5,11,23,47,95")
            fh.write("/* KP nodes mapped to standard-normal space: z = Phi^-1((1+u)/2) = sqrt(2) erfinv(u), so that\n"
                     "   int g(z) phi(z) dz = (1/2) int_{-1}^{1} g(z(u)) du and both rule families share one code path */\n")
            fh.write("static const double jp_kp_znodes[%d] = {\n" % len(kp_order))
            for u in kp_order:
                z = mp.sqrt(2) * mp.erfinv(u)
                fh.write("  %s, /* %s */\n" % (hexf(z), mp.nstr(z, 20)))
            fh.write("};\n")
    print("GK nodes:", [mp.nstr(x, 17) for x in gk_order[:9]])
    print("GK 35 max node:", mp.nstr(max(gk_order), 17), "min weight:", mp.nstr(min(gk_rules[-1][1], key=abs), 5))
    print("negative GK weights per level:", [sum(1 for x in r[1] if x < 0) for r in gk_rules])
    print("KP nodes:", [mp.nstr(x, 17) for x in kp_order[:7]])


if __name__ == "__main__":
    main()
