#!/usr/bin/env python
"""Generate the nested 1-D quadrature tables used by the Smolyak builder.

Genz-Keister (1996) nested extensions of Gauss-Hermite for the standard-normal
weight phi(x) = exp(-x^2/2)/sqrt(2 pi): 1 -> 3 -> 9 -> 19 -> 35 points with
polynomial exactness 1, 5, 15, 29, 51, and the three further published members of the
family with 37, 41 and 43 points (exactness 55, 63, 67).  Those three extend the
19-POINT rule by 18 / 22 / 24 nodes -- they are nested with levels 1-4 but NOT with the
35-point rule: a Kronrod-type extension of the 35-point rule by one node pair is
degenerate (its two new nodes get weight zero and the exactness stays 51; checked below),
by two pairs it has complex roots.  The tables therefore carry, per level, the list of
master-node indices the level uses (a prefix of the master list for the nested levels).  The reference reaches these rules through
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

mp.mp.dps = 400   # the 43-point Vandermonde system loses ~90 digits


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
    # every level so far is a prefix of the master list
    index = [list(range(len(nodes))) for nodes, _ in rules]
    return order, rules, index


def side_extensions(order, rules, index, base_level, adds, moment, expect_deg):
    """Further rules that extend the rule of `base_level` (1-based) by m new nodes each, m in adds.  Their new nodes
    are appended to the master list as (+g, -g) pairs; a rule is (master indices, weights per member)."""
    base_idx = index[base_level - 1]
    base = [order[i] for i in base_idx]
    for m, ed in zip(adds, expect_deg):
        pos = symmetrise(extend(base, m, moment))
        assert 2 * len(pos) == m
        idx = list(base_idx)
        for g in pos:
            idx += [len(order), len(order) + 1]
            order += [g, -g]
        nodes = [order[i] for i in idx]
        w = weights(nodes, moment)
        deg = exactness(nodes, w, moment, ed + 2)
        assert deg == ed, ("exactness", len(nodes), deg, ed)
        assert abs(sum(w) - 1) < mp.mpf(10) ** (-90)
        rules.append((nodes, w))
        index.append(idx)
    return order, rules, index


def degenerate_35_plus_2(order, rules, moment):
    """The claim of the module docstring: extending the 35-point rule by one node pair gains nothing."""
    nodes35 = list(rules[4][0])
    pos = symmetrise(extend(nodes35, 2, moment))
    nodes = nodes35 + [pos[0], -pos[0]]
    w = weights(nodes, moment)
    return exactness(nodes, w, moment, 60), max(abs(w[-1]), abs(w[-2]))


def hexf(x):
    return float(x).hex()


def emit(fh, prefix, order, rules, index, comment):
    n = len(order)
    L = len(rules)
    fh.write("/* %s */\n" % comment)
    fh.write("#define %s_LEVELS %d\n" % (prefix, L))
    fh.write("#define %s_NMAX %d\n" % (prefix, n))
    fh.write("static const int %s_npts[%d] = {%s};\n" % (prefix.lower(), L, ", ".join(str(len(r[0])) for r in rules)))
    fh.write("/* master node list: index 0 = centre, then (+g,-g) pairs in order of first appearance;\n"
             "   the level-l rule uses the master indices index[l][0 .. npts[l]-1] (a prefix for the nested levels) */\n")
    fh.write("static const double %s_nodes[%d] = {\n" % (prefix.lower(), n))
    for x in order:
        fh.write("  %s, /* %s */\n" % (hexf(x), mp.nstr(x, 20)))
    fh.write("};\n")
    fh.write("/* weights[l][j]: weight of MASTER node j in the level-(l+1) rule (0 for nodes the level does not use) */\n")
    fh.write("static const double %s_weights[%d][%d] = {\n" % (prefix.lower(), L, n))
    for (nodes, w), idx in zip(rules, index):
        row = ["0.0"] * n
        for i, x in zip(idx, w):
            row[i] = hexf(x)
        fh.write("  { %s },\n" % ", ".join(row))
    fh.write("};\n")
    fh.write("/* index[l][pos]: master index of the pos-th node of the level-(l+1) rule (generation order of the grid build) */\n")
    fh.write("static const unsigned char %s_index[%d][%d] = {\n" % (prefix.lower(), L, n))
    for idx in index:
        fh.write("  { %s },\n" % ", ".join(str(i) for i in idx + [0] * (n - len(idx))))
    fh.write("};\n\n")


def main():
    gk_order, gk_rules, gk_index = build_family([2, 6, 10, 16], gauss_moment, [1, 5, 15, 29, 51])
    deg37, w37 = degenerate_35_plus_2(gk_order, gk_rules, gauss_moment)
    assert deg37 == 51 and w37 < mp.mpf(10) ** (-100), (deg37, w37)
    gk_order, gk_rules, gk_index = side_extensions(gk_order, gk_rules, gk_index, 4, [18, 22, 24], gauss_moment, [55, 63, 67])
    kp_order, kp_rules, kp_index = build_family([2, 4, 8, 16, 32], unif_moment, [1, 5, 11, 23, 47, 95])
    for path in sys.argv[1:]:
        with open(path, "w") as fh:
            fh.write("/* GENERATED by tools/gen_rules.py -- do not edit.\n"
                     " * Nested 1-D quadrature tables (hex-float literals: bit-identical on CPU and GPU). */\n"
                     "#pragma once\n\n")
            emit(fh, "JP_GK", gk_order, gk_rules, gk_index,
                 "Genz-Keister rules for the weight exp(-x^2/2)/sqrt(2 pi): 1,3,9,19,35 points (nested; exactness 1,5,15,29,51) and "
                 "the 37-, 41-, 43-point extensions of the 19-point rule (exactness 55,63,67)")
            emit(fh, "JP_KP", kp_order, kp_rules, kp_index,
                 "Gauss-Kronrod-Patterson nested rules for the weight 1/2 on [-1,1]: 1,3,7,15,31,63 points; exactness 1,5,11,23,47,95")
            fh.write("/* KP nodes mapped to standard-normal space: z = Phi^-1((1+u)/2) = sqrt(2) erfinv(u), so that\n"
                     "   int g(z) phi(z) dz = (1/2) int_{-1}^{1} g(z(u)) du and both rule families share one code path */\n")
            fh.write("static const double jp_kp_znodes[%d] = {\n" % len(kp_order))
            for u in kp_order:
                z = mp.sqrt(2) * mp.erfinv(u)
                fh.write("  %s, /* %s */\n" % (hexf(z), mp.nstr(z, 20)))
            fh.write("};\n")
    print("GK nodes:", [mp.nstr(x, 17) for x in gk_order[:9]])
    print("GK 35 max node:", mp.nstr(max(gk_order[:35]), 17), "min weight:", mp.nstr(min(gk_rules[4][1], key=abs), 5))
    print("GK 35 + one node pair: exactness", deg37, "weight of the new pair", mp.nstr(w37, 3))
    print("GK points per level:", [len(r[0]) for r in gk_rules], "master nodes:", len(gk_order))
    print("negative GK weights per level:", [sum(1 for x in r[1] if x < 0) for r in gk_rules])
    print("KP nodes:", [mp.nstr(x, 17) for x in kp_order[:7]])


if __name__ == "__main__":
    main()
