"""CPU, property-based: the symbolic template compiler (scptoolbox.jl_b200/parser.py).  For random small programs whose
coefficients are random linear combinations of "sources", the compiled arrays  [Avals; Gvals; c; b; h] = W @ src  must
reproduce a direct numeric evaluation of every expression:  zero rows  A z - b = expr(z),  nonpos rows
G z - h = expr(z),  SOC rows  h - G z = expr(z),  cost  c'z + c0 = cost(z)  -- for any z and any source vector."""
import numpy as np
import scipy.sparse as sp
from hypothesis import given, settings, strategies as st


def _rand_expr(pkg, rng, nvar, nsrc, vars_):
    """random affine expression: sum_j (Lin coefficient) * z_j + Lin constant; coefficients mix constants and sources"""
    P = pkg.parser
    e = P.Expr()
    for j in rng.choice(nvar, size=rng.integers(1, min(4, nvar) + 1), replace=False):
        if rng.random() < 0.5:
            e = e + vars_[j] * float(rng.normal())
        else:
            L = P.Lin.src(int(rng.integers(1, nsrc)), float(rng.normal())) + P.Lin.const(float(rng.normal()))
            e = e + vars_[j].scale_lin(L)
    e = e + P.Expr(None, P.Lin.src(int(rng.integers(1, nsrc)), float(rng.normal())) + P.Lin.const(float(rng.normal())))
    return e


def _eval(e, z, src):
    lin = lambda L: sum(w * src[k] for k, w in L.t.items())
    return sum(lin(c) * z[v] for v, c in e.t.items()) + lin(e.c)


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10_000), l1_block=st.sampled_from([0, 2, 4]))
def test_compiled_template_reproduces_direct_evaluation(pkg, seed, l1_block):
    rng = np.random.default_rng(seed)
    P = pkg.parser
    nsrc = 6
    prg = P.ConicTemplate(nsrc, l1_block=l1_block)
    x = list(prg.new_variable(int(rng.integers(2, 6)), "x", stage="idx"))
    y = list(prg.new_variable((2, 2), "y", stage="col").ravel(order="F"))
    vars_ = x + y
    nv0 = len(vars_)
    zero = [_rand_expr(pkg, rng, nv0, nsrc, vars_) for _ in range(rng.integers(0, 4))]
    nonpos = [_rand_expr(pkg, rng, nv0, nsrc, vars_) for _ in range(rng.integers(1, 5))]
    socs = [[_rand_expr(pkg, rng, nv0, nsrc, vars_) for _ in range(rng.integers(2, 5))] for _ in range(rng.integers(0, 3))]
    l1 = [_rand_expr(pkg, rng, nv0, nsrc, vars_) for _ in range(rng.integers(2, 7))]
    linf = [_rand_expr(pkg, rng, nv0, nsrc, vars_) for _ in range(rng.integers(2, 4))]
    cost = _rand_expr(pkg, rng, nv0, nsrc, vars_)
    prg.zero(zero); prg.nonpos(nonpos)
    for c in socs:
        prg.soc(c)
    prg.l1(l1, stage=0); prg.linf(linf)
    prg.add_cost(cost)
    cp = prg.compile()
    n, p, m, l = cp["n"], cp["p"], cp["m"], cp["l"]
    src = np.concatenate([[1.0], rng.normal(size=nsrc - 1)])
    vals = cp["W"] @ src
    A = sp.csr_matrix((vals[:cp["nnzA"]], cp["A"].indices, cp["A"].indptr), shape=(p, n))
    G = sp.csr_matrix((vals[cp["nnzA"]:cp["nnzA"] + cp["nnzG"]], cp["G"].indices, cp["G"].indptr), shape=(m, n))
    c = vals[cp["off_c"]:cp["off_c"] + n]; b = vals[cp["off_b"]:cp["off_b"] + p]; h = vals[cp["off_h"]:cp["off_h"] + m]
    z = rng.normal(size=n)
    tol = 1e-12
    got = A @ z - b
    for i, e in enumerate(zero):
        assert abs(got[i] - _eval(e, z, src)) <= tol * (1 + abs(got[i]))
    got = G @ z - h
    for i, e in enumerate(nonpos):
        assert abs(got[i] - _eval(e, z, src)) <= tol * (1 + abs(got[i]))
    # SOC rows: s = h - G z = expr(z); they follow the l nonpos-type rows in cone order
    assert list(cp["soc_dims"]) == [len(c_) for c_ in socs] and m == l + sum(len(c_) for c_ in socs)
    srow = h - G @ z
    o = l
    for c_ in socs:
        for e in c_:
            assert abs(srow[o] - _eval(e, z, src)) <= tol * (1 + abs(srow[o]))
            o += 1
    assert abs(c @ z + sum(w * src[k] for k, w in cp["cost_const"].t.items()) - _eval(cost, z, src)) <= tol * (1 + abs(c @ z))


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 10_000), l1_block=st.sampled_from([0, 2, 3, 4]), k=st.integers(1, 9))
def test_l1_and_linf_lowerings_are_epigraphs(pkg, seed, l1_block, k):
    """min t subject to the lowered cone rows with x fixed equals |x|_1 (L1, any partial-sum block size) and |x|_inf."""
    from scipy.optimize import linprog
    rng = np.random.default_rng(seed)
    P = pkg.parser
    xv = rng.normal(size=k)
    for kind, want in (("l1", np.abs(xv).sum()), ("linf", np.abs(xv).max())):
        prg = P.ConicTemplate(1, l1_block=l1_block)
        t = prg.new_variable(1, "t")[0]
        x = list(prg.new_variable(k, "x"))
        (prg.l1([t] + x, stage=0) if kind == "l1" else prg.linf([t] + x))
        prg.add_cost(t)
        cp = prg.compile()
        vals = cp["W"] @ np.ones(1)
        n, m = cp["n"], cp["m"]
        G = sp.csr_matrix((vals[cp["nnzA"]:cp["nnzA"] + cp["nnzG"]], cp["G"].indices, cp["G"].indptr), shape=(m, n))
        h = vals[cp["off_h"]:cp["off_h"] + m]
        c = vals[cp["off_c"]:cp["off_c"] + n]
        bounds = [(None, None)] * n
        for i in range(k):
            bounds[1 + i] = (xv[i], xv[i])
        res = linprog(c, A_ub=G.toarray(), b_ub=h, bounds=bounds, method="highs")
        assert res.status == 0 and abs(res.fun - want) <= 1e-9 * max(1.0, want), (kind, l1_block, res.fun, want)
