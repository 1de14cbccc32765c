"""CPU: the oracle's two cone solvers (oracle/conic.py: HiGHS for LPs, the CVXOPT-style Nesterov-Todd interior point for
LP + SOC programs) pinned against ANALYTIC optima and against each other, so that the SOC parity of the CUDA solver
(tests/test_conic_gpu.py compares with `solve_ipm`) does not rest on "the same algorithm twice".

  * LP over a simplex                      -> the smallest cost coefficient;
  * linear cost over a Euclidean ball      -> x* = -r c / |c|,  value -r |c|;
  * projection onto the second-order cone  -> the closed-form Lorentz-cone projection;
  * least squares as a cone program        -> numpy's lstsq residual;
  * an LP with L1 / LINF epigraph lowerings (the NormOne / NormInfinity bridges the reference goes through) -> both solvers
    and the analytic value agree.
The reference's own conic known answer (LCvx double integrator vs the maximum principle) is tests/test_lcvx_known_answer.py."""
import numpy as np
import pytest

from oracle import conic


def _var(P, n, name):
    return P.new_variable(n, name)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_lp_over_a_simplex(seed):
    rng = np.random.default_rng(seed)
    n = 7
    c = rng.standard_normal(n)
    P = conic.ConeProgram()
    x = _var(P, n, "x")
    tot = conic.Aff()
    for i in range(n):
        P.nonpos([-x[i]])
        tot = tot + x[i]
    P.zero([tot - 1.0])
    cost = conic.Aff()
    for i in range(n):
        cost = cost + x[i] * float(c[i])
    P.add_cost(cost)
    cp = P.compile()
    a, b = conic.solve_highs(cp), conic.solve_ipm(cp, tol=1e-10)
    assert a["status"] == b["status"] == "OPTIMAL"
    assert abs(a["obj"] - c.min()) <= 1e-9 and abs(b["obj"] - c.min()) <= 1e-8
    assert abs(b["z"][np.argmin(c)] - 1.0) <= 1e-6


@pytest.mark.parametrize("seed,r", [(0, 1.0), (1, 2.5), (2, 0.3)])
def test_linear_cost_over_a_ball(seed, r):
    rng = np.random.default_rng(seed)
    n = 5
    c = rng.standard_normal(n)
    P = conic.ConeProgram()
    x = _var(P, n, "x")
    P.soc([conic.Aff(None, r)] + [x[i] for i in range(n)])       # |x|_2 <= r
    cost = conic.Aff()
    for i in range(n):
        cost = cost + x[i] * float(c[i])
    P.add_cost(cost)
    out = conic.solve_ipm(P.compile(), tol=1e-10)
    assert out["status"] == "OPTIMAL"
    want = -r * c / np.linalg.norm(c)
    assert abs(out["obj"] + r * np.linalg.norm(c)) <= 1e-8 * max(1.0, r * np.linalg.norm(c))
    assert np.abs(out["z"][:n] - want).max() <= 1e-6


def _project_soc(p):
    x0, t0 = p[1:], p[0]
    nx = np.linalg.norm(x0)
    if nx <= t0:
        return p.copy()
    if nx <= -t0:
        return np.zeros_like(p)
    a = 0.5 * (nx + t0)
    return np.concatenate([[a], a * x0 / nx])


@pytest.mark.parametrize("p", [[1.0, 2.0, -1.0, 0.5], [-0.2, 0.3, 0.4, 0.0], [3.0, 1.0, 1.0, 1.0], [-5.0, 1.0, 0.0, 2.0]])
def test_projection_onto_the_second_order_cone(p):
    """min d  s.t.  |z - p|_2 <= d,  z = (t, x) in SOC: z* is the closed-form Lorentz projection, d* the distance."""
    p = np.array(p)
    n = len(p)
    P = conic.ConeProgram()
    z = _var(P, n, "z"); d = _var(P, 1, "d")
    P.soc([z[i] for i in range(n)])
    P.soc([d[0]] + [z[i] - float(p[i]) for i in range(n)])
    P.add_cost(d[0])
    out = conic.solve_ipm(P.compile(), tol=1e-10)
    assert out["status"] == "OPTIMAL"
    want = _project_soc(p)
    dist = np.linalg.norm(want - p)
    assert abs(out["obj"] - dist) <= 1e-7 * max(1.0, dist)
    if dist > 1e-6:       # the minimiser is unique when the point is outside the cone
        assert np.abs(out["z"][:n] - want).max() <= 1e-5


@pytest.mark.parametrize("seed", [0, 1])
def test_least_squares_as_a_cone_program(seed):
    rng = np.random.default_rng(seed)
    m, n = 9, 4
    A, b = rng.standard_normal((m, n)), rng.standard_normal(m)
    P = conic.ConeProgram()
    x = _var(P, n, "x"); t = _var(P, 1, "t")
    rows = []
    for i in range(m):
        e = conic.Aff(None, -float(b[i]))
        for j in range(n):
            e = e + x[j] * float(A[i, j])
        rows.append(e)
    P.soc([t[0]] + rows)
    P.add_cost(t[0])
    out = conic.solve_ipm(P.compile(), tol=1e-10)
    xs, res, *_ = np.linalg.lstsq(A, b, rcond=None)
    assert out["status"] == "OPTIMAL"
    assert abs(out["obj"] - np.sqrt(res[0])) <= 1e-8 and np.abs(out["z"][:n] - xs).max() <= 1e-6


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_l1_and_linf_epigraphs_agree_across_solvers(seed):
    """min |x - a|_1 + 2 |x - b|_inf over a box: polyhedral (both solvers apply), value checked by brute force on the
    vertices of the piecewise-linear objective in one dimension per coordinate (separable L1 part) via HiGHS == IPM."""
    rng = np.random.default_rng(seed)
    n = 4
    a, b = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    P = conic.ConeProgram()
    x = _var(P, n, "x"); t1 = _var(P, 1, "t1"); t2 = _var(P, 1, "t2")
    for i in range(n):
        P.nonpos([x[i] - 2.0]); P.nonpos([-2.0 - x[i]])
    P.l1([t1[0]] + [x[i] - float(a[i]) for i in range(n)])
    P.linf([t2[0]] + [x[i] - float(b[i]) for i in range(n)])
    P.add_cost(t1[0] + t2[0] * 2.0)
    cp = P.compile()
    h, ip = conic.solve_highs(cp), conic.solve_ipm(cp, tol=1e-9)
    # (the LINF part makes the optimal face degenerate: the interior point may stop at its reduced tolerances)
    assert h["status"] == "OPTIMAL" and ip["status"] in ("OPTIMAL", "ALMOST_OPTIMAL")
    assert abs(h["obj"] - ip["obj"]) <= 1e-7 * max(1.0, abs(h["obj"]))
    f = lambda v: np.abs(v - a).sum() + 2.0 * np.abs(v - b).max()
    assert abs(f(ip["z"][:n]) - ip["obj"]) <= 1e-6           # the epigraph variables are tight at the optimum
    # no feasible point does better (random search around the optimum + the two anchors)
    cand = [a, b, ip["z"][:n]] + [ip["z"][:n] + 0.05 * rng.standard_normal(n) for _ in range(200)]
    assert min(f(np.clip(v, -2, 2)) for v in cand) >= ip["obj"] - 1e-6


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_geom_cone_is_the_squared_norm_epigraph(seed):
    """The q_tr = 4 trust region of ptr.jl:604-630: |d|_2 <= d_lq (SOC), |d_lq| <= w (SOC), geomean(eta, 1) >= w (GEOM)
    makes eta the epigraph of |d|_2^2.  With d pinned to a given vector, min eta must return its squared norm."""
    rng = np.random.default_rng(seed)
    n = 5
    a = rng.standard_normal(n)
    P = conic.ConeProgram()
    d = _var(P, n, "d"); d_lq = _var(P, 1, "d_lq"); w = _var(P, 1, "w"); eta = _var(P, 1, "eta")
    P.zero([d[i] - float(a[i]) for i in range(n)])
    P.soc([d_lq[0]] + [d[i] for i in range(n)])
    P.soc([w[0], d_lq[0]])
    P.geom([w[0], eta[0], 1.0])
    P.add_cost(eta[0])
    out = conic.solve_ipm(P.compile(), tol=1e-10)
    assert out["status"] in ("OPTIMAL", "ALMOST_OPTIMAL")
    assert abs(out["obj"] - a @ a) <= 1e-7 * max(1.0, a @ a)
