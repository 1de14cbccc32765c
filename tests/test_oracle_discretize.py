"""Pin the CPU oracle (oracle/scp_oracle.c) with known-answer checks.

The reference ships no numeric fixtures for this path, so the oracle
is pinned by properties the reference algorithm must satisfy:
  * analytic Jacobians A, B, F == central finite differences of f (every model);
  * LTI models: A_k == expm(p0 * A_c * dt) (what helper.jl:248-265 `c2d` computes);
  * the DLTV reproduces the nonlinear propagation at the reference point:
        x_prop = A_k x_k + B-_k u_k + B+_k u_{k+1} + F_k p + r_k
  * a trajectory that IS an RK4 roll-out has zero defect and feas == true;
  * freeflyer: the integration action keeps |q| = 1.
"""
import numpy as np
import pytest
from scipy.linalg import expm

from oracle import orc, problems

MODELS = ["dblint", "rocket", "starship", "quadrotor", "freeflyer"]


def _fd(fun, z, rel=1e-6):
    J = np.zeros((fun(z).size, z.size))
    for i in range(z.size):
        h = rel * max(1.0, abs(z[i]))
        e = np.zeros(z.size); e[i] = h
        J[:, i] = (fun(z + e) - fun(z - e)) / (2 * h)
    return J


@pytest.mark.parametrize("name", MODELS)
def test_jacobians_match_finite_differences(name):
    N = 6
    pb = problems.make_problem(name, N)
    m = pb.orc_model()
    xd, ud, p = problems.test_trajectory(pb, 1, N, seed=3)
    x, u, p = xd[0, 2], ud[0, 2], p[0]
    for t in (0.2, 0.8):
        f, A, B, F = orc.dyn_eval(m, t, 1, x, u, p)
        Afd = _fd(lambda z: orc.dyn_eval(m, t, 1, z, u, p)[0], x)
        Bfd = _fd(lambda z: orc.dyn_eval(m, t, 1, x, z, p)[0], u)
        scale = max(np.abs(A).max(), 1e-12)
        assert np.abs(A - Afd).max() <= 1e-7 * scale + 1e-9
        assert np.abs(B - Bfd).max() <= 1e-7 * max(np.abs(B).max(), 1e-12) + 1e-9
        # F: only the first few parameters enter the dynamics
        npd = min(pb.np, 10)
        Ffd = _fd(lambda z: orc.dyn_eval(m, t, 1, x, u, np.concatenate([z, p[npd:]]))[0], p[:npd])
        assert np.abs(F[:, :npd] - Ffd).max() <= 1e-7 * max(np.abs(F).max(), 1e-12) + 1e-9
        assert np.abs(F[:, npd:]).max(initial=0.0) == 0.0


@pytest.mark.parametrize("name", ["dblint", "rocket", "quadrotor"])
def test_lti_transition_matrix_is_expm(name):
    N, Nsub = 8, 15
    pb = problems.make_problem(name, N)
    m = pb.orc_model()
    xd, ud, p = problems.test_trajectory(pb, 1, N, seed=1)
    d = orc.discretize(m, xd[0], ud[0], p[0], Nsub, np.ones(pb.nx), 1e-3)
    _, Ac, _, _ = orc.dyn_eval(m, 0.5, 1, xd[0, 0], ud[0, 0], p[0])  # already scaled by p0
    dt = 1.0 / (N - 1)
    Ak = expm(Ac * dt)
    for k in range(N - 1):
        assert np.allclose(d.A[k], Ak, rtol=0, atol=1e-9 * max(1.0, np.abs(Ak).max()))


@pytest.mark.parametrize("name", MODELS)
def test_dltv_reproduces_propagation_at_reference(name):
    # starship has a stiff actuator-lag mode (lambda = -tdil/rate_delay ~ -500): RK4 needs a fine sub-grid
    N, Nsub = 7, (400 if name == "starship" else 12)
    pb = problems.make_problem(name, N)
    m = pb.orc_model()
    xd, ud, p = problems.test_trajectory(pb, 1, N, seed=5)
    xd, ud, p = xd[0], ud[0], p[0]
    d = orc.discretize(m, xd, ud, p, Nsub, np.ones(pb.nx), 1e-3)
    for k in range(N - 1):
        xprop = xd[k + 1] - d.defect[k]
        xlin = d.A[k] @ xd[k] + d.Bm[k] @ ud[k] + d.Bp[k] @ ud[k + 1] + d.F[k] @ p + d.r[k]
        # freeflyer: the quaternion renormalisation is outside the linearisation; starship: RK4
        # truncation on the stiff lag mode and the O(h) kink at t = tau_s (inherent to the reference)
        tol = {"freeflyer": 1e-6, "starship": 2e-5}.get(name, 1e-7)
        assert np.abs(xprop - xlin).max() <= tol * max(1.0, np.abs(xprop).max())


def test_phi_starts_at_identity_and_E_is_integral_of_inverse():
    # with Nsub = 2 (one RK4 step of tiny dynamics: p ~ 0) Phi ~ I and E_k ~ dt * I
    N = 5
    pb = problems.make_problem("quadrotor", N)
    m = pb.orc_model()
    xd, ud, p = problems.test_trajectory(pb, 1, N, seed=2)
    p[:] = 1e-9
    d = orc.discretize(m, xd[0], ud[0], p[0], 2, np.ones(6), 1e-3)
    dt = 1.0 / (N - 1)
    assert np.allclose(d.A[0], np.eye(6), atol=1e-8)
    assert np.allclose(d.E[0], dt * np.eye(6), atol=1e-8)


@pytest.mark.parametrize("name", ["starship", "freeflyer"])
def test_rollout_has_zero_defect(name):
    N, Nsub = 6, (100 if name == "starship" else 10)
    pb = problems.make_problem(name, N)
    m = pb.orc_model()
    xd, ud, p = problems.test_trajectory(pb, 1, N, seed=9)
    xd, ud, p = xd[0].copy(), ud[0], p[0]
    # build xd as the RK4 roll-out segment by segment using the oracle's own defect
    for k in range(N - 1):
        d = orc.discretize(m, xd, ud, p, Nsub, np.ones(pb.nx), 1e-3)
        xd[k + 1] = xd[k + 1] - d.defect[k]
    d = orc.discretize(m, xd, ud, p, Nsub, np.ones(pb.nx), 1e-9)
    assert np.abs(d.defect).max() < 1e-12 * max(1.0, np.abs(xd).max())
    assert d.feas
    if name == "freeflyer":
        assert np.allclose(np.linalg.norm(xd[:, 6:10], axis=1), 1.0, atol=1e-12)


def test_feasibility_flag_and_nan_quirk():
    N = 5
    pb = problems.make_problem("dblint", N)
    m = pb.orc_model()
    xd, ud, p = problems.test_trajectory(pb, 1, N, seed=4)
    d = orc.discretize(m, xd[0], ud[0], p[0], 5, np.ones(2), 1e-12)
    assert not d.feas
    d2 = orc.discretize(m, xd[0], ud[0], p[0], 5, np.ones(2), 1e9)
    assert d2.feas


def test_linrange_matches_julia_lerp():
    # Julia: LinRange(0,1,N)[i] = (1 - j/d)*0 + (j/d)*1, j = i-1, d = N-1
    for N in (31, 100):
        tg = orc.t_grid(N)
        assert tg[0] == 0.0 and tg[-1] == 1.0
        j = np.arange(N) / (N - 1)
        assert np.array_equal(tg, j)
