"""CPU: known-answer pins of the oracle's whole SCP chain (discretize -> formulate -> conic solve -> PTR logic).

The reference holds no numeric fixtures for its SCP solvers (status-only asserts), so the oracle PTR is
pinned against (i) the closed-form maximum-principle optimum of the double integrator (the reference's own analytic
cross-check, double_integrator/definition.jl:137-294, is the same idea for its LCvx program) and (ii) the
status/iteration budget of the reference tests."""
import numpy as np
import pytest

from oracle import problems, ptr as optr


@pytest.mark.parametrize("choice", [1, 2])
def test_double_integrator_min_time_matches_maximum_principle(choice):
    N = 30                                    # BASELINE config C1
    pb = problems.DoubleIntegratorProblem(N, choice)
    pars = optr.Parameters(N=N, Nsub=10, iter_max=30, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=1e-4, feas_tol=1e-3,
                           solver_tol=1e-9)
    out = optr.PTR(pb, pars).solve(pb.guess(N))
    assert out["status"] == "SCP_SOLVED" and out["sol"].feas
    T, t1 = pb.t_opt()
    tf = out["sol"].p[0]
    # first-order-hold inputs cannot switch inside a segment: the discrete optimum is slightly slower, never faster
    assert T * (1 - 1e-6) <= tf <= T * (1 + 5e-3), (tf, T)
    u = out["sol"].ud[:, 0]
    tau_sw = t1 / T
    tau = np.arange(N) / (N - 1)
    assert (u[tau < tau_sw - 1.0 / (N - 1)] > pb.u_max * (1 - 1e-5)).all()      # bang ...
    assert (u[tau > tau_sw + 1.0 / (N - 1)] < -pb.u_max * (1 - 1e-5)).all()     # ... bang
    x = out["sol"].xd
    assert abs(x[-1, 0] - pb.s) <= 1e-6 * pb.s and abs(x[-1, 1]) <= 1e-6 and np.abs(x[0]).max() <= 1e-9


def test_rocket_landing_ptr_converges_and_is_physical():
    N = 12
    pb = problems.RocketProblem(N)
    pars = optr.Parameters(N=N, Nsub=15, iter_max=20, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=1e-4, feas_tol=1e-3,
                           solver_tol=1e-9)
    out = optr.PTR(pb, pars).solve(pb.guess(N), prefer="ipm")
    s = out["sol"]
    assert out["status"] == "SCP_SOLVED" and out["iterations"] <= 15 and s.feas
    a, xi = s.ud[:, 0:3], s.ud[:, 3]
    assert (np.linalg.norm(a, axis=1) <= xi * (1 + 1e-6) + 1e-9).all()           # SOC slack holds
    assert pb.m_dry <= np.exp(s.xd[-1, 6]) <= pb.m_wet
    assert pb.tf_min <= s.p[0] <= pb.tf_max and np.abs(s.xd[-1, 0:6]).max() <= 1e-5


def test_oracle_scvx_solves_the_reference_starship_configuration():
    """starship_flip/tests.jl:69-121: SCvx must end SCP_SOLVED within iter_max = 100 (the reference's own assertion),
    with a dynamically feasible trajectory; the trust region only ever shrinks by beta_sh or grows by beta_gr."""
    from oracle import scvx as oscvx
    N = 31
    pb = problems.StarshipProblem(N)
    pars = oscvx.Parameters(N=N, Nsub=100, iter_max=100, lam=5e2, rho_0=0.0, rho_1=0.1, rho_2=0.7, beta_sh=2.0,
                            beta_gr=2.0, eta_init=1.0, eta_lb=1e-8, eta_ub=10.0, eps_abs=1e-5, eps_rel=1e-4, feas_tol=5e-3)
    out = oscvx.SCvx(pb, pars).solve(pb.guess(N))
    assert out["status"] == "SCP_SOLVED" and out["iterations"] < 100 and out["sol"].feas
    etas = [h.eta for h in out["history"]]
    for a, b in zip(etas[:-1], etas[1:]):
        assert b in (a, a / 2.0, min(10.0, 2.0 * a))
    J = [h.J_aug for h in out["history"]]
    assert J[-1] < 1e-3 * J[0]                     # the penalised cost drops by three orders of magnitude
