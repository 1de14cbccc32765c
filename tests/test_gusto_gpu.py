"""GPU: BASELINE config C4 -- quadrotor obstacle avoidance with GuSTO (scpb_gusto_* through the host API,
scptoolbox.jl_b200/gusto.py) against the oracle's GuSTO loop (oracle/gusto.py restating src/solvers/gusto.jl) at the
reference's own test constants (test/examples/quadrotor/tests.jl:81-145: N = 30, Nsub = 15, iter_max = 15, lambda_init =
1e4, lambda_max = 1e9, rho in (0.1, 0.9), beta = 2, gamma_fail = 5, eta in [1e-3, 10] from 10, mu = 0.8 from iteration 6,
eps_abs = eps_rel = 0  =>  exactly iter_max iterations, pen = :quad, q_tr = q_exit = Inf).

Asserted: same status, same iteration count, the same accept / reject / penalty history as far as it shows in the final
trust-region radius and penalty weight (exactly), final cost and physical trajectory within the stated tolerance with both
cone solvers at 1e-11."""
import numpy as np
import pytest

from oracle import gusto as ogusto, problems, ptr as optr

pytestmark = pytest.mark.gpu

TOL = dict(feastol=1e-11, abstol=1e-11, reltol=1e-11)
KW = dict(lam_init=1e4, lam_max=1e9, rho_0=0.1, rho_1=0.9, beta_sh=2.0, beta_gr=2.0, gamma_fail=5.0, eta_init=10.0,
          eta_lb=1e-3, eta_ub=10.0, mu=0.8, iter_mu=6, feas_tol=1e-3)


def _setup(pkg, handle, N, iter_max, eps=0.0):
    ex = pkg.examples.quadrotor
    mdl = ex.QuadrotorProblem()
    traj = pkg.problem.TrajectoryProblem(mdl)
    ex.define_problem(traj, "gusto", handle=handle)
    pars = pkg.gusto.Parameters(N, 15, iter_max, pkg.ptr.FOH, KW["lam_init"], KW["lam_max"], KW["rho_0"], KW["rho_1"],
                                KW["beta_sh"], KW["beta_gr"], KW["gamma_fail"], KW["eta_init"], KW["eta_lb"], KW["eta_ub"],
                                KW["mu"], KW["iter_mu"], eps, eps, KW["feas_tol"], "quad", 100.0, np.inf, np.inf, None,
                                {"verbose": 0, "maxit": 100})
    return mdl, traj, pars


def test_correct_convex_matches_oracle(pkg, handle):
    """correct_convex! (scp.jl:275-361) of a guess that violates the input set (thrust below u_min and outside the tilt
    cone, flight time beyond tf_max): the batched GPU projection equals the oracle's, seed by seed."""
    N, nb = 12, 4
    mdl, traj, pars = _setup(pkg, handle, N, 3)
    pbm = pkg.gusto.create(pars, traj, handle)
    pbo = problems.QuadrotorProblem(N)
    so = optr.Scaling(pbo, N)
    sc = pbm.scale
    assert np.abs(sc.Su - so.Su).max() <= 1e-6 * so.Su.max() and np.abs(sc.cu - so.cu).max() <= 1e-6 * so.Su.max()
    g = pbo.guess(N)
    rng = np.random.default_rng(0)
    X0 = np.array([g[0] + 0.1 * rng.standard_normal(g[0].shape) for b in range(nb)])
    U0 = np.array([g[1] + 3.0 * rng.standard_normal(g[1].shape) for b in range(nb)])
    U0[0, :, 3] = 0.1                       # below u_min and inside |a|
    P0 = np.array([g[2] * (1 + b) for b in range(nb)])      # 1.25, 2.5, 3.75 (> tf_max), 5.0
    xd, ud, p = pkg.ptr.correct_convex(pbm, (X0, U0, P0), **TOL)
    pbm.close()
    for b in range(nb):
        xo, uo, po = optr.correct_convex(pbo, so, N, X0[b], U0[b], P0[b], tol=1e-11)
        ex = np.abs(xd[b] - xo).max(); eu = np.abs((ud[b] - uo) / so.Su).max(); ep = np.abs(p[b] - po).max()
        print("correct_convex seed", b, "ex", ex, "eu", eu, "ep", ep, "moved u by", np.abs(uo - U0[b]).max())
        assert ex <= 1e-7 and eu <= 1e-7 and ep <= 1e-7
        assert po[0] <= pbo.tf_max + 1e-9


def test_batched_gusto_matches_oracle_gusto(pkg, handle):
    N, nb, K = 30, 3, 15
    mdl, traj, pars = _setup(pkg, handle, N, K)
    pbo = problems.QuadrotorProblem(N)
    P = ogusto.GuSTO(pbo, ogusto.Parameters(N=N, Nsub=15, iter_max=K, eps_abs=0.0, eps_rel=0.0, solver_tol=1e-11, **KW))
    sc = P.scale
    g = pbo.guess(N)
    rng = np.random.default_rng(7)
    X0 = np.array([g[0] + (0.02 * rng.standard_normal(g[0].shape) if b else 0.0) for b in range(nb)])
    U0 = np.array([g[1] + (0.2 * rng.standard_normal(g[1].shape) if b else 0.0) for b in range(nb)])
    P0 = np.array([g[2] * (1 + 0.1 * b) for b in range(nb)])
    pbm = pkg.gusto.create(pars, traj, handle)
    sol = pkg.gusto.solve(pbm, (X0, U0, P0), **TOL)
    info = pbm.cone.info()
    pbm.close()
    print("quadrotor GuSTO KKT", info["nk"], info["nnzL"], info["levels"], "timing", sol.timing)
    for b in range(nb):
        guess = optr.correct_convex(pbo, sc, N, X0[b], U0[b], P0[b], tol=1e-11)      # generate_initial_guess, gusto.jl:517-526
        ref = P.solve(guess)
        rs = ref["sol"]
        ex = np.abs(sol.xd[b] - rs.xd).max(); eu = np.abs((sol.ud[b] - rs.ud) / sc.Su).max()
        ep = np.abs((sol.p[b] - rs.p) / sc.Sp).max(); dJ = abs(sol.cost[b] - rs.J_aug) / max(1.0, abs(rs.J_aug))
        print("gusto parity seed", b, "iters", sol.iterations[b], ref["iterations"], "eta", sol.eta[b], ref["eta"], "lam",
              sol.lam[b], ref["lam"], "ex", ex, "eu", eu, "ep", ep, "dJ", dJ, "J_aug", sol.cost[b], rs.J_aug, "feas", sol.feas[b], rs.feas)
        assert sol.status[b] == ref["status"] == "SCP_SOLVED", (sol.status, sol.raw_status, ref["status"])
        assert int(sol.iterations[b]) == ref["iterations"] == K
        assert sol.eta[b] == pytest.approx(ref["eta"], rel=1e-12) and sol.lam[b] == pytest.approx(ref["lam"], rel=1e-12)
        assert bool(sol.feas[b]) == bool(rs.feas)
        assert dJ <= 1e-6 and max(ex, eu, ep) <= 1e-5
        # the keep-out zones are respected (nonconvex feasibility of the converged trajectory)
        assert max(pbo.s(0, k + 1, sol.xd[b][k], None, sol.p[b]).max() for k in range(N)) <= 1e-3


def test_gusto_c4_batch(pkg, handle):
    """C4 at the bench node count (N = 60) on a 64-seed sweep of perturbed guesses with the reference's stopping
    tolerances switched on (eps_abs = 1e-5, eps_rel = 1e-4): every seed ends SCP_SOLVED, dynamically feasible and clear of
    the obstacles; seeds stop on their own schedule (frozen seeds are skipped by the device loop)."""
    N, nb = 60, 64
    mdl, traj, pars = _setup(pkg, handle, N, 30, eps=0.0)
    pars.eps_abs, pars.eps_rel = 1e-5, 1e-4
    pbo = problems.QuadrotorProblem(N)
    g = pbo.guess(N)
    rng = np.random.default_rng(3)
    X0 = np.array([g[0] + 0.05 * rng.standard_normal(g[0].shape) for b in range(nb)])
    U0 = np.array([g[1] + 0.5 * rng.standard_normal(g[1].shape) for b in range(nb)])
    P0 = np.array([g[2] * rng.uniform(0.8, 1.2) for b in range(nb)])
    pbm = pkg.gusto.create(pars, traj, handle)
    sol = pkg.gusto.solve(pbm, (X0, U0, P0))
    pbm.close()
    its = np.asarray(sol.iterations)
    print("quadrotor C4 batch: iterations min/med/max", its.min(), int(np.median(its)), its.max(), "feas", int(np.sum(sol.feas)),
          "timing", sol.timing)
    assert all(s == "SCP_SOLVED" for s in sol.status), sol.status
    assert np.all(np.asarray(sol.feas) == 1)
    for b in range(nb):
        assert max(pbo.s(0, k + 1, sol.xd[b][k], None, sol.p[b]).max() for k in range(N)) <= 1e-3
    print("cost spread", np.min(sol.cost), np.median(sol.cost), np.max(sol.cost))
