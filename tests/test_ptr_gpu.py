"""GPU parity of the whole SCP hot path: batched PTR (scpb_ptr_solve through the host API) vs the oracle's
PTR loop (oracle/ptr.py with the oracle IPM standing in for ECOS) on the same initial guesses.

Stated tolerance (measured floor of round 1): final augmented cost 1e-6 relative,
iteration counts equal (+-1), states/parameters 2e-3 and thrust/gimbal inputs 2e-2 of their ranges (the LP
subproblem has flat directions, so two interior-point codes that stop at different duality gaps agree on the cost
far better than on the minimiser); both must report SCP_SOLVED.  North-star target is 1e-6 on the trajectory."""
import numpy as np
import pytest

from oracle import problems, ptr as optr

pytestmark = pytest.mark.gpu


def _setup(pkg, handle, N, Nsub, iter_max=15):
    ex = pkg.examples.starship
    mdl = ex.StarshipProblem()
    traj = pkg.problem.TrajectoryProblem(mdl)
    ex.define_problem(traj, "ptr", handle=handle)
    pars = pkg.ptr.Parameters(N=N, Nsub=Nsub, iter_max=iter_max, disc_method=pkg.ptr.FOH, wvc=1e3, wtr=0.1,
                              eps_abs=1e-5, eps_rel=0.01 / 100, feas_tol=5e-3, q_tr=np.inf, q_exit=np.inf,
                              solver_opts={"verbose": 0, "maxit": 100})
    return mdl, traj, pars


def test_initial_guess_matches_oracle(pkg, handle):
    N = 12
    mdl, traj, pars = _setup(pkg, handle, N, 40)
    xg, ug, pg = traj.guess(N)          # phase-2 SOCPs solved as one batch on the GPU cone solver
    pbo = problems.StarshipProblem(N)
    xo, uo, po = pbo.guess(N)
    assert abs(pg[1] - po[1]) <= 1.0 and abs(pg[0] - po[0]) < 1e-9      # same flip time; descent time within one 1-s candidate step
    assert abs(mdl.hs - pbo.hs) < 1e-9
    sx = np.array([r[1] - r[0] for r in pbo.ranges()[0]])
    # the terminal-descent SOCP has flat directions: two interior-point codes agree to ~3e-3 of the ranges
    assert np.abs((xg - xo) / sx).max() < 1e-2
    assert np.abs((ug[:, 0] - uo[:, 0]) / 1e6).max() < 5e-2


@pytest.mark.parametrize("N,Nsub,nb", [(12, 60, 4), (31, 100, 3)])
def test_batched_ptr_matches_oracle_ptr(pkg, handle, N, Nsub, nb):
    mdl, traj, pars = _setup(pkg, handle, N, Nsub)
    pbo = problems.StarshipProblem(N)
    g = pbo.guess(N)
    mdl.hs = pbo.hs                      # same cost normalisation on both sides
    opars = optr.Parameters(N=N, Nsub=Nsub, iter_max=15, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=0.01 / 100,
                            feas_tol=5e-3, solver_tol=1e-9)
    P = optr.PTR(pbo, opars)
    rng = np.random.default_rng(N)
    sc = P.scale
    X0 = np.array([g[0] + (0.01 * sc.Sx * rng.standard_normal(g[0].shape) if b else 0.0) for b in range(nb)])
    U0 = np.array([g[1] + (0.01 * sc.Su * rng.standard_normal(g[1].shape) if b else 0.0) for b in range(nb)])
    P0 = np.array([g[2] * (1 + (0.02 * rng.uniform(-1, 1, g[2].shape) if b else 0.0)) for b in range(nb)])
    pbm = pkg.ptr.create(pars, traj, handle)
    sol = pkg.ptr.solve(pbm, (X0, U0, P0))
    pbm.close()
    assert all(s == "SCP_SOLVED" for s in sol.status), sol.status
    for b in range(nb):
        ref = P.solve((X0[b], U0[b], P0[b]), prefer="ipm")
        assert ref["status"] == "SCP_SOLVED"
        rs = ref["sol"]
        ex = np.abs((sol.xd[b] - rs.xd) / sc.Sx).max()
        eu = np.abs((sol.ud[b] - rs.ud) / sc.Su).max()
        ep = np.abs((sol.p[b] - rs.p) / sc.Sp).max()
        # states and parameters are pinned by the dynamics; the gimbal-rate input has flat directions in this LP
        # (it only enters two-sided rate constraints), so inputs are compared on thrust and gimbal angle
        eu2 = np.abs((sol.ud[b][:, :2] - rs.ud[:, :2]) / sc.Su[:2]).max()
        print("parity seed", b, "ex", ex, "eu(T,delta)", eu2, "eu(all)", eu, "ep", ep)
        assert max(ex, ep) <= 2e-3 and eu2 <= 2e-2, (b, ex, eu, ep, sol.iterations[b], ref["iterations"])
        assert abs(sol.cost[b] - rs.J_aug) <= 1e-6 * max(1.0, abs(rs.J_aug))
        assert abs(int(sol.iterations[b]) - ref["iterations"]) <= 1
        assert bool(sol.feas[b]) == rs.feas
