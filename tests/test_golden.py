"""Committed fixtures (tests/golden/, written by scripts/make_golden.py): small fixed inputs with the ORACLE's outputs.

They are regression fixtures of the oracle, not reference outputs (the reference ships none and cannot run here).  The
CPU tests pin the oracle against them; the GPU tests give the CUDA path the same file-based targets."""
import os

import numpy as np
import pytest

from oracle import orc, problems, ptr as optr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = ["dblint", "rocket", "starship", "quadrotor", "freeflyer"]
KEYS = ("A", "Bm", "Bp", "F", "r", "E", "defect")


@pytest.mark.parametrize("name", MODELS)
def test_oracle_discretize_and_propagate_match_golden(name):
    g = np.load(os.path.join(GOLD, f"oracle_discretize_{name}.npz"))
    N, Nsub = int(g["N"]), int(g["Nsub"])
    pb = problems.make_problem(name, N)
    m = pb.orc_model()
    for b in range(2):
        r = orc.discretize(m, g["xd"][b], g["ud"][b], g["p"][b], Nsub, g["iSx"], float(g["feas_tol"]))
        for k in KEYS:
            want = g[f"{k}_{b}"]
            assert np.abs(getattr(r, k) - want).max() <= 1e-13 * max(1.0, np.abs(want).max()), (name, k)
        assert bool(r.feas) == bool(g[f"feas_{b}"])
        xc = orc.propagate(m, g["xd"][b], g["ud"][b], g["p"][b], 2 * Nsub * (N - 1))
        assert np.abs(xc - g[f"xc_{b}"]).max() <= 1e-13 * max(1.0, np.abs(g[f"xc_{b}"]).max())


def test_oracle_ptr_matches_golden():
    g = np.load(os.path.join(GOLD, "oracle_ptr_dblint.npz"))
    N = 30
    pb = problems.DoubleIntegratorProblem(N, 1)
    pars = optr.Parameters(N=N, Nsub=10, iter_max=30, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=1e-4, feas_tol=1e-3,
                           solver_tol=1e-9)
    out = optr.PTR(pb, pars).solve(pb.guess(N))
    assert out["iterations"] == int(g["iterations"])
    assert np.abs(out["sol"].xd - g["xd"]).max() <= 1e-7 * pb.s and abs(out["sol"].p[0] - g["p"][0]) <= 1e-7
    assert abs(g["p"][0] - g["t_opt"][0]) <= 5e-3 * g["t_opt"][0]          # the fixture itself sits on the analytic optimum


@pytest.mark.gpu
@pytest.mark.parametrize("name", MODELS)
def test_cuda_discretize_and_propagate_match_golden(handle, name):
    g = np.load(os.path.join(GOLD, f"oracle_discretize_{name}.npz"))
    N, Nsub = int(g["N"]), int(g["Nsub"])
    pb = problems.make_problem(name, N)
    handle.model_set(pb.model_id, pb.par(), pb.nx, pb.nu, pb.np)
    out = handle.discretize(orc.t_grid(N), g["xd"], g["ud"], g["p"], g["iSx"], float(g["feas_tol"]), Nsub)
    M = N - 1
    col = lambda a, r, c: a.reshape(M, c, r).transpose(0, 2, 1)
    shp = {"A": (pb.nx, pb.nx), "Bm": (pb.nx, pb.nu), "Bp": (pb.nx, pb.nu), "F": (pb.nx, pb.np), "E": (pb.nx, pb.nx)}
    for b in range(2):
        for k in KEYS:
            got = col(out[k][b], *shp[k]) if k in shp else out[k][b]
            want = g[f"{k}_{b}"]
            assert np.abs(got - want).max() <= 1e-11 * max(np.abs(want).max(), 1e-300), (name, k)
        assert bool(out["feas"][b]) == bool(g[f"feas_{b}"])
    tc, xc, _ = handle.propagate(orc.t_grid(N), g["xd"], g["ud"], g["p"], 2 * Nsub * (N - 1))
    for b in range(2):
        assert np.abs(xc[b] - g[f"xc_{b}"]).max() <= 1e-8 * max(1.0, np.abs(g[f"xc_{b}"]).max())


@pytest.mark.gpu
def test_cuda_rocket_ptr_matches_golden(pkg, handle):
    g = np.load(os.path.join(GOLD, "oracle_ptr_rocket.npz"))
    ex = pkg.examples.rocket_landing
    traj = pkg.problem.TrajectoryProblem(ex.RocketProblem())
    ex.define_problem(traj, "ptr", handle=handle)
    pars = pkg.ptr.Parameters(N=12, Nsub=15, iter_max=20, disc_method=pkg.ptr.FOH, wvc=1e3, wtr=0.1, eps_abs=1e-5,
                              eps_rel=1e-4, feas_tol=1e-3, q_tr=np.inf, q_exit=np.inf)
    pbm = pkg.ptr.create(pars, traj, handle)
    sol = pkg.ptr.solve(pbm)
    pbm.close()
    Sx = pbm.scale.Sx
    assert sol.status[0] == "SCP_SOLVED" and abs(int(sol.iterations[0]) - int(g["iterations"])) <= 1
    assert np.abs((sol.xd[0] - g["xd"]) / Sx).max() <= 1e-4 and abs(sol.cost[0] - float(g["J"])) <= 1e-6
