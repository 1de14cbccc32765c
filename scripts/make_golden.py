"""Generate tests/golden/*.npz: small fixed inputs and the ORACLE's outputs for them.

These are regression fixtures of the CPU oracle (oracle/), not outputs of the reference: SCPToolbox.jl ships no numeric
fixtures and neither Julia nor ECOS can run in this image ("parity unpinned").  They pin the oracle
against accidental change between rounds and give the CUDA path a second, file-based target.
    python scripts/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import orc, problems, ptr as optr, scvx as oscvx  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = [("dblint", 9, 10), ("rocket", 8, 15), ("starship", 9, 120), ("quadrotor", 10, 15), ("freeflyer", 6, 15)]


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, N, Nsub in CASES:
        pb = problems.make_problem(name, N)
        m = pb.orc_model()
        xd, ud, p = problems.test_trajectory(pb, 2, N, seed=11)
        iS = np.ones(pb.nx)
        d = {"N": N, "Nsub": Nsub, "xd": xd, "ud": ud, "p": p, "iSx": iS, "feas_tol": 1e-3}
        for b in range(2):
            r = orc.discretize(m, xd[b], ud[b], p[b], Nsub, iS, 1e-3)
            for k in ("A", "Bm", "Bp", "F", "r", "E", "defect"):
                d[f"{k}_{b}"] = getattr(r, k)
            d[f"feas_{b}"] = np.array(r.feas)
            d[f"xc_{b}"] = orc.propagate(m, xd[b], ud[b], p[b], 2 * Nsub * (N - 1))
        np.savez_compressed(os.path.join(OUT, f"oracle_discretize_{name}.npz"), **d)
    # whole-loop fixtures: double-integrator minimum time (closed-form answer next to it) and three SCvx iterations
    N = 30
    pb = problems.DoubleIntegratorProblem(N, 1)
    pars = optr.Parameters(N=N, Nsub=10, iter_max=30, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=1e-4, feas_tol=1e-3,
                           solver_tol=1e-9)
    out = optr.PTR(pb, pars).solve(pb.guess(N))
    s = out["sol"]
    np.savez_compressed(os.path.join(OUT, "oracle_ptr_dblint.npz"), xd=s.xd, ud=s.ud, p=s.p, J=s.J_aug,
                        iterations=out["iterations"], t_opt=np.array(pb.t_opt()))
    N = 12
    pb = problems.RocketProblem(N)
    pars = optr.Parameters(N=N, Nsub=15, iter_max=20, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=1e-4, feas_tol=1e-3,
                           solver_tol=1e-9)
    out = optr.PTR(pb, pars).solve(pb.guess(N), prefer="ipm")
    s = out["sol"]
    np.savez_compressed(os.path.join(OUT, "oracle_ptr_rocket.npz"), xd=s.xd, ud=s.ud, p=s.p, J=s.J_aug,
                        iterations=out["iterations"])


if __name__ == "__main__":
    main()
