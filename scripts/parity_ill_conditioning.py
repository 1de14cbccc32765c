"""How well do two CONVERGED interior-point solves determine the bench-configuration trajectory?  The oracle PTR (N = 100,
Nsub = 100, bench seeds) is run twice at tolerance 1e-12 -- with and without Ruiz equilibration in its interior point, i.e.
two different but equally valid solvers of the same subproblems -- and the final trajectories are compared in the units of
tests/test_ptr_gpu.py::test_bench_configuration_parity.  Usage: python scripts/parity_ill_conditioning.py [seed ...]"""
import multiprocessing as mp
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import conic, problems, ptr as optr  # noqa: E402

N, NSUB, TOL = 100, 100, 1e-12


def worker(args):
    b, equil, hs, xd, ud, p = args
    warnings.filterwarnings("ignore")
    if not equil and not getattr(conic, "_noequil", False):      # variant B: the same interior point on the unequilibrated program
        orig = conic.equilibrate
        conic.equilibrate = lambda cp, iters=5, _o=orig: _o(cp, iters=0)
        conic._noequil = True
    pb = problems.StarshipProblem(N); pb.hs = hs
    P = optr.PTR(pb, optr.Parameters(N=N, Nsub=NSUB, iter_max=15, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=0.01 / 100,
                                     feas_tol=5e-3, solver_tol=TOL))
    r = P.solve((xd, ud, p), prefer="ipm")
    s = r["sol"]
    return b, equil, r["status"], r["iterations"], s.xd, s.ud, s.p, s.J_aug


def main():
    seeds = [int(a) for a in sys.argv[1:]] or [1, 2, 5, 7]
    pb = problems.StarshipProblem(N)
    sc = optr.Scaling(pb, N)
    X0, U0, P0 = bench.make_seeds(pb.guess(N), sc.Sx, sc.Su, 8, 0, sc.cx, sc.cu)
    jobs = [(b, e, pb.hs, X0[b], U0[b], P0[b]) for b in seeds for e in (True, False)]
    with mp.get_context("fork").Pool(min(len(jobs), os.cpu_count() or 1), maxtasksperchild=1) as pool:
        out = pool.map(worker, jobs, chunksize=1)
    res = {(b, e): r for (b, e, *r) in out}
    print("oracle PTR (equilibrated interior point) vs oracle PTR (no equilibration), both at 1e-12, N=100 bench seeds")
    print("seed  iterations  states(phys)  inputs(T,delta)  parameters  J_aug(rel)")
    for b in seeds:
        st1, it1, x1, u1, p1, J1 = res[(b, True)]
        st2, it2, x2, u2, p2, J2 = res[(b, False)]
        ex = np.abs((x1[:, :7] - x2[:, :7]) / sc.Sx[:7]).max(); eu = np.abs((u1[:, :2] - u2[:, :2]) / sc.Su[:2]).max()
        ep = np.abs((p1 - p2) / sc.Sp).max(); dJ = abs(J1 - J2) / max(1.0, abs(J1))
        print(f"{b:4d}  {it1:3d}/{it2:3d} {st1[:10]:>10s}  {ex:9.2e}  {eu:9.2e}  {ep:9.2e}  {dJ:9.2e}")


if __name__ == "__main__":
    main()
