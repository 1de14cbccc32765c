"""Generate tests/golden/oracle_ptr_bench_seeds.npz: the ORACLE's PTR solutions (oracle/ptr.py, its interior point at
1e-11) for the first 8 seeds of the bench workload (bench.make_seeds: starship PTR, N = 100, Nsub = 100, SURVEY 8(d)
perturbations).  tests/test_ptr_gpu.py::test_bench_configuration_parity compares scpb_ptr_solve with these seed by seed;
the fixture stores the seeds too, and the test falls back to running the oracle when its own seeds differ.
Like the other files in tests/golden these are fixtures of the CPU oracle, not outputs of the reference.
    python scripts/make_golden_bench.py          (about 10 minutes on 8 cores)
"""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import problems, ptr as optr  # noqa: E402

N, NSUB, NB, TOL = 100, 100, 8, 1e-12


def worker(args):
    hs, xd, ud, p = args
    import warnings
    warnings.filterwarnings("ignore")
    pb = problems.StarshipProblem(N); pb.hs = hs
    P = optr.PTR(pb, optr.Parameters(N=N, Nsub=NSUB, iter_max=15, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=0.01 / 100,
                                     feas_tol=5e-3, solver_tol=TOL))
    r = P.solve((xd, ud, p), prefer="ipm")
    s = r["sol"]
    return r["status"], r["iterations"], s.xd, s.ud, s.p, s.J_aug, bool(s.feas)


def main():
    pb = problems.StarshipProblem(N)
    sc = optr.Scaling(pb, N)
    X0, U0, P0 = bench.make_seeds(pb.guess(N), sc.Sx, sc.Su, NB, 0, sc.cx, sc.cu)
    with mp.get_context("fork").Pool(min(NB, os.cpu_count() or 1)) as pool:
        refs = pool.map(worker, [(pb.hs, X0[b], U0[b], P0[b]) for b in range(NB)], chunksize=1)
    d = dict(N=N, Nsub=NSUB, tol=TOL, X0=X0, U0=U0, P0=P0,
             status=np.array([r[0] for r in refs]), iterations=np.array([r[1] for r in refs]),
             xd=np.array([r[2] for r in refs]), ud=np.array([r[3] for r in refs]), p=np.array([r[4] for r in refs]),
             J_aug=np.array([r[5] for r in refs]), feas=np.array([r[6] for r in refs]))
    out = os.path.join(ROOT, "tests", "golden", "oracle_ptr_bench_seeds.npz")
    np.savez_compressed(out, **d)
    print("wrote", out, "iterations", d["iterations"], "status", d["status"])


if __name__ == "__main__":
    main()
