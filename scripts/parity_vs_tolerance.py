"""CPU only: how much does the ORACLE's own PTR result at the bench configuration (starship, N = 100, Nsub = 100, SURVEY 8(d)
seeds) move when only the subproblem-solver tolerance changes?  This is the floor under any cross-implementation
trajectory comparison at that size: two exact-arithmetic-equivalent solvers stopped at tolerance tol agree no better than
the oracle agrees with itself at tol vs tol/10.   python scripts/parity_vs_tolerance.py [nb] > profiles/r2_parity_vs_tolerance.txt"""
import sys; sys.path.insert(0, '.')
import multiprocessing as mp
import warnings; warnings.filterwarnings("ignore")
import numpy as np


def work(a):
    import warnings; warnings.filterwarnings("ignore")
    from oracle import problems as pr, ptr as op
    N, Nsub, hs, xd, ud, p, tol = a
    pb = pr.StarshipProblem(N); pb.hs = hs
    P = op.PTR(pb, op.Parameters(N=N, Nsub=Nsub, iter_max=15, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=0.01 / 100,
                                 feas_tol=5e-3, solver_tol=tol))
    r = P.solve((xd, ud, p), prefer="ipm")
    s = r["sol"]
    return r["status"], r["iterations"], s.xd, s.ud, s.p, s.J_aug


if __name__ == "__main__":
    from oracle import problems, ptr as optr
    import bench
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    N, Nsub = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (100, 100)
    pbo = problems.StarshipProblem(N)
    g = pbo.guess(N)
    sc = optr.Scaling(pbo)
    X0, U0, P0 = bench.make_seeds(g, sc.Sx, sc.Su, nb, 0, sc.cx, sc.cu)
    tols = [1e-8, 1e-9, 1e-10, 1e-11, 1e-12]
    tasks = [(N, Nsub, pbo.hs, X0[b], U0[b], P0[b], t) for t in tols for b in range(nb)]
    with mp.get_context("fork").Pool(min(len(tasks), mp.cpu_count())) as pool:
        res = pool.map(work, tasks, chunksize=1)
    R = {(t, b): res[i * nb + b] for i, t in enumerate(tols) for b in range(nb)}
    print(f"oracle PTR vs oracle PTR, starship N={N} Nsub={Nsub}, bench seeds 0..{nb - 1}; reference = tol 1e-12")
    print("columns: seed: iterations(tol)/iterations(1e-12)  max|dx/Sx| (x[0:7])  max|du/Su| (T, delta)  max|dp/Sp|  |dJ|/max(1,|J|)")
    for t in tols[:-1]:
        row = []
        for b in range(nb):
            st, it, xd, ud, p, J = R[(t, b)]
            st0, it0, xd0, ud0, p0, J0 = R[(1e-12, b)]
            ex = np.abs((xd[:, :7] - xd0[:, :7]) / sc.Sx[:7]).max(); eu = np.abs((ud[:, :2] - ud0[:, :2]) / sc.Su[:2]).max()
            ep = np.abs((p - p0) / sc.Sp).max(); dJ = abs(J - J0) / max(1, abs(J0))
            row.append(f"{b}: {it}/{it0} {ex:.1e} {eu:.1e} {ep:.1e} {dJ:.1e}")
        print(f"tol {t:.0e} | " + " | ".join(row))
