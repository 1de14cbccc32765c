"""GPU: every subproblem the ORACLE GuSTO loop builds for one C4 seed is also handed to the GPU cone solver (same data,
RCM ordering); prints both solvers' status / iterations / objective.  python scripts/dbg_gusto_cone.py N seed [opts k=v ...]"""
import sys; sys.path.insert(0, '.')
import warnings; warnings.filterwarnings("ignore")
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
import bench
from oracle import gusto as ogusto, problems, ptr as optr, conic
N, b = int(sys.argv[1]), int(sys.argv[2])
opts = {}
for a in sys.argv[3:]:
    k, v = a.split("="); opts[k] = float(v) if "." in v or "e" in v else int(v)
h = pkg.Handle(0)
pb = problems.QuadrotorProblem(N); gq = pb.guess(N)
X, U, P0 = bench.make_seeds_c4(gq, b + 1, 0, pb.r0, pb.rf)
c = dict(bench.GUSTO)
P = ogusto.GuSTO(pb, ogusto.Parameters(N=N, Nsub=15, solver_tol=1e-9, **c))
guess = optr.correct_convex(pb, P.scale, N, X[b], U[b], P0[b], tol=1e-9)
pars = P.pars
lam, eta = pars.lam_init, pars.eta_init
ref = P.make_solution(*guess)
for k in range(1, pars.iter_max + 1):
    prg, hh = P.build(ref, lam, eta)
    cp = prg.compile()
    A, G = cp["A"].tocsr(), cp["G"].tocsr()
    A.sort_indices(); G.sort_indices()
    cone = pkg.lib.ConeProblem(h, A, G, cp["l"], cp["q"], perm=pkg.ordering.rcm_order(A, G))
    out = cone.solve(A.data[None], G.data[None], cp["c"][None], cp["b"][None], cp["h"][None], **opts)
    inf = cone.info()
    cone.close()
    sol, _, res = P.solve_subproblem(ref, lam, eta)
    print(f"it {k} lam {lam:.2e} eta {eta:.3g} | oracle {res['status']} {res['iters']} obj {res['obj']:.9e} | gpu "
          f"{pkg.lib.CONE_STATUS[int(out['status'][0])]} {int(out['iters'][0])} pobj {out['pobj'][0] + cp['c0']:.9e} dobj {out['dobj'][0] + cp['c0']:.9e} "
          f"retries {inf['cycles']['factor_retries']}", flush=True)
    if sol.status not in ("OPTIMAL", "ALMOST_OPTIMAL"):
        break
    sol.lam, sol.eta = lam, eta
    if P.check_stop(k, ref, sol, lam):
        print("stop"); break
    sol.rho, sol.cost_error, sol.dyn_error = P.rho(ref, sol)
    ref, eta, lam, rej = P.update_rule(k, ref, sol, lam, eta)
    print(f"    rho {sol.rho:.3e} rej {rej} J_aug {sol.J_aug:.6e} dev {sol.deviation:.3g}")
