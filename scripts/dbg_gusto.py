"""GPU: the C4 batch of tests/test_gusto_gpu.py with per-seed diagnostics; the first failing seed is re-run through the oracle."""
import sys; sys.path.insert(0, '.')
import warnings; warnings.filterwarnings("ignore")
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from oracle import gusto as ogusto, problems, ptr as optr
sys.path.insert(0, 'tests')
import test_gusto_gpu as T
N, nb = int(sys.argv[1]), int(sys.argv[2])
h = pkg.Handle(0)
mdl, traj, pars = T._setup(pkg, h, N, 30, eps=0.0)
pars.eps_abs, pars.eps_rel = 1e-5, 1e-4
pbo = problems.QuadrotorProblem(N)
gq = pbo.guess(N)
import bench
X0, U0, P0 = bench.make_seeds_c4(gq, nb, 0, pbo.r0, pbo.rf)
pbm = pkg.gusto.create(pars, traj, h)
kw = {}
if len(sys.argv) > 3:
    t = float(sys.argv[3]); kw = dict(feastol=t, abstol=t, reltol=t)
sol = pkg.gusto.solve(pbm, (X0, U0, P0), **kw)
for b in range(nb):
    print(b, "p0 %.3f" % P0[b][0], sol.status[b], "it", sol.iterations[b], "eta %.3g lam %.3g J %.6g feas %d" % (sol.eta[b], sol.lam[b], sol.cost[b], sol.feas[b]))
print(sol.timing)
bad = [b for b in range(nb) if sol.status[b] != "SCP_SOLVED" and P0[b][0] > 1.2]
print("solved", sum(s_ == "SCP_SOLVED" for s_ in sol.status), "of", nb)
if bad and len(sys.argv) > 4:
    b = bad[0]
    P = ogusto.GuSTO(pbo, ogusto.Parameters(N=N, Nsub=15, iter_max=30, eps_abs=1e-5, eps_rel=1e-4, solver_tol=1e-9, **T.KW))
    guess = optr.correct_convex(pbo, P.scale, N, X0[b], U0[b], P0[b], tol=1e-10)
    ref = P.solve(guess, verbose=True)
    print("oracle on seed", b, ref["status"], ref["iterations"], ref["eta"], ref["lam"])
