"""GPU: does a failing seed disturb its group partner?  [nominal, tdil = 0.8] in one CTA (group = 2) vs separate CTAs (group = 1)."""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import warnings; warnings.filterwarnings("ignore")
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from oracle import problems
import test_gusto_gpu as T
N = 30
h = pkg.Handle(0)
mdl, traj, pars = T._setup(pkg, h, N, 15)
pbo = problems.QuadrotorProblem(N)
gq = pbo.guess(N)
pbm = pkg.gusto.create(pars, traj, h)
for name, P0 in (("nominal+0.8", [gq[2], [0.8]]), ("nominal+nominal", [gq[2], gq[2]]), ("0.8+nominal", [[0.8], gq[2]]), ("nominal+1.0", [gq[2], [1.0]])):
    X0 = np.array([gq[0], gq[0]]); U0 = np.array([gq[1], gq[1]]); P0 = np.array(P0, dtype=float)
    for grp in (2, 1):
        for tol in (1e-8, 1e-10):
            sol = pkg.gusto.solve(pbm, (X0, U0, P0), group=grp, feastol=tol, abstol=tol, reltol=tol)
            print(name, "group", grp, "tol", tol, sol.status, sol.iterations, sol.timing["ipm_iterations"], flush=True)
