"""GPU, under compute-sanitizer: two identical nominal quadrotor GuSTO seeds, few iterations."""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import warnings; warnings.filterwarnings("ignore")
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from oracle import problems
import test_gusto_gpu as T
N = int(sys.argv[1]); K = int(sys.argv[2]); grp = int(sys.argv[3]); reps = int(sys.argv[4])
h = pkg.Handle(0)
mdl, traj, pars = T._setup(pkg, h, N, K)
pbo = problems.QuadrotorProblem(N)
gq = pbo.guess(N)
pbm = pkg.gusto.create(pars, traj, h)
X0 = np.array([gq[0], gq[0]]); U0 = np.array([gq[1], gq[1]]); P0 = np.array([gq[2], gq[2]], dtype=float)
for r in range(reps):
    sol = pkg.gusto.solve(pbm, (X0, U0, P0), group=grp, project_guess=False)
    print(sol.status, sol.iterations, sol.timing["ipm_iterations"], sol.cost, flush=True)
