"""GPU: IPM residual trace (SCPB_IPM_TRACE) of the first quadrotor GuSTO subproblem, repeated until a run fails."""
import os, sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
os.environ["SCPB_IPM_TRACE"] = "0"
import warnings; warnings.filterwarnings("ignore")
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from oracle import problems
import test_gusto_gpu as T
N = int(sys.argv[1]); reps = int(sys.argv[2])
h = pkg.Handle(0)
mdl, traj, pars = T._setup(pkg, h, N, 1)
pbo = problems.QuadrotorProblem(N)
gq = pbo.guess(N)
pbm = pkg.gusto.create(pars, traj, h)
X0 = np.array([gq[0]]); U0 = np.array([gq[1]]); P0 = np.array([gq[2]], dtype=float)
np.set_printoptions(linewidth=250, precision=3)
shown_ok = False
for r in range(reps):
    sol = pkg.gusto.solve(pbm, (X0, U0, P0), group=1, project_guess=False)
    tr = pbm.cone.ipm_trace()
    ok = sol.status[0] == "SCP_SOLVED"
    print(r, sol.status, sol.timing["ipm_iterations"], flush=True)
    if (not ok) or (not shown_ok):
        shown_ok = shown_ok or ok
        print("   it      pres      dres       gap         pcost         dcost    ap    ad   delta   sigmu")
        for row in tr:
            print("  %3d %9.2e %9.2e %9.2e %13.6e %13.6e %5.3f %5.3f %7.1e %8.1e" % tuple(row))
    if not ok:
        break
