"""One batched PTR iteration (1 k_ipm_solve launch) on the bench workload, for ncu."""
import sys; sys.path.insert(0, '.')
import os, numpy as np
import __graft_entry__ as g
pkg = g.load_package()
import bench
N, Nsub, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
h = pkg.Handle(0)
ex = pkg.examples.starship
mdl = ex.StarshipProblem(); traj = pkg.problem.TrajectoryProblem(mdl); ex.define_problem(traj, "ptr", handle=h)
P = dict(bench.PTR); P["iter_max"] = int(os.environ.get("ITER_MAX", "1"))
pars = pkg.ptr.Parameters(N=N, Nsub=Nsub, disc_method=pkg.ptr.FOH, q_tr=np.inf, q_exit=np.inf, **P)
base = traj.guess(N)
pbm = pkg.ptr.create(pars, traj, h)
X, U, Pp = bench.make_seeds(base, pbm.scale.Sx, pbm.scale.Su, B, 0, pbm.scale.cx, pbm.scale.cu)
opts = eval(os.environ.get('CONE_OPTS', '{}'))
sol = pkg.ptr.solve(pbm, (X, U, Pp), **opts)
print("timing", sol.timing)
