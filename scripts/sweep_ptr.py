"""Sweep solver launch options on the bench workload (one process, one template)."""
import sys; sys.path.insert(0, '.')
import os, numpy as np
import __graft_entry__ as g
pkg = g.load_package()
import bench
N, Nsub, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
h = pkg.Handle(0)
ex = pkg.examples.starship
mdl = ex.StarshipProblem(); traj = pkg.problem.TrajectoryProblem(mdl); ex.define_problem(traj, "ptr", handle=h)
pars = pkg.ptr.Parameters(N=N, Nsub=Nsub, disc_method=pkg.ptr.FOH, q_tr=np.inf, q_exit=np.inf, **bench.PTR)
base = traj.guess(N)
pbm = pkg.ptr.create(pars, traj, h)
print(pbm.cone.info(), flush=True)
X, U, Pp = bench.make_seeds(base, pbm.scale.Sx, pbm.scale.Su, B, 0)
for opts in eval(sys.argv[4]):
    sol = pkg.ptr.solve(pbm, (X, U, Pp), **opts)
    sol = pkg.ptr.solve(pbm, (X, U, Pp), **opts)
    t = sol.timing
    print(opts, "solved", sum(s == "SCP_SOLVED" for s in sol.status), "iters", int(sol.iterations.sum()), "ipm", t["ipm_iterations"],
          f"solve {t['solve']:.3f}s total {t['total']:.3f}s  it/s {sol.iterations.sum()/t['total']:.0f}  J0 {sol.cost[0]:.9f}", flush=True)
    cy = pbm.cone.info()["cycles"]; tot = max(cy["total"], 1)
    print("   last-launch cycle shares:", {k: round(100 * v / tot, 1) for k, v in cy.items() if k != "ldl_count"}, "total Mcyc", round(tot / 1e6, 1), "ldl solves", cy["ldl_count"], flush=True)
