import sys, time; sys.path.insert(0, '.')
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from oracle import problems, ptr as optr
h = pkg.Handle(0)
N, Nsub, nb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ex = pkg.examples.starship
mdl = ex.StarshipProblem(); traj = pkg.problem.TrajectoryProblem(mdl); ex.define_problem(traj, "ptr", handle=h)
pars = pkg.ptr.Parameters(N=N, Nsub=Nsub, iter_max=15, disc_method=pkg.ptr.FOH, wvc=1e3, wtr=0.1, eps_abs=1e-5,
                          eps_rel=0.01/100, feas_tol=5e-3, q_tr=np.inf, q_exit=np.inf)
t0 = time.time(); xg, ug, pg = traj.guess(N); print("gpu guess", round(time.time()-t0,2), "t1,t2", pg[:2], "hs", mdl.hs, flush=True)
pbo = problems.StarshipProblem(N); go = pbo.guess(N)
print("guess diff", np.abs(xg-go[0]).max(), np.abs(ug-go[1]).max(), np.abs(pg-go[2]).max(), flush=True)
mdl.hs = pbo.hs
t0 = time.time(); pbm = pkg.ptr.create(pars, traj, h); print("create", round(time.time()-t0,2), pbm.cone.info(), flush=True)
rng = np.random.default_rng(0)
sc = pbm.scale
X0 = np.array([go[0] + (0.01*sc.Sx*rng.standard_normal(go[0].shape) if b else 0.0) for b in range(nb)])
U0 = np.array([go[1] + (0.01*sc.Su*rng.standard_normal(go[1].shape) if b else 0.0) for b in range(nb)])
P0 = np.array([go[2]*(1+(0.02*rng.uniform(-1,1,go[2].shape) if b else 0.0)) for b in range(nb)])
import os
opts = eval(os.environ.get('CONE_OPTS', '{}'))
sol = pkg.ptr.solve(pbm, (X0, U0, P0), **opts)
print("opts", opts, "status", np.unique(sol.raw_status, return_counts=True), "iters", np.unique(sol.iterations, return_counts=True), "J", sol.cost[:4], "dev", sol.deviation[:4], "feas", sol.feas[:8], flush=True)
print("timing", sol.timing, flush=True)
sol = pkg.ptr.solve(pbm, (X0, U0, P0), **opts)
print("timing2", sol.timing, "SCP it/s", sol.iterations.sum()/sol.timing["total"], flush=True)
if nb <= 8:
    opars = optr.Parameters(N=N, Nsub=Nsub, iter_max=15, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=0.01/100, feas_tol=5e-3, solver_tol=1e-9)
    P = optr.PTR(pbo, opars)
    for b in range(min(nb, 3)):
        t0 = time.time(); ref = P.solve((X0[b], U0[b], P0[b]), prefer="ipm"); rs = ref["sol"]
        print(b, "oracle", ref["status"], ref["iterations"], rs.J_aug, "gpu J", sol.cost[b], "ex", np.abs((sol.xd[b]-rs.xd)/sc.Sx).max(),
              "eu", np.abs((sol.ud[b]-rs.ud)/sc.Su).max(), "ep", np.abs((sol.p[b]-rs.p)/sc.Sp).max(), round(time.time()-t0,1), flush=True)
