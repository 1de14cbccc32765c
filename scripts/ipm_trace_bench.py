"""GPU diagnostic: interior-point residual trace (SCPB_IPM_TRACE) of bench subproblems (starship PTR, N=100, Nsub=100).
For each (seed, PTR iteration k) pair the batched PTR is run with iter_max = k, so the traced launch is the k-th subproblem.
Usage: python scripts/ipm_trace_bench.py [B] [seed:k ...]"""
import sys; sys.path.insert(0, '.')
import os, numpy as np
import __graft_entry__ as g
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pairs = [tuple(int(v) for v in a.split(":")) for a in sys.argv[2:]] or [(1, 1), (1, 4), (3, 2)]
opts = eval(os.environ.get('CONE_OPTS', '{}'))
for seed, k in pairs:
    os.environ["SCPB_IPM_TRACE"] = str(seed)
    pkg = g.load_package()
    h = pkg.Handle(0)
    ex = pkg.examples.starship
    mdl = ex.StarshipProblem(); traj = pkg.problem.TrajectoryProblem(mdl); ex.define_problem(traj, "ptr", handle=h)
    P = dict(bench.PTR); P["iter_max"] = k
    pars = pkg.ptr.Parameters(N=100, Nsub=100, disc_method=pkg.ptr.FOH, q_tr=np.inf, q_exit=np.inf, **P)
    pbm = pkg.ptr.create(pars, traj, h)
    X, U, Pp = bench.make_seeds(traj.guess(100), pbm.scale.Sx, pbm.scale.Su, B, 0, pbm.scale.cx, pbm.scale.cu)
    sol = pkg.ptr.solve(pbm, (X, U, Pp), **opts)
    tr = pbm.cone.ipm_trace()
    print(f"== seed {seed} PTR subproblem {k}: ipm_iterations(all seeds) {sol.timing['ipm_iterations']}  rows {len(tr)}")
    print("  it      pres      dres       gap         pcost         dcost  alpha_p  alpha_d     delta  sigma*mu")
    for r in tr:
        print(f"{int(r[0]):4d} {r[1]:9.2e} {r[2]:9.2e} {r[3]:9.2e} {r[4]:13.6e} {r[5]:13.6e} {r[6]:8.4f} {r[7]:8.4f} {r[8]:9.1e} {r[9]:9.2e}")
    pbm.close()
