"""GPU: how close is the batched PTR loop to the oracle PTR loop as a function of the subproblem tolerance?
python scripts/exp_parity.py N Nsub nb"""
import sys; sys.path.insert(0, '.')
import warnings; warnings.filterwarnings("ignore")
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from oracle import problems, ptr as optr
N, Nsub, nb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
h = pkg.Handle(0)
ex = pkg.examples.starship
mdl = ex.StarshipProblem(); traj = pkg.problem.TrajectoryProblem(mdl); ex.define_problem(traj, "ptr", handle=h)
pars = pkg.ptr.Parameters(N=N, Nsub=Nsub, iter_max=15, disc_method=pkg.ptr.FOH, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=1e-4,
                          feas_tol=5e-3, q_tr=np.inf, q_exit=np.inf, solver_opts={"verbose": 0, "maxit": 100})
pbo = problems.StarshipProblem(N); gq = pbo.guess(N); mdl.hs = pbo.hs
sc = optr.Scaling(pbo)
rng = np.random.default_rng(N)
X0 = np.array([gq[0] + (0.01 * sc.Sx * rng.standard_normal(gq[0].shape) if b else 0.0) for b in range(nb)])
U0 = np.array([gq[1] + (0.01 * sc.Su * rng.standard_normal(gq[1].shape) if b else 0.0) for b in range(nb)])
P0 = np.array([gq[2] * (1 + (0.02 * rng.uniform(-1, 1, gq[2].shape) if b else 0.0)) for b in range(nb)])
pbm = pkg.ptr.create(pars, traj, h)
refs = {}
for otol in (1e-9, 1e-11):
    P = optr.PTR(pbo, optr.Parameters(N=N, Nsub=Nsub, iter_max=15, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=1e-4, feas_tol=5e-3, solver_tol=otol))
    refs[otol] = [P.solve((X0[b], U0[b], P0[b]), prefer="ipm") for b in range(nb)]
for gtol in (1e-8, 1e-9, 1e-10, 1e-11):
    sol = pkg.ptr.solve(pbm, (X0, U0, P0), feastol=gtol, abstol=gtol, reltol=gtol)
    for otol in (1e-9, 1e-11):
        out = []
        for b in range(nb):
            rs = refs[otol][b]["sol"]
            ex7 = np.abs((sol.xd[b][:, :7] - rs.xd[:, :7]) / sc.Sx[:7]).max(); eu2 = np.abs((sol.ud[b][:, :2] - rs.ud[:, :2]) / sc.Su[:2]).max()
            ep = np.abs((sol.p[b] - rs.p) / sc.Sp).max(); dJ = abs(sol.cost[b] - rs.J_aug) / max(1, abs(rs.J_aug))
            out.append("%d/%d %.0e %.0e %.0e dJ %.0e" % (sol.iterations[b], refs[otol][b]["iterations"], ex7, eu2, ep, dJ))
        print("gpu tol", gtol, "oracle tol", otol, sol.status.count("SCP_SOLVED"), "|", " | ".join(out), flush=True)
    print("   ipm its/solve", sol.timing["ipm_iterations"] / max(1, sol.timing["lockstep_iterations"] * nb))
