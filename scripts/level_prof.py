"""Per-level cycle profile of k_ipm_solve (CTA 0) on the bench workload: SCPB_LEVEL_PROFILE=1 python scripts/level_prof.py 100 100 256"""
import sys; sys.path.insert(0, '.')
import os, json, numpy as np
os.environ["SCPB_LEVEL_PROFILE"] = "1"
import __graft_entry__ as g
pkg = g.load_package()
import bench
N, Nsub, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
h = pkg.Handle(0)
ex = pkg.examples.starship
mdl = ex.StarshipProblem(); traj = pkg.problem.TrajectoryProblem(mdl); ex.define_problem(traj, "ptr", handle=h)
P = dict(bench.PTR); P["iter_max"] = 1
pars = pkg.ptr.Parameters(N=N, Nsub=Nsub, disc_method=pkg.ptr.FOH, q_tr=np.inf, q_exit=np.inf, **P)
base = traj.guess(N)
pbm = pkg.ptr.create(pars, traj, h)
X, U, Pp = bench.make_seeds(base, pbm.scale.Sx, pbm.scale.Su, B, 0, pbm.scale.cx, pbm.scale.cu)
opts = eval(os.environ.get('CONE_OPTS', '{}'))
sol = pkg.ptr.solve(pbm, (X, U, Pp), **opts)
info = pbm.cone.info()
lp = pbm.cone.level_profile()
cyc = info["cycles"]
print("cycles", cyc)
nf, ns = max(cyc["factor_count"], 1), max(cyc["ldl_count"], 1)
print("lvl  factor_us  fw_us  bw_us   (per call, 1.965 GHz)")
for l in range(lp.shape[1] if not info["hybrid"]["cut_used"] else 0):
    print(f"{l:3d} {lp[0, l] / nf / 1965:9.2f} {lp[1, l] / ns / 1965:7.2f} {lp[2, l] / ns / 1965:7.2f}")
hy = info["hybrid"]
if hy["cut_used"]:   # hybrid program: [3][scalar levels incl. bridge] then [3][top levels], both packed at the front
    raw = lp.reshape(-1)
    nl, ntl = hy["levels"], hy["top_levels"]
    sc = raw[:3 * nl].reshape(3, nl); tp = raw[3 * nl:3 * (nl + ntl)].reshape(3, ntl)
    print("hybrid cut", hy["cut_used"], "scalar levels (last = bridge)", nl, "top levels", ntl)
    for l in range(nl):
        print(f"s{l:3d} {sc[0, l] / nf / 1965:9.2f} {sc[1, l] / ns / 1965:7.2f} {sc[2, l] / ns / 1965:7.2f}")
    for l in range(ntl):
        print(f"t{l:3d} {tp[0, l] / nf / 1965:9.2f} {tp[1, l] / ns / 1965:7.2f} {tp[2, l] / ns / 1965:7.2f}")
    print("sum scalar", sc[0].sum() / nf / 1965, sc[1].sum() / ns / 1965, sc[2].sum() / ns / 1965,
          "top", tp[0].sum() / nf / 1965, tp[1].sum() / ns / 1965, tp[2].sum() / ns / 1965)
else:
    print("sum", lp[0].sum() / nf / 1965, lp[1].sum() / ns / 1965, lp[2].sum() / ns / 1965)
