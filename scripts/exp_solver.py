"""Bench workload (starship PTR, N=100, Nsub=100, B seeds): one full batched PTR solve per solver variant.
Prints per variant: wall seconds of the solve phase, lock-step iterations, interior-point iterations, statuses and
the cycle shares of CTA 0 (scpb_cone_info).  Usage: python scripts/exp_solver.py [N] [Nsub] [B] [variants...]"""
import sys; sys.path.insert(0, '.')
import os, json, time, numpy as np
import __graft_entry__ as g
pkg = g.load_package()
import bench
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
Nsub = int(sys.argv[2]) if len(sys.argv) > 2 else 100
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
h = pkg.Handle(0)
ex = pkg.examples.starship
mdl = ex.StarshipProblem(); traj = pkg.problem.TrajectoryProblem(mdl); ex.define_problem(traj, "ptr", handle=h)
P = dict(bench.PTR); P["iter_max"] = int(os.environ.get("ITER_MAX", str(P["iter_max"])))
pars = pkg.ptr.Parameters(N=N, Nsub=Nsub, disc_method=pkg.ptr.FOH, q_tr=np.inf, q_exit=np.inf, **P)
base = traj.guess(N)
variants = [("sn0 t1024", dict(SCPB_SUPERNODAL="0", SCPB_HYBRID="0"), dict()),
            ("sn1 t1024", dict(SCPB_SUPERNODAL="1", SCPB_HYBRID="0"), dict()),
            ("sn0 t512", dict(SCPB_SUPERNODAL="0", SCPB_HYBRID="0"), dict(threads=512)),
            ("sn0 tol1e-10", dict(SCPB_SUPERNODAL="0", SCPB_HYBRID="0"), dict(feastol=1e-10, abstol=1e-10, reltol=1e-10))]
for cut in range(2, 13):
    variants.append((f"hy{cut} t1024", dict(SCPB_SUPERNODAL="0", SCPB_HYBRID=str(cut)), dict()))
want = sys.argv[4:]
ref = None
seeds = None
for name, env, opts in variants:
    if want and not any(w == name.split()[0] or w == name for w in want):
        continue
    os.environ.update(env)        # SCPB_HYBRID is read when the cone problem is set up, SCPB_SUPERNODAL at every launch
    pbm = pkg.ptr.create(pars, traj, h)
    if seeds is None:
        seeds = bench.make_seeds(base, pbm.scale.Sx, pbm.scale.Su, B, 0, pbm.scale.cx, pbm.scale.cu)
    X, U, Pp = seeds
    best = None
    for rep in range(2):
        t = time.time()
        sol = pkg.ptr.solve(pbm, (X, U, Pp), **opts)
        dt = time.time() - t
        if best is None or sol.timing["solve"] < best[0].timing["solve"]:
            best = (sol, dt)
    sol, dt = best
    info = pbm.cone.info()
    cyc = info["cycles"]
    tot = max(1, cyc["total"])
    shares = {k: round(v / tot, 3) for k, v in cyc.items() if k not in ("total", "ldl_count", "factor_count", "factor_retries")}
    st = {s: sol.status.count(s) for s in set(sol.status)}
    out = dict(variant=name, wall_s=round(dt, 3), timing={k: (round(v, 4) if isinstance(v, float) else v) for k, v in sol.timing.items()},
               scp_iters=[int(sol.iterations.min()), float(np.median(sol.iterations)), int(sol.iterations.max())], status=st,
               ipm_per_solve=round(sol.timing["ipm_iterations"] / max(1, sol.iterations.sum()), 2),
               solves_per_ipm=round(cyc["ldl_count"] / max(1, cyc["factor_count"]), 2), shares=shares,
               retries_cta0=cyc["factor_retries"], ipm_its_cta0=cyc["factor_count"],
               ms_per_ipm_it_cta0=round(cyc["total"] / 1.965e6 / max(1, cyc["factor_count"]), 3), hybrid=info["hybrid"])
    if ref is None:
        ref = sol
    else:
        out["max_dx_vs_first"] = float(np.abs((sol.xd - ref.xd) / pbm.scale.Sx).max()); out["max_dJ_vs_first"] = float(np.abs(sol.cost - ref.cost).max())
        out["iters_equal_first"] = bool((sol.iterations == ref.iterations).all())
    print(json.dumps(out), flush=True)
    pbm.close()
