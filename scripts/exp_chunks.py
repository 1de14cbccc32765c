"""Bench workload through scpb_ptr_solve with the streamed chains: SCPB_PTR_CHUNKS in argv, CUDA_DEVICE_MAX_CONNECTIONS
from the environment (it must be set before the CUDA context exists).  Usage: python scripts/exp_chunks.py 0 8 16 32"""
import sys; sys.path.insert(0, '.')
import os, json, time, numpy as np
import __graft_entry__ as g
pkg = g.load_package()
import bench
N, Nsub, B = 100, 100, int(os.environ.get("B", "256"))
h = pkg.Handle(0)
ex = pkg.examples.starship
mdl = ex.StarshipProblem(); traj = pkg.problem.TrajectoryProblem(mdl); ex.define_problem(traj, "ptr", handle=h)
PT = dict(bench.PTR); PT["iter_max"] = int(os.environ.get("ITER_MAX", str(PT["iter_max"])))
pars = pkg.ptr.Parameters(N=N, Nsub=Nsub, disc_method=pkg.ptr.FOH, q_tr=np.inf, q_exit=np.inf, **PT)
pbm = pkg.ptr.create(pars, traj, h)
X, U, Pp = bench.make_seeds(traj.guess(N), pbm.scale.Sx, pbm.scale.Su, B, 0, pbm.scale.cx, pbm.scale.cu)
ref = None
for ch in sys.argv[1:]:
    if ch.endswith("n"):            # "<chunks>n": warm start of the interior-point method off
        os.environ["SCPB_NO_WARM"] = "1"; ch = ch[:-1]
    else:
        os.environ.pop("SCPB_NO_WARM", None)
    os.environ["SCPB_PTR_CHUNKS"] = ch
    best = None
    for rep in range(2):
        sol = pkg.ptr.solve(pbm, (X, U, Pp))
        if best is None or sol.timing["total"] < best.timing["total"]:
            best = sol
    sol = best
    out = dict(k1_mb=os.environ.get("SCPB_K1_MB"), chunks=ch, warm=("SCPB_NO_WARM" not in os.environ), conn=os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"), total_s=round(sol.timing["total"], 4),
               it_per_s=round(float(sol.iterations.sum()) / sol.timing["total"], 1),
               timing={k: (round(v, 4) if isinstance(v, float) else v) for k, v in sol.timing.items()},
               iters=[int(sol.iterations.min()), float(np.median(sol.iterations)), int(sol.iterations.max())],
               solved=sol.status.count("SCP_SOLVED"))
    if ref is None:
        ref = sol
    else:
        out["same_iterations"] = bool((sol.iterations == ref.iterations).all())
        out["max_dx"] = float(np.abs((sol.xd - ref.xd) / pbm.scale.Sx).max())
    print(json.dumps(out), flush=True)
