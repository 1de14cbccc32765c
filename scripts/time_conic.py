"""Timing probe: batched IPM on the starship PTR subproblem replicated over B seeds."""
import sys, time; sys.path.insert(0, '.')
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from tests import helpers
from oracle import conic
h = pkg.Handle(0)
for N in (31, 100):
    pb, P, subs = helpers.starship_subproblems(N, 2, seed=N, Nsub=30 if N == 100 else 100)
    Apat, Avals = helpers.union_pattern([s["cp"]["A"] for s in subs])
    Gpat, Gvals = helpers.union_pattern([s["cp"]["G"] for s in subs])
    lab = helpers.labels_from_program(subs[0]["prg"], N)
    t0 = time.time(); perm = pkg.ordering.stage_order(Apat, Gpat, lab, N); t1 = time.time()
    cone = pkg.lib.ConeProblem(h, Apat, Gpat, subs[0]["cp"]["l"], [], perm=perm); t2 = time.time()
    print("N", N, "info", cone.info(), "order s", round(t1 - t0, 2), "setup s", round(t2 - t1, 2), flush=True)
    ref = [conic.solve_highs(s["cp"]) for s in subs]
    for B in (2, 256):
        idx = np.arange(B) % 2
        c = np.array([subs[i]["cp"]["c"] for i in idx]); b = np.array([subs[i]["cp"]["b"] for i in idx]); hh = np.array([subs[i]["cp"]["h"] for i in idx])
        for G in ([1, 2, 4] if B >= 64 else [1]):
            if B // G > 1200: continue
            out = cone.solve(Avals[idx], Gvals[idx], c, b, hh, group=G)
            out = cone.solve(Avals[idx], Gvals[idx], c, b, hh, group=G)
            err = max(abs(out["pobj"][k] - (ref[idx[k]]["obj"] - subs[idx[k]]["cp"]["c0"])) for k in range(B))
            print(f"  B {B:5d} G {G} solve {1e3*out['seconds']:8.2f} ms  iters {out['iters'].min()}-{out['iters'].max()} status {np.unique(out['status'])} objerr {err:.2e}", flush=True)
    cone.close()
