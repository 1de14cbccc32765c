"""Timing/accuracy probe of solver options on the product's own PTR template (starship)."""
import sys, time; sys.path.insert(0, '.')
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from tests import helpers
from oracle import conic
h = pkg.Handle(0)
N = int(sys.argv[1]); B = int(sys.argv[2])
pb, P, subs = helpers.starship_subproblems(N, 2, seed=N, Nsub=30 if N == 100 else 100)
Apat, Avals = helpers.union_pattern([s["cp"]["A"] for s in subs])
Gpat, Gvals = helpers.union_pattern([s["cp"]["G"] for s in subs])
lab = helpers.labels_from_program(subs[0]["prg"], N)
perm = pkg.ordering.stage_order(Apat, Gpat, lab, N)
cone = pkg.lib.ConeProblem(h, Apat, Gpat, subs[0]["cp"]["l"], [], perm=perm)
print("N", N, cone.info(), flush=True)
ref = [conic.solve_highs(s["cp"]) for s in subs]
idx = np.arange(B) % 2
c = np.array([subs[i]["cp"]["c"] for i in idx]); b = np.array([subs[i]["cp"]["b"] for i in idx]); hh = np.array([subs[i]["cp"]["h"] for i in idx])
for kw in (dict(nref=3), dict(nref=2), dict(nref=1), dict(nref=1, delta_dyn=1e-8), dict(nref=0), dict(nref=1, group=1), dict(nref=1, group=4)):
    out = cone.solve(Avals[idx], Gvals[idx], c, b, hh, **kw)
    err = max(abs(out["pobj"][k] - (ref[idx[k]]["obj"] - subs[idx[k]]["cp"]["c0"])) for k in range(B))
    print(f"  {kw}: {1e3*out['seconds']:8.2f} ms iters {out['iters'].min()}-{out['iters'].max()} status {np.unique(out['status'])} objerr {err:.2e}", flush=True)
