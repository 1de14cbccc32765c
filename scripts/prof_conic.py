import sys; sys.path.insert(0, '.')
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from tests import helpers
h = pkg.Handle(0)
N = int(sys.argv[1]); B = int(sys.argv[2])
pb, P, subs = helpers.starship_subproblems(N, 2, seed=N, Nsub=30)
Apat, Avals = helpers.union_pattern([s["cp"]["A"] for s in subs])
Gpat, Gvals = helpers.union_pattern([s["cp"]["G"] for s in subs])
lab = helpers.labels_from_program(subs[0]["prg"], N)
perm = pkg.ordering.stage_order(Apat, Gpat, lab, N)
cone = pkg.lib.ConeProblem(h, Apat, Gpat, subs[0]["cp"]["l"], [], perm=perm)
idx = np.arange(B) % 2
c = np.array([subs[i]["cp"]["c"] for i in idx]); b = np.array([subs[i]["cp"]["b"] for i in idx]); hh = np.array([subs[i]["cp"]["h"] for i in idx])
out = cone.solve(Avals[idx], Gvals[idx], c, b, hh)
print("ms", 1e3 * out["seconds"], out["iters"].max(), np.unique(out["status"]))
