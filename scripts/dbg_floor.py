"""Residual floor of the cone solver on the starship N=12 LP batch (tests/test_conic_gpu.py case) for a library variant."""
import sys; sys.path.insert(0, '.')
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
from tests import helpers
N, nb = int(sys.argv[1]), 6
h = pkg.Handle(0)
pb, P, subs = helpers.starship_subproblems(N, nb, seed=N)
Apat, Avals = helpers.union_pattern([s["cp"]["A"] for s in subs])
Gpat, Gvals = helpers.union_pattern([s["cp"]["G"] for s in subs])
lab = helpers.labels_from_program(subs[0]["prg"], N)
perm = pkg.ordering.stage_order(Apat, Gpat, lab, N)
cone = pkg.lib.ConeProblem(h, Apat, Gpat, subs[0]["cp"]["l"], [], perm=perm)
c = np.array([s["cp"]["c"] for s in subs]); b = np.array([s["cp"]["b"] for s in subs]); hh = np.array([s["cp"]["h"] for s in subs])
for rep in range(3):
    out = cone.solve(Avals, Gvals, c, b, hh)
    res = []
    for k, sub in enumerate(subs):
        cp = sub["cp"]; x = out["x"][k]; y = out["y"][k]; z = out["z"][k]; s = out["s"][k]
        pres = max(np.abs(cp["A"] @ x - cp["b"]).max() / max(1, np.abs(cp["b"]).max()), np.abs(cp["G"] @ x + s - cp["h"]).max() / max(1, np.abs(cp["h"]).max()))
        dres = np.abs(cp["c"] + cp["A"].T @ y + cp["G"].T @ z).max() / max(1, np.abs(cp["c"]).max())
        gap = abs(out["pobj"][k] - out["dobj"][k]) / max(1, abs(out["pobj"][k]))
        res.append((pres, dres, gap))
    print("status", out["status"], "iters", out["iters"])
    print("  pres", ["%.1e" % r[0] for r in res], "dres", ["%.1e" % r[1] for r in res], "gap", ["%.1e" % r[2] for r in res])
