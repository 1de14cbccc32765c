import sys, time; sys.path.insert(0, '.')
import numpy as np, scipy.sparse as sp
import __graft_entry__ as g
pkg = g.load_package()
from oracle import conic
from tests.test_conic_gpu import _feasible_program
h = pkg.Handle(0)
for (seed, n, p, l, soc) in [(0, 10, 3, 8, []), (1, 16, 5, 10, [3, 4])]:
    rng = np.random.default_rng(seed)
    A, G, l2, c0, b0, h0 = _feasible_program(np.random.default_rng(1000 * seed + 7), n, p, l, soc)
    t0 = time.time()
    cone = pkg.lib.ConeProblem(h, A, G, l2, soc, perm=pkg.ordering.rcm_order(A, G))
    t1 = time.time()
    for maxit in (3, 6, 9, 10, 11, 12, 20):
        out = cone.solve(A.data[None], G.data[None], c0[None], b0[None], h0[None], maxit=maxit)
        print(seed, "maxit", maxit, "status", out["status"], "iters", out["iters"], "pobj", out["pobj"], "dobj", out["dobj"], "sec", out["seconds"], "setup", t1 - t0, flush=True)
    cp = dict(c=c0, c0=0.0, A=A, b=b0, G=G, h=h0, l=l2, q=list(soc))
    ref = conic.solve_ipm(cp, tol=1e-9)
    print("oracle", ref["status"], ref["obj"], ref["iters"])
