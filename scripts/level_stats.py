"""CPU: level structure of the bench KKT factorisation (rows per level, L nnz per level, ops per level)."""
import sys; sys.path.insert(0, '.')
import numpy as np, scipy.sparse as sp
import __graft_entry__ as g
pkg = g.load_package()
sys.path.insert(0, 'proto')
from symbolic_proto import etree

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ex = pkg.examples.starship
mdl = ex.StarshipProblem(); mdl.hs = 100.0
traj = pkg.problem.TrajectoryProblem(mdl)
ex.define_problem(traj, "ptr", handle=None)
pars = pkg.ptr.Parameters(N=N, Nsub=100, iter_max=5, disc_method=pkg.ptr.FOH, wvc=1e3, wtr=0.1, eps_abs=1e-5,
                          eps_rel=1e-4, feas_tol=5e-3, q_tr=np.inf, q_exit=np.inf)
pkg.lib.ConeProblem = lambda *a, **k: type("C", (), {"c": None, "close": lambda s: None})()
class FH:
    def model_set(self, *a): pass
fake = FH(); fake.lib = type("L", (), {"scpb_ptr_setup": staticmethod(lambda *a: 0)})(); fake.h = None
fake._check = lambda rc, what: None
pbm = pkg.ptr.SCPProblem(pars, traj, fake, l1_block=4)
cp = pbm.cp
A, G = cp["A"], cp["G"]
n, p = A.shape[1], A.shape[0]
A1 = sp.csr_matrix((np.ones(A.nnz), A.indices, A.indptr), shape=A.shape)
G1 = sp.csr_matrix((np.ones(G.nnz), G.indices, G.indptr), shape=G.shape)
M = sp.bmat([[sp.eye(n) + G1.T @ G1, A1.T], [A1, sp.eye(p)]], format="csr")
perm = pbm.perm
M = M[perm][:, perm].tocsc()
nk = n + p
U = sp.triu(M, format="csc")
parent = etree(U.indptr, U.indices, nk)
# row structures of L via column merge
L = sp.tril(M, format="csc")
children = [[] for _ in range(nk)]
for j in range(nk):
    if parent[j] >= 0: children[parent[j]].append(j)
struct = [None] * nk
height = np.zeros(nk, dtype=int)
cols = []
for j in range(nk):
    s = set(L.indices[L.indptr[j]:L.indptr[j + 1]].tolist()); s.discard(j)
    for c in children[j]:
        s |= struct[c]; height[j] = max(height[j], height[c] + 1)
    s.discard(j)
    struct[j] = s
    cols.append(np.array(sorted(s), dtype=int))
    for c in children[j]:
        struct[c] = None
cnt = np.array([len(c) for c in cols])
rowlen = np.zeros(nk, dtype=int)
for j in range(nk):
    rowlen[cols[j]] += 1
print("nk", nk, "nnzL", cnt.sum(), "levels", height.max() + 1)
# factor ops per target row i (row-oriented): sum over k<i in row i of |{j in col k, j<=i...}| -- approximate by colcount based
ops_col = cnt * (cnt + 1) // 2
print("ops total (col-based)", ops_col.sum())
print("lvl rows  nnzLrow maxrow  colnnz maxcol  ops")
for l in range(height.max() + 1):
    idx = np.where(height == l)[0]
    print(f"{l:3d} {idx.size:5d} {rowlen[idx].sum():7d} {rowlen[idx].max():5d} {cnt[idx].sum():7d} {cnt[idx].max():5d} {ops_col[idx].sum():8d}")
