"""CPU study for round 2: how many dependent steps would a SUPERNODAL schedule of the bench KKT factorisation need?
(scalar level schedule today: 88 levels).  Fundamental supernodes = chains j -> parent(j) = j+1 with
struct(L[:, j]) = {j+1} + struct(L[:, j+1]); relaxed: also merge a child chain into its parent when the padding
(explicit zeros) stays below a bound."""
import sys; sys.path.insert(0, '.')
import numpy as np, scipy.sparse as sp
import __graft_entry__ as g
pkg = g.load_package()
sys.path.insert(0, 'proto')
from symbolic_proto import etree

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
FUND = int(sys.argv[2]) if len(sys.argv) > 2 else 1
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
L = sp.tril(M, format="csc")
children = [[] for _ in range(nk)]
for j in range(nk):
    if parent[j] >= 0: children[parent[j]].append(j)
struct = [None] * nk
for j in range(nk):
    s = set(L.indices[L.indptr[j]:L.indptr[j + 1]].tolist()); s.discard(j)
    for c in children[j]:
        s |= struct[c]
    s.discard(j)
    struct[j] = s
cnt = np.array([len(s) for s in struct])
height = np.zeros(nk, dtype=int)
for j in range(nk):
    for c in children[j]:
        height[j] = max(height[j], height[c] + 1)
print("scalar: nk", nk, "nnzL", cnt.sum(), "levels", height.max() + 1)
# fundamental supernodes
sn = np.arange(nk)           # supernode id = first column
first = []
j = 0
while j < nk:
    k = j
    while k + 1 < nk and parent[k] == k + 1 and (FUND == 0 or len(children[k + 1]) == 1) and cnt[k] == cnt[k + 1] + 1:
        k += 1
    first.append((j, k))
    j = k + 1
sid = np.zeros(nk, dtype=int)
for i, (a, b) in enumerate(first):
    sid[a:b + 1] = i
ns = len(first)
sparent = np.array([sid[parent[b]] if parent[b] >= 0 else -1 for (a, b) in first])
sh = np.zeros(ns, dtype=int)
for i in range(ns):
    if sparent[i] >= 0:
        sh[sparent[i]] = max(sh[sparent[i]], sh[i] + 1)
width = np.array([b - a + 1 for a, b in first])
rows = np.array([cnt[a] + 1 for a, b in first])     # rows of the supernode's dense panel (incl. diagonal block)
print("fundamental supernodes:", ns, "supernodal levels", sh.max() + 1, "max width", width.max(), "max panel rows", rows.max())
for lv in range(sh.max() + 1):
    idx = np.where(sh == lv)[0]
    print(f"  slevel {lv:3d}: {idx.size:5d} supernodes, width {width[idx].min()}..{width[idx].max()}, "
          f"panel rows {rows[idx].min()}..{rows[idx].max()}, columns {width[idx].sum()}")
flops = sum(w * r * r for w, r in zip(width, rows))
print("dense panel flops ~ sum w*r^2 =", flops, "(scalar factor ops x2 =", 2 * int((cnt * (cnt + 1) // 2).sum()), ")")
