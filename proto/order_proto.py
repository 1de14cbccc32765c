import numpy as np, scipy.sparse as sp, pickle, sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/proto')
from symbolic_proto import symbolic
from oracle import problems, ptr

def labels_from_program(prg, N):
    lab = np.full(prg.nvar, -2, dtype=np.int64)   # -2 unknown, -1 global
    for name, (off, shape) in prg.blocks.items():
        n = int(np.prod(shape))
        if name.startswith("_l1aux"): continue
        if len(shape) == 2:
            for j in range(shape[1]):
                lab[off + j*shape[0]: off + (j+1)*shape[0]] = j
        elif shape[0] in (N, N-1) and name not in ("p",):
            lab[off:off+n] = np.arange(n)
        else:
            lab[off:off+n] = -1
    return lab

N = int(sys.argv[1]) if len(sys.argv) > 1 else 31
pb = problems.StarshipProblem(N); g = pb.guess(N)
pars = ptr.Parameters(N=N, Nsub=30, iter_max=15, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=1e-4, feas_tol=5e-3)
P = ptr.PTR(pb, pars); ref = P.make_solution(*g)
prg, h = P.build(ref); cp = prg.compile()
A, G = cp["A"].tocsr(), cp["G"].tocsr()
n, p = A.shape[1], A.shape[0]
lab = labels_from_program(prg, N)
# vic/vtc/Pf: attach to stage via neighbours; aux: propagate labels through inequality rows
Gb = (abs(G) > 0).astype(np.int8)
for _ in range(3):
    unk = np.where(lab == -2)[0]
    if unk.size == 0: break
    for v in unk:
        rows = Gb[:, v].nonzero()[0]
        nb = np.unique(Gb[rows].nonzero()[1])
        l = lab[nb]; l = l[l >= 0]
        if l.size: lab[v] = l.min()
        elif np.any(lab[nb] == -1): lab[v] = -1
# globals that touch a single stage -> that stage (vic, vtc, Pf)
H = (Gb.T @ Gb).tocsr()
Ab = (abs(A) > 0).astype(np.int8).tocsr()
for v in np.where(lab == -1)[0]:
    rows = Ab[:, v].nonzero()[0]
    nb = np.unique(np.concatenate([Ab[rows].nonzero()[1], H[v].nonzero()[1]]))
    l = lab[nb]; l = np.unique(l[l >= 0])
    if l.size == 1 and (lab[nb] != -1).sum() >= 1 and nb.size < 40: lab[v] = l[0]
print("globals:", (lab == -1).sum(), "unknown", (lab == -2).sum())
# eq row intervals
rlo = np.zeros(p, dtype=np.int64); rhi = np.zeros(p, dtype=np.int64)
for r in range(p):
    l = lab[Ab[r].nonzero()[1]]; l = l[l >= 0]
    rlo[r], rhi[r] = (l.min(), l.max()) if l.size else (-1, -1)
order = []
local_rows = [[] for _ in range(N)]; strad = {}
for r in range(p):
    if rlo[r] == rhi[r] and rlo[r] >= 0: local_rows[rlo[r]].append(n + r)
    elif rlo[r] >= 0: strad.setdefault((rlo[r], rhi[r]), []).append(n + r)
for k in range(N):
    order += np.where(lab == k)[0].tolist() + local_rows[k]
def nd(lo, hi):   # stage range; returns straddling rows in ND order
    if hi <= lo: return []
    mid = (lo + hi) // 2
    return nd(lo, mid) + nd(mid + 1, hi) + [r for (a, b), rr in strad.items() if a <= mid < b and a >= lo and b <= hi for r in rr]
chain = nd(0, N - 1)
rest_rows = [n + r for r in range(p) if (n + r) not in set(order) and (n + r) not in set(chain)]
order += chain + np.where(lab == -1)[0].tolist() + rest_rows
order = np.array(order); assert len(set(order.tolist())) == n + p, (len(order), n + p)
Hn = (H + sp.eye(n)).tocsr()
K = sp.bmat([[Hn, Ab.T], [Ab, -sp.eye(p)]], format="csc")
Pm = K[order][:, order].tocsc()
parent, cnt, height = symbolic(Pm)
print("reduced KKT dim", n + p, "nnz", K.nnz, "nnzL", cnt.sum(), "ops", int((cnt*(cnt+1)//2).sum()), "height", height.max(), "maxcol", cnt.max())
lv = np.bincount(height)
print("levels", len(lv), "median width", np.median(lv), "min", lv.min())
