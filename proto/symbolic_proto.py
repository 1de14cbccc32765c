"""Prototype: fill / etree height of the IPM KKT matrix under different orderings."""
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla, pickle, sys, time
from scipy.sparse.csgraph import reverse_cuthill_mckee

def kkt_pattern(cp):
    A, G = cp["A"], cp["G"]
    n = A.shape[1]; p = A.shape[0]; m = G.shape[0]
    K = sp.bmat([[sp.eye(n), A.T, G.T], [A, -sp.eye(p), None], [G, None, -sp.eye(m)]], format="csc")
    return K

def etree(Ap, Ai, n):
    parent = -np.ones(n, dtype=np.int64); anc = -np.ones(n, dtype=np.int64)
    for j in range(n):
        for p in range(Ap[j], Ap[j+1]):
            i = Ai[p]
            while i != -1 and i < j:
                nxt = anc[i]; anc[i] = j
                if nxt == -1: parent[i] = j
                i = nxt
    return parent

def symbolic(M):
    """M symmetric csc (full). returns parent, colcounts, height."""
    n = M.shape[0]
    U = sp.triu(M, format="csc")
    parent = etree(U.indptr, U.indices, n)
    # column structures via children merge
    L = sp.tril(M, format="csc")
    struct = [None]*n
    children = [[] for _ in range(n)]
    for j in range(n):
        if parent[j] >= 0: children[parent[j]].append(j)
    cnt = np.zeros(n, dtype=np.int64)
    height = np.zeros(n, dtype=np.int64)
    for j in range(n):
        s = set(L.indices[L.indptr[j]:L.indptr[j+1]].tolist())
        s.discard(j)
        for c in children[j]:
            s |= struct[c]; struct[c] = None
            height[j] = max(height[j], height[c]+1)
        s.discard(j)
        struct[j] = s
        cnt[j] = len(s)
    return parent, cnt, height

if __name__ == "__main__":
    d = pickle.load(open(sys.argv[1], "rb")); cp = d["cp"]
    K = kkt_pattern(cp); n = K.shape[0]
    print("KKT dim", n, "nnz", K.nnz)
    for name in ("natural", "rcm", "mmd"):
        t0 = time.time()
        if name == "natural": perm = np.arange(n)
        elif name == "rcm": perm = reverse_cuthill_mckee(K.tocsr(), symmetric_mode=True)
        else:
            lu = spla.splu(K.tocsc(), permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
            perm = lu.perm_c
        P = K[perm][:, perm].tocsc()
        parent, cnt, height = symbolic(P)
        print(name, "nnzL", cnt.sum(), "flops/2", int((cnt*(cnt+1)//2).sum()), "height", height.max(), "maxcol", cnt.max(), "t", round(time.time()-t0,1))
