"""Shared test helpers (oracle-side problem construction for parity tests)."""
import numpy as np
import scipy.sparse as sp

from oracle import conic, problems, ptr


def labels_from_program(prg, N):
    """Stage label per program variable from the oracle's block layout: column index for (.., N) blocks,
    position for length-N vectors, -1 for globals, -2 for L1 auxiliaries (derived from neighbours)."""
    lab = np.full(prg.nvar, -2, dtype=np.int64)
    for name, (off, shape) in prg.blocks.items():
        n = int(np.prod(shape))
        if name.startswith("_l1aux"):
            continue
        if len(shape) == 2:
            for j in range(shape[1]):
                lab[off + j * shape[0]: off + (j + 1) * shape[0]] = j
        elif shape[0] in (N, N - 1) and name != "p":
            lab[off:off + n] = np.arange(n)
        else:
            lab[off:off + n] = -1
    return lab


def union_pattern(mats):
    """Common CSR pattern of a list of same-shape sparse matrices + per-matrix value arrays on it."""
    pat = None
    for M in mats:
        Mb = sp.csr_matrix((np.ones(M.nnz), M.indices, M.indptr), shape=M.shape)
        pat = Mb if pat is None else pat + Mb
    pat = pat.tocsr(); pat.sort_indices()
    pat.data[:] = 1.0
    vals = []
    for M in mats:
        # position of every union entry inside M (or zero)
        Z = (pat * 0.0 + M).tocsr() if False else None
        Md = M.tocsr(); Md.sort_indices()
        v = np.zeros(pat.nnz)
        for r in range(pat.shape[0]):
            a0, a1 = pat.indptr[r], pat.indptr[r + 1]
            b0, b1 = Md.indptr[r], Md.indptr[r + 1]
            idx = np.searchsorted(pat.indices[a0:a1], Md.indices[b0:b1])
            v[a0 + idx] = Md.data[b0:b1]
        vals.append(v)
    return pat, np.array(vals)


def starship_subproblems(N, nb, seed=0, Nsub=30, perturb=0.02):
    """nb PTR subproblems (iteration 1) around perturbed initial guesses; returns oracle objects + cps."""
    pb = problems.StarshipProblem(N)
    g = pb.guess(N)
    pars = ptr.Parameters(N=N, Nsub=Nsub, iter_max=15, wvc=1e3, wtr=0.1, eps_abs=1e-5, eps_rel=1e-4, feas_tol=5e-3)
    P = ptr.PTR(pb, pars)
    rng = np.random.default_rng(seed)
    sc = P.scale
    out = []
    for b in range(nb):
        xd = g[0] + (perturb * sc.Sx * rng.standard_normal(g[0].shape) if b else 0.0)
        ud = g[1] + (perturb * sc.Su * rng.standard_normal(g[1].shape) if b else 0.0)
        p = g[2] * (1 + (0.05 * rng.uniform(-1, 1, g[2].shape) if b else 0.0))
        ref = P.make_solution(xd, ud, p)
        prg, h = P.build(ref)
        out.append(dict(ref=ref, prg=prg, h=h, cp=prg.compile()))
    return pb, P, out
