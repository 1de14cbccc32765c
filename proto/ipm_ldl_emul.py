"""CPU emulation of the GPU IPM numerics: oracle IPM loop with the KKT solves done by the product's
static-order LDL' programs (host interpreter), to tune delta / refinement / equilibration."""
import sys, pickle, numpy as np, scipy.sparse as sp
sys.path.insert(0, '/root/repo')
import __graft_entry__ as g
pkg = g.load_package()
from oracle import conic
from tests import helpers

def solve(cp, perm, delta, nref, equil, tol=1e-8, maxit=60, verbose=False, ddyn=0.0):
    if equil:
        cp, D, Ea, Eg = conic.equilibrate(cp)
    c, A, b, G, h = cp["c"], cp["A"].tocsr(), cp["b"], cp["G"].tocsr(), cp["h"]
    A.sort_indices(); G.sort_indices()
    n, p, m, l = c.size, A.shape[0], G.shape[0], cp["l"]
    def kkt(w2, bx, by, bz):
        wm = 1.0 / w2                       # reduced form uses W^-2
        t = wm * bz
        r1 = bx + G.T @ t
        rhs = np.concatenate([r1, by])
        sol = np.zeros(n + p)
        r = rhs.copy()
        for it in range(nref + 1):
            d, _ = pkg.lib.debug_kkt_solve(A, G, l, [], perm, A.data, G.data, wm, delta, r, ddyn)
            sol += d
            dx, dy = sol[:n], sol[n:]
            r = np.concatenate([r1 - (G.T @ (wm * (G @ dx)) + A.T @ dy), by - A @ dx])
        dx, dy = sol[:n], sol[n:]
        return dx, dy, wm * (G @ dx - bz), None, np.abs(r).max()
    one = np.ones(m)
    x, _, dz, _, _ = kkt(one, np.zeros(n), b, h)
    s = -dz
    ts = np.max(-s)
    if ts >= -1e-8 * max(1, np.linalg.norm(s)): s = s + (1 + ts)
    _, y, z, _, _ = kkt(one, -c, np.zeros(p), np.zeros(m))
    tz = np.max(-z)
    if tz >= -1e-8 * max(1, np.linalg.norm(z)): z = z + (1 + tz)
    nb, nh, nc = max(1, np.linalg.norm(b)), max(1, np.linalg.norm(h)), max(1, np.linalg.norm(c))
    last = np.nan
    for it in range(maxit):
        rx = c + A.T @ y + G.T @ z; ry = A @ x - b; rz = G @ x + s - h
        gap = s @ z; pc = c @ x; dc = -b @ y - h @ z
        pres = max(np.linalg.norm(ry) / nb, np.linalg.norm(rz) / nh); dres = np.linalg.norm(rx) / nc
        relgap = gap / max(abs(pc), abs(dc), 1)
        if verbose: print(it, f"pc {pc:+.6e} dc {dc:+.6e} gap {gap:.1e} pres {pres:.1e} dres {dres:.1e}")
        if pres <= tol and dres <= tol and (gap <= tol or relgap <= tol): return "OPTIMAL", it, pc
        if not np.isfinite(pres + dres + gap): return "NUMERICAL", it, last
        last = pc
        wm = s / z; mu = gap / m
        dxa, dya, dza, gm, e = kkt(wm, -rx, -ry, -rz + s)
        dsa = -rz - G @ dxa
        def amax(u, du):
            neg = du < 0
            return np.min(-u[neg] / du[neg]) if neg.any() else np.inf
        al = min(1, amax(s, dsa), amax(z, dza)); sig = (1 - al) ** 3
        tmp = -s + (sig * mu - dsa * dza) / z
        dx, dy, dz, gm, e = kkt(wm, -(1 - sig) * rx, -(1 - sig) * ry, -(1 - sig) * rz - tmp)
        ds = -(1 - sig) * rz - G @ dx
        al = min(1, 0.99 * min(amax(s, ds), amax(z, dz)))
        x, y, z, s = x + al * dx, y + al * dy, z + al * dz, s + al * ds
    return "MAXIT", it, pc

if __name__ == "__main__":
    N = int(sys.argv[1])
    pb, P, subs = helpers.starship_subproblems(N, 1, seed=N)
    sub = subs[0]
    lab = helpers.labels_from_program(sub["prg"], N)
    perm = pkg.ordering.stage_order(sub["cp"]["A"], sub["cp"]["G"], lab, N)

    ref = conic.solve_highs(sub["cp"])
    print("ref obj", ref["obj"] - sub["cp"]["c0"])
    for equil in (True, False):
        for delta in (1e-10, 1e-9):
            for nref in (2,):
                for ddyn in (1e-10, 1e-12, 1e-14):
                    st, it, pc = solve(sub["cp"], perm, delta, nref, equil, ddyn=ddyn)
                    print(f"equil {equil} delta {delta:.0e} ddyn {ddyn:.0e} nref {nref}: {st} it {it} obj {pc:.10f} err {abs(pc-(ref['obj']-sub['cp']['c0'])):.1e}", flush=True)
