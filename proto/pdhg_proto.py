"""numpy prototype of the batched first-order conic solver (design study for the CUDA kernels)."""
import numpy as np, scipy.sparse as sp, pickle, sys, time

def ruiz_pc(K, iters=10, alpha=1.0):
    m, n = K.shape
    d1 = np.ones(m); d2 = np.ones(n)
    Ks = K.copy().tocsr()
    for _ in range(iters):
        rn = np.sqrt(np.abs(Ks).max(axis=1).toarray().ravel()); rn[rn == 0] = 1
        cn = np.sqrt(np.abs(Ks).max(axis=0).toarray().ravel()); cn[cn == 0] = 1
        d1 /= rn; d2 /= cn
        Ks = sp.diags(1/rn) @ Ks @ sp.diags(1/cn)
    # Pock-Chambolle alpha=1
    rn = np.sqrt(np.asarray(np.abs(Ks).sum(axis=1)).ravel()); rn[rn == 0] = 1
    cn = np.sqrt(np.asarray(np.abs(Ks).sum(axis=0)).ravel()); cn[cn == 0] = 1
    d1 /= rn; d2 /= cn
    Ks = sp.diags(1/rn) @ Ks @ sp.diags(1/cn)
    return Ks.tocsr(), d1, d2

def solve(cp, tol=1e-8, maxit=200000, method="avg", verbose=True, x0=None, y0=None, check=64):
    A, G = cp["A"], cp["G"]
    K = sp.vstack([A, G]).tocsr(); q = np.concatenate([cp["b"], cp["h"]]); c = cp["c"]
    meq = A.shape[0]; m, n = K.shape
    Ks, d1, d2 = ruiz_pc(K)
    cs = c * d2; qs = q * d1
    KsT = Ks.T.tocsr()
    # power iteration
    v = np.random.default_rng(0).standard_normal(n)
    for _ in range(60):
        v = KsT @ (Ks @ v); nv = np.linalg.norm(v); v /= nv
    Knorm = np.sqrt(nv)
    eta = 0.95 / Knorm
    w = np.linalg.norm(cs) / max(np.linalg.norm(qs), 1e-12) if np.linalg.norm(qs) > 0 else 1.0
    x = np.zeros(n) if x0 is None else x0 / d2
    y = np.zeros(m) if y0 is None else y0 / d1
    def proj_y(y):
        y = y.copy(); y[meq:] = np.maximum(y[meq:], 0); return y
    def kkt(x, y):
        # unscaled residuals
        xu = x * d2; yu = y * d1
        r = K @ xu - q
        pr = np.concatenate([r[:meq], np.maximum(r[meq:], 0)])
        dr = c + K.T @ yu
        pobj = c @ xu; dobj = -q @ yu
        ep = np.linalg.norm(pr) / (1 + np.linalg.norm(q))
        ed = np.linalg.norm(dr) / (1 + np.linalg.norm(c))
        eg = abs(pobj - dobj) / (1 + abs(pobj) + abs(dobj))
        return ep, ed, eg, pobj
    def kkt_s(x, y):
        r = Ks @ x - qs
        pr = np.concatenate([r[:meq], np.maximum(r[meq:], 0)])
        dr = cs + KsT @ y
        gap = cs @ x + qs @ y
        return np.sqrt(w * (pr @ pr) + (dr @ dr) / w + gap * gap)
    xs, ys = x.copy(), y.copy()  # restart anchors
    xa, ya = np.zeros(n), np.zeros(m); na = 0
    k_last = 0; err_restart = kkt_s(x, y); err_prev = err_restart
    nrestart = 0
    t0 = time.time()
    for it in range(1, maxit + 1):
        tau, sig = eta / w, eta * w
        if method == "avg":
            xn = x - tau * (cs + KsT @ y)
            yn = proj_y(y + sig * (Ks @ (2 * xn - x) - qs))
            x, y = xn, yn
            xa += x; ya += y; na += 1
        elif method == "halpern":
            # reflected halpern: z+ = (k+1)/(k+2) (2T(z)-z)... with anchor z0
            xn = x - tau * (cs + KsT @ y)
            yn = proj_y(y + sig * (Ks @ (2 * xn - x) - qs))
            kk = it - k_last
            rho = 2.0
            xr = x + rho * (xn - x); yr = y + rho * (yn - y)
            lam = (kk) / (kk + 1.0)
            x = lam * xr + (1 - lam) * xs; y = lam * yr + (1 - lam) * ys
        if it % check == 0:
            if method == "avg":
                xc, yc = xa / na, ya / na
                e_avg = kkt_s(xc, yc); e_cur = kkt_s(x, y)
                if e_cur < e_avg: xc, yc, e_c = x, y, e_cur
                else: e_c = e_avg
            else:
                xc, yc = xn, yn   # last T(z)
                e_c = kkt_s(xc, yc)
            ep, ed, eg, pobj = kkt(xc, yc)
            if max(ep, ed, eg) < tol:
                if verbose: print(f"converged it {it} restarts {nrestart} ep {ep:.1e} ed {ed:.1e} eg {eg:.1e} obj {pobj:.9e} t {time.time()-t0:.1f}s")
                return dict(x=xc * d2, y=yc * d1, it=it, obj=pobj)
            do = False
            if e_c <= 0.2 * err_restart: do = True
            elif e_c <= 0.8 * err_restart and e_c > err_prev: do = True
            elif (it - k_last) >= 0.36 * it: do = True
            err_prev = e_c
            if do:
                dx = np.linalg.norm(xc - xs); dy = np.linalg.norm(yc - ys)
                if dx > 1e-10 and dy > 1e-10:
                    w = np.exp(0.5 * np.log(dy / dx) + 0.5 * np.log(w))
                x, y = xc.copy(), yc.copy(); xs, ys = x.copy(), y.copy()
                xa[:] = 0; ya[:] = 0; na = 0; k_last = it; err_restart = kkt_s(x, y); err_prev = err_restart
                nrestart += 1
            if verbose and it % (check * 64) == 0:
                print(f"it {it} ep {ep:.1e} ed {ed:.1e} eg {eg:.1e} obj {pobj:.9e} w {w:.2e} restarts {nrestart}")
    return dict(x=xc * d2, y=yc * d1, it=it, obj=pobj)

if __name__ == "__main__":
    d = pickle.load(open(sys.argv[1], "rb"))
    cp = d["cp"]
    for meth in sys.argv[2:]:
        for tol in (1e-4, 1e-6, 1e-8):
            r = solve(cp, tol=tol, method=meth, verbose=False)
            print(meth, tol, "iters", r["it"], "obj", r["obj"], "ref", d["obj"], "xerr", np.abs(r["x"] - d["z"]).max())
