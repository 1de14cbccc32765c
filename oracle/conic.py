"""Oracle cone-program modelling layer + CPU solvers.  TEST INFRASTRUCTURE ONLY.

Plays the role of the reference's parser + JuMP + ECOS stack for the oracle:
  * `ConeProgram`  ~ ConicProgram (src/parser/program.jl:63-76) with @new_variable /
    @add_constraint / @add_cost semantics; cones as in src/parser/cone.jl:36-47
    (ZERO z==0, NONPOS z<=0, L1 |x|_1<=t, SOC |x|_2<=t, LINF |x|_inf<=t, GEOM geomean(x1,x2)>=t).
  * L1 / LINF are lowered exactly like MathOptInterface's NormOneBridge / NormInfinityBridge
    (the path JuMP takes for ECOS): aux y_i >= |x_i|, sum(y) <= t;  -t <= x_i <= t.
  * compiled form (ECOS / CVXOPT convention):   min c'z  s.t.  A z = b,  G z + s = h,  s in K,
    K = R+^l x SOC(q_1) x ... .
  * `solve_highs`  : scipy.optimize.linprog (HiGHS) for polyhedral programs.
  * `solve_ipm`    : primal-dual path-following IPM with Nesterov-Todd scaling and Mehrotra
    correction (the published CVXOPT `conelp` / ECOS algorithm family), fp64, sparse KKT via
    SuperLU.  ECOS itself is an unvendored third-party dependency of the reference (ECOS.jl ->
    libecos, version unpinned: Project.toml:6-12), so this restates its published algorithm.

Parity status: unpinned against ECOS (not installable here); the two solvers are pinned against
each other on LPs and against analytic optima in tests/test_oracle_conic.py.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


# ------------------------------------------------------------------ affine expressions
class Aff:
    """const + sum_j coef_j * z_j over program variables (sparse dict)."""
    __slots__ = ("t", "c")

    def __init__(self, terms=None, const=0.0):
        self.t = terms if terms is not None else {}
        self.c = float(const)

    @staticmethod
    def lift(v):
        return v if isinstance(v, Aff) else Aff(None, float(v))

    def __add__(self, o):
        o = Aff.lift(o)
        t = dict(self.t)
        for k, v in o.t.items():
            t[k] = t.get(k, 0.0) + v
        return Aff(t, self.c + o.c)

    __radd__ = __add__

    def __neg__(self):
        return Aff({k: -v for k, v in self.t.items()}, -self.c)

    def __sub__(self, o):
        return self + (-Aff.lift(o))

    def __rsub__(self, o):
        return Aff.lift(o) + (-self)

    def __mul__(self, a):
        a = float(a)
        if a == 0.0:
            return Aff(None, 0.0)
        return Aff({k: v * a for k, v in self.t.items()}, self.c * a)

    __rmul__ = __mul__

    def __truediv__(self, a):
        return self * (1.0 / float(a))

    def value(self, z):
        return self.c + sum(v * z[k] for k, v in self.t.items())


def avec(exprs):
    return np.array([Aff.lift(e) for e in exprs], dtype=object)


def matvec(M, v):
    """numeric matrix times vector of Aff -> vector of Aff (skips structural zeros)."""
    M = np.asarray(M, dtype=float)
    out = []
    for i in range(M.shape[0]):
        e = Aff()
        for j in range(M.shape[1]):
            if M[i, j] != 0.0:
                e = e + v[j] * M[i, j]
        out.append(e)
    return np.array(out, dtype=object)


# ------------------------------------------------------------------ program
class ConeProgram:
    def __init__(self):
        self.nvar = 0
        self.blocks = {}         # name -> (offset, shape)
        self.eq = []             # Aff == 0
        self.ineq = []           # Aff <= 0
        self.socs = []           # list of [t, x...] Aff with |x|_2 <= t
        self.cost = Aff()
        self.names = []

    # @new_variable: returns an array of Aff; optional affine scaling x = S*xh + c (block.jl:368-397)
    def new_variable(self, shape, name, S=None, c=None):
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        n = int(np.prod(shape))
        off = self.nvar
        self.nvar += n
        self.blocks[name] = (off, shape)
        # column-major like Julia arrays: index = i + nrow*j
        flat = [Aff({off + i: 1.0}) for i in range(n)]
        arr = np.array(flat, dtype=object).reshape(shape, order="F")
        if S is not None:
            S = np.asarray(S, dtype=float)
            c = np.zeros_like(S) if c is None else np.asarray(c, dtype=float)
            if arr.ndim == 2:
                for i in range(shape[0]):
                    for j in range(shape[1]):
                        arr[i, j] = arr[i, j] * S[i] + c[i]
            else:
                for i in range(shape[0]):
                    arr[i] = arr[i] * S[i] + c[i]
        return arr

    def raw_index(self, name):
        off, shape = self.blocks[name]
        return off, shape

    # cones -------------------------------------------------------
    def zero(self, exprs, name=""):
        for e in exprs:
            self.eq.append(Aff.lift(e))

    def nonpos(self, exprs, name=""):
        for e in exprs:
            self.ineq.append(Aff.lift(e))

    def l1(self, exprs, name=""):
        """[t, x...]: |x|_1 <= t  (NormOneBridge)."""
        t, xs = Aff.lift(exprs[0]), [Aff.lift(e) for e in exprs[1:]]
        y = self.new_variable(len(xs), f"_l1aux{len(self.blocks)}")
        tot = Aff()
        for xi, yi in zip(xs, y):
            self.ineq.append(xi - yi)
            self.ineq.append(-xi - yi)
            tot = tot + yi
        self.ineq.append(tot - t)

    def linf(self, exprs, name=""):
        """[t, x...]: |x|_inf <= t  (NormInfinityBridge)."""
        t, xs = Aff.lift(exprs[0]), [Aff.lift(e) for e in exprs[1:]]
        for xi in xs:
            self.ineq.append(xi - t)
            self.ineq.append(-xi - t)

    def soc(self, exprs, name=""):
        self.socs.append([Aff.lift(e) for e in exprs])

    def geom(self, exprs, name=""):
        """[t, x1, x2]: geomean(x1, x2) >= t (cone.jl:45, MOI.GeometricMeanCone(3)) -- every GEOM use of the reference has
        two entries (ptr.jl:615, scvx.jl:659, gusto.jl:1129, double_integrator/definition.jl:80).  MathOptInterface's
        bridges hand ECOS the rotated second-order cone x1 x2 >= t^2, x >= 0, i.e. |(2 t, x1 - x2)|_2 <= x1 + x2."""
        if len(exprs) != 3:
            raise NotImplementedError("GEOM cone over more than two entries")
        t, x1, x2 = (Aff.lift(e) for e in exprs)
        self.socs.append([x1 + x2, t * 2.0, x1 - x2])

    def add_cost(self, expr):
        self.cost = self.cost + Aff.lift(expr)
        return Aff.lift(expr)

    # compile -----------------------------------------------------
    def compile(self):
        n = self.nvar

        def rows(exprs):
            r, cidx, v, const = [], [], [], np.zeros(len(exprs))
            for i, e in enumerate(exprs):
                const[i] = e.c
                for k, a in e.t.items():
                    if a != 0.0:
                        r.append(i); cidx.append(k); v.append(a)
            return sp.csr_matrix((v, (r, cidx)), shape=(len(exprs), n)), const

        A, ca = rows(self.eq)
        Gl, cl = rows(self.ineq)          # Gl z + cl <= 0  ->  Gl z + s = -cl, s >= 0
        soc_rows, q = [], []
        for cone in self.socs:
            q.append(len(cone))
            soc_rows.extend(cone)
        Gq, cq = rows(soc_rows)           # s = (t, x) = Gq z + cq in SOC  ->  -Gq z + s = cq
        G = sp.vstack([Gl, -Gq], format="csr") if len(soc_rows) else Gl
        h = np.concatenate([-cl, cq]) if len(soc_rows) else -cl
        c = np.zeros(n)
        for k, a in self.cost.t.items():
            c[k] += a
        return dict(c=c, c0=self.cost.c, A=A.tocsr(), b=-ca, G=G.tocsr(), h=h, l=len(self.ineq), q=q)


# ------------------------------------------------------------------ HiGHS (LP only)
def solve_highs(cp, tol=1e-9):
    from scipy.optimize import linprog
    assert not cp["q"], "HiGHS path handles polyhedral programs only"
    res = linprog(cp["c"], A_ub=cp["G"], b_ub=cp["h"], A_eq=cp["A"], b_eq=cp["b"],
                  bounds=[(None, None)] * cp["c"].size, method="highs",
                  options=dict(primal_feasibility_tolerance=tol, dual_feasibility_tolerance=tol,
                               ipm_optimality_tolerance=tol, presolve=True))
    status = {0: "OPTIMAL", 1: "ITERATION_LIMIT", 2: "INFEASIBLE", 3: "DUAL_INFEASIBLE"}.get(res.status, "NUMERICAL_ERROR")
    z = res.x if res.x is not None else np.full(cp["c"].size, np.nan)
    out = dict(status=status, z=z, obj=(float(res.fun) + cp["c0"]) if res.fun is not None else np.nan,
               iters=int(getattr(res, "nit", 0)))
    if res.status == 0:
        out["y_eq"] = -np.asarray(res.eqlin.marginals)
        out["z_ineq"] = -np.asarray(res.ineqlin.marginals)
    return out


# ------------------------------------------------------------------ IPM (LP + SOC)
class _Cones:
    def __init__(self, l, q):
        self.l, self.q = l, list(q)
        self.m = l + sum(q)
        self.deg = l + len(q)
        self.off = [l + int(np.sum(self.q[:i])) for i in range(len(self.q))]

    def e(self):
        v = np.zeros(self.m)
        v[:self.l] = 1.0
        for o in self.off:
            v[o] = 1.0
        return v

    def max_step(self, u):
        """smallest t >= 0 such that u + t*e is in K (0 if u interior)."""
        t = 0.0
        if self.l:
            t = max(t, float(np.max(-u[:self.l])))
        for o, q in zip(self.off, self.q):
            t = max(t, float(np.linalg.norm(u[o + 1:o + q]) - u[o]))
        return t

    def max_alpha(self, lam, v):
        """largest alpha (possibly inf) with lam + alpha*v in K, lam in int K."""
        a_max = np.inf
        if self.l:
            neg = v[:self.l] < 0
            if np.any(neg):
                a_max = min(a_max, float(np.min(-lam[:self.l][neg] / v[:self.l][neg])))
        for o, q in zip(self.off, self.q):
            l0, l1, v0, v1 = lam[o], lam[o + 1:o + q], v[o], v[o + 1:o + q]
            a = v0 * v0 - v1 @ v1
            bq = 2.0 * (l0 * v0 - l1 @ v1)
            cq = l0 * l0 - l1 @ l1
            if abs(a) < 1e-300:
                if bq < 0:
                    a_max = min(a_max, -cq / bq)
                continue
            disc = bq * bq - 4 * a * cq
            if a < 0:
                a_max = min(a_max, (-bq - np.sqrt(max(disc, 0.0))) / (2 * a))
            elif disc >= 0 and bq < 0:
                a_max = min(a_max, (-bq - np.sqrt(disc)) / (2 * a))
        return a_max

    def prod(self, u, v):
        """Jordan product u o v."""
        w = np.empty(self.m)
        w[:self.l] = u[:self.l] * v[:self.l]
        for o, q in zip(self.off, self.q):
            w[o] = u[o:o + q] @ v[o:o + q]
            w[o + 1:o + q] = u[o] * v[o + 1:o + q] + v[o] * u[o + 1:o + q]
        return w

    def div(self, lam, v):
        """solve lam o w = v."""
        w = np.empty(self.m)
        w[:self.l] = v[:self.l] / lam[:self.l]
        for o, q in zip(self.off, self.q):
            l0, l1 = lam[o], lam[o + 1:o + q]
            v0, v1 = v[o], v[o + 1:o + q]
            det = l0 * l0 - l1 @ l1
            w0 = (l0 * v0 - l1 @ v1) / det
            w[o] = w0
            w[o + 1:o + q] = (v1 - w0 * l1) / l0
        return w

    def nt(self, s, z):
        """Nesterov-Todd scaling: returns (W, Winv) as sparse block-diagonal matrices and lam = W z."""
        blocks, iblocks = [], []
        if self.l:
            w = np.sqrt(s[:self.l] / z[:self.l])
            blocks.append(sp.diags(w)); iblocks.append(sp.diags(1.0 / w))
        for o, q in zip(self.off, self.q):
            sk, zk = s[o:o + q], z[o:o + q]
            sn = np.sqrt(sk[0] ** 2 - sk[1:] @ sk[1:])
            zn = np.sqrt(zk[0] ** 2 - zk[1:] @ zk[1:])
            sb, zb = sk / sn, zk / zn
            gam = np.sqrt((1.0 + sb @ zb) / 2.0)
            wb = np.empty(q)
            wb[0] = (sb[0] + zb[0]) / (2 * gam)
            wb[1:] = (sb[1:] - zb[1:]) / (2 * gam)
            eta = np.sqrt(sn / zn)
            # W = eta * [[w0, w1'], [w1, I + w1 w1'/(1+w0)]]
            Wk = np.empty((q, q))
            Wk[0, 0] = wb[0]; Wk[0, 1:] = wb[1:]; Wk[1:, 0] = wb[1:]
            Wk[1:, 1:] = np.eye(q - 1) + np.outer(wb[1:], wb[1:]) / (1.0 + wb[0])
            Wi = np.empty((q, q))
            Wi[0, 0] = wb[0]; Wi[0, 1:] = -wb[1:]; Wi[1:, 0] = -wb[1:]
            Wi[1:, 1:] = np.eye(q - 1) + np.outer(wb[1:], wb[1:]) / (1.0 + wb[0])
            blocks.append(sp.csr_matrix(eta * Wk)); iblocks.append(sp.csr_matrix(Wi / eta))
        W = sp.block_diag(blocks, format="csr") if blocks else sp.csr_matrix((0, 0))
        Wi = sp.block_diag(iblocks, format="csr") if iblocks else sp.csr_matrix((0, 0))
        return W, Wi, W @ z


def equilibrate(cp, iters=5):
    """Ruiz equilibration of [A; G] (ECOS applies the same kind of preprocessing, its `equil` step):
    A^ = Ea A D, G^ = Eg G D with one common factor per second-order cone; b^ = Ea b, h^ = Eg h, c^ = D c.
    Solution map: x = D x^, y = Ea y^, z = Eg z^, s = s^ / Eg."""
    A, G = cp["A"].tocsr(), cp["G"].tocsr()
    p, n = A.shape
    m, l, q = G.shape[0], cp["l"], list(cp["q"])
    D, Ea, Eg = np.ones(n), np.ones(p), np.ones(m)
    As, Gs = A.copy(), G.copy()
    for _ in range(iters):
        ra = np.sqrt(np.abs(As).max(axis=1).toarray().ravel()) if p else np.zeros(0)
        rg = np.sqrt(np.abs(Gs).max(axis=1).toarray().ravel()) if m else np.zeros(0)
        off = l
        for qk in q:  # one factor per cone
            rg[off:off + qk] = rg[off:off + qk].mean()
            off += qk
        ra[ra == 0] = 1.0; rg[rg == 0] = 1.0
        cn = np.zeros(n)
        if p: cn = np.maximum(cn, np.abs(As).max(axis=0).toarray().ravel())
        if m: cn = np.maximum(cn, np.abs(Gs).max(axis=0).toarray().ravel())
        cn = np.sqrt(cn); cn[cn == 0] = 1.0
        Ea /= ra; Eg /= rg; D /= cn
        As = sp.diags(1 / ra) @ As @ sp.diags(1 / cn) if p else As
        Gs = sp.diags(1 / rg) @ Gs @ sp.diags(1 / cn) if m else Gs
    out = dict(cp)
    out.update(A=sp.csr_matrix(As), G=sp.csr_matrix(Gs), b=cp["b"] * Ea, h=cp["h"] * Eg, c=cp["c"] * D)
    return out, D, Ea, Eg


def solve_ipm(cp, tol=1e-9, maxit=100, verbose=False, equil=True):
    """min c'x s.t. Ax=b, Gx+s=h, s in K.  Returns dict(status, z, obj, iters, y_eq, z_ineq)."""
    if equil:
        cps, D, Ea, Eg = equilibrate(cp)
        r = solve_ipm(cps, tol=tol, maxit=maxit, verbose=verbose, equil=False)
        r["z"] = r["z"] * D
        if "y_eq" in r and r["y_eq"] is not None:
            r["y_eq"] = r["y_eq"] * Ea
            r["z_ineq"] = r["z_ineq"] * Eg
            r["s"] = r["s"] / Eg
        return r
    c, A, b, G, h = cp["c"], cp["A"].tocsc(), cp["b"], cp["G"].tocsc(), cp["h"]
    n, p, K = c.size, A.shape[0], _Cones(cp["l"], cp["q"])
    m = K.m
    reg = 1e-10

    def kkt_factor(Wi):
        # [H A'; A -reg I],  H = G' Winv' Winv G + reg I  (normal-equations form of the CVXOPT KKT)
        WG = (Wi @ G).tocsc() if m else sp.csc_matrix((0, n))
        H = (WG.T @ WG + reg * sp.eye(n)).tocsc()
        Kmat = sp.bmat([[H, A.T], [A, -reg * sp.eye(p)]], format="csc") if p else H
        return (spla.splu(Kmat), Kmat), WG

    def kkt_solve(fac, WG, Wi, bx, by, bz):
        """solve [0 A' G'; A 0 0; G 0 -W'W][dx;dy;dz] = [bx;by;bz]."""
        # dz = (W'W)^-1 (G dx - bz)  ->  (G'(W'W)^-1 G) dx + A'dy = bx + G'(W'W)^-1 bz
        t = Wi @ bz if m else bz
        rhs = np.concatenate([bx + (WG.T @ t if m else 0.0), by])
        lu, Kmat = fac
        sol = lu.solve(rhs)
        for _ in range(3):  # iterative refinement against the unregularised system
            res = rhs - (Kmat @ sol - reg * np.concatenate([sol[:n], -sol[n:]]))
            sol = sol + lu.solve(res)
        dx, dy = sol[:n], sol[n:]
        dz = Wi.T @ (WG @ dx - t) if m else np.zeros(0)
        return dx, dy, dz

    # ---- starting point (CVXOPT conelp section 7.1) ----
    I_m = sp.eye(m, format="csr")
    fac, WG = kkt_factor(I_m)
    x, y, zz = kkt_solve(fac, WG, I_m, np.zeros(n), b, h)
    s = -zz
    ts = K.max_step(s)
    if ts >= -1e-8 * max(1.0, np.linalg.norm(s)):
        s = s + (1.0 + ts) * K.e()
    _, y, z = kkt_solve(fac, WG, I_m, -c, np.zeros(p), np.zeros(m))
    tz = K.max_step(z)
    if tz >= -1e-8 * max(1.0, np.linalg.norm(z)):
        z = z + (1.0 + tz) * K.e()

    nb, nh, nc = max(1.0, np.linalg.norm(b)), max(1.0, np.linalg.norm(h)), max(1.0, np.linalg.norm(c))
    status = "ITERATION_LIMIT"
    it = 0
    best, stall = None, 0
    for it in range(maxit):
        rx = c + A.T @ y + G.T @ z
        ry = A @ x - b
        rz = G @ x + s - h
        gap = float(s @ z)
        pcost = float(c @ x)
        dcost = float(-b @ y - h @ z)
        pres = max(np.linalg.norm(ry) / nb, np.linalg.norm(rz) / nh)
        dres = np.linalg.norm(rx) / nc
        relgap = gap / max(1e-300, max(abs(pcost), abs(dcost), 1.0))
        if verbose:
            print(f"{it:3d} pcost {pcost:+.8e} dcost {dcost:+.8e} gap {gap:.2e} pres {pres:.2e} dres {dres:.2e}")
        if pres <= tol and dres <= tol and (gap <= tol or relgap <= tol):
            status = "OPTIMAL"
            break
        if not np.isfinite(pres + dres + gap):
            status = "NUMERICAL_ERROR"
            break
        acc = max(pres, dres, min(gap, relgap))
        if best is None or acc < best[4]:
            best, stall = (x.copy(), y.copy(), z.copy(), s.copy(), acc), 0
        else:
            stall += 1
            if stall >= 3 and best[4] <= 1e-6:  # numerical floor reached: return the best iterate
                x, y, z, s, acc = best
                status = "OPTIMAL" if acc <= 10 * tol else ("ALMOST_OPTIMAL" if acc <= 1e-6 else "NUMERICAL_ERROR")
                break
        if m == 0:
            fac, WG = kkt_factor(I_m)
            dx, dy, _ = kkt_solve(fac, WG, I_m, -rx, -ry, np.zeros(0))
            x, y = x + dx, y + dy
            continue
        with np.errstate(all="ignore"):
            W, Wi, lam = K.nt(s, z)
        ok = bool(np.all(np.isfinite(lam))) and bool(np.all(np.isfinite(Wi.data)))
        if ok:
            try:
                fac, WG = kkt_factor(Wi)
            except RuntimeError:
                ok = False
        if not ok:
            # iterate touched the cone boundary: keep the last good iterate (ECOS reports
            # ALMOST_OPTIMAL in the same situation when the reduced tolerances hold)
            x, y, z, s, acc = best
            status = "OPTIMAL" if acc <= 10 * tol else ("ALMOST_OPTIMAL" if acc <= 1e-6 else "NUMERICAL_ERROR")
            break
        mu = gap / K.deg

        def direction(ds_rhs, scale):
            # lam o (W dz + W^-T ds) = ds_rhs  =>  ds = W'(lam\ds_rhs) - W'W dz,  bz = -scale*rz - W'(lam\ds_rhs)
            tmp = W.T @ K.div(lam, ds_rhs)
            dx, dy, dz = kkt_solve(fac, WG, Wi, -scale * rx, -scale * ry, -scale * rz - tmp)
            ds = -scale * rz - G @ dx   # == tmp - W'W dz, but keeps G dx + ds = -scale*rz to rounding
            return dx, dy, dz, ds

        def raw_step(ds, dz):
            return min(K.max_alpha(lam, Wi.T @ ds), K.max_alpha(lam, W @ dz))

        dxa, dya, dza, dsa = direction(-K.prod(lam, lam), 1.0)
        alpha = min(1.0, raw_step(dsa, dza))
        sigma = (1.0 - alpha) ** 3
        ds_rhs = -K.prod(lam, lam) - K.prod(Wi.T @ dsa, W @ dza) + sigma * mu * K.e()
        dx, dy, dz, ds = direction(ds_rhs, 1.0 - sigma)
        alpha = min(1.0, 0.99 * raw_step(ds, dz))
        x, y, z, s = x + alpha * dx, y + alpha * dy, z + alpha * dz, s + alpha * ds
    return dict(status=status, z=x, obj=float(c @ x) + cp["c0"], iters=it, y_eq=y, z_ineq=z, s=s)


def solve(prog_or_cp, tol=1e-9, prefer="auto"):
    """Solve with HiGHS when polyhedral (the oracle's fast path), else the IPM."""
    cp = prog_or_cp.compile() if isinstance(prog_or_cp, ConeProgram) else prog_or_cp
    if prefer == "ipm" or cp["q"]:
        return solve_ipm(cp, tol=tol)
    return solve_highs(cp, tol=tol)
