"""Oracle GuSTO: CPU restatement of the reference's GuSTO loop (pen = :quad).  TEST INFRASTRUCTURE ONLY.

Follows src/solvers/gusto.jl call by call (shared pieces come from oracle/ptr.py):
  Parameters                        gusto.jl:58-85
  Subproblem ctor, kappa            gusto.jl:218-287   (x, u, p scaled; kappa = mu^(1 + iter - iter_mu) from iter_mu on)
  solve loop                        gusto.jl:425-502   (add_cost!, dynamics / boundary conditions UN-relaxed, hard U)
  generate_initial_guess            gusto.jl:517-526   (correct_convex! -> oracle.ptr.correct_convex)
  original_cost                     gusto.jl:570-707   (terminal cost + trapz of u'Su, S convex; nonconvex mode identical)
  state_penalty_cost                gusto.jl:725-867   (convex X sets through their indicators, nonconvex s linearised)
  soft_penalty, quadratic           gusto.jl:936-995   (u >= 0, f + u - v <= 0, cost lambda v^2)
  trust_region_cost                 gusto.jl:1056-1190 (q in {1, 2, Inf}: dx_lq[k] + dp_lq - (eta + tr[k]) <= 0, penalty on tr)
  check_stopping_criterion!         gusto.jl:1203-1231
  update_trust_region! / rule       gusto.jl:1245-1427
The quadratic terms of the objective reach the (linear-objective) cone solver as rotated second-order cones, one
epigraph variable per time node and cost group: q >= sum_i (sqrt(lambda) v_i)^2.  JuMP does the same for ECOS through its
quadratic-objective bridge; the minimiser is identical.
"""
from __future__ import annotations

import time
from dataclasses import dataclass

import numpy as np

from . import conic, orc
from .ptr import PTR, Solution, trapz


@dataclass
class Parameters:           # gusto.jl:58-85
    N: int
    Nsub: int
    iter_max: int
    lam_init: float
    lam_max: float
    rho_0: float
    rho_1: float
    beta_sh: float
    beta_gr: float
    gamma_fail: float
    eta_init: float
    eta_lb: float
    eta_ub: float
    mu: float
    iter_mu: int
    eps_abs: float
    eps_rel: float
    feas_tol: float
    q_tr: float = np.inf
    q_exit: float = np.inf
    solver_tol: float = 1e-9


def _sumsq(prg, exprs, name):
    """q >= sum e_i^2 as |(2 e, q - 1)| <= q + 1; returns q (Aff)"""
    q = prg.new_variable(1, f"_q{len(prg.blocks)}")[0]
    prg.soc([q + 1.0] + [conic.Aff.lift(e) * 2.0 for e in exprs] + [q - 1.0], name)
    return q


class GuSTO(PTR):
    """pb supplies: S(t, k, p) (convex input quadratic), optional phi_quad(p) terms, s / C / G of the nonconvex path
    constraints as functions of (t, k, x, p) [the GuSTO flavour has no input dependence, quadrotor/definition.jl:264-268],
    emit_U, gic / gtc with Jacobians."""

    def __init__(self, pb, pars):
        super().__init__(pb, pars)

    # ------------------------------------------------------------------ costs at a numeric trajectory
    def original_cost(self, xd, ud, p) -> float:          # gusto.jl:680-707 (nonconvex mode; S convex => same formula)
        pb, t = self.pb, self.t
        run = [float(ud[k] @ pb.S(t[k], k + 1, p) @ ud[k]) for k in range(self.pars.N)]
        term = float(pb.phi(xd[-1], p)) if hasattr(pb, "phi") else 0.0
        return term + trapz(run, t)

    def state_penalty_nonconvex(self, xd, p, lam) -> float:   # gusto.jl:846-864
        pb, t = self.pb, self.t
        pen = []
        for k in range(self.pars.N):
            s = pb.s(t[k], k + 1, xd[k], None, p)
            pen.append(float(np.sum(lam * np.maximum(0.0, s) ** 2)))
        return trapz(pen, t)

    def tr_lhs(self, ref, xd, p, eta):                    # trust_region_cost, :nonconvex (gusto.jl:1167-1187)
        sc, q = self.scale, self.pars.q_tr
        dx = (xd - sc.cx) * sc.iSx - (ref.xd - sc.cx) * sc.iSx
        dp = (p - sc.cp) * sc.iSp - (ref.p - sc.cp) * sc.iSp
        dpn = np.linalg.norm(dp, q)
        return np.array([np.linalg.norm(dx[k], q) + dpn - eta for k in range(self.pars.N)])

    def make_solution(self, xd, ud, p) -> Solution:
        sol = super().make_solution(xd, ud, p)
        sol.J_aug = float("nan")
        return sol

    # ------------------------------------------------------------------ subproblem (gusto.jl:218-287, 534-1190)
    def build(self, ref: Solution, lam: float = 1e4, eta: float = 1.0):
        pb, pars, sc, t = self.pb, self.pars, self.scale, self.t
        N, nx, nu, np_ = pars.N, pb.nx, pb.nu, pb.np
        prg = conic.ConeProgram()
        x = prg.new_variable((nx, N), "x", sc.Sx, sc.cx)
        u = prg.new_variable((nu, N), "u", sc.Su, sc.cu)
        p = prg.new_variable(np_, "p", sc.Sp, sc.cp)
        sl = float(np.sqrt(lam))
        # ---- original cost (gusto.jl:604-677): trapz_k u_k' S u_k with S = diag-like PSD -> Cholesky rows
        L_run = []
        for k in range(N):
            S = pb.S(t[k], k + 1, ref.p)
            w, V = np.linalg.eigh(S)
            rows = [sum((u[j, k] * (np.sqrt(w[i]) * V[j, i]) for j in range(nu)), conic.Aff()) for i in range(nu) if w[i] > 1e-14]
            L_run.append(_sumsq(prg, rows, "run_cost") if rows else conic.Aff())
        L = trapz(L_run, t)
        if hasattr(pb, "phi_emit"):
            L = L + pb.phi_emit(prg, x[:, N - 1], p)
        # ---- state penalty (gusto.jl:742-842): nonconvex s linearised at the reference, quadratic soft penalty
        L_st_nodes = []
        for k in range(N):
            a = (t[k], k + 1, ref.xd[k], None, ref.p)
            s = pb.s(*a)
            C, G = pb.C(*a), pb.G(*a)
            vs = []
            for i in range(s.size):
                uu = prg.new_variable(1, f"_su{k}_{i}")[0]
                vv = prg.new_variable(1, f"_sv{k}_{i}")[0]
                prg.nonpos([-uu])
                f_lin = conic.Aff(None, float(s[i] - C[i] @ ref.xd[k] - G[i] @ ref.p))
                for j in range(nx):
                    if C[i, j] != 0.0:
                        f_lin = f_lin + x[j, k] * float(C[i, j])
                for j in range(np_):
                    if G[i, j] != 0.0:
                        f_lin = f_lin + p[j] * float(G[i, j])
                prg.nonpos([f_lin + uu - vv])
                vs.append(vv * sl)
            L_st_nodes.append(_sumsq(prg, vs, "state_penalty") if vs else conic.Aff())
        L_st = trapz(L_st_nodes, t)
        # ---- trust region (gusto.jl:1075-1164)
        q = pars.q_tr
        cone = {1: prg.l1, 2: prg.soc, np.inf: prg.linf}[q]
        xh_ref = (ref.xd - sc.cx) * sc.iSx
        ph_ref = (ref.p - sc.cp) * sc.iSp
        tr = prg.new_variable(N, "tr")
        dx_lq = prg.new_variable(N, "dx_lq")
        dp_lq = prg.new_variable(1, "dp_lq")
        cone([dp_lq[0]] + [(p[i] - sc.cp[i]) * sc.iSp[i] - ph_ref[i] for i in range(np_)], "parameter_trust_region")
        L_tr_nodes = []
        for k in range(N):
            cone([dx_lq[k]] + [(x[i, k] - sc.cx[i]) * sc.iSx[i] - xh_ref[k, i] for i in range(nx)], "state_trust_region")
            prg.nonpos([dx_lq[k] + dp_lq[0] - (tr[k] + eta)], "trust_region_bound")
            uu = prg.new_variable(1, f"_tu{k}")[0]
            vv = prg.new_variable(1, f"_tv{k}")[0]
            prg.nonpos([-uu])
            prg.nonpos([tr[k] + uu - vv])
            L_tr_nodes.append(_sumsq(prg, [vv * sl], "trust_penalty"))
        L_tr = trapz(L_tr_nodes, t)
        prg.add_cost(L); prg.add_cost(L_st); prg.add_cost(L_tr)
        # ---- dynamics, un-relaxed (gusto.jl:452, scp.jl:657-674)
        dyn = ref.dyn
        for k in range(N - 1):
            rhs = (conic.matvec(dyn.A[k], x[:, k]) + conic.matvec(dyn.Bm[k], u[:, k]) +
                   conic.matvec(dyn.Bp[k], u[:, k + 1]) + conic.matvec(dyn.F[k], p))
            prg.zero([x[i, k + 1] - (rhs[i] + dyn.r[k][i]) for i in range(nx)], "dynamics")
        # ---- convex input constraints (hard), boundary conditions (un-relaxed)
        if hasattr(pb, "emit_U"):
            for k in range(N):
                pb.emit_U(prg, t[k], k + 1, u[:, k], p)
        for g_, H_, K_, node in ((getattr(pb, "gic", None), getattr(pb, "H0", None), getattr(pb, "K0", None), 0),
                                 (getattr(pb, "gtc", None), getattr(pb, "Hf", None), getattr(pb, "Kf", None), N - 1)):
            if g_ is None:
                continue
            xr = ref.xd[node]
            g = g_(xr, ref.p)
            H = H_(xr, ref.p)
            K = K_(xr, ref.p) if K_ else np.zeros((g.size, np_))
            l0 = g - H @ xr - K @ ref.p
            lhs = conic.matvec(H, x[:, node]) + conic.matvec(K, p)
            prg.zero([lhs[i] + l0[i] for i in range(g.size)], "boundary_condition")
        h = dict(x=x, u=u, p=p, L=L, L_st=L_st, L_tr=L_tr, tr=tr)
        return prg, h

    def solve_subproblem(self, ref: Solution, lam, eta, prefer="ipm"):
        t0 = time.perf_counter()
        prg, h = self.build(ref, lam, eta)
        cp = prg.compile()
        t_form = time.perf_counter() - t0
        t0 = time.perf_counter()
        res = conic.solve_ipm(cp, tol=self.pars.solver_tol)
        t_solve = time.perf_counter() - t0
        z = res["z"]
        if res["status"] not in ("OPTIMAL", "ALMOST_OPTIMAL"):
            sol = Solution(xd=ref.xd, ud=ref.ud, p=ref.p, status=res["status"])
            sol.timing.update(formulate=t_form, solve=t_solve)
            return sol, cp, res
        val = np.vectorize(lambda e: e.value(z), otypes=[float])
        xd, ud, p = val(h["x"]).T.copy(), val(h["u"]).T.copy(), val(h["p"])
        sol = self.make_solution(xd, ud, p)          # discretize! of the new iterate
        sol.status = res["status"]
        # SubproblemSolution(spbm), gusto.jl:399-418
        sol.J = self.original_cost(xd, ud, p)
        sol.J_st = self.state_penalty_nonconvex(xd, p, lam)
        sol.J_tr = conic.Aff.lift(h["L_tr"]).value(z)
        sol.J_aug = sol.J + sol.J_st + sol.J_tr
        sol.L = conic.Aff.lift(h["L"]).value(z)
        sol.L_st = conic.Aff.lift(h["L_st"]).value(z)
        sol.L_aug = res["obj"]
        sol.timing.update(formulate=t_form, solve=t_solve, solver_iters=res["iters"])
        return sol, cp, res

    # ------------------------------------------------------------------ update_trust_region! (gusto.jl:1245-1293)
    def rho(self, ref: Solution, sol: Solution):
        pb, t, m = self.pb, self.t, self.model
        N = self.pars.N
        cost_error = abs(sol.J_aug - sol.L_aug)
        cost_nrml = abs(sol.L_aug)
        df, dxdt = np.zeros(N), np.zeros(N)
        for k in range(N):
            f, A, B, F = orc.dyn_eval(m, t[k], k + 1, ref.xd[k], ref.ud[k], ref.p)
            r = f - A @ ref.xd[k] - B @ ref.ud[k] - F @ ref.p
            f_lin = A @ sol.xd[k] + B @ sol.ud[k] + F @ sol.p + r
            f_nl = orc.dyn_eval(m, t[k], k + 1, sol.xd[k], sol.ud[k], sol.p)[0]
            df[k] = np.linalg.norm(f_nl - f_lin)
            dxdt[k] = np.linalg.norm(f_lin)
        dyn_error, dyn_nrml = trapz(df, t), trapz(dxdt, t)
        return (cost_error + dyn_error) / (cost_nrml + dyn_nrml), cost_error, dyn_error

    def update_rule(self, it, ref, sol, lam, eta):         # gusto.jl:1310-1427
        pars, pb, t = self.pars, self.pb, self.t
        kappa = 1.0 if it < pars.iter_mu else pars.mu ** (1 + it - pars.iter_mu)
        tr = self.tr_lhs(ref, sol.xd, sol.p, eta)
        trust_viol = bool(np.any(tr > 1e-3))
        feasible = True
        if not trust_viol:
            for k in range(pars.N):
                if np.any(pb.s(t[k], k + 1, sol.xd[k], None, sol.p) > 1e-3):
                    feasible = False
                    break
        if trust_viol:
            nxt = (ref, eta, pars.gamma_fail * lam, True)
        elif sol.rho < pars.rho_1:
            n_eta = min(pars.eta_ub, pars.beta_gr * eta) if sol.rho < pars.rho_0 else eta
            n_lam = pars.lam_init if feasible else pars.gamma_fail * lam
            nxt = (sol, n_eta, n_lam, False)
        else:
            nxt = (ref, max(pars.eta_lb, eta / pars.beta_sh), lam, True)
        n_ref, n_eta, n_lam, rej = nxt
        if kappa < 1:
            n_eta *= kappa
        return n_ref, n_eta, n_lam, rej

    def check_stop(self, it, ref, sol, lam):               # gusto.jl:1203-1231
        pars = self.pars
        sol.deviation = self.deviation(ref, sol)
        with np.errstate(all="ignore"):
            dJ = abs(ref.J_aug - sol.J_aug) / abs(ref.J_aug)
        infeas = lam > pars.lam_max
        return (it > 1) and ((sol.feas and (dJ <= pars.eps_rel or sol.deviation <= pars.eps_abs)) or infeas)

    # ------------------------------------------------------------------ GuSTO.solve (gusto.jl:425-502)
    def solve(self, guess, verbose=False):
        pars = self.pars
        lam, eta = pars.lam_init, pars.eta_init
        ref = self.make_solution(*guess)
        history, status = [], "SCP_SOLVED"
        last = None
        k = 0
        for k in range(1, pars.iter_max + 1):
            sol, cp, res = self.solve_subproblem(ref, lam, eta)
            last = sol
            if sol.status not in ("OPTIMAL", "ALMOST_OPTIMAL"):
                status = f"SCP_FAILED ({sol.status})"
                break
            sol.lam, sol.eta = lam, eta
            stop = self.check_stop(k, ref, sol, lam)
            history.append(sol)
            if stop:
                break
            sol.rho, sol.cost_error, sol.dyn_error = self.rho(ref, sol)
            ref, eta, lam, sol.reject = self.update_rule(k, ref, sol, lam, eta)
            if verbose:
                print(f"it {k:2d} J_aug {sol.J_aug:.6e} L_aug {sol.L_aug:.6e} rho {sol.rho:.3e} eta {eta:.3e} lam {lam:.1e} "
                      f"rej {sol.reject} feas {sol.feas} dev {sol.deviation:.2e}")
        return dict(status=status, iterations=k, sol=last, ref=ref, eta=eta, lam=lam, history=history)
