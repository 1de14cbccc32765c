"""Oracle SCvx: CPU restatement of the reference's successive-convexification loop.  TEST INFRASTRUCTURE ONLY.

Follows, call by call (everything shared with PTR is inherited from oracle/ptr.py):
  Parameters                        src/solvers/scvx.jl:57-81
  Subproblem ctor                   src/solvers/scvx.jl:225-303   (no trust-region variables: eta is data)
  add_trust_region!                 src/solvers/scvx.jl:578-678   (q in {1, 2, Inf}; dx_lq[k] + du_lq[k] + dp_lq <= eta)
  add_cost! / linear penalty        src/solvers/scvx.jl:689-701, 804-901 (lambda (trapz(P) + Pf1 + Pf2))
  check_stopping_criterion!         src/solvers/scvx.jl:711-733
  update_trust_region! / rule       src/solvers/scvx.jl:745-770, 1000-1045
  actual_cost_penalty!              src/solvers/scvx.jl:919-951
  solution_cost!                    src/solvers/scvx.jl:968-988  (kind :linear returns the ORIGINAL cost sol.L --
                                    the predicted improvement is J_ref - L(sol), as the reference computes it)
  solve loop                        src/solvers/scvx.jl:460-546
Not restated: correct_convex! of the initial guess (scvx.jl:563, scp.jl:275-361); callers pass guesses that are used
as they are (the same on the CUDA side).
"""
from __future__ import annotations

import time
from dataclasses import dataclass

import numpy as np

from . import conic
from .ptr import PTR, Solution, trapz


@dataclass
class Parameters:           # scvx.jl:57-81
    N: int
    Nsub: int
    iter_max: int
    lam: float
    rho_0: float
    rho_1: float
    rho_2: float
    beta_sh: float
    beta_gr: float
    eta_init: float
    eta_lb: float
    eta_ub: float
    eps_abs: float
    eps_rel: float
    feas_tol: float
    q_tr: float = np.inf
    q_exit: float = np.inf
    solver_tol: float = 1e-9


class SCvx(PTR):
    # ------------------------------------------------------------------ nonlinear cost (scvx.jl:919-988)
    def original_cost(self, sol: Solution) -> float:
        pb = self.pb
        if not hasattr(pb, "cost_aff"):
            return 0.0
        v = pb.cost_aff(sol.xd.T, sol.ud.T, sol.p, self.t)
        return float(conic.Aff.lift(v).value(np.zeros(0))) if isinstance(v, conic.Aff) else float(v)

    def actual_cost_penalty(self, sol: Solution) -> float:
        pb, pars, t = self.pb, self.pars, self.t
        N = pars.N
        P = np.zeros(N)
        for k in range(N):
            dk = sol.defect[k] if k < N - 1 else np.zeros(pb.nx)
            sk = pb.s(t[k], k + 1, sol.xd[k], sol.ud[k], sol.p) if getattr(pb, "ns", 0) else np.zeros(1)
            P[k] = pars.lam * (np.abs(dk).sum() + np.maximum(sk, 0.0).sum())
        gic = pb.gic(sol.xd[0], sol.p) if getattr(pb, "gic", None) is not None else np.zeros(1)
        gtc = pb.gtc(sol.xd[-1], sol.p) if getattr(pb, "gtc", None) is not None else np.zeros(1)
        return trapz(P, t) + pars.lam * (np.abs(gic).sum() + np.abs(gtc).sum())

    def make_solution(self, xd, ud, p) -> Solution:      # SubproblemSolution(x,u,p,iter,pbm): discretize! + J_aug
        sol = super().make_solution(xd, ud, p)
        sol.L = self.original_cost(sol)
        sol.J_aug = sol.L + self.actual_cost_penalty(sol)
        return sol

    # ------------------------------------------------------------------ subproblem
    def build(self, ref: Solution, eta: float = 1.0):
        pb, pars, sc, t = self.pb, self.pars, self.scale, self.t
        N, nx, nu, np_ = pars.N, pb.nx, pb.nu, pb.np
        prg = conic.ConeProgram()
        x = prg.new_variable((nx, N), "x", sc.Sx, sc.cx)
        u = prg.new_variable((nu, N), "u", sc.Su, sc.cu)
        p = prg.new_variable(np_, "p", sc.Sp, sc.cp)
        vd = prg.new_variable((nx, N - 1), "vd")
        P = prg.new_variable(N, "P")
        Pf = prg.new_variable(2, "Pf")
        dyn = ref.dyn
        for k in range(N - 1):
            rhs = (conic.matvec(dyn.A[k], x[:, k]) + conic.matvec(dyn.Bm[k], u[:, k]) +
                   conic.matvec(dyn.Bp[k], u[:, k + 1]) + conic.matvec(dyn.F[k], p) +
                   conic.matvec(dyn.E[k], vd[:, k]))
            prg.zero([x[i, k + 1] - (rhs[i] + dyn.r[k][i]) for i in range(nx)], "dynamics")
        if hasattr(pb, "emit_X"):
            for k in range(N):
                pb.emit_X(prg, t[k], k + 1, x[:, k], p)
        if hasattr(pb, "emit_U"):
            for k in range(N):
                pb.emit_U(prg, t[k], k + 1, u[:, k], p)
        vs = None
        if getattr(pb, "ns", 0):
            ns = pb.ns
            vs = prg.new_variable((ns, N), "vs")
            for k in range(N):
                a = (t[k], k + 1, ref.xd[k], ref.ud[k], ref.p)
                s = pb.s(*a)
                C, D, G = pb.C(*a), pb.D(*a), pb.G(*a)
                r = s - C @ ref.xd[k] - D @ ref.ud[k] - G @ ref.p
                lhs = conic.matvec(C, x[:, k]) + conic.matvec(D, u[:, k]) + conic.matvec(G, p)
                prg.nonpos([lhs[i] + r[i] - vs[i, k] for i in range(ns)], "path_ncvx")
        vic = vtc = None
        if getattr(pb, "gic", None) is not None:
            g = pb.gic(ref.xd[0], ref.p)
            H0 = pb.H0(ref.xd[0], ref.p)
            K0 = pb.K0(ref.xd[0], ref.p) if getattr(pb, "K0", None) else np.zeros((g.size, np_))
            l0 = g - H0 @ ref.xd[0] - K0 @ ref.p
            vic = prg.new_variable(g.size, "vic")
            lhs = conic.matvec(H0, x[:, 0]) + conic.matvec(K0, p)
            prg.zero([lhs[i] + l0[i] + vic[i] for i in range(g.size)], "initial_condition")
        if getattr(pb, "gtc", None) is not None:
            g = pb.gtc(ref.xd[-1], ref.p)
            Hf = pb.Hf(ref.xd[-1], ref.p)
            Kf = pb.Kf(ref.xd[-1], ref.p) if getattr(pb, "Kf", None) else np.zeros((g.size, np_))
            lf = g - Hf @ ref.xd[-1] - Kf @ ref.p
            vtc = prg.new_variable(g.size, "vtc")
            lhs = conic.matvec(Hf, x[:, N - 1]) + conic.matvec(Kf, p)
            prg.zero([lhs[i] + lf[i] + vtc[i] for i in range(g.size)], "terminal_condition")
        # trust region (scvx.jl:578-678)
        q = pars.q_tr
        cone = {1: prg.l1, 2: prg.soc, 4: prg.soc, np.inf: prg.linf}[q]
        xh_ref = (ref.xd - sc.cx) * sc.iSx
        uh_ref = (ref.ud - sc.cu) * sc.iSu
        ph_ref = (ref.p - sc.cp) * sc.iSp
        dp_lq = prg.new_variable(1, "dp_lq")
        cone([dp_lq[0]] + [(p[i] - sc.cp[i]) * sc.iSp[i] - ph_ref[i] for i in range(np_)], "parameter_trust_region")
        dx_lq = prg.new_variable(N, "dx_lq")
        for k in range(N):
            cone([dx_lq[k]] + [(x[i, k] - sc.cx[i]) * sc.iSx[i] - xh_ref[k, i] for i in range(nx)], "state_trust_region")
        du_lq = prg.new_variable(N, "du_lq")
        for k in range(N):
            cone([du_lq[k]] + [(u[i, k] - sc.cu[i]) * sc.iSu[i] - uh_ref[k, i] for i in range(nu)], "input_trust_region")
        for k in range(N):
            if q == 4:      # scvx.jl:646-662: |(dx_lq, du_lq, dp_lq)|_2 <= w, geomean(eta, 1) >= w
                w = prg.new_variable(1, f"w_tr_{k}")
                prg.soc([w[0], dx_lq[k], du_lq[k], dp_lq[0]], "trust_region_bound")
                prg.geom([w[0], eta, 1.0], "trust_region_bound")
            else:
                prg.nonpos([dx_lq[k] + du_lq[k] + dp_lq[0] - eta], "trust_region_bound")
        # cost (scvx.jl:689-701, 804-901)
        L = pb.cost_aff(x, u, p, t) if hasattr(pb, "cost_aff") else conic.Aff()
        prg.add_cost(L)
        for k in range(N):
            if k < N - 1:
                Ev = conic.matvec(dyn.E[k], vd[:, k])
                prg.l1([P[k]] + list(Ev) + (list(vs[:, k]) if vs is not None else []), "vd_vs_penalty")
            elif vs is not None:
                prg.l1([P[k]] + list(vs[:, k]), "vd_vs_penalty")
            else:
                prg.zero([P[k]])
        if vic is not None:
            prg.l1([Pf[0]] + list(vic), "vic_penalty")
        else:
            prg.zero([Pf[0]])
        if vtc is not None:
            prg.l1([Pf[1]] + list(vtc), "vtc_penalty")
        else:
            prg.zero([Pf[1]])
        L_pen = (trapz(P, t) + Pf[0] + Pf[1]) * pars.lam
        prg.add_cost(L_pen)
        return prg, dict(x=x, u=u, p=p, vd=vd, vs=vs, vic=vic, vtc=vtc, L=L, L_pen=L_pen)

    def solve_subproblem(self, ref: Solution, eta: float, prefer="auto"):
        t0 = time.perf_counter()
        prg, h = self.build(ref, eta)
        cp = prg.compile()
        t_form = time.perf_counter() - t0
        t0 = time.perf_counter()
        res = conic.solve(cp, tol=self.pars.solver_tol, prefer=prefer)
        t_solve = time.perf_counter() - t0
        z = res["z"]
        val = lambda a: np.vectorize(lambda e: conic.Aff.lift(e).value(z), otypes=[float])(a)
        sol = self.make_solution(val(h["x"]).T.copy(), val(h["u"]).T.copy(), val(h["p"]))
        sol.status = res["status"]
        sol.vd = val(h["vd"]).T
        sol.L = conic.Aff.lift(h["L"]).value(z)          # sol.L = value(spbm.L), scvx.jl:437
        sol.L_pen = h["L_pen"].value(z)
        sol.L_aug = res["obj"]
        sol.timing.update(formulate=t_form, solve=t_solve, solver_iters=res["iters"])
        return sol, cp, res

    # ------------------------------------------------------------------ loop (scvx.jl:460-546)
    def solve(self, guess, verbose=False, prefer="auto"):
        pars = self.pars
        eta = pars.eta_init
        ref = self.make_solution(*guess)
        history = []
        k = 1
        status = "SCP_FAILED"
        while True:
            sol, cp, res = self.solve_subproblem(ref, eta, prefer=prefer)
            sol.eta = eta
            history.append(sol)
            if sol.status not in ("OPTIMAL", "ALMOST_OPTIMAL"):
                status = f"SCP_FAILED ({sol.status})"
                break
            # check_stopping_criterion!
            sol.deviation = self.deviation(ref, sol)
            sol.pre_improv = ref.J_aug - sol.L
            with np.errstate(all="ignore"):
                pre_rel = sol.pre_improv / abs(ref.J_aug)
            stop = bool(k > 1 and (sol.feas and (pre_rel <= pars.eps_rel or sol.deviation <= pars.eps_abs)))
            status = "SCP_SOLVED"
            if verbose:
                print(f"{k:3d} {sol.status:8s} eta {eta:.2e} J_ref {ref.J_aug:+.6e} J_sol {sol.J_aug:+.6e} L {sol.L:+.6e} "
                      f"dev {sol.deviation:.2e} feas {sol.feas}")
            if stop:
                break
            # update_trust_region! + update_rule
            sol.act_improv = ref.J_aug - sol.J_aug
            with np.errstate(all="ignore"):
                sol.rho = sol.act_improv / sol.pre_improv
            rho = sol.rho
            if rho < pars.rho_0:
                eta = max(pars.eta_lb, eta / pars.beta_sh); sol.reject = True
            elif rho < pars.rho_1:
                eta = max(pars.eta_lb, eta / pars.beta_sh); ref = sol; sol.reject = False
            elif rho < pars.rho_2:
                ref = sol; sol.reject = False
            else:
                eta = min(pars.eta_ub, pars.beta_gr * eta); ref = sol; sol.reject = False
            k += 1
            if k > pars.iter_max:
                break
        return dict(status=status, iterations=len(history), sol=history[-1], history=history, eta=eta)
