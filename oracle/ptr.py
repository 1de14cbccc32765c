"""Oracle PTR: CPU restatement of the reference's penalized-trust-region loop.  TEST INFRASTRUCTURE ONLY.

Follows, call by call:
  SCPProblem / compute_scaling      src/solvers/scp.jl:140-157, 376-517 (user-advised ranges only)
  Subproblem ctor                   src/solvers/ptr.jl:213-293
  add_dynamics! / state_update!     src/solvers/scp.jl:657-674, discretization.jl:424-497
  add_convex_state/input_constr.    src/solvers/scp.jl:685-734
  add_nonconvex_constraints!        src/solvers/scp.jl:744-794
  add_bcs!                          src/solvers/scp.jl:808-895
  add_trust_region!                 src/solvers/ptr.jl:565-743   (q in {1, 2, Inf})
  add_cost! (+ vc / tr penalties)   src/solvers/ptr.jl:753-895, scp.jl:552-601
  solve loop, stopping rule         src/solvers/ptr.jl:448-532, 908-932; scp.jl:909-931
The conic solve (JuMP -> ECOS in the reference) is oracle/conic.py.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

from . import conic, orc


@dataclass
class Parameters:           # ptr.jl:57-71
    N: int
    Nsub: int
    iter_max: int
    wvc: float
    wtr: float
    eps_abs: float
    eps_rel: float
    feas_tol: float
    q_tr: float = np.inf
    q_exit: float = np.inf
    solver_tol: float = 1e-9


def compute_bbox(pb, N):
    """compute_scaling's bounding boxes (scp.jl:376-481): advised ranges where given (entries of pb.ranges() that are
    not None), otherwise min / max of the variable over the convex set X (states, parameters) or U (inputs, parameters)
    imposed at every node on ONE (x, u, p); unbounded (DUAL_INFEASIBLE) or failed (NUMERICAL_ERROR) programs keep the
    default box [0, 1].  HiGHS for polyhedral sets, the oracle IPM for sets with second-order cones."""
    from . import orc
    xrg, urg, prg_ = pb.ranges()
    t = orc.t_grid(N)
    box = {"x": [list(r) if r is not None else [0.0, 1.0] for r in xrg],
           "u": [list(r) if r is not None else [0.0, 1.0] for r in urg],
           "p": [list(r) if r is not None else [0.0, 1.0] for r in prg_]}
    adv = {"x": xrg, "u": urg, "p": prg_}
    passes = [("x", "emit_X"), ("u", "emit_U"), ("p", "emit_X"), ("p", "emit_U")]
    for blk, emit in passes:
        miss = [i for i, r in enumerate(adv[blk]) if r is None]
        if not miss or not hasattr(pb, emit):
            continue
        for i in miss:
            for j in (0, 1):
                prg = conic.ConeProgram()
                x = prg.new_variable(pb.nx, "x"); u = prg.new_variable(pb.nu, "u"); p = prg.new_variable(pb.np, "p")
                for k in range(N):
                    if emit == "emit_X":
                        pb.emit_X(prg, t[k], k + 1, x, p)
                    else:
                        pb.emit_U(prg, t[k], k + 1, u, p)
                z = {"x": x, "u": u, "p": p}[blk][i]
                sgn = 1.0 if j == 0 else -1.0
                prg.add_cost(z * sgn)
                cp = prg.compile()
                if cp["G"].shape[0] == 0 and cp["A"].shape[0] == 0:
                    continue
                if cp["q"]:
                    with np.errstate(all="ignore"):
                        res = conic.solve_ipm(cp, tol=1e-9, maxit=60)
                    if res["status"] in ("OPTIMAL", "ALMOST_OPTIMAL") and np.isfinite(res["obj"]) and abs(res["obj"]) < 1e7:
                        box[blk][i][j] = sgn * res["obj"]
                else:
                    res = conic.solve_highs(cp)
                    if res["status"] == "OPTIMAL":
                        box[blk][i][j] = sgn * res["obj"]
                    elif res["status"] not in ("DUAL_INFEASIBLE", "NUMERICAL_ERROR", "INFEASIBLE"):
                        raise RuntimeError(f"Solver failed during variable scaling ({res['status']})")
    return box["x"], box["u"], box["p"]


def correct_convex(pb, sc, N, xd, ud, p, tol=1e-9):
    """correct_convex! (scp.jl:275-361): the closest trajectory, in the scaled L1 sense, that satisfies the convex path
    constraints X and U at every node.  Returns the projected (xd, ud, p)."""
    from . import orc
    t = orc.t_grid(N)
    prg = conic.ConeProgram()
    x = prg.new_variable((pb.nx, N), "x", sc.Sx, sc.cx)
    u = prg.new_variable((pb.nu, N), "u", sc.Su, sc.cu)
    pp = prg.new_variable(pb.np, "p", sc.Sp, sc.cp)
    for k in range(N):
        if hasattr(pb, "emit_X"):
            pb.emit_X(prg, t[k], k + 1, x[:, k], pp)
        if hasattr(pb, "emit_U"):
            pb.emit_U(prg, t[k], k + 1, u[:, k], pp)
    ex = prg.new_variable(N, "tau_x"); eu = prg.new_variable(N, "tau_u"); ep = prg.new_variable(1, "tau_p")
    for k in range(N):
        prg.l1([ex[k]] + [(x[i, k] - xd[k, i]) * sc.iSx[i] for i in range(pb.nx)], "x_variation")
        prg.l1([eu[k]] + [(u[i, k] - ud[k, i]) * sc.iSu[i] for i in range(pb.nu)], "u_variation")
    prg.l1([ep[0]] + [(pp[i] - p[i]) * sc.iSp[i] for i in range(pb.np)], "p_variation")
    J = conic.Aff()
    for k in range(N):
        J = J + ex[k] + eu[k]
    prg.add_cost(J + ep[0])
    res = conic.solve(prg.compile(), tol=tol, prefer="ipm")
    if res["status"] not in ("OPTIMAL", "ALMOST_OPTIMAL"):
        raise RuntimeError(f"Solver failed to find the closest initial guess that satisfies the convex constraints ({res['status']})")
    z = res["z"]
    val = np.vectorize(lambda e: e.value(z), otypes=[float])
    return val(x).T.copy(), val(u).T.copy(), val(pp)


class Scaling:              # scp.jl:483-516
    def __init__(self, pb, N=None):
        xrg, urg, prg = pb.ranges()
        if any(r is None for r in list(xrg) + list(urg) + list(prg)):
            xrg, urg, prg = compute_bbox(pb, N if N is not None else pb.N)
        zero_tol = np.sqrt(np.finfo(float).eps)

        def mk(rg):
            lo = np.array([r[0] for r in rg], dtype=float)
            hi = np.array([r[1] for r in rg], dtype=float)
            S = (hi - lo) / 1.0
            S[S < zero_tol] = 1.0
            return S, lo - S * 0.0

        self.Sx, self.cx = mk(xrg)
        self.Su, self.cu = mk(urg)
        self.Sp, self.cp = mk(prg)
        self.iSx, self.iSu, self.iSp = 1.0 / self.Sx, 1.0 / self.Su, 1.0 / self.Sp


@dataclass
class Solution:             # ptr.jl:74-102
    xd: np.ndarray
    ud: np.ndarray
    p: np.ndarray
    dyn: object = None
    feas: bool = False
    defect: np.ndarray = None
    J: float = np.nan
    J_tr: float = np.nan
    J_vc: float = np.nan
    J_aug: float = np.nan
    vd: np.ndarray = None
    vs: np.ndarray = None
    vic: np.ndarray = None
    vtc: np.ndarray = None
    eta_x: np.ndarray = None
    eta_u: np.ndarray = None
    eta_p: float = np.nan
    status: str = "OPTIMIZE_NOT_CALLED"
    deviation: float = np.nan
    improv_rel: float = np.nan
    timing: dict = field(default_factory=dict)


def trapz(f, grid):         # helper.jl:560-568
    F = 0.0
    for k in range(len(grid) - 1):
        d = grid[k + 1] - grid[k]
        F = F + (f[k + 1] + f[k]) * (0.5 * d)
    return F


class PTR:
    def __init__(self, pb, pars: Parameters):
        self.pb, self.pars = pb, pars
        self.scale = Scaling(pb, pars.N)
        self.t = orc.t_grid(pars.N)
        self.model = pb.orc_model()

    # SubproblemSolution(x,u,p,iter,pbm) -> discretize!   (ptr.jl:313-383)
    def make_solution(self, xd, ud, p) -> Solution:
        t0 = time.perf_counter()
        d = orc.discretize(self.model, xd, ud, p, self.pars.Nsub, self.scale.iSx, self.pars.feas_tol, self.t)
        sol = Solution(xd=np.array(xd, dtype=float), ud=np.array(ud, dtype=float), p=np.array(p, dtype=float),
                       dyn=d, feas=d.feas, defect=d.defect)
        sol.timing["discretize"] = time.perf_counter() - t0
        return sol

    # ------------------------------------------------------------------ subproblem
    def build(self, ref: Solution):
        pb, pars, sc, t = self.pb, self.pars, self.scale, self.t
        N, nx, nu, np_ = pars.N, pb.nx, pb.nu, pb.np
        prg = conic.ConeProgram()
        x = prg.new_variable((nx, N), "x", sc.Sx, sc.cx)
        u = prg.new_variable((nu, N), "u", sc.Su, sc.cu)
        p = prg.new_variable(np_, "p", sc.Sp, sc.cp)
        vd = prg.new_variable((nx, N - 1), "vd")
        eta_x = prg.new_variable(N, "eta_x")
        eta_u = prg.new_variable(N, "eta_u")
        eta_p = prg.new_variable(1, "eta_p")
        dyn = ref.dyn

        # add_dynamics!
        for k in range(N - 1):
            rhs = (conic.matvec(dyn.A[k], x[:, k]) + conic.matvec(dyn.Bm[k], u[:, k]) +
                   conic.matvec(dyn.Bp[k], u[:, k + 1]) + conic.matvec(dyn.F[k], p) +
                   conic.matvec(dyn.E[k], vd[:, k]))
            prg.zero([x[i, k + 1] - (rhs[i] + dyn.r[k][i]) for i in range(nx)], "dynamics")
        # convex state / input constraints
        if hasattr(pb, "emit_X"):
            for k in range(N):
                pb.emit_X(prg, t[k], k + 1, x[:, k], p)
        if hasattr(pb, "emit_U"):
            for k in range(N):
                pb.emit_U(prg, t[k], k + 1, u[:, k], p)
        # nonconvex constraints
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
        # boundary conditions (relaxed)
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
        # trust region
        q = pars.q_tr
        cone = {1: prg.l1, 2: prg.soc, 4: prg.soc, np.inf: prg.linf}[q]     # q2cone, ptr.jl:582
        xh_ref = (ref.xd - sc.cx) * sc.iSx
        uh_ref = (ref.ud - sc.cu) * sc.iSu
        ph_ref = (ref.p - sc.cp) * sc.iSp

        def bound(d_lq, eta, name):
            if q == 4:      # ptr.jl:604-630: w with |d_lq| <= w (SOC) and geomean(eta, 1) >= w (GEOM), i.e. d_lq^2 <= eta
                w = prg.new_variable(1, f"w_{name}_{prg.nvar}")
                prg.soc([w[0], d_lq], name)
                prg.geom([w[0], eta, 1.0], name)
            else:
                prg.nonpos([d_lq - eta])

        dp_lq = prg.new_variable(1, "dp_lq")
        cone([dp_lq[0]] + [(p[i] - sc.cp[i]) * sc.iSp[i] - ph_ref[i] for i in range(np_)], "parameter_trust_region")
        bound(dp_lq[0], eta_p[0], "parameter_trust_region")
        dx_lq = prg.new_variable(N, "dx_lq")
        for k in range(N):
            cone([dx_lq[k]] + [(x[i, k] - sc.cx[i]) * sc.iSx[i] - xh_ref[k, i] for i in range(nx)], "state_trust_region")
            bound(dx_lq[k], eta_x[k], "state_trust_region")
        du_lq = prg.new_variable(N, "du_lq")
        for k in range(N):
            cone([du_lq[k]] + [(u[i, k] - sc.cu[i]) * sc.iSu[i] - uh_ref[k, i] for i in range(nu)], "input_trust_region")
            bound(du_lq[k], eta_u[k], "input_trust_region")
        # cost
        if hasattr(pb, "cost_emit"):        # convex non-affine cost: the problem adds its epigraph cones to the program
            J = pb.cost_emit(prg, x, u, p, t)
        else:
            J = pb.cost_aff(x, u, p, t) if hasattr(pb, "cost_aff") else conic.Aff()
        prg.add_cost(J)
        J_tr = (trapz(eta_x, t) + trapz(eta_u, t) + eta_p[0]) * pars.wtr
        prg.add_cost(J_tr)
        P = prg.new_variable(N, "P")
        Pf = prg.new_variable(2, "Pf")
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
        J_vc = (trapz(P, t) + Pf[0] + Pf[1]) * pars.wvc
        prg.add_cost(J_vc)
        h = dict(x=x, u=u, p=p, vd=vd, vs=vs, vic=vic, vtc=vtc, eta_x=eta_x, eta_u=eta_u, eta_p=eta_p,
                 J=J, J_tr=J_tr, J_vc=J_vc)
        return prg, h

    def solve_subproblem(self, ref: Solution, prefer="auto"):
        t0 = time.perf_counter()
        prg, h = self.build(ref)
        cp = prg.compile()
        t_form = time.perf_counter() - t0
        t0 = time.perf_counter()
        res = conic.solve(cp, tol=self.pars.solver_tol, prefer=prefer)
        t_solve = time.perf_counter() - t0
        z = res["z"]
        val = np.vectorize(lambda e: e.value(z), otypes=[float])
        if res["status"] not in ("OPTIMAL", "ALMOST_OPTIMAL"):
            sol = Solution(xd=ref.xd, ud=ref.ud, p=ref.p, status=res["status"])
            sol.timing.update(formulate=t_form, solve=t_solve)
            return sol, cp, res
        xd = val(h["x"]).T.copy()
        ud = val(h["u"]).T.copy()
        p = val(h["p"])
        sol = self.make_solution(xd, ud, p)
        sol.status = res["status"]
        sol.vd = val(h["vd"]).T
        sol.vs = val(h["vs"]).T if h["vs"] is not None else None
        sol.vic = val(h["vic"]) if h["vic"] is not None else None
        sol.vtc = val(h["vtc"]) if h["vtc"] is not None else None
        sol.J = conic.Aff.lift(h["J"]).value(z)
        sol.J_tr = h["J_tr"].value(z)
        sol.J_vc = h["J_vc"].value(z)
        sol.J_aug = res["obj"]
        sol.eta_x, sol.eta_u, sol.eta_p = val(h["eta_x"]), val(h["eta_u"]), float(val(h["eta_p"])[0])
        sol.timing.update(formulate=t_form, solve=t_solve, solver_iters=res["iters"],
                          n=cp["c"].size, m_eq=cp["A"].shape[0], m_ineq=cp["G"].shape[0])
        return sol, cp, res

    # solution_deviation, scp.jl:909-931
    def deviation(self, ref: Solution, sol: Solution):
        sc, q = self.scale, self.pars.q_exit
        xh, xr = (sol.xd - sc.cx) * sc.iSx, (ref.xd - sc.cx) * sc.iSx
        ph, pr = (sol.p - sc.cp) * sc.iSp, (ref.p - sc.cp) * sc.iSp
        dp = np.linalg.norm(ph - pr, q)
        dx = max(np.linalg.norm(xh[k] - xr[k], q) for k in range(self.pars.N))
        return dp + dx

    # check_stopping_criterion!, ptr.jl:908-932
    def check_stop(self, it, ref, sol):
        sol.deviation = self.deviation(ref, sol)
        with np.errstate(all="ignore"):
            sol.improv_rel = (ref.J_aug - sol.J_aug) / abs(ref.J_aug)
        return bool(it > 1 and (sol.feas and (abs(sol.improv_rel) <= self.pars.eps_rel or
                                              sol.deviation <= self.pars.eps_abs)))

    # solve, ptr.jl:448-532
    def solve(self, guess, verbose=False, prefer="auto"):
        xd, ud, p = guess
        ref = self.make_solution(xd, ud, p)
        history = []
        k = 1
        status = "SCP_FAILED"
        while True:
            sol, cp, res = self.solve_subproblem(ref, prefer=prefer)
            history.append(sol)
            if sol.status not in ("OPTIMAL", "ALMOST_OPTIMAL"):
                status = f"SCP_FAILED ({sol.status})"
                break
            stop = self.check_stop(k, ref, sol)
            if verbose:
                mvd = np.abs(sol.vd).max()
                mvs = max(sol.vs.max(), 0.0) if sol.vs is not None else 0.0
                print(f"{k:3d} {sol.status:8s} vd {mvd:.1e} vs {mvs:.1e} J {sol.J_aug:+.6e} dJ% "
                      f"{100 * sol.improv_rel:+.3f} dev {sol.deviation:.2e} feas {sol.feas} "
                      f"[form {sol.timing['formulate']:.2f}s solve {sol.timing['solve']:.2f}s "
                      f"it {sol.timing['solver_iters']} disc {sol.timing['discretize']:.3f}s]")
            status = "SCP_SOLVED"
            if stop:
                break
            ref = sol
            k += 1
            if k > self.pars.iter_max:
                break
        return dict(status=status, iterations=len(history), sol=history[-1], history=history)
