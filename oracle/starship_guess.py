"""Oracle restatement of starship_initial_guess (starship_flip/definition.jl:97-445).
TEST INFRASTRUCTURE ONLY.  Phase 1: bang-bang gimbal flip propagated by RK4 (no aero torques);
phase 2: double-integrator terminal descent found by a small SOCP (oracle IPM instead of ECOS)."""
from __future__ import annotations

import math

import numpy as np

from . import conic


def _dynamics(pb, t, x, u, p, no_aero_torques=False):
    """definition.jl:498-550."""
    v = x[2:4]; th = x[4]; om = x[5]; dd = x[7]
    T, de = u[0], u[1]
    tdil = p[0] / pb.tau_s if t <= pb.tau_s else p[1] / (1 - pb.tau_s)
    leng = -pb.lcg
    lcp = pb.lcp - pb.lcg
    ei = np.array([math.cos(th), math.sin(th)])
    ej = np.array([-math.sin(th), math.cos(th)])
    Tv = T * (-math.sin(de) * ei + math.cos(de) * ej)
    MT = leng * T * math.sin(de)
    D = -pb.CD * np.linalg.norm(v) * v
    MD = 0.0 if no_aero_torques else -lcp * (D @ ei)
    f = np.zeros(8)
    f[0:2] = v
    f[2:4] = (Tv + D) / pb.m + np.array([0.0, -pb.g0])
    f[4] = om
    f[5] = (MT + MD) / pb.J
    f[6] = pb.alpha_e * T
    f[7] = (de - dd) / pb.rate_delay
    return f * tdil


def _rk4_full(f, x0, tspan):
    X = np.zeros((len(tspan), x0.size))
    X[0] = x0
    for k in range(1, len(tspan)):
        t, tp = tspan[k - 1], tspan[k]
        h = tp - t
        x = X[k - 1]
        k1 = f(t, x); k2 = f(t + h / 2, x + h / 2 * k1); k3 = f(t + h / 2, x + h / 2 * k2); k4 = f(t + h, x + h * k3)
        X[k] = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    return X


def _linrange(a, b, n):
    j = np.arange(n) / (n - 1)
    return (1 - j) * a + j * b


def _sample_linear(tg, X, t):
    t = max(tg[0], min(tg[-1], t))
    k = int(np.sum(t > tg))
    k = max(k, 1)
    k = min(k, len(tg) - 1)
    c = (tg[k] - t) / (tg[k] - tg[k - 1])
    return c * X[k - 1] + (1 - c) * X[k]


def starship_initial_guess(pb, N):
    tau = _linrange(0.0, 1.0, N)
    id1 = np.where(tau <= pb.tau_s)[0]
    id2 = np.arange(id1[-1], N)
    xg = np.zeros((N, 8)); ug = np.zeros((N, 3))
    # ---- phase 1: flip ----
    flip_ac = pb.lcg / pb.J * pb.T_min3 * math.sin(pb.delta_max)
    flip_ts = math.sqrt((pb.theta0 - pb.theta_s) / flip_ac)

    def ctrl(t):
        if t <= flip_ts:
            d = pb.delta_max
        elif t <= 2 * flip_ts:
            d = -pb.delta_max
        else:
            d = 0.0
        return np.array([pb.T_min3, d, 0.0])

    pflip = np.zeros(10); pflip[0] = pb.tau_s; pflip[1] = 1 - pb.tau_s
    f = lambda t, x: _dynamics(pb, t, x, ctrl(t), pflip, no_aero_torques=True)
    x10 = np.zeros(8)
    x10[0:2] = pb.r0; x10[2:4] = pb.v0; x10[4] = pb.theta0; x10[7] = pb.delta_max
    tf = 2 * flip_ts + 10.0
    tt = _linrange(0.0, tf, 5000)
    x1 = _rk4_full(f, x10, tt)
    vs = pb.vs @ pb.ey
    k0 = int(np.argmax(x1[:, 3] >= vs))
    if not (x1[k0, 3] >= vs):
        raise RuntimeError("no terminal velocity crossing")
    tt = tt[:k0 + 1]; x1 = x1[:k0 + 1]
    t1 = tt[-1]
    tau2t = lambda ta: ta / pb.tau_s * t1
    for i in id1:
        xg[i] = _sample_linear(tt, x1, tau2t(tau[i]))
        ug[i] = ctrl(tau2t(tau[i]))
    # ---- phase 2: terminal descent ----
    xs = _sample_linear(tt, x1, tau2t(tau[id1[-1]]))
    pb.hs = float(xs[0:2] @ pb.ey)
    tau2 = tau[id2] - tau[id2[0]]
    N2 = len(tau2)
    tdil = lambda t2: t2 / (1 - pb.tau_s)
    nx, nu = 4, 2
    A_lti = np.zeros((4, 4)); A_lti[0, 2] = A_lti[1, 3] = 1.0
    B_lti = np.zeros((4, 2)); B_lti[2, 0] = B_lti[3, 1] = 1.0 / pb.m
    r_lti = np.array([0.0, 0.0, 0.0, -pb.g0])

    def discretize(t2):
        dt = tau2[1] - tau2[0]
        td = tdil(t2)
        iA, iBm, iBp, ir = 0, 16, 24, 32

        def derivs(t, V):
            Phi = V[iA:iA + 16].reshape(4, 4, order="F")
            sm, sp_ = (dt - t) / dt, t / dt
            iPhi = np.linalg.solve(Phi, np.eye(4))
            return np.concatenate([((td * A_lti) @ Phi).flatten(order="F"),
                                   (iPhi @ (td * B_lti) * sm).flatten(order="F"),
                                   (iPhi @ (td * B_lti) * sp_).flatten(order="F"), iPhi @ (td * r_lti)])

        V0 = np.zeros(36); V0[iA:iA + 16] = np.eye(4).flatten(order="F")
        V = _rk4_full(derivs, V0, _linrange(0.0, dt, 100))[-1]
        A = V[iA:iA + 16].reshape(4, 4, order="F")
        return A, A @ V[iBm:iBm + 8].reshape(4, 2, order="F"), A @ V[iBp:iBp + 8].reshape(4, 2, order="F"), A @ V[ir:ir + 4]

    zero_tol = math.sqrt(np.finfo(float).eps)
    Tmax_x = pb.T_max1 * math.sin(pb.thetamax2)
    Sx, cx, Su, cu = np.ones(4), np.zeros(4), np.ones(2), np.zeros(2)

    def upd(S, c, i, lo, hi):
        if lo > hi:
            lo, hi = hi, lo
        if hi - lo > zero_tol:
            S[i] = hi - lo; c[i] = lo

    upd(Sx, cx, 0, 0, xs[0]); upd(Sx, cx, 1, 0, xs[1]); upd(Sx, cx, 2, 0, xs[2]); upd(Sx, cx, 3, 0, xs[3])
    upd(Su, cu, 0, -Tmax_x, Tmax_x); upd(Su, cu, 1, pb.T_min1, pb.T_max1)

    def solve_traj(t2):
        cvx = conic.ConeProgram()
        x = cvx.new_variable((nx, N2), "x", Sx, cx)
        u = cvx.new_variable((nu, N2), "u", Su, cu)
        x0 = np.array([xs[0], xs[1], xs[2], xs[3]])
        xf = np.array([0.0, 0.0, pb.vf[0], pb.vf[1]])
        cvx.zero([x[i, 0] - x0[i] for i in range(4)])
        cvx.zero([x[i, N2 - 1] - xf[i] for i in range(4)])
        A, Bm, Bp, r = discretize(t2)
        for k in range(N2 - 1):
            rhs = conic.matvec(A, x[:, k]) + conic.matvec(Bm, u[:, k]) + conic.matvec(Bp, u[:, k + 1])
            cvx.zero([x[i, k + 1] - (rhs[i] + r[i]) for i in range(4)])
        for k in range(N2):
            uk = u[:, k]
            cvx.soc([pb.T_max1, uk[0], uk[1]])
            cvx.nonpos([pb.T_min1 - uk[1]])
            cvx.soc([uk[1] / math.cos(pb.thetamax2), uk[0], uk[1]])
        for k in range(N2):
            cvx.nonpos([-x[1, k]])
        res = conic.solve_ipm(cvx.compile(), tol=1e-8)
        val = np.vectorize(lambda e: e.value(res["z"]), otypes=[float])
        return val(x), val(u), res["status"]

    t2 = 10.0
    while True:
        x2, T2, status = solve_traj(t2)
        if status in ("OPTIMAL", "ALMOST_OPTIMAL"):
            break
        t2 += 1.0
        if t2 > 40.0:
            raise RuntimeError("could not find a terminal descent time of flight")
    xg[id2, 0:2] = x2[0:2].T
    xg[id2, 2:4] = x2[2:4].T
    td = tdil(t2)
    m20 = xg[id2[0], 6]
    for k in range(N2):
        Tk = T2[:, k]; j = id2[k]
        xg[j, 4] = -math.atan2(Tk[0], Tk[1])
        ug[j, 0] = np.linalg.norm(Tk)
        if k > 0:
            dth = xg[j, 4] - xg[j - 1, 4]
            dt = (tau2[k] - tau2[k - 1]) * td
            xg[j - 1, 5] = dth / dt
            f_ = pb.alpha_e * ug[id2[:k + 1], 0]; g_ = tau2[:k + 1] * td
            xg[j, 6] = m20 + sum(0.5 * (g_[i + 1] - g_[i]) * (f_[i + 1] + f_[i]) for i in range(k))
    pg = np.zeros(10)
    pg[0] = t1; pg[1] = t2; pg[2:10] = xs
    return xg, ug, pg
