"""Oracle-side problem data (numpy restatement of the example parameter files).
TEST INFRASTRUCTURE ONLY -- see oracle/scp_oracle.h.

Each problem object carries
  * the dimensions and the dynamics parameter vector `par` (layout shared with the
    product's device model packs, documented in include/scpb.h),
  * the variable ranges handed to problem_advise_scale! (-> compute_scaling, scp.jl:376-517),
  * numpy closures for the nonconvex constraints s/C/D/G, boundary conditions and cost,
  * `emit_X` / `emit_U`: the convex path constraints as rows added to an oracle cone program.
"""
from __future__ import annotations

import math

import numpy as np

from . import orc


def deg2rad(d):
    return d * math.pi / 180.0


class StarshipProblem:
    """starship_flip/parameters.jl:100-212 and definition.jl (PTR / SCvx flavour)."""

    name = "starship"
    model_id = orc.MODEL_STARSHIP
    nx, nu, np = 8, 3, 10

    def __init__(self, N: int):
        self.N = N
        g0 = 9.81
        self.g0 = g0
        rs, ls = 4.5, 50.0
        self.m = 120e3
        self.lcg = 0.4 * ls
        self.lcp = 0.45 * ls
        self.J = 1 / 12 * self.m * (6 * rs ** 2 + ls ** 2)
        vterm = 85
        CD = self.m * g0 / vterm ** 2
        CD *= 1.2
        self.CD = CD
        Isp = 330
        self.T_min1 = 880e3
        self.T_max1 = 2210e3
        self.T_min3 = 3 * self.T_min1
        self.T_max3 = 3 * self.T_max1
        self.alpha_e = -1 / (Isp * g0)
        self.delta_max = deg2rad(10.0)
        self.deltadot_max = 2 * self.delta_max
        self.rate_delay = 0.05
        # trajectory
        self.r0 = np.array([100.0, 600.0])
        self.v0 = np.array([0.0, -float(vterm)])
        self.theta0 = deg2rad(90.0)
        self.theta_s = deg2rad(-10.0)
        self.vs = np.array([0.0, -10.0])
        self.vf = np.array([0.0, -0.1])
        self.tf_min = 0.0
        self.tf_max = 40.0
        self.gamma_gs = deg2rad(27.0)
        self.thetamax2 = deg2rad(15.0)
        self.tau_s = 0.5
        self.hs = 100.0  # overwritten by the initial guess (definition.jl:195)
        self.ey = np.array([0.0, 1.0])

    # -- dynamics parameter vector (indices: scp_oracle.c enum SS_*) --
    def par(self):
        return np.array([self.m, self.J, self.lcg, self.lcp, self.CD, self.alpha_e, self.rate_delay,
                         self.g0, self.tau_s])

    def orc_model(self):
        return orc.make_model(self.model_id, self.nx, self.nu, self.np, self.par())

    # -- set_scale!, definition.jl:50-79 --
    def ranges(self):
        xrg = [(-100.0, 100.0), (0.0, self.r0[1]), (-10.0, 10.0), (self.v0[1], 0.0),
               (0.0, self.theta0), (deg2rad(-10.0), deg2rad(10.0)), (self.m - 1e3, self.m),
               (-self.delta_max, self.delta_max)]
        urg = [(self.T_min1, self.T_max3), (-self.delta_max, self.delta_max),
               (-self.deltadot_max, self.deltadot_max)]
        prg = [(0.0, self.tf_max), (0.0, self.tf_max)] + list(xrg)
        return xrg, urg, prg

    # -- terminal cost, definition.jl:454-478: phi(x, p) = c_x'x + c_p'p --
    def cost_terminal_lin(self):
        cx = np.zeros(self.nx)
        cp = np.zeros(self.np)
        mu = 0.3
        cp[2 + 1] = -mu / self.hs          # -mu * alt / hs, alt = p[id_xs][id_r][2]
        cx[6] = -1.0 / 10e3                # (0 - mf)/1e4
        return cx, cp, 0.0

    has_running_cost = False

    def cost_aff(self, x, u, p, t):
        """Original cost as an affine expression (terminal only; no running cost)."""
        cx, cp, c0 = self.cost_terminal_lin()
        e = c0
        for i in range(self.nx):
            if cx[i] != 0.0:
                e = e + x[i, -1] * cx[i]
        for i in range(self.np):
            if cp[i] != 0.0:
                e = e + p[i] * cp[i]
        return e

    def guess(self, N):
        from .starship_guess import starship_initial_guess
        return starship_initial_guess(self, N)

    # -- phase helpers, definition.jl:707-721 --
    def phase_switch(self, t):
        dt = 1 / (self.N - 1)
        tol = 1e-3
        return (self.tau_s - dt) + tol <= t and t <= self.tau_s + tol

    def phase2(self, t):
        return self.phase_switch(t) or t > self.tau_s

    ns = 7 + 2 * 8

    # -- nonconvex path constraints, definition.jl:725-807 --
    def s(self, t, k, x, u, p):
        s = np.zeros(self.ns)
        r = x[0:2]
        th = x[4]
        dd = x[7]
        de, ddot = u[1], u[2]
        s[0] = (de - dd) - ddot * self.rate_delay
        s[1] = ddot * self.rate_delay - (de - dd)
        s[2] = ddot - self.deltadot_max
        s[3] = -self.deltadot_max - ddot
        s[4] = np.linalg.norm(r) * math.cos(self.gamma_gs) - r @ self.ey
        if self.phase_switch(t):
            s[5:13] = p[2:10] - x
            s[13:21] = x - p[2:10]
        if self.phase2(t):
            s[-2] = th - self.thetamax2
            s[-1] = -self.thetamax2 - th
        return s

    def C(self, t, k, x, u, p):
        Cm = np.zeros((self.ns, self.nx))
        r = x[0:2]
        nr = np.linalg.norm(r)
        gn = np.zeros(2) if nr < math.sqrt(np.finfo(float).eps) else r / nr
        Cm[0, 7] = -1.0
        Cm[1, 7] = 1.0
        Cm[4, 0:2] = gn * math.cos(self.gamma_gs) - self.ey
        if self.phase_switch(t):
            Cm[5:13, :] = -np.eye(8)
            Cm[13:21, :] = np.eye(8)
        if self.phase2(t):
            Cm[-2, 4] = 1.0
            Cm[-1, 4] = -1.0
        return Cm

    def D(self, t, k, x, u, p):
        Dm = np.zeros((self.ns, self.nu))
        Dm[0, 1] = 1.0
        Dm[0, 2] = -self.rate_delay
        Dm[1, 1] = -1.0
        Dm[1, 2] = self.rate_delay
        Dm[2, 2] = 1.0
        Dm[3, 2] = -1.0
        return Dm

    def G(self, t, k, x, u, p):
        Gm = np.zeros((self.ns, self.np))
        if self.phase_switch(t):
            Gm[5:13, 2:10] = np.eye(8)
            Gm[13:21, 2:10] = -np.eye(8)
        return Gm

    # -- boundary conditions, definition.jl:812-873 --
    def gic(self, x, p):
        rhs = np.array([self.r0[0], self.r0[1], self.v0[0], self.v0[1], self.theta0, 0.0, 0.0])
        return x[0:7] - rhs

    def H0(self, x, p):
        H = np.zeros((7, self.nx))
        H[np.arange(7), np.arange(7)] = 1.0
        return H

    K0 = None

    def gtc(self, x, p):
        rhs = np.array([0.0, 0.0, self.vf[0], self.vf[1], 0.0, 0.0])
        return x[0:6] - rhs

    def Hf(self, x, p):
        H = np.zeros((6, self.nx))
        H[np.arange(6), np.arange(6)] = 1.0
        return H

    Kf = None

    # -- convex constraints, definition.jl:639-702.  `prg` is an oracle ConeProgram;
    #    x/u/p are affine-expression vectors in physical units --
    def emit_X(self, prg, t, k, x, p):
        v = x[2:4]
        tf1, tf2 = p[0], p[1]
        prg.nonpos([v[1] * 1.0], "no_climb")
        prg.nonpos([tf1 + tf2 - self.tf_max], "max_time")
        prg.nonpos([self.tf_min - (tf1 + tf2)], "min_time")

    def emit_U(self, prg, t, k, u, p):
        T, de = u[0], u[1]
        flip = t <= self.tau_s
        T_max = self.T_max3 if flip else self.T_max1
        T_min = self.T_min3 if flip else self.T_min1
        prg.nonpos([T - T_max], "max_thrust")
        prg.nonpos([T_min - T], "min_thrust")
        prg.l1([self.delta_max + 0.0 * de, de], "gimbal")


class _DynOnly:
    """Dynamics-only problem data (used by the discretization parity tests)."""

    def orc_model(self):
        return orc.make_model(self.model_id, self.nx, self.nu, self.np, self.par())


class DoubleIntegratorProblem(_DynOnly):
    """Double integrator with friction as a free-final-time, minimum-time PTR problem (BASELINE config C1).

    The reference ships this plant only as a fixed-time LCvx program (double_integrator/definition.jl:38-118 on
    parameters.jl:50-64: f = [x2; u - g], travel distance s, two parameter choices); the SCP form is a NEW definition
    on the same data: x = [position, velocity], |u| <= u_max = 2 (the outer bound of definition.jl:60-65),
    p = [tf], f = tf [x2; u - g], rest-to-rest over the distance s, cost tf.  Its continuous-time optimum is the
    bang-bang law of the maximum principle (cf. solve_mp, definition.jl:137-294), available in closed form:
    see t_opt()."""
    name = "dblint"
    model_id = orc.MODEL_DBLINT
    nx, nu, np = 2, 1, 1
    ns = 0

    def __init__(self, N: int, choice: int = 1):
        self.N = N
        self.g = 0.1 if choice == 1 else 0.6
        self.s = 47.0 if choice == 1 else 30.0
        self.T = 10.0
        self.u_max = 2.0
        self.tf_min, self.tf_max = 1.0, 30.0

    def par(self):
        return np.array([self.g])

    def t_opt(self):
        """minimum time: accelerate with a1 = u_max - g until t1, brake with a2 = u_max + g"""
        a1, a2 = self.u_max - self.g, self.u_max + self.g
        t1 = math.sqrt(2 * self.s * a2 / (a1 * (a1 + a2)))
        return t1 * (1 + a1 / a2), t1

    def ranges(self):
        vmax = 2.0 * self.s / 8.0
        return [(0.0, self.s), (0.0, vmax)], [(-self.u_max, self.u_max)], [(self.tf_min, self.tf_max)]

    has_running_cost = False

    def cost_aff(self, x, u, p, t):
        return p[0] * (1.0 / self.T)

    def guess(self, N):
        tau = np.arange(N) / (N - 1)
        xg = np.zeros((N, 2)); ug = np.zeros((N, 1))
        xg[:, 0] = tau * self.s
        xg[:, 1] = self.s / self.T
        return xg, ug, np.array([self.T])

    def gic(self, x, p):
        return x[0:2] - np.zeros(2)

    def H0(self, x, p):
        return np.eye(2)

    K0 = None

    def gtc(self, x, p):
        return x[0:2] - np.array([self.s, 0.0])

    def Hf(self, x, p):
        return np.eye(2)

    Kf = None

    def emit_X(self, prg, t, k, x, p):
        prg.nonpos([p[0] - self.tf_max], "max_time")
        prg.nonpos([self.tf_min - p[0]], "min_time")

    def emit_U(self, prg, t, k, u, p):
        prg.l1([self.u_max + 0.0 * u[0], u[0]], "input_bound")


class RocketProblem(_DynOnly):
    """Mars powered-descent guidance as a free-final-time PTR problem (BASELINE config C2).

    The reference ships this vehicle only as a single-shot LCvx program (rocket_landing/definition.jl:33-140); the
    SCP form below is a NEW definition on the same data (parameters.jl:78-150): state x = [r(3) v(3) z = ln m],
    input u = [a(3) xi] (thrust acceleration and its slack), parameter p = [tf], dynamics f = tf (A_c x + B_c u + p_c)
    (parameters.jl:110-121), constraints of definition.jl:93-131 with the mass profile z0 evaluated on a NOMINAL
    flight time (so that every constraint stays convex in (x, u, p): thrust-magnitude bounds mu = rho exp(-z0) are
    node constants), terminal cost -z_N (maximum final mass).  There is no nonconvex path constraint."""
    name = "rocket"
    model_id = orc.MODEL_ROCKET
    nx, nu, np = 7, 4, 1
    ns = 0

    def __init__(self, N: int):
        self.N = N
        ex, ey, ez = np.eye(3)
        self.g = -3.7114 * ez
        th = 30 * math.pi / 180
        T_sidereal_mars = 24.6229 * 3600
        self.omega = (2 * math.pi / T_sidereal_mars) * (ex * math.cos(th) + ey * 0 + ez * math.sin(th))
        self.m_dry, self.m_wet, self.Isp = 1505.0, 1905.0, 225.0
        n_eng = 6
        self.phi = 27 * math.pi / 180
        T_max = 3.1e3
        self.rho_min = n_eng * 0.3 * T_max * math.cos(self.phi)
        self.rho_max = n_eng * 0.8 * T_max * math.cos(self.phi)
        self.gamma_gs = 86 * math.pi / 180
        self.gamma_p = 40 * math.pi / 180
        self.v_max = 500 * 1e3 / 3600
        self.r0 = (2 * ex + 0 * ey + 1.5 * ez) * 1e3
        self.v0 = 80 * ex + 30 * ey - 75 * ez
        self.alpha = 1 / (self.Isp * 9.807 * math.cos(self.phi))
        self.tf_min, self.tf_max, self.tf_nom = 50.0, 110.0, 75.0
        cg, sg = math.cos(self.gamma_gs), math.sin(self.gamma_gs)
        self.H_gs = np.array([[cg, 0, -sg], [-cg, 0, -sg], [0, cg, -sg], [0, -cg, -sg]])

    def par(self):
        return np.concatenate([self.g, self.omega, [self.alpha]])

    # node constants of the thrust / mass bounds (definition.jl:93-104) on the nominal time grid
    def z0(self, tau):
        return math.log(self.m_wet - self.alpha * self.rho_max * tau * self.tf_nom)

    def z1(self, tau):
        return math.log(self.m_wet - self.alpha * self.rho_min * tau * self.tf_nom)

    def ranges(self):
        amax = self.rho_max / self.m_dry
        xrg = [(-500.0, 2500.0), (-500.0, 500.0), (0.0, 1600.0), (-100.0, 100.0), (-100.0, 100.0), (-100.0, 100.0),
               (math.log(self.m_dry), math.log(self.m_wet))]
        urg = [(-amax * math.sin(self.gamma_p), amax * math.sin(self.gamma_p))] * 2 + [(0.0, amax), (0.0, amax)]
        prg = [(self.tf_min, self.tf_max)]
        return xrg, urg, prg

    has_running_cost = False

    def cost_aff(self, x, u, p, t):
        dz = math.log(self.m_wet) - math.log(self.m_dry)
        return (x[6, -1] - math.log(self.m_wet)) * (-1.0 / dz)      # fraction of the propellant budget spent

    def guess(self, N):
        tau = np.arange(N) / (N - 1)
        xg = np.zeros((N, 7)); ug = np.zeros((N, 4))
        for k in range(N):
            xg[k, 0:3] = (1 - tau[k]) * self.r0
            xg[k, 3:6] = (1 - tau[k]) * self.v0
            xg[k, 6] = (1 - tau[k]) * math.log(self.m_wet) + tau[k] * math.log(0.5 * (self.m_dry + self.m_wet))
            ug[k, 0:3] = -self.g
            ug[k, 3] = np.linalg.norm(self.g)
        return xg, ug, np.array([self.tf_nom])

    # boundary conditions (definition.jl:124-131); the final-mass inequality is a state constraint at node N
    def gic(self, x, p):
        rhs = np.concatenate([self.r0, self.v0, [math.log(self.m_wet)]])
        return x[0:7] - rhs

    def H0(self, x, p):
        return np.eye(7)

    K0 = None

    def gtc(self, x, p):
        return x[0:6] - np.zeros(6)

    def Hf(self, x, p):
        H = np.zeros((6, 7)); H[np.arange(6), np.arange(6)] = 1.0
        return H

    Kf = None

    def emit_X(self, prg, t, k, x, p):
        r, v, z = x[0:3], x[3:6], x[6]
        prg.nonpos([self.z0(t) - z], "mass_lower")
        prg.nonpos([z - self.z1(t)], "mass_upper")
        for i in range(4):
            prg.nonpos([r[0] * self.H_gs[i, 0] + r[1] * self.H_gs[i, 1] + r[2] * self.H_gs[i, 2]], "glide_slope")
        prg.soc([self.v_max + 0.0 * v[0], v[0], v[1], v[2]], "max_speed")
        if k == self.N:
            prg.nonpos([math.log(self.m_dry) - z], "dry_mass")
        prg.nonpos([p[0] - self.tf_max], "max_time")
        prg.nonpos([self.tf_min - p[0]], "min_time")

    def emit_U(self, prg, t, k, u, p):
        a, xi = u[0:3], u[3]
        prg.soc([xi, a[0], a[1], a[2]], "lcvx_equality")
        prg.nonpos([xi * math.cos(self.gamma_p) - a[2]], "pointing")
        # thrust magnitude bounds: the mass entering mu = rho/m is the node constant exp(z0) (nominal profile)
        mu_min, mu_max = self.rho_min * math.exp(-self.z0(t)), self.rho_max * math.exp(-self.z0(t))
        prg.nonpos([mu_min - xi], "min_thrust")
        prg.nonpos([xi - mu_max], "max_thrust")


class QuadrotorProblem(_DynOnly):
    """quadrotor/parameters.jl:96-135 and definition.jl (scaling advice :48-58, guess :61-91, cost :94-138, input set
    :188-235, obstacles :237-270, boundary conditions :272-330) in their GuSTO flavour (BASELINE config C4)."""
    name = "quadrotor"
    model_id = orc.MODEL_QUADROTOR
    nx, nu, np = 6, 4, 1
    ns = 2

    def __init__(self, N: int):
        self.N = N
        self.g = np.array([0.0, 0.0, -9.81])   # parameters.jl:83-84: g = -gnrm * e_z
        self.u_max, self.u_min, self.tilt_max = 23.2, 0.6, deg2rad(60)
        self.tf_min, self.tf_max = 0.0, 2.5
        self.gamma = 0.0
        self.obs_H = [np.diag([2.0, 2.0, 0.0]), np.diag([1.5, 1.5, 0.0])]
        self.obs_c = [np.array([1.0, 2.0, 0.0]), np.array([2.0, 5.0, 0.0])]
        self.r0, self.v0 = np.zeros(3), np.zeros(3)
        self.rf, self.vf = np.array([2.5, 6.0, 0.0]), np.zeros(3)

    def par(self):
        ob = []
        for H, c in zip(self.obs_H, self.obs_c):
            ob += list(H.flatten(order="F")) + list(c)
        return np.concatenate([self.g, ob])

    def ranges(self):
        return [None] * 6, [None] * 4, [(self.tf_min, self.tf_min + 1.0 * (self.tf_max - self.tf_min))]

    def guess(self, N):
        x0 = np.concatenate([self.r0, self.v0]); xf = np.concatenate([self.rf, self.vf])
        x = np.array([(1 - k / (N - 1)) * x0 + (k / (N - 1)) * xf for k in range(N)])
        hover = np.concatenate([-self.g, [np.linalg.norm(self.g)]])
        return x, np.tile(hover, (N, 1)), np.array([0.5 * (self.tf_min + self.tf_max)])

    def S(self, t, k, p):                       # definition.jl:122-133
        S = np.zeros((4, 4))
        S[3, 3] = (1 - self.gamma) / np.linalg.norm(self.g) ** 2
        return S

    def phi(self, x, p):                        # definition.jl:96-103
        return self.gamma * (p[0] / self.tf_max) ** 2

    def s(self, t, k, x, u, p):                 # definition.jl:240-250
        return np.array([1 - np.linalg.norm(H @ (x[0:3] - c)) for H, c in zip(self.obs_H, self.obs_c)])

    def C(self, t, k, x, u, p):                 # definition.jl:253-262
        C = np.zeros((2, 6))
        for i, (H, c) in enumerate(zip(self.obs_H, self.obs_c)):
            y = H @ (x[0:3] - c)
            C[i, 0:3] = -(H.T @ y) / np.linalg.norm(y)
        return C

    def D(self, t, k, x, u, p):
        return np.zeros((2, 4))

    def G(self, t, k, x, u, p):
        return np.zeros((2, 1))

    def gic(self, x, p):
        return x[0:6] - np.concatenate([self.r0, self.v0])

    def H0(self, x, p):
        return np.eye(6)

    K0 = None

    def gtc(self, x, p):
        return x[0:6] - np.concatenate([self.rf, self.vf])

    def Hf(self, x, p):
        return np.eye(6)

    Kf = None

    def emit_U(self, prg, t, k, u, p):          # definition.jl:188-235
        a, sg = u[0:3], u[3]
        prg.nonpos([self.u_min - sg], "min_accel")
        prg.nonpos([sg - self.u_max], "max_accel")
        prg.soc([sg, a[0], a[1], a[2]], "lcvx_equality")
        prg.nonpos([sg * math.cos(self.tilt_max) - a[2]], "max_tilt")
        prg.nonpos([p[0] - self.tf_max], "max_duration")
        prg.nonpos([self.tf_min - p[0]], "min_duration")


class FreeFlyerProblem(_DynOnly):
    """freeflyer/parameters.jl:105-190 and definition.jl (scaling advice :52-67, guess :84-167, cost :170-222, convex
    sets :286-375, nonconvex constraints :376-442, boundary conditions :455-520); np = 1 + 6N (room SDF slack per node).
    The PTR instance (BASELINE config C5) uses the SCvx flavour of the running cost (definition.jl:190-204)."""
    name = "freeflyer"
    model_id = orc.MODEL_FREEFLYER
    nx, nu = 13, 6
    ns = 4

    def __init__(self, N: int):
        self.N = N
        self.n_iss, self.n_obs = 6, 3
        self.np = 1 + self.n_iss * N
        self.mass = 7.2
        self.J = np.diag([0.1083, 0.1083, 0.1083])
        self.v_max, self.omega_max = 0.4, deg2rad(1)
        self.T_max, self.M_max = 20e-3, 1e-4
        self.tf_min, self.tf_max = 60.0, 200.0
        self.gamma, self.hom, self.eps_sdf = 0.0, 50.0, 1e-4
        z_iss = 4.75
        self.obs_H = [np.diag([1.0, 1.0, 1.0]) / 0.3] * 3
        self.obs_c = [np.array([8.5, -0.15, 5.0]), np.array([11.2, 1.84, 5.0]), np.array([11.3, 3.8, 4.8])]
        self.rooms = [self._room([6.0, 0.0, z_iss], 1.0, 1.0, 1.5, 0.0, 90.0),
                      self._room([7.5, 0.0, z_iss], 2.0, 2.0, 4.0, 0.0, 90.0),
                      self._room([11.5, 0.0, z_iss], 1.25, 1.25, 0.5, 0.0, 90.0),
                      self._room([10.75, -1.0, z_iss], 1.5, 1.5, 1.5, -90.0, 90.0),
                      self._room([10.75, 1.0, z_iss], 1.5, 1.5, 1.5, 90.0, 90.0),
                      self._room([10.75, 2.5, z_iss], 2.5, 2.5, 4.5, 90.0, 90.0)]
        self.r0 = np.array([6.5, -0.2, 5.0]); self.v0 = np.array([0.035, 0.035, 0.0])
        self.q0 = self._qaa(deg2rad(-40), [0.0, 1.0, 1.0]); self.w0 = np.zeros(3)
        self.rf = np.array([11.3, 6.0, 4.5]); self.vf = np.zeros(3)
        self.qf = self._qaa(0.0, [0.0, 0.0, 1.0]); self.wf = np.zeros(3)

    # ---- geometry / quaternion helpers (src/utils/hyperrectangle.jl:85-138, quaternion.jl) ----
    @staticmethod
    def _room(offset, width, height, depth, yaw, pitch):
        cd = lambda a: math.cos(math.radians(a)); sd = lambda a: math.sin(math.radians(a))
        lo = np.array([-width / 2, -height / 2, 0.0]); hi = np.array([width / 2, height / 2, depth])
        Rz = np.array([[cd(yaw), -sd(yaw), 0], [sd(yaw), cd(yaw), 0], [0, 0, 1]])
        Ry = np.array([[cd(pitch), 0, sd(pitch)], [0, 1, 0], [-sd(pitch), 0, cd(pitch)]])
        R = Rz @ Ry
        lr, ur = R @ lo, R @ hi
        l = np.minimum(lr, ur) + np.asarray(offset, float); u = np.maximum(lr, ur) + np.asarray(offset, float)
        return (u + l) / 2, (u - l) / 2

    @staticmethod
    def _qaa(alpha, axis):
        a = np.asarray(axis, float); a = a / np.linalg.norm(a)
        return np.concatenate([a * math.sin(alpha / 2), [math.cos(alpha / 2)]])

    @staticmethod
    def _qmul(q, p):
        return np.concatenate([q[3] * p[:3] + np.cross(q[:3], p[:3]) + q[:3] * p[3], [q[3] * p[3] - q[:3] @ p[:3]]])

    @staticmethod
    def _qlog(q):
        nv = np.linalg.norm(q[:3])
        return 2 * math.atan2(nv, q[3]), q[:3] / nv

    def idd(self, i, k):
        return 1 + i + self.n_iss * k

    def par(self):
        ob = []
        for H, c in zip(self.obs_H, self.obs_c):
            ob += list(H.flatten(order="F")) + list(c)
        return np.concatenate([[self.mass], self.J.flatten(order="F"), np.linalg.inv(self.J).flatten(order="F"),
                               [self.hom], ob])

    def ranges(self):
        lo, hi = np.minimum(self.r0, self.rf), np.maximum(self.r0, self.rf)
        xrg = [(lo[i], hi[i]) for i in range(3)] + [None] * 10
        urg = [None] * 6
        prg = [(self.tf_min, self.tf_max)] + [(-100.0, 1.0)] * (self.np - 1)
        return xrg, urg, prg

    def cost_emit(self, prg, x, u, p, t):
        """terminal eps_sdf * sum(-delta) (+ gamma (tdil/tdil_max)^2) and the trapezoid rule of the convex running cost
        (1 - gamma) (|T|^2 / T_max^2 + |M|^2 / M_max^2), one rotated second-order cone per node"""
        from . import conic
        from .ptr import trapz
        J = conic.Aff()
        for i in range(1, self.np):
            J = J + p[i] * (-self.eps_sdf)
        run = []
        for k in range(self.N):
            q = prg.new_variable(1, f"_q{k}")[0]
            es = [u[i, k] * (1.0 / self.T_max) for i in range(3)] + [u[3 + i, k] * (1.0 / self.M_max) for i in range(3)]
            prg.soc([q + 1.0] + [e * 2.0 for e in es] + [q - 1.0], "input_energy")
            run.append(q * (1.0 - self.gamma))
        return J + trapz(run, t)

    def guess(self, N):
        p = np.zeros(self.np)
        ft = 0.5 * (self.tf_min + self.tf_max)
        p[0] = ft
        x = np.zeros((N, 13))
        speed = np.abs(self.rf - self.r0).sum() / ft
        times = np.array([(1 - k / (N - 1)) * 0.0 + (k / (N - 1)) * ft for k in range(N)])
        cum = np.cumsum(np.abs(self.rf - self.r0) / speed)
        for k in range(N):
            tk = min(times[k], cum[2])
            for i in range(3):
                if tk <= cum[i]:
                    t0 = cum[i - 1] if i > 0 else 0.0
                    tf = cum[i]
                    r0 = self.r0.copy(); r0[:i] = self.rf[:i]
                    rf = r0.copy(); rf[i] = self.rf[i]
                    c = (tf - tk) / (tf - t0)
                    x[k, 0:3] = c * r0 + (1 - c) * rf
                    d = rf - r0
                    x[k, 3:6] = speed * d / np.linalg.norm(d)
                    break
        qc = np.concatenate([-self.q0[:3], [self.q0[3]]])
        da, dax = self._qlog(self._qmul(qc, self.qf))
        for k in range(N):
            tau = max(0.0, min(1.0, k / (N - 1)))
            x[k, 6:10] = self._qmul(self.q0, self._qaa(tau * da, dax))
        ang, ax = self._qlog(self._qmul(self.qf, qc))
        x[:, 10:13] = (ang / ft) * ax
        for i in range(self.n_iss):
            c, s_ = self.rooms[i]
            for k in range(N):
                p[self.idd(i, k)] = 1 - np.abs((x[k, 0:3] - c) / s_).max()
        return x, np.zeros((N, 6)), p

    # nonconvex path constraints (definition.jl:376-442); k is 1-based
    def _ell(self, r):
        s = np.zeros(self.n_obs); g = np.zeros((self.n_obs, 3))
        for i in range(self.n_obs):
            y = self.obs_H[i] @ (r - self.obs_c[i])
            E = np.linalg.norm(y)
            s[i] = 1 - E
            g[i] = -(self.obs_H[i].T @ y) / E
        return s, g

    def _lse(self, d):
        a = np.max(self.hom * d)
        e = np.exp(self.hom * d - a)
        return (a + math.log(e.sum())) / self.hom, e / e.sum()

    def s(self, t, k, x, u, p):
        s, _ = self._ell(x[0:3])
        d, _ = self._lse(np.array([p[self.idd(i, k - 1)] for i in range(self.n_iss)]))
        return np.concatenate([s, [-d]])

    def C(self, t, k, x, u, p):
        C = np.zeros((self.ns, 13))
        C[:self.n_obs, 0:3] = self._ell(x[0:3])[1]
        return C

    def D(self, t, k, x, u, p):
        return np.zeros((self.ns, 6))

    def G(self, t, k, x, u, p):
        G = np.zeros((self.ns, self.np))
        _, w = self._lse(np.array([p[self.idd(i, k - 1)] for i in range(self.n_iss)]))
        for i in range(self.n_iss):
            G[self.ns - 1, self.idd(i, k - 1)] = -w[i]
        return G

    def gic(self, x, p):
        return x[0:13] - np.concatenate([self.r0, self.v0, self.q0, self.w0])

    def H0(self, x, p):
        return np.eye(13)

    K0 = None

    def gtc(self, x, p):
        return x[0:13] - np.concatenate([self.rf, self.vf, self.qf, self.wf])

    def Hf(self, x, p):
        return np.eye(13)

    Kf = None

    def emit_X(self, prg, t, k, x, p):
        r, v, w = x[0:3], x[3:6], x[10:13]
        prg.soc([self.v_max + 0.0 * v[0], v[0], v[1], v[2]], "max_lin_vel")
        prg.soc([self.omega_max + 0.0 * w[0], w[0], w[1], w[2]], "max_ang_vel")
        prg.nonpos([p[0] - self.tf_max], "max_duration")
        prg.nonpos([self.tf_min - p[0]], "min_duration")
        for i in range(self.n_iss):
            c, s_ = self.rooms[i]
            d = p[self.idd(i, k - 1)]
            prg.linf([1.0 - d] + [(r[j] - c[j]) * (1.0 / s_[j]) for j in range(3)], f"room_sdf_{i + 1}")

    def emit_U(self, prg, t, k, u, p):
        prg.soc([self.T_max + 0.0 * u[0], u[0], u[1], u[2]], "max_thrust")
        prg.soc([self.M_max + 0.0 * u[3], u[3], u[4], u[5]], "max_torque")


def test_trajectory(pb, nb: int, N: int, seed: int = 0):
    """Seeded, physically plausible random trajectories for parity tests: (xd, ud, p)."""
    rng = np.random.default_rng(seed)
    nx, nu, np_ = pb.nx, pb.nu, pb.np
    if pb.name == "starship":
        x0 = np.array([50., 300., 3., -40., 0.7, 0.02, -100., 0.05])
        u0 = np.array([3e6, 0.05, 0.01])
        p0 = np.concatenate([[12., 15.], x0])
    elif pb.name == "dblint":
        x0 = np.array([10., 2.]); u0 = np.array([0.5]); p0 = np.array([10.])
    elif pb.name == "rocket":
        x0 = np.array([1500., 100., 1200., 60., 20., -50., math.log(1800.)])
        u0 = np.array([1., 0.5, 5., 6.]); p0 = np.array([60.])
    elif pb.name == "quadrotor":
        x0 = np.array([1., 2., 0.5, 1., -0.5, 0.2]); u0 = np.array([0.5, -0.3, 9.9, 10.]); p0 = np.array([2.])
    elif pb.name == "freeflyer":
        q = np.array([0.1, -0.3, 0.2, 0.9]); q /= np.linalg.norm(q)
        x0 = np.concatenate([[7., 0.5, 4.8], [0.03, 0.02, -0.01], q, [0.005, -0.004, 0.003]])
        u0 = np.array([5e-3, -3e-3, 2e-3, 2e-5, -1e-5, 3e-5])
        p0 = np.concatenate([[100.], -1.0 + 0.1 * np.arange(np_ - 1) / max(np_ - 1, 1)])
    else:
        raise KeyError(pb.name)
    xd = x0 * (1 + 0.03 * rng.standard_normal((nb, N, nx))) + 1e-3 * rng.standard_normal((nb, N, nx))
    ud = u0 * (1 + 0.05 * rng.standard_normal((nb, N, nu)))
    p = p0 * (1 + 0.05 * rng.standard_normal((nb, np_)))
    if pb.name == "freeflyer":
        xd[..., 6:10] /= np.linalg.norm(xd[..., 6:10], axis=-1, keepdims=True)
    return xd, ud, p


def make_problem(name: str, N: int):
    cls = {"starship": StarshipProblem, "dblint": DoubleIntegratorProblem, "rocket": RocketProblem,
           "quadrotor": QuadrotorProblem, "freeflyer": FreeFlyerProblem}[name]
    return cls(N)
