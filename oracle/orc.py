"""ctypes front-end of the CPU oracle (oracle/scp_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.

Arrays follow Julia's column-major memory: a trajectory x[nx, N] is passed as a
C-contiguous numpy array of shape (N, nx) (node-major == Julia column-major).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODEL_DBLINT, MODEL_ROCKET, MODEL_STARSHIP, MODEL_QUADROTOR, MODEL_FREEFLYER, MODEL_RENDEZVOUS2D = 1, 2, 3, 4, 5, 6
MAX_PAR = 64


class OrcModel(C.Structure):
    _fields_ = [("model_id", C.c_int), ("nx", C.c_int), ("nu", C.c_int), ("np", C.c_int),
                ("par", C.c_double * MAX_PAR)]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "scp_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        mp = C.POINTER(OrcModel)
        for name in ("orc_f", "orc_A", "orc_B", "orc_F"):
            fn = getattr(_LIB, name)
            fn.argtypes = [mp, C.c_double, C.c_int, dp, dp, dp, dp]
            fn.restype = None
        _LIB.orc_discretize_foh.argtypes = [mp, C.c_int, C.c_int, dp, dp, dp, dp, dp, C.c_double,
                                            dp, dp, dp, dp, dp, dp, dp, C.POINTER(C.c_int)]
        _LIB.orc_discretize_foh.restype = C.c_int
        _LIB.orc_discretize_foh_batch.argtypes = [mp, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp,
                                                  C.c_double, dp, dp, dp, dp, dp, dp, dp,
                                                  C.POINTER(C.c_int), C.c_int]
        _LIB.orc_discretize_foh_batch.restype = C.c_int
        _LIB.orc_propagate_foh.argtypes = [mp, C.c_int, C.c_int, dp, dp, dp, dp, dp]
        _LIB.orc_propagate_foh.restype = C.c_int
        _LIB.orc_discretize_impulse.argtypes = [mp, C.c_int, C.c_int, dp, dp, dp, dp, dp, C.c_double,
                                                dp, dp, dp, dp, dp, dp, C.POINTER(C.c_int)]
        _LIB.orc_discretize_impulse.restype = C.c_int
        _LIB.orc_propagate_impulse.argtypes = [mp, C.c_int, C.c_int, dp, dp, dp, dp, dp]
        _LIB.orc_propagate_impulse.restype = C.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_model(model_id: int, nx: int, nu: int, np_: int, par) -> OrcModel:
    m = OrcModel()
    m.model_id, m.nx, m.nu, m.np = model_id, nx, nu, np_
    par = np.asarray(par, dtype=np.float64)
    assert par.size <= MAX_PAR
    for i, v in enumerate(par):
        m.par[i] = float(v)
    return m


def dyn_eval(m: OrcModel, t, k, x, u, p):
    """f, A, B, F at one point; matrices returned as (rows, cols) numpy arrays."""
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    f = np.zeros(m.nx)
    A = np.zeros(m.nx * m.nx)
    B = np.zeros(m.nx * m.nu)
    F = np.zeros(m.nx * m.np)
    L.orc_f(C.byref(m), t, k, _p(x), _p(u), _p(p), _p(f))
    L.orc_A(C.byref(m), t, k, _p(x), _p(u), _p(p), _p(A))
    L.orc_B(C.byref(m), t, k, _p(x), _p(u), _p(p), _p(B))
    L.orc_F(C.byref(m), t, k, _p(x), _p(u), _p(p), _p(F))
    return (f, A.reshape(m.nx, m.nx, order="F"), B.reshape(m.nx, m.nu, order="F"),
            F.reshape(m.nx, m.np, order="F"))


def t_grid(N: int) -> np.ndarray:
    """RealVector(LinRange(0, 1, N)) with Julia's lerpi arithmetic (scp.jl:147)."""
    j = np.arange(N, dtype=np.float64) / float(N - 1)
    return (1.0 - j) * 0.0 + j * 1.0


class DLTV:
    """discretization.jl:28-84 -- arrays indexed [k] over segments, matrices (rows, cols)."""

    def __init__(self, A, Bm, Bp, F, r, E, defect, feas):
        self.A, self.Bm, self.Bp, self.F, self.r, self.E = A, Bm, Bp, F, r, E
        self.defect, self.feas = defect, feas


def discretize(m: OrcModel, xd, ud, p, Nsub, iSx_diag, feas_tol, tg=None) -> DLTV:
    """discretize! (FOH) for one trajectory; xd (N, nx), ud (N, nu), p (np,)."""
    xd = np.ascontiguousarray(xd, dtype=np.float64)
    ud = np.ascontiguousarray(ud, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    N = xd.shape[0]
    nx, nu, np_ = m.nx, m.nu, m.np
    tg = t_grid(N) if tg is None else np.ascontiguousarray(tg, dtype=np.float64)
    iS = np.ascontiguousarray(iSx_diag, dtype=np.float64)
    A = np.zeros((N - 1, nx * nx)); Bm = np.zeros((N - 1, nx * nu)); Bp = np.zeros((N - 1, nx * nu))
    F = np.zeros((N - 1, nx * np_)); r = np.zeros((N - 1, nx)); E = np.zeros((N - 1, nx * nx))
    dfc = np.zeros((N - 1, nx))
    feas = C.c_int(0)
    rc = lib().orc_discretize_foh(C.byref(m), N, Nsub, _p(tg), _p(xd), _p(ud), _p(p), _p(iS), feas_tol,
                                  _p(A), _p(Bm), _p(Bp), _p(F), _p(r), _p(E), _p(dfc), C.byref(feas))
    if rc != 0:
        raise RuntimeError("oracle discretize failed (singular Phi)")
    col = lambda a, rr, cc: a.reshape(N - 1, cc, rr).transpose(0, 2, 1)
    return DLTV(col(A, nx, nx), col(Bm, nx, nu), col(Bp, nx, nu), col(F, nx, np_), r, col(E, nx, nx),
                dfc, bool(feas.value))


def discretize_batch(m: OrcModel, xd, ud, p, Nsub, iSx_diag, feas_tol, nthreads=0):
    """Raw batched call (CPU baseline): xd (B, N, nx) etc.; returns flat Julia-layout arrays."""
    xd = np.ascontiguousarray(xd, dtype=np.float64)
    ud = np.ascontiguousarray(ud, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    nb, N = xd.shape[0], xd.shape[1]
    nx, nu, np_ = m.nx, m.nu, m.np
    tg = t_grid(N)
    iS = np.ascontiguousarray(iSx_diag, dtype=np.float64)
    A = np.zeros((nb, N - 1, nx * nx)); Bm = np.zeros((nb, N - 1, nx * nu)); Bp = np.zeros((nb, N - 1, nx * nu))
    F = np.zeros((nb, N - 1, nx * np_)); r = np.zeros((nb, N - 1, nx)); E = np.zeros((nb, N - 1, nx * nx))
    dfc = np.zeros((nb, N - 1, nx))
    feas = np.zeros(nb, dtype=np.int32)
    rc = lib().orc_discretize_foh_batch(C.byref(m), nb, N, Nsub, _p(tg), _p(xd), _p(ud), _p(p), _p(iS),
                                        feas_tol, _p(A), _p(Bm), _p(Bp), _p(F), _p(r), _p(E), _p(dfc),
                                        feas.ctypes.data_as(C.POINTER(C.c_int)), nthreads)
    if rc != 0:
        raise RuntimeError("oracle discretize failed")
    return dict(A=A, Bm=Bm, Bp=Bp, F=F, r=r, E=E, defect=dfc, feas=feas)


def propagate(m: OrcModel, xd, ud, p, res):
    xd = np.ascontiguousarray(xd, dtype=np.float64)
    ud = np.ascontiguousarray(ud, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    N = xd.shape[0]
    tg = t_grid(N)
    xc = np.zeros((res, m.nx))
    lib().orc_propagate_foh(C.byref(m), N, res, _p(tg), _p(xd), _p(ud), _p(p), _p(xc))
    return xc


def discretize_impulse(m: OrcModel, xd, ud, p, Nsub, iSx_diag, feas_tol, tg=None) -> DLTV:
    """discretize! (IMPULSE, discretization.jl:186-193, 304-340, 384-390) for one trajectory; Bp is None."""
    xd = np.ascontiguousarray(xd, dtype=np.float64)
    ud = np.ascontiguousarray(ud, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    N = xd.shape[0]
    nx, nu, np_ = m.nx, m.nu, m.np
    tg = t_grid(N) if tg is None else np.ascontiguousarray(tg, dtype=np.float64)
    iS = np.ascontiguousarray(iSx_diag, dtype=np.float64)
    A = np.zeros((N - 1, nx * nx)); Bm = np.zeros((N - 1, nx * nu))
    F = np.zeros((N - 1, nx * np_)); r = np.zeros((N - 1, nx)); E = np.zeros((N - 1, nx * nx))
    dfc = np.zeros((N - 1, nx))
    feas = C.c_int(0)
    rc = lib().orc_discretize_impulse(C.byref(m), N, Nsub, _p(tg), _p(xd), _p(ud), _p(p), _p(iS), feas_tol,
                                      _p(A), _p(Bm), _p(F), _p(r), _p(E), _p(dfc), C.byref(feas))
    if rc != 0:
        raise RuntimeError("oracle discretize failed (singular Phi)")
    col = lambda a, rr, cc: a.reshape(N - 1, cc, rr).transpose(0, 2, 1)
    return DLTV(col(A, nx, nx), col(Bm, nx, nu), None, col(F, nx, np_), r, col(E, nx, nx), dfc, bool(feas.value))


def propagate_impulse(m: OrcModel, xd, ud, p, res):
    """propagate, IMPULSE branch (discretization.jl:539-558): (1 + (N-1)*ceil(res/(N-1)), nx)."""
    xd = np.ascontiguousarray(xd, dtype=np.float64)
    ud = np.ascontiguousarray(ud, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    N = xd.shape[0]
    tg = t_grid(N)
    sub = -(-res // (N - 1))
    xc = np.zeros((1 + (N - 1) * sub, m.nx))
    lib().orc_propagate_impulse(C.byref(m), N, res, _p(tg), _p(xd), _p(ud), _p(p), _p(xc))
    return xc
