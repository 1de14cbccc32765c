"""ctypes binding of libscpb.so -- the C ABI declared in include/scpb.h.

This is the only way the Python host reaches the GPU path: there is no CPU fallback.
If the shared library is missing, or no CUDA device is present when a handle is created,
the call raises -- loudly -- instead of computing anything on the host.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libscpb.so")

MODEL_DBLINT, MODEL_ROCKET, MODEL_STARSHIP, MODEL_QUADROTOR, MODEL_FREEFLYER = 1, 2, 3, 4, 5
FOH, IMPULSE = 0, 1

_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

# name -> (restype, argtypes); must list EVERY symbol include/scpb.h declares
SIGNATURES = {
    "scpb_create": (C.c_int32, [C.c_int32, C.POINTER(C.c_void_p)]),
    "scpb_destroy": (C.c_int32, [C.c_void_p]),
    "scpb_last_error": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "scpb_version": (C.c_int32, []),
    "scpb_launch_count": (C.c_int64, [C.c_void_p]),
    "scpb_stream": (C.c_void_p, [C.c_void_p]),
    "scpb_sync": (C.c_int32, [C.c_void_p]),
    "scpb_model_set": (C.c_int32, [C.c_void_p, C.c_int32, _dp, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "scpb_discretize": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    _dp, _dp, _dp, _dp, _dp, C.c_double,
                                    _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _dp]),
    "scpb_discretize_dev": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
}


class ScpbError(RuntimeError):
    pass


def load():
    """dlopen libscpb.so and bind every exported symbol; raises if the library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise ScpbError(f"{SO} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        lib = C.CDLL(SO)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


class Handle:
    """Opaque library handle bound to one CUDA device + stream (include/scpb.h)."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.scpb_create(device, C.byref(h))
        if rc != 0 or not h.value:
            raise ScpbError(f"scpb_create(device={device}) failed with status {rc}: no usable CUDA device "
                            "(the product path has no CPU fallback)")
        self.h = h
        self.nx = self.nu = self.np = 0
        self.model_id = 0

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.scpb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            buf = C.create_string_buffer(512)
            self.lib.scpb_last_error(self.h, buf, 512)
            raise ScpbError(f"{what} failed ({rc}): {buf.value.decode(errors='replace')}")

    @property
    def launches(self) -> int:
        return int(self.lib.scpb_launch_count(self.h))

    @property
    def stream(self) -> int:
        return int(self.lib.scpb_stream(self.h) or 0)

    def sync(self):
        self._check(self.lib.scpb_sync(self.h), "scpb_sync")

    def model_set(self, model_id: int, par, nx: int, nu: int, np_: int):
        par, pp = _f64(par)
        self._check(self.lib.scpb_model_set(self.h, model_id, pp, par.size, nx, nu, np_), "scpb_model_set")
        self.model_id, self.nx, self.nu, self.np = model_id, nx, nu, np_

    def discretize(self, t_grid, xd, ud, p, iSx_diag, feas_tol, Nsub, method=FOH):
        """discretize! for a batch: xd (B, N, nx), ud (B, N, nu), p (B, np) host arrays.

        Returns a dict of Julia-layout arrays A (B, N-1, nx*nx) ..., feas (B,), seconds.
        """
        xd, pxd = _f64(xd)
        ud, pud = _f64(ud)
        p, pp = _f64(p)
        tg, ptg = _f64(t_grid)
        iS, piS = _f64(iSx_diag)
        B, N = xd.shape[0], xd.shape[1]
        nx, nu, np_ = self.nx, self.nu, self.np
        assert xd.shape == (B, N, nx) and ud.shape == (B, N, nu) and p.shape == (B, np_), \
            (xd.shape, ud.shape, p.shape)
        M = N - 1
        out = dict(A=np.empty((B, M, nx * nx)), Bm=np.empty((B, M, nx * nu)), Bp=np.empty((B, M, nx * nu)),
                   F=np.empty((B, M, nx * np_)), r=np.empty((B, M, nx)), E=np.empty((B, M, nx * nx)),
                   defect=np.empty((B, M, nx)), feas=np.zeros(B, dtype=np.int32))
        sec = C.c_double(0.0)
        g = lambda k: out[k].ctypes.data_as(_dp)
        rc = self.lib.scpb_discretize(self.h, method, B, N, Nsub, ptg, pxd, pud, pp, piS, float(feas_tol),
                                      g("A"), g("Bm"), g("Bp"), g("F"), g("r"), g("E"), g("defect"),
                                      out["feas"].ctypes.data_as(_ip), C.byref(sec))
        self._check(rc, "scpb_discretize")
        out["seconds"] = sec.value
        return out

    def discretize_dev(self, t_grid, xd, ud, p, iSx_diag, feas_tol, Nsub, A, Bm, Bp, F, r, E, defect, feas,
                       B, N, method=FOH):
        """Device-pointer variant; every array argument is an integer device address."""
        rc = self.lib.scpb_discretize_dev(self.h, method, B, N, Nsub, t_grid, xd, ud, p, iSx_diag,
                                          float(feas_tol), A, Bm, Bp, F, r, E, defect, feas)
        self._check(rc, "scpb_discretize_dev")
