"""ctypes front-end of the scalar C oracle (TEST INFRASTRUCTURE ONLY -- see oracle/ens_oracle.c).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this module; the product (``assimulo_b200``) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libens_oracle.so")

MAX_N = 16
SOLVERS = {"ExplicitEuler": 0, "RungeKutta4": 1, "RungeKutta34": 2, "Dopri5": 3, "Radau5ODE": 4, "ImplicitEuler": 5,
           "RodasODE": 6}
MODELS = {"vdp": 0, "ball": 1, "robertson": 2, "gyro": 3, "linear": 4, "linear_ev": 5, "extended": 6, "time_ev": 7}
STATS = ["nsteps", "nfcns", "nerrfails", "njacs", "nlus", "nstatefcns", "nstateevents", "nstepstotal",
         "ntimeevents", "nniters", "nnfails"]


class Opts(C.Structure):
    _fields_ = [("solver", C.c_int), ("maxsteps", C.c_int), ("newt", C.c_int), ("usejac", C.c_int),
                ("rtol", C.c_double), ("atol", C.c_double * MAX_N), ("h", C.c_double),
                ("inith", C.c_double), ("maxh", C.c_double), ("safe", C.c_double), ("fac1", C.c_double),
                ("fac2", C.c_double), ("beta", C.c_double), ("thet", C.c_double), ("fnewt", C.c_double),
                ("quot1", C.c_double), ("quot2", C.c_double)]


def build(force=False):
    src = [os.path.join(HERE, "ens_oracle.c"), os.path.join(HERE, "ens_oracle.h")]
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= max(os.path.getmtime(s) for s in src)):
        return LIB
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-strict-aliasing", "-fPIC", "-shared",
                           "-fopenmp", src[0], "-o", LIB, "-lm"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ora_default_opts.argtypes = [C.c_int, C.c_int, C.POINTER(Opts)]
        _lib.ora_model_dims.argtypes = [C.c_int] + [C.POINTER(C.c_int)] * 4
        _lib.ora_simulate.restype = C.c_int
        _lib.ora_simulate.argtypes = [C.c_int, C.POINTER(Opts), C.c_long, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p]
    return _lib


def model_dims(model):
    v = [C.c_int() for _ in range(4)]
    lib().ora_model_dims(MODELS[model], *[C.byref(x) for x in v])
    return tuple(x.value for x in v)  # n, np, ng, nsw


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def simulate(model, solver, y0, p=None, sw0=None, t0=0.0, tf=1.0, t_grid=None, max_ev=0, **options):
    """Run the C oracle on N members.  y0: [N,n]; p: [N,np]; sw0: [N,nsw] (bool).

    options: rtol, atol (scalar or [n]), h, inith, maxh, maxsteps, safe, fac1, fac2, beta, newt, thet, fnewt,
    quot1, quot2, usejac.  Returns a dict with y [N,n], t [N], sw, status [N], stats {name: [N]},
    dense [N,n_grid,n] (or None), ev_t [N,max_ev], ev_info [N,max_ev].
    """
    n, npar, ng, nsw = model_dims(model)
    y0 = np.ascontiguousarray(np.atleast_2d(np.asarray(y0, dtype=np.float64)))
    N = y0.shape[0]
    assert y0.shape[1] == n, (y0.shape, n)
    if npar:
        p = np.ascontiguousarray(np.broadcast_to(np.asarray(p, dtype=np.float64).reshape(-1, npar), (N, npar)))
    else:
        p = None
    if nsw:
        sw0 = np.ascontiguousarray(np.broadcast_to(np.asarray(sw0, dtype=np.uint8).reshape(-1, nsw), (N, nsw)))
    else:
        sw0 = None
    o = Opts()
    lib().ora_default_opts(SOLVERS[solver], n, C.byref(o))
    for k, v in options.items():
        if k == "atol":
            a = np.broadcast_to(np.asarray(v, dtype=np.float64), (n,))
            for i in range(n):
                o.atol[i] = a[i]
        elif k in ("maxsteps", "newt", "usejac"):
            setattr(o, k, int(v))
        else:
            setattr(o, k, float(v))
    ng_rows = 0 if t_grid is None else len(t_grid)
    tg = None if t_grid is None else np.ascontiguousarray(t_grid, dtype=np.float64)
    out = {
        "y": np.zeros((N, n)), "t": np.zeros(N), "sw": np.zeros((N, max(nsw, 1)), dtype=np.uint8),
        "status": np.zeros(N, dtype=np.int32), "stats": np.zeros((N, len(STATS)), dtype=np.int32),
        "dense": np.full((N, ng_rows, n), np.nan) if ng_rows else None,
        "ev_t": np.full((N, max(max_ev, 1)), np.nan), "ev_info": np.zeros((N, max(max_ev, 1)), dtype=np.int32),
    }
    rc = lib().ora_simulate(MODELS[model], C.byref(o), N, _ptr(y0), _ptr(p), _ptr(sw0), float(t0), float(tf),
                            ng_rows, _ptr(tg), _ptr(out["y"]), _ptr(out["t"]),
                            _ptr(out["sw"]) if nsw else None, _ptr(out["status"]), _ptr(out["stats"]),
                            _ptr(out["dense"]), int(max_ev), _ptr(out["ev_t"]) if max_ev else None,
                            _ptr(out["ev_info"]) if max_ev else None)
    assert rc == 0
    st = out.pop("stats")
    out["stats"] = {name: st[:, i].copy() for i, name in enumerate(STATS)}
    return out
