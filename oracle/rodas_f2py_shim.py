"""Stand-in for the f2py module ``assimulo.lib.rodas`` (TEST INFRASTRUCTURE ONLY).

The reference builds ``assimulo.lib.rodas`` from Fortran with f2py (thirdparty/hairer/rodas_decsol.pyf); no
Fortran compiler exists in this image.  ``oracle/build_ref.py`` installs this file as
``oracle/_ref/assimulo/lib/rodas.py``.  It presents the two callables RodasODE uses
(src/solvers/rosenbrock.py:307-311, 407-408), with the argument order and return tuple of rodas_decsol.pyf, on
top of ``oracle/rodas_c.c`` (our C restatement of the RODAS part of rodas_decsol.f):

    x, y, h, iwork, idid = rodas(fcn, ifcn, x, y, xend, h, rtol, atol, itol, jac, ijac, mljac, mujac, dfx, idfx,
                                 mas, imas, mlmas, mumas, solout, iout, work, iwork)
    v = contro(i, x, cont)

Callback conventions: ``dy = fcn(x, y)``, ``fjac = jac(x, y)`` and ``irtrn = solout(nr, xold, x, y, cont, lrc, irtrn)``.
"""
import ctypes as C
import os

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_lib = None
for _cand in (os.path.join(_here, "..", "..", "librodas_c.so"), os.path.join(_here, "librodas_c.so")):
    if os.path.exists(_cand):
        _lib = C.CDLL(os.path.abspath(_cand))
        break
if _lib is None:
    raise ImportError("librodas_c.so not found; run oracle/build_ref.py")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_FCN = C.CFUNCTYPE(None, C.c_int, C.c_double, _dp, _dp, C.c_void_p)
_JAC = C.CFUNCTYPE(None, C.c_int, C.c_double, _dp, _dp, C.c_int, C.c_void_p)
_SOLOUT = C.CFUNCTYPE(C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_void_p)
_lib.rodas_c.restype = C.c_int
_lib.rodas_c.argtypes = [C.c_int, _FCN, C.c_void_p, C.c_int, _dp, _dp, C.c_double, _dp, _dp, _dp, C.c_int,
                         _JAC, C.c_void_p, C.c_int, _SOLOUT, C.c_void_p, C.c_int, _dp, C.c_int, _ip, C.c_int]
_lib.contro_c.restype = C.c_double
_lib.contro_c.argtypes = [C.c_int, C.c_double, _dp, C.c_int]


def rodas(fcn, ifcn, x, y, xend, h, rtol, atol, itol, jac, ijac, mljac, mujac, dfx, idfx, mas, imas, mlmas, mumas,
          solout, iout, work, iwork):
    if imas != 0 or idfx != 0 or mljac < len(y):
        raise NotImplementedError("rodas_c.c restates the explicit, full-Jacobian, IDFX = 0 case Assimulo uses")
    y = np.array(y, dtype=np.float64, order="C")
    n = y.size
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    work = np.array(work, dtype=np.float64, order="C")
    iwork = np.array(iwork, dtype=np.intc, order="C")
    err = []

    def _fcn(n_, x_, y_, f_, _ext):
        if err:
            return
        try:
            yy = np.ctypeslib.as_array(y_, shape=(n_,)).copy()
            np.ctypeslib.as_array(f_, shape=(n_,))[:] = np.asarray(fcn(x_, yy), dtype=np.float64)
        except BaseException as e:
            err.append(e)

    def _jac(n_, x_, y_, j_, ld, _ext):
        if err:
            return
        try:
            yy = np.ctypeslib.as_array(y_, shape=(n_,)).copy()
            J = np.asarray(jac(x_, yy), dtype=np.float64)           # J[i, j] = d f_i / d y_j
            np.ctypeslib.as_array(j_, shape=(n_ * ld,))[:] = J.T.reshape(-1)  # Fortran (column-major) storage
        except BaseException as e:
            err.append(e)

    def _solout(nr, xold, x_, y_, cont_, lrc, n_, irtrn, _ext):
        if err:
            return -1
        try:
            yy = np.ctypeslib.as_array(y_, shape=(n_,)).copy()
            cont = np.ctypeslib.as_array(cont_, shape=(lrc,)).copy()
            return int(solout(float(nr), xold, x_, yy, cont, lrc, irtrn))
        except BaseException as e:
            err.append(e)
            return -1

    xc, hc = C.c_double(x), C.c_double(h)
    idid = _lib.rodas_c(n, _FCN(_fcn), None, int(ifcn), C.byref(xc), y.ctypes.data_as(_dp), float(xend), C.byref(hc),
                        rtol.ctypes.data_as(_dp), atol.ctypes.data_as(_dp), int(itol), _JAC(_jac), None, int(ijac),
                        _SOLOUT(_solout), None, int(iout), work.ctypes.data_as(_dp), work.size,
                        iwork.ctypes.data_as(_ip), iwork.size)
    if err:
        raise err[0]
    return xc.value, y, hc.value, iwork, idid


def contro(i, x, cont):
    cont = np.ascontiguousarray(cont, dtype=np.float64)
    return _lib.contro_c(int(i), float(x), cont.ctypes.data_as(_dp), cont.size)
