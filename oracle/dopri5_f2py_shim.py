"""Stand-in for the f2py module ``assimulo.lib.dopri5`` (TEST INFRASTRUCTURE ONLY).

The reference builds ``assimulo.lib.dopri5`` from Fortran with f2py (thirdparty/hairer/dopri5.pyf); no
Fortran compiler exists in this image.  ``oracle/build_ref.py`` installs this file as
``oracle/_ref/assimulo/lib/dopri5.py``.  It presents the same two callables, with the argument order and
return tuple of dopri5.pyf:31-61, on top of ``oracle/dopri5_c.c`` (our C restatement of dopri5.f):

    x, y, iwork, idid = dopri5(fcn, x, y, xend, rtol, atol, itol, solout, iout, work, iwork)
    v = contd5(ii, x, con, icomp)

Callback conventions (dopri5.pyf:6-26): ``k1 = fcn(x, y)`` and
``irtrn = solout(nr, xold, x, y, cont, icomp, irtrn)``.
"""
import ctypes as C
import os

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_lib = None
for _cand in (os.path.join(_here, "..", "..", "libdopri5_c.so"), os.path.join(_here, "libdopri5_c.so")):
    if os.path.exists(_cand):
        _lib = C.CDLL(os.path.abspath(_cand))
        break
if _lib is None:
    raise ImportError("libdopri5_c.so not found; run oracle/build_ref.py")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_FCN = C.CFUNCTYPE(None, C.c_int, C.c_double, _dp, _dp, C.c_void_p)
_SOLOUT = C.CFUNCTYPE(C.c_int, C.c_int, C.c_double, C.c_double, _dp, C.c_int, _dp, _ip, C.c_int, C.c_int,
                      C.c_void_p)
_lib.dopri5_c.restype = C.c_int
_lib.dopri5_c.argtypes = [C.c_int, _FCN, C.c_void_p, _dp, _dp, C.c_double, _dp, _dp, C.c_int,
                          _SOLOUT, C.c_void_p, C.c_int, _dp, C.c_int, _ip, C.c_int]
_lib.contd5_c.restype = C.c_double
_lib.contd5_c.argtypes = [C.c_int, C.c_double, _dp, _ip, C.c_int]


def dopri5(fcn, x, y, xend, rtol, atol, itol, solout, iout, work, iwork):
    y = np.array(y, dtype=np.float64, order="C")
    n = y.size
    rtol = np.ascontiguousarray(rtol, dtype=np.float64)
    atol = np.ascontiguousarray(atol, dtype=np.float64)
    work = np.array(work, dtype=np.float64, order="C")       # intent(in): private copy
    iwork = np.array(iwork, dtype=np.intc, order="C")        # intent(in,out): returned copy
    err = []

    def _fcn(n_, x_, y_, f_, _ext):
        if err:
            return
        try:
            yy = np.ctypeslib.as_array(y_, shape=(n_,)).copy()
            ff = np.asarray(fcn(x_, yy), dtype=np.float64)
            np.ctypeslib.as_array(f_, shape=(n_,))[:] = ff
        except BaseException as e:  # f2py would propagate; remember and stop at the next solout
            err.append(e)

    def _solout(nr, xold, x_, y_, n_, cont_, icomp_, nrd, irtrn, _ext):
        if err:
            return -1
        try:
            yy = np.ctypeslib.as_array(y_, shape=(n_,)).copy()
            cont = np.ctypeslib.as_array(cont_, shape=(5 * nrd,)).copy()
            icomp = np.ctypeslib.as_array(icomp_, shape=(nrd,)).copy()
            return int(solout(float(nr), xold, x_, yy, cont, icomp, irtrn))
        except BaseException as e:
            err.append(e)
            return -1

    xc = C.c_double(x)
    idid = _lib.dopri5_c(n, _FCN(_fcn), None, C.byref(xc), y.ctypes.data_as(_dp), float(xend),
                         rtol.ctypes.data_as(_dp), atol.ctypes.data_as(_dp), int(itol),
                         _SOLOUT(_solout), None, int(iout), work.ctypes.data_as(_dp), work.size,
                         iwork.ctypes.data_as(_ip), iwork.size)
    if err:
        raise err[0]
    return xc.value, y, iwork, idid


def contd5(ii, x, con, icomp):
    con = np.ascontiguousarray(con, dtype=np.float64)
    icomp = np.ascontiguousarray(icomp, dtype=np.intc)
    return _lib.contd5_c(int(ii), float(x), con.ctypes.data_as(_dp), icomp.ctypes.data_as(_ip), icomp.size)
