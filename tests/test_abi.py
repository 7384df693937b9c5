"""The C-ABI shared library: loads without a GPU, exports every symbol include/assimulo_cuda.h declares, reports
the reference's default options, fails loudly (no CPU fallback) when no device exists, and compiles user models
with NVRTC for sm_100a (compile only -- no compute calls without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, _cuda_available

LIB = os.path.join(ROOT, "assimulo_b200", "lib", "libassimulo_cuda.so")
HDR = os.path.join(ROOT, "include", "assimulo_cuda.h")


def declared_functions():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(aens_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = C.CDLL(LIB)
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libassimulo_cuda.so does not export %s" % n
    lib.aens_version.restype = C.c_char_p
    assert lib.aens_version().startswith(b"assimulo_cuda")


def test_kernels_are_sm_100a_cubins():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", LIB], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out
    assert "sm_90" not in out and "sm_80" not in out


def test_default_opts_match_the_reference():
    from assimulo_b200.gpu import _core
    d = _core.default_opts("Dopri5")          # runge_kutta.py:56-64
    assert (d["safe"], d["fac1"], d["fac2"], d["beta"], d["maxsteps"], d["rtol"], d["inith"]) == \
        (0.9, 0.2, 10.0, 0.04, 100000, 1e-6, 0.0)
    d = _core.default_opts("RungeKutta34")    # runge_kutta.py:449-452
    assert (d["inith"], d["maxsteps"], d["rtol"], d["atol"][0]) == (0.01, 10000, 1e-6, 1e-6)
    d = _core.default_opts("Radau5ODE")       # radau5.py:99-113
    assert (d["inith"], d["newt"], d["thet"], d["quot1"], d["quot2"], d["fac1"], d["fac2"], d["safe"]) == \
        (0.01, 7, 1e-3, 1.0, 1.2, 0.2, 8.0, 0.9)
    assert _core.default_opts("RungeKutta4")["h"] == 0.01 and _core.default_opts("ExplicitEuler")["h"] == 0.01
    assert _core.default_opts("RungeKutta4")["arith"] == 0    # strict: fixed-step parity bar is 1e-13


@pytest.mark.skipif(_cuda_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_device():
    from assimulo_b200.gpu import _core, Batched_Explicit_Problem, Dopri5
    with pytest.raises(_core.AensError) as e:
        _core.Handle(0)
    assert e.value.code == -2
    sim = Dopri5(Batched_Explicit_Problem(model="vdp", y0=[2.0, -0.6], p0=[[1.0], [2.0]]))
    with pytest.raises(_core.AensError):
        sim.simulate(2.0)
    lib = C.CDLL(LIB)
    h = C.c_void_p()
    assert lib.aens_create(0, C.byref(h)) == -2 and not h.value


USER_SRC = r"""
// damped oscillator with an event at y0 = 0
__device__ void aens_rhs(double t, const double* y, const unsigned char* sw, const double* p, double* yd) {
    yd[0] = y[1];
    yd[1] = -p[0] * y[0] - p[1] * y[1];
}
__device__ void aens_jac(double t, const double* y, const unsigned char* sw, const double* p, double* J) {
    J[0] = 0.; J[1] = -p[0]; J[2] = 1.; J[3] = -p[1];
}
__device__ void aens_events(double t, const double* y, const unsigned char* sw, const double* p, double* g) {
    g[0] = y[0];
}
__device__ int aens_handle_event(double t, double* y, unsigned char* sw, const int* info, const double* p) {
    sw[0] = !sw[0];
    return 0;
}
"""


def test_nvrtc_compiles_user_models_for_sm_100a():
    from assimulo_b200.gpu import _core
    for strict in (False, True):
        size = _core.nvrtc_check(USER_SRC, 2, 2, 1, 1, True, strict)
        assert size > 10000
    # flag AENS_MODEL_RHS_H: the fused hk = h * f hook must exist when announced, and is compiled in when it does
    fused = USER_SRC + """
__device__ void aens_rhs_h(double h, double t, const double* y, const unsigned char* sw, const double* p, double* hk) {
    double yd[2]; aens_rhs(t, y, sw, p, yd); hk[0] = h * yd[0]; hk[1] = h * yd[1];
}
"""
    assert _core.nvrtc_check(fused, 2, 2, 1, 1, 1 | _core.MODEL_RHS_H, False) > 10000
    with pytest.raises(_core.AensError, match="aens_rhs_h"):
        _core.nvrtc_check(USER_SRC, 2, 2, 1, 1, 1 | _core.MODEL_RHS_H, False)
    with pytest.raises(_core.AensError) as e:
        _core.nvrtc_check("__device__ void aens_rhs(double t) { this is not C }", 2, 1)
    assert e.value.code == -3 and "user_model.cu" in str(e.value)


def test_opts_struct_layout_matches_header():
    """ctypes mirror of aens_opts_t has the size the header implies (6 ints + 28 doubles)."""
    class Opts(C.Structure):
        _fields_ = [("i", C.c_int * 6), ("rtol", C.c_double), ("atol", C.c_double * 16), ("rest", C.c_double * 11)]
    lib = C.CDLL(LIB)
    o = Opts()
    assert lib.aens_default_opts(3, C.byref(o)) == 0
    assert o.i[0] == 3 and o.i[2] == 100000 and o.rtol == 1e-6 and o.atol[15] == 1e-6
    assert list(o.rest)[:6] == [0.01, 0.0, 0.0, 0.9, 0.2, 10.0]
    assert lib.aens_default_opts(99, C.byref(o)) == -1


def test_every_entry_point_is_documented_in_integration_md():
    """INTEGRATION.md section 1 maps each C entry point to the reference interface it replaces: none may be missing."""
    import re
    hdr = open(os.path.join(ROOT, "include", "assimulo_cuda.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    names = set(re.findall(r"^[a-z][^\n(]*?\**(aens_[a-z0-9_]+)\(", hdr, flags=re.M))
    assert len(names) > 30
    missing = sorted(n for n in names if n not in doc)
    assert not missing, missing
