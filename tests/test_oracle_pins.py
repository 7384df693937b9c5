"""Independent pins of the two Fortran restatements (dopri5.f, rodas_decsol.f cannot be compiled here: no Fortran
compiler in the image, and the reference's sources do not travel to the GPU box).

* DOPRI5: SciPy ships its own build of Hairer's code (scipy.integrate.ode('dopri5'), beta set to Assimulo's 0.04).  The
  ensemble oracle -- which the golden vectors tie bit for bit to the reference's Dopri5 class over oracle/dopri5_c.c --
  must reproduce SciPy member by member over the WHOLE cfg1 sweep (4096 mu values, 0.1 .. 100) and on the 12-state
  gyroscope: the same number of right-hand-side evaluations, the same number of accepted steps, final states to 1e-12.
  (thirdparty/hairer/dopri5.f:356-556 DOPCOR, :558-619 HINIT.)
* RODAS: nothing independent executes the same algorithm here, and the reference holds ONE known answer for it
  (tests/solvers/test_rosenbrock.py:91, examples/rodasode_vanderpol.py:88).  What can be pinned without the Fortran is
  the METHOD: a Rosenbrock scheme with a wrong coefficient loses its order, so the restatement's global error against
  a tight Radau5ODE solution (the reference's own C code, bit-exact golden vectors) must fall like tol^(~3/4 .. 1) over
  four decades of tolerance on a stiff and a non-stiff problem, with step counts growing accordingly.
"""
import os
import re

import numpy as np
import pytest

import cases
import ens_oracle as O
from conftest import ROOT


def _scipy_dopri5(f, y0, tf, rtol, atol):
    from scipy.integrate import ode
    nf, nacc = [0], [0]

    def rhs(t, y):
        nf[0] += 1
        return f(t, y)

    def solout(t, y):
        nacc[0] += 1
    r = ode(rhs).set_integrator("dopri5", rtol=rtol, atol=atol, beta=0.04, nsteps=100000)
    r.set_solout(solout)
    r.set_initial_value(list(y0), 0.0)
    y = r.integrate(tf)
    assert r.successful()
    return np.array(y), nf[0], nacc[0] - 1          # solout is also called once for the initial point


def test_dopri5_oracle_equals_scipy_on_the_whole_cfg1_sweep():
    N = 4096
    y0, mu = cases.vdp_sweep(N)
    o = O.simulate("vdp", "Dopri5", y0, mu, tf=2.0, rtol=1e-6, atol=1e-6)
    worst = 0.0
    for i in range(N):
        m = float(mu[i, 0])
        ys, nf, nacc = _scipy_dopri5(lambda t, y: [y[1], m * ((1. - y[0] * y[0]) * y[1] - y[0])], y0[i], 2.0, 1e-6, 1e-6)
        assert nf == o["stats"]["nfcns"][i], (i, m, nf, o["stats"]["nfcns"][i])
        assert nacc == o["stats"]["nsteps"][i], (i, m, nacc, o["stats"]["nsteps"][i])
        worst = max(worst, float(np.max(np.abs(ys - o["y"][i]) / np.maximum(np.abs(o["y"][i]), 1.0))))
    assert worst <= 1e-12, worst
    assert int(o["stats"]["nsteps"].sum()) > 250000        # the sweep is not trivial: mu = 100 takes hundreds of steps


def test_dopri5_oracle_equals_scipy_on_the_gyroscope():
    N = 4096
    idx = np.arange(0, N, 128)
    y0, p = cases.gyro_sweep(N, idx)
    o = O.simulate("gyro", "Dopri5", y0, p, tf=0.1, rtol=1e-8, atol=1e-8)
    I1, I2, I3 = 1000., 5000., 6000.

    def f(t, u):
        om0, om1, om2 = u[0] / I1, u[1] / I2, u[2] / I3
        ud = [0.0] * 12
        ud[0] = om2 * u[1] + (-om1) * u[2]
        ud[1] = (-om2) * u[0] + om0 * u[2]
        ud[2] = om1 * u[0] + (-om0) * u[1]
        for c in range(3):
            q0, q1, q2 = u[3 + c], u[6 + c], u[9 + c]
            ud[3 + c] = om2 * q1 + (-om1) * q2
            ud[6 + c] = (-om2) * q0 + om0 * q2
            ud[9 + c] = om1 * q0 + (-om0) * q1
        return ud
    for k in range(len(idx)):
        ys, nf, nacc = _scipy_dopri5(f, y0[k], 0.1, 1e-8, 1e-8)
        assert nf == o["stats"]["nfcns"][k] and nacc == o["stats"]["nsteps"][k], (k, nf, nacc)
        assert np.max(np.abs(ys - o["y"][k]) / np.maximum(np.abs(o["y"][k]), 1.0)) <= 1e-12
    assert o["stats"]["nsteps"].min() >= 16


@pytest.mark.parametrize("model,tf,atol", [("vdp", 2.0, 1e-10), ("robertson", 40.0, [1e-12, 1e-16, 1e-12])])
def test_rodas_restatement_has_the_order_of_the_method(model, tf, atol):
    if model == "vdp":
        y0, p = cases.vdp_sweep(64)
        p = np.minimum(p * 10.0, 1000.0)                   # mu = 1 .. 1000: non-stiff to stiff
    else:
        y0, p = cases.robertson_sweep(64)
    ref = O.simulate(model, "Radau5ODE", y0, p, tf=tf, rtol=1e-12, atol=np.asarray(atol) * 1e-2, usejac=1)
    assert (ref["status"] == 0).all()
    errs, steps = [], []
    for rtol in (1e-4, 1e-6, 1e-8):
        a = np.asarray(atol) * (rtol / 1e-8)               # absolute tolerances shrink with the relative one
        o = O.simulate(model, "RodasODE", y0, p, tf=tf, rtol=rtol, atol=a, usejac=1)
        assert (o["status"] == 0).all()
        scale = np.maximum(np.abs(ref["y"]), np.asarray(a) / rtol)
        errs.append(float(np.median(np.max(np.abs(o["y"] - ref["y"]) / scale, axis=1))))
        steps.append(int(o["stats"]["nsteps"].sum()))
    # error proportional to tol^(3/4) at least (a 4th-order pair with local error control), and well below 10 tol
    assert errs[0] / errs[1] > 30 and errs[1] / errs[2] > 30, errs
    assert errs[0] < 1e-3 and errs[1] < 1e-5 and errs[2] < 1e-7, errs
    # steps grow like tol^(-1/4): 100x tighter -> about 3.2x more, never less than 1.8x or more than 6x
    for s0, s1 in zip(steps, steps[1:]):
        assert 1.8 < s1 / s0 < 6.0, steps


# ---------------------------------------------------------------------------------------------------------
# The coefficient tables, read from the Fortran text itself (build container only: /root/reference does not travel)
# ---------------------------------------------------------------------------------------------------------
REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")


def _fortran_assignments(text):
    """NAME = <fortran double literal or rational> pairs of a block of Fortran source"""
    out = {}
    for name, val in re.findall(r"\b([A-Z][A-Z0-9]*)\s*=\s*(-?\s*[0-9.]+D[+-]?[0-9]+(?:\s*/\s*-?[0-9.]+D[+-]?[0-9]+)?)", text):
        num = val.replace(" ", "").replace("D", "e")
        out[name] = eval(num)  # noqa: S307 -- "a/b" of two float literals
    return out


def _c_table(path, names):
    """NAME = <literal> pairs of a C / CUDA source (case-insensitive names, first occurrence)"""
    text = open(path).read()
    out = {}
    for n in names:
        m = re.search(r"\b%s\s*=\s*(-?[0-9.]+(?:e[+-]?[0-9]+)?(?:\s*/\s*-?[0-9.]+)?)" % n, text, flags=re.I)
        if m:
            out[n] = eval(m.group(1))  # noqa: S307
    return out


@needs_ref
def test_rodas_coefficients_equal_the_fortran_source():
    """ROCOE, METH = 1 (thirdparty/hairer/rodas_decsol.f:889-941): every coefficient of the C restatement, of the ensemble
    oracle and of the device table (in declaration order) is the Fortran literal."""
    src = open(os.path.join(REF, "thirdparty", "hairer", "rodas_decsol.f")).read()
    block = src[src.index("SUBROUTINE ROCOE"):]
    block = block[block.index(" 1      C2="):block.index(" 2      C2=")]
    f = _fortran_assignments(block)
    names = ["C2", "C3", "C4", "D1", "D2", "D3", "D4", "A21", "A31", "A32", "A41", "A42", "A43", "A51", "A52", "A53", "A54",
             "C21", "C31", "C32", "C41", "C42", "C43", "C51", "C52", "C53", "C54", "C61", "C62", "C63", "C64", "C65", "GAMMA",
             "D21", "D22", "D23", "D24", "D25", "D31", "D32", "D33", "D34", "D35"]
    assert set(names) <= set(f)
    for path in ("oracle/rodas_c.c", "oracle/ens_oracle.c"):
        c = _c_table(os.path.join(ROOT, path), names)
        assert {n: c.get(n) for n in names} == {n: f[n] for n in names}, path
    # the device table: `static __constant__ Consts K = { ... }` of namespace ro, members in the order of `names`
    cu = open(os.path.join(ROOT, "assimulo_b200", "csrc", "aens_radau5.cuh")).read()
    ro = cu[cu.index("namespace ro {"):cu.index("}  // namespace ro")]
    decl = re.findall(r"double ([^;]+);", ro[ro.index("struct Consts"):ro.index("};")])
    members = [m.strip().upper() for d in decl for m in d.split(",")]
    assert members == names
    init = ro[ro.index("Consts K = {") + len("Consts K = {"):]
    vals = [float(v) for v in re.findall(r"-?[0-9]+\.[0-9]+(?:e[+-][0-9]+)?", init[:init.index("};")])]
    assert vals == [f[n] for n in names]


@needs_ref
def test_dopri5_coefficients_equal_the_fortran_source():
    """CDOPRI (thirdparty/hairer/dopri5.f:646-692): the Dormand-Prince tableau of the restatement, the oracle and the
    device's constant table are the Fortran's rational literals, evaluated in double."""
    src = open(os.path.join(REF, "thirdparty", "hairer", "dopri5.f")).read()
    block = src[src.index("SUBROUTINE CDOPRI"):]
    f = _fortran_assignments(block)
    names = ["C2", "C3", "C4", "C5", "A21", "A31", "A32", "A41", "A42", "A43", "A51", "A52", "A53", "A54", "A61", "A62", "A63",
             "A64", "A65", "A71", "A73", "A74", "A75", "A76", "E1", "E3", "E4", "E5", "E6", "E7", "D1", "D3", "D4", "D5", "D6", "D7"]
    assert set(names) <= set(f) and abs(f["A71"] - 35.0 / 384.0) == 0.0 and f["E7"] == -1.0 / 40.0
    c = _c_table(os.path.join(ROOT, "oracle", "ens_oracle.c"), ["dp_" + n.lower() for n in names])
    assert [c["dp_" + n.lower()] for n in names] == [f[n] for n in names]
    cu = open(os.path.join(ROOT, "assimulo_b200", "csrc", "aens_device.cuh")).read()
    init = cu[cu.index("static __constant__ Tab T = {") + len("static __constant__ Tab T = {"):]
    init = init[:init.index("};")]
    vals = [eval(v) for v in re.findall(r"-?[0-9.]+(?:\s*/\s*[0-9.]+)?", init)]  # noqa: S307
    assert vals[:len(names)] == [f[n] for n in names] and vals[len(names)] == 1.0 / 1.01
