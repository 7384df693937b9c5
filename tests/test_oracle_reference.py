"""The REAL reference (oracle/_ref, built from /root/reference by oracle/build_ref.py) against its own known
answers and against the committed golden vectors.  Skipped where oracle/_ref is absent.  Dopri5 here is the
reference's runge_kutta.py::Dopri5 class over oracle/dopri5_c.c (the Fortran core cannot be built: no compiler)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden, ref_available

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import io
    import contextlib
    with contextlib.redirect_stderr(io.StringIO()):
        import assimulo.solvers as S
        from assimulo.problem import Explicit_Problem
    import ref_models
    return S, Explicit_Problem, ref_models


def _run(S, name, prob, tf, ncp=0, **opts):
    sim = getattr(S, name)(prob)
    sim.verbosity = 50
    for k, v in opts.items():
        setattr(sim, k, v)
    t, y = sim.simulate(tf, ncp)
    return sim, t, y


def test_reference_known_answers(ref):
    S, EP, M = ref
    f = lambda t, y: -y  # noqa: E731
    assert abs(_run(S, "Dopri5", EP(f, [4.0]), 5.0)[2][-1][0] - 0.02695199) < 1e-5       # dopri5_basic.py:58
    assert abs(_run(S, "RungeKutta34", EP(f, [4.0]), 5.0, inith=0.1)[2][-1][0] - 0.02695199) < 1e-5
    assert abs(_run(S, "RungeKutta4", EP(f, [4.0]), 5.0, h=0.01)[2][-1][0] - 0.02695179) < 1e-6
    sim, t, y = _run(S, "Dopri5", EP(lambda t, y: 1.0, 1), 1.0)                            # test_rungekutta.py:40-47
    assert sim.t_sol[-1] == pytest.approx(1.0) and sim.y_sol[-1][0] == pytest.approx(2.0)
    sim, t, y = _run(S, "Dopri5", M.extended(), 10.0)                                      # dopri5_with_disc.py:157-159
    np.testing.assert_allclose(y[-1], [8.0, 3.0, 2.0], atol=1e-6)
    sim, t, y = _run(S, "Dopri5", M.linear_ev(), 2.0, 33)                                  # test_solvers.py:52-57
    assert y[-1][0] == pytest.approx(0.135, abs=1e-3)


def test_dopri5_restatement_against_scipy(ref):
    """SciPy's independent C translation of the same Hairer code: VdP mu=1 -> 12 accepted steps, 86 f-evals,
    y(2) = [-0.35758069, -2.50277519]."""
    S, EP, M = ref
    sim, t, y = _run(S, "Dopri5", M.vdp(1.0), 2.0, rtol=1e-6, atol=1e-6)
    assert sim.statistics["nsteps"] == 12 and sim.statistics["nfcns"] == 86
    np.testing.assert_allclose(y[-1], [-0.35758069, -2.50277519], atol=1e-8)
    from scipy.integrate import ode
    nf = [0]

    def f(t, y):
        nf[0] += 1
        return [y[1], 1.0 * ((1. - y[0] * y[0]) * y[1] - y[0])]
    r = ode(f).set_integrator("dopri5", rtol=1e-6, atol=1e-6, beta=0.04)
    r.set_initial_value([2.0, -0.6], 0.0)
    ys = r.integrate(2.0)
    np.testing.assert_allclose(ys, y[-1], rtol=0, atol=1e-12)
    assert nf[0] == sim.statistics["nfcns"]


def test_golden_vectors_are_reproducible(ref):
    """Regenerate a few golden members from the live reference and compare with the committed files."""
    S, EP, M = ref
    g = golden("cfg1_dopri5_vdp")
    for k in (0, 17, 63):
        sim, t, y = _run(S, "Dopri5", M.vdp(float(g["mu"][k])), 2.0, rtol=1e-6, atol=1e-6)
        assert np.array_equal(y[-1], g["y"][k]) and sim.statistics["nsteps"] == g["nsteps"][k]
    g = golden("cfg4_radau5_robertson_jac")
    sim, t, y = _run(S, "Radau5ODE", M.robertson(*g["p"][3]), 40.0, rtol=1e-6, atol=[1e-8, 1e-14, 1e-6])
    assert np.array_equal(y[-1], g["y"][3]) and sim.statistics["nsteps"] == g["nsteps"][3]
    g = golden("cfg3_ball_dopri5")
    prob = M.ball(float(g["h0"][5]))
    sim, t, y = _run(S, "Dopri5", prob, 3.0, rtol=1e-6, atol=1e-6)
    assert np.array_equal([e[0] for e in prob.events], g["ev_t"][5][: len(prob.events)])
