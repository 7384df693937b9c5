"""Parity tests proper (-m gpu): the CUDA path, called through the Cython module over the C ABI, against the CPU
oracle on the same inputs, against the committed golden vectors of the reference, and -- at BASELINE.json's full
sizes -- through size-independent properties.

Bars (BASELINE.json north_star):
  * fixed-step RungeKutta4 / ExplicitEuler: <= 1e-13 relative in fp64 (default arithmetic = strict: bit-exact);
  * adaptive Dopri5 / RungeKutta34 / Radau5ODE: final states within 10 x rtol of the reference, accepted-step
    counts within +-2 %, event times within the event tolerance (1e-13 relative of the locator, we allow 1e-9);
    the strict build reproduces the oracle's step counts exactly.
"""
import numpy as np
import pytest

import cases
import ens_oracle as O
from conftest import golden, rel_err

pytestmark = pytest.mark.gpu

RTOL = 1e-6


def gpu(model, solver, y0, p=None, sw0=None, tf=1.0, arith=None, t_grid=None, event_slots=24, order=None,
        refill=32, reverse=0, **opts):
    from assimulo_b200.gpu import _core
    h = _core.Handle(0)
    h.set_model_builtin(model)
    n = h.n
    y0 = np.ascontiguousarray(np.atleast_2d(np.asarray(y0, dtype=float)))
    N = y0.shape[0]
    if h.n_p:
        p = np.ascontiguousarray(np.broadcast_to(np.asarray(p, dtype=float).reshape(-1, h.n_p), (N, h.n_p)))
    if h.n_sw:
        sw0 = np.ascontiguousarray(np.broadcast_to(np.asarray(sw0, dtype=np.uint8).reshape(-1, h.n_sw), (N, h.n_sw)))
    h.set_ensemble(y0, p if h.n_p else None, sw0 if h.n_sw else None, 0.0)
    d = {"solver": _core.SOLVER_IDS[solver]}
    if arith is not None:
        d["arith"] = {"strict": 0, "fast": 1}[arith]
    d.update(opts)
    h.set_opts(d)
    h.set_output(t_grid, event_slots)
    h.set_schedule(refill, order)
    h.set_fetch_reverse(reverse)
    h.simulate(tf)
    t, y, sw, st = h.get_final()
    out = {"t": t, "y": y, "sw": sw, "status": st, "stats": {k: h.get_stats(k) for k in _core.STAT_NAMES},
           "sums": h.get_stats_sum(), "n": n}
    if t_grid is not None:
        out["dense"] = np.transpose(h.get_dense(), (1, 0, 2))  # -> [N, rows, n] like the oracle
    if h.n_g or (h.has_jac & 2):
        out["nev"], out["ev_t"], out["ev_info"] = h.get_events()
    h.close()
    return out


def steps_within(g, o, frac=0.02):
    a, b = int(g["stats"]["nsteps"].astype(np.int64).sum()), int(o["stats"]["nsteps"].astype(np.int64).sum())
    return abs(a - b) <= frac * b


# ---------------------------------------------------------------------------------------------------------
# fixed step: <= 1e-13 (strict default is bit-exact)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("solver", ["RungeKutta4", "ExplicitEuler"])
def test_fixed_step_vdp_1e13(solver):
    N = 4096
    y0, p = cases.vdp_sweep(N)
    o = O.simulate("vdp", solver, y0, p, tf=2.0, h=1e-3)
    g = gpu("vdp", solver, y0, p, tf=2.0, h=1e-3)                  # default arithmetic of the fixed-step solvers
    assert (g["status"] == 0).all()
    assert rel_err(g["y"], o["y"]) <= 1e-13
    assert np.array_equal(g["y"], o["y"])                           # in fact identical doubles
    assert np.array_equal(g["t"], o["t"])
    assert (g["stats"]["nsteps"] == 2001).all() and np.array_equal(g["stats"]["nfcns"], o["stats"]["nfcns"])
    # against the reference's own vectors
    gold = golden("cfg2_rk4_vdp" if solver == "RungeKutta4" else "cfg2_euler_vdp")
    y0g, pg = cases.vdp_sweep(int(gold["N"]), gold["idx"])
    gg = gpu("vdp", solver, y0g, pg, tf=2.0, h=1e-3)
    assert rel_err(gg["y"], gold["y"]) <= 1e-13
    assert np.array_equal(gg["stats"]["nsteps"], gold["nsteps"])


@pytest.mark.parametrize("solver", ["RungeKutta4", "ExplicitEuler"])
def test_fixed_step_fast_arithmetic_is_close_but_not_the_default(solver):
    y0, p = cases.vdp_sweep(1024)
    o = O.simulate("vdp", solver, y0, p, tf=2.0, h=1e-3)
    g = gpu("vdp", solver, y0, p, tf=2.0, h=1e-3, arith="fast")
    assert rel_err(g["y"], o["y"]) <= 1e-10                         # FMA contraction: ~1e-12, above the 1e-13 bar


def test_known_answers_linear():
    p = [[-1.0, 0.0, 0.0]]
    assert abs(gpu("linear", "Dopri5", [[4.0]], p, tf=5.0)["y"][0, 0] - 0.02695199) < 1e-5
    assert abs(gpu("linear", "RungeKutta34", [[4.0]], p, tf=5.0, inith=0.1)["y"][0, 0] - 0.02695199) < 1e-5
    assert abs(gpu("linear", "RungeKutta4", [[4.0]], p, tf=5.0, h=0.01)["y"][0, 0] - 0.02695179) < 1e-6
    assert abs(gpu("linear", "ExplicitEuler", [[4.0]], p, tf=5.0, h=0.01)["y"][0, 0] - 0.02628193) < 1e-6
    for s in ("Dopri5", "RungeKutta34", "RungeKutta4", "ExplicitEuler", "Radau5ODE"):
        out = gpu("linear", s, [[1.0]], [[0.0, 1.0, 0.0]], tf=1.0)
        assert out["y"][0, 0] == pytest.approx(2.0, abs=1e-9) and out["t"][0] == pytest.approx(1.0, abs=1e-12)


# ---------------------------------------------------------------------------------------------------------
# adaptive solvers on the VdP sweep (cfg 1 / cfg 2)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("solver,opts", [("Dopri5", dict(rtol=RTOL, atol=1e-6)),
                                          ("RungeKutta34", dict(rtol=RTOL, atol=1e-6, inith=0.01)),
                                          ("Radau5ODE", dict(rtol=RTOL, atol=1e-6, usejac=1)),
                                          ("Radau5ODE", dict(rtol=RTOL, atol=1e-6, usejac=0))])
@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_adaptive_vdp_sweep(solver, opts, arith):
    N = 4096
    y0, p = cases.vdp_sweep(N)
    o = O.simulate("vdp", solver, y0, p, tf=2.0, **opts)
    g = gpu("vdp", solver, y0, p, tf=2.0, arith=arith, **opts)
    assert (g["status"] == 0).all()
    assert rel_err(g["y"], o["y"]) <= 10 * RTOL
    assert steps_within(g, o)
    per_member = np.abs(g["stats"]["nsteps"].astype(int) - o["stats"]["nsteps"]) <= np.maximum(2, 0.02 * o["stats"]["nsteps"])
    assert per_member.all()
    if arith == "strict":
        assert rel_err(g["y"], o["y"]) <= 1e-10
        for k in ("nsteps", "nfcns", "nerrfails", "njacs", "nlus"):
            assert np.array_equal(g["stats"][k], o["stats"][k]), k
    assert np.array_equal(g["t"], o["t"])


def test_cfg1_dopri5_against_reference_golden():
    gold = golden("cfg1_dopri5_vdp")
    y0, p = cases.vdp_sweep(int(gold["N"]), gold["idx"])
    for arith in ("strict", "fast"):
        g = gpu("vdp", "Dopri5", y0, p, tf=2.0, arith=arith, rtol=1e-6, atol=1e-6)
        assert rel_err(g["y"], gold["y"]) <= 10 * RTOL
        assert np.abs(g["stats"]["nsteps"] - gold["nsteps"]).max() <= (0 if arith == "strict" else 2)
        assert abs(int(g["stats"]["nsteps"].sum()) - int(gold["nsteps"].sum())) <= 0.02 * gold["nsteps"].sum()
    gold = golden("cfg2_rk34_vdp")
    y0, p = cases.vdp_sweep(int(gold["N"]), gold["idx"])
    g = gpu("vdp", "RungeKutta34", y0, p, tf=2.0, rtol=1e-6, atol=1e-6, inith=0.01)
    assert rel_err(g["y"], gold["y"]) <= 10 * RTOL
    assert abs(int(g["stats"]["nsteps"].sum()) - int(gold["nsteps"].sum())) <= 0.02 * gold["nsteps"].sum()


# ---------------------------------------------------------------------------------------------------------
# state events on device (cfg 3) + the reference tests' event problems
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("solver,opts", [("Dopri5", dict(rtol=1e-6, atol=1e-6)),
                                          ("RungeKutta34", dict(rtol=1e-6, atol=1e-6)),
                                          ("ExplicitEuler", dict(h=1e-3)),
                                          ("Radau5ODE", dict(rtol=1e-6, atol=1e-6))])
@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_ball_events(solver, opts, arith):
    N = 2048
    y0, p, sw0 = cases.ball_sweep(N)
    o = O.simulate("ball", solver, y0, p, sw0, tf=3.0, max_ev=24, **opts)
    g = gpu("ball", solver, y0, p, sw0, tf=3.0, arith=arith, **opts)
    assert (g["status"] == 0).all()
    assert np.array_equal(g["nev"], o["stats"]["nstateevents"])            # every event found, none invented
    m = np.isfinite(o["ev_t"])
    assert np.max(np.abs(g["ev_t"][m] - o["ev_t"][m])) <= 1e-9              # event times
    assert np.array_equal(g["ev_info"][m], o["ev_info"][m])                 # which indicator, which direction
    assert np.array_equal(g["sw"], o["sw"])
    assert rel_err(g["y"], o["y"]) <= 10 * RTOL
    if solver != "ExplicitEuler":
        assert steps_within(g, o)
    if arith == "strict":
        for k in ("nsteps", "nfcns"):
            assert np.array_equal(g["stats"][k], o["stats"][k]), k
        # The locator itself is reproduced operation for operation: with the fixed-step solver, whose step sizes involve
        # no library function, every member makes exactly the oracle's probes.  The adaptive solvers size their steps
        # with pow() (HINIT's fifth root, the controller's err^expo1 / facold^beta), where the device library and glibc
        # differ by an ulp now and then; such a step ends an ulp away, the secant iterates shift with it and an
        # iteration more or less meets the 1e-13 bracket -- a few probes per thousand, never an event.
        if solver == "ExplicitEuler":
            assert np.array_equal(g["stats"]["nstatefcns"], o["stats"]["nstatefcns"])
        assert abs(int(g["stats"]["nstatefcns"].sum()) - int(o["stats"]["nstatefcns"].sum())) <= 0.02 * o["stats"]["nstatefcns"].sum()
        assert np.max(np.abs(g["ev_t"][m] - o["ev_t"][m])) <= 2e-12
    # first impact against the closed form t = sqrt(2 h0 / g)
    if solver != "ExplicitEuler":
        assert np.max(np.abs(g["ev_t"][:, 0] - np.sqrt(2 * y0[:, 0] / 9.81))) < 1e-5


@pytest.mark.parametrize("solver", ["Dopri5", "RungeKutta34", "ExplicitEuler", "Radau5ODE"])
def test_ball_against_reference_golden(solver):
    gold = golden("cfg3_ball_" + solver.lower())
    y0, p, sw0 = cases.ball_sweep(int(gold["N"]), gold["idx"])
    opts = dict(h=1e-3) if solver == "ExplicitEuler" else dict(rtol=1e-6, atol=1e-6)
    g = gpu("ball", solver, y0, p, sw0, tf=3.0, **opts)
    assert np.array_equal(g["nev"], gold["nev"])
    m = np.isfinite(gold["ev_t"])
    assert np.max(np.abs(g["ev_t"][m] - gold["ev_t"][m])) <= 1e-9
    assert rel_err(g["y"], gold["y"]) <= 10 * RTOL


@pytest.mark.parametrize("solver", ["Dopri5", "RungeKutta34", "ExplicitEuler", "Radau5ODE"])
def test_extended_problem_event_iteration(solver):
    """tests/conftest.py Extended_Problem: expected end state [8,3,2], one state event, event iteration inside
    handle_event on the device."""
    g = gpu("extended", solver, [[0.0, -1.0, 0.0]], None, [[0, 1, 1]], tf=10.0)
    np.testing.assert_allclose(g["y"][0], [8.0, 3.0, 2.0], atol=1e-6)
    assert g["nev"][0] == 1 and g["sw"][0].tolist() == [1, 0, 0]
    gold = golden("extended_" + solver.lower())
    assert rel_err(g["y"], gold["y"]) <= 1e-9
    assert abs(g["ev_t"][0, 0] - gold["ev_t"][0, 0]) <= 1e-9
    g = gpu("linear_ev", solver, [[1.0]], [[-1.0, 0.0, 0.5]], [[1]], tf=2.0)
    assert g["y"][0, 0] == pytest.approx(0.135, abs=2e-3) and g["sw"][0, 0] == 0
    assert abs(g["ev_t"][0, 0] - np.log(2.0)) < (5e-3 if solver == "ExplicitEuler" else 1e-5)
    gold = golden("linear_ev_" + solver.lower())
    assert abs(g["ev_t"][0, 0] - gold["ev_t"][0, 0]) <= 1e-9 and rel_err(g["y"], gold["y"]) <= 1e-6


@pytest.mark.parametrize("solver,opts", [("Dopri5", dict(rtol=1e-6, atol=1e-6)),
                                          ("Radau5ODE", dict(rtol=1e-6, atol=1e-6)), ("ExplicitEuler", dict(h=1e-3)),
                                          ("RungeKutta34", dict(rtol=1e-6, atol=1e-6))])
def test_dense_output_with_events(solver, opts):
    grid = np.linspace(0.0, 3.0, 31)[1:]
    N = 512
    y0, p, sw0 = cases.ball_sweep(N)
    o = O.simulate("ball", solver, y0, p, sw0, tf=3.0, t_grid=grid, max_ev=24, **opts)
    g = gpu("ball", solver, y0, p, sw0, tf=3.0, t_grid=grid, arith="strict", **opts)
    assert np.array_equal(np.isnan(g["dense"]), np.isnan(o["dense"]))
    assert np.nanmax(np.abs(g["dense"] - o["dense"])) <= 1e-9
    if solver != "RungeKutta34":  # reference vectors exist for the clean report_continuously=True modes only
        gold = golden("ball_dense_" + solver.lower())
        gg = gpu("ball", solver, [[2.0, 0.0]], [[-9.81, 0.88]], [[1, 0]], tf=3.0, t_grid=gold["grid"], **opts)
        have = np.isfinite(gold["dense"][:, 0])
        assert np.max(np.abs(gg["dense"][0][have] - gold["dense"][have])) <= 1e-5


# ---------------------------------------------------------------------------------------------------------
# Radau5ODE Robertson (cfg 4), gyro with dense output (cfg 5)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("usejac", [1, 0])
@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_radau5_robertson(usejac, arith):
    N = 4096
    y0, p = cases.robertson_sweep(1 << 18, np.arange(0, 1 << 18, 64))
    opts = dict(rtol=1e-6, atol=cases.ROBERTSON_ATOL, usejac=usejac)
    o = O.simulate("robertson", "Radau5ODE", y0, p, tf=40.0, **opts)
    g = gpu("robertson", "Radau5ODE", y0, p, tf=40.0, arith=arith, **opts)
    assert g["y"].shape == (N, 3) and (g["status"] == 0).all()
    assert rel_err(g["y"], o["y"]) <= 10 * RTOL
    assert steps_within(g, o)
    if arith == "strict":
        for k in ("nsteps", "nfcns", "nerrfails", "njacs", "nlus"):
            assert np.array_equal(g["stats"][k], o["stats"][k]), k
    assert np.max(np.abs(g["y"].sum(axis=1) - 1.0)) < 1e-6        # mass conservation of the kinetics
    gold = golden("cfg4_radau5_robertson_" + ("jac" if usejac else "fd"))
    gg = gpu("robertson", "Radau5ODE", np.tile([1.0, 0.0, 0.0], (len(gold["p"]), 1)), gold["p"], tf=40.0,
             arith=arith, **opts)
    assert rel_err(gg["y"], gold["y"]) <= 10 * RTOL
    assert np.abs(gg["stats"]["nsteps"] - gold["nsteps"]).max() <= (0 if arith == "strict" else 2)


def test_radau5_stiff_vdp_reference_case():
    """tests/solvers/test_radau5.py:438-446: mu=1e6, y1(2)=1.7061680350 +-1e-4, nsteps < 300."""
    gold = golden("radau5_vdp_mu1e6")
    g = gpu("vdp", "Radau5ODE", [[2.0, -0.6]], [[1e6]], tf=2.0, rtol=1e-4, atol=1e-4, inith=1e-4, usejac=1,
            arith="strict")
    assert abs(g["y"][0, 0] - 1.7061680350) < 1e-4
    assert g["stats"]["nsteps"][0] == gold["nsteps"][0] < 300


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_gyro_dense_output(arith):
    N = 4096                                                          # every 1024th member of config 5
    y0, p = cases.gyro_sweep(1 << 22, np.arange(0, 1 << 22, 1024))
    grid = np.linspace(0.0, 0.1, 101)[1:]
    o = O.simulate("gyro", "Dopri5", y0, p, tf=0.1, t_grid=grid, rtol=1e-8, atol=1e-8)
    g = gpu("gyro", "Dopri5", y0, p, tf=0.1, t_grid=grid, arith=arith, rtol=1e-8, atol=1e-8)
    assert g["dense"].shape == (N, 100, 12) and not np.isnan(g["dense"]).any()
    scale = np.maximum(np.abs(o["dense"]), 1.0)
    assert np.max(np.abs(g["dense"] - o["dense"]) / scale) <= 10 * 1e-8
    assert rel_err(g["y"], o["y"]) <= 10 * 1e-8
    assert steps_within(g, o)
    if arith == "strict":
        assert np.array_equal(g["stats"]["nsteps"], o["stats"]["nsteps"])
    Q = g["y"][:, 3:].reshape(N, 3, 3)                               # rotation matrix stays orthogonal
    assert np.max(np.abs(np.einsum("nij,nkj->nik", Q, Q) - np.eye(3))) < 1e-6
    gold = golden("cfg5_gyro_dopri5")
    gg = gpu("gyro", "Dopri5", gold["y0"], np.tile([1000.0, 5000.0, 6000.0], (3, 1)), tf=0.1,
             t_grid=gold["t_rows"][0][1:], arith=arith, rtol=1e-8, atol=1e-8)
    sc = np.maximum(np.abs(gold["y_rows"][:, 1:, :]), 1.0)
    assert np.max(np.abs(gg["dense"] - gold["y_rows"][:, 1:, :]) / sc) <= 10 * 1e-8


# ---------------------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------------------
def test_ragged_sizes_and_schedules():
    """N not a multiple of the warp/block size, N = 1, every refill threshold and a permuted fetch order give the
    same per-member results (members are independent)."""
    for N in (1, 31, 33, 129, 1000):
        y0, p = cases.vdp_sweep(4096, np.arange(N) * (4096 // max(N, 1)))
        o = O.simulate("vdp", "Dopri5", y0, p, tf=2.0, rtol=1e-6, atol=1e-6)
        base = gpu("vdp", "Dopri5", y0, p, tf=2.0, arith="strict", rtol=1e-6, atol=1e-6)
        assert np.array_equal(base["stats"]["nsteps"], o["stats"]["nsteps"])
        for refill in (1, 8, 32):
            rng = np.random.RandomState(N + refill)
            g = gpu("vdp", "Dopri5", y0, p, tf=2.0, arith="strict", rtol=1e-6, atol=1e-6, refill=refill,
                    order=rng.permutation(N).astype(np.intc))
            assert np.array_equal(g["y"], base["y"]) and np.array_equal(g["stats"]["nsteps"], base["stats"]["nsteps"])
            for reverse in (1, 2):      # descending and two-ended fetch: every member exactly once, any size
                g = gpu("vdp", "Dopri5", y0, p, tf=2.0, arith="strict", rtol=1e-6, atol=1e-6, refill=refill, reverse=reverse)
                assert np.array_equal(g["y"], base["y"]) and np.array_equal(g["stats"]["nsteps"], base["stats"]["nsteps"])


def test_per_member_failure_status_does_not_abort_the_ensemble():
    y0, p = cases.vdp_sweep(64)
    g = gpu("vdp", "Dopri5", y0, p, tf=2.0, rtol=1e-6, atol=1e-6, maxsteps=50)
    o = O.simulate("vdp", "Dopri5", y0, p, tf=2.0, rtol=1e-6, atol=1e-6, maxsteps=50)
    assert np.array_equal(g["status"], o["status"])
    assert (g["status"] == 0).any() and (g["status"] == -2).any()
    g = gpu("vdp", "RungeKutta34", y0, p, tf=2.0, rtol=1e-6, atol=1e-6, maxsteps=100)
    o = O.simulate("vdp", "RungeKutta34", y0, p, tf=2.0, rtol=1e-6, atol=1e-6, maxsteps=100)
    assert np.array_equal(g["status"], o["status"]) and (g["status"] == -2).any()
    g = gpu("vdp", "Radau5ODE", y0, p, tf=2.0, rtol=1e-6, atol=1e-6, maxsteps=20, arith="strict")
    o = O.simulate("vdp", "Radau5ODE", y0, p, tf=2.0, rtol=1e-6, atol=1e-6, maxsteps=20)
    assert np.array_equal(g["status"], o["status"]) and (g["status"] == -5).any()


def test_tfinal_equal_to_t0_is_a_noop():
    g = gpu("vdp", "Dopri5", [[2.0, -0.6]], [[1.0]], tf=0.0)
    assert g["y"][0].tolist() == [2.0, -0.6] and g["stats"]["nsteps"][0] == 0 and g["status"][0] == 0


# ---------------------------------------------------------------------------------------------------------
# time events: Explicit_Problem.time_events + handle_event(event_info=[[], True])
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("solver,opts", [("Dopri5", {}), ("RungeKutta34", {}), ("RungeKutta4", dict(h=0.01)),
                                          ("ExplicitEuler", dict(h=0.01)), ("Radau5ODE", {})])
def test_time_events(solver, opts):
    gold = golden("time_ev_" + solver.lower())
    # the reference's vectors, member by member (each has its own tfinal)
    for k in range(len(gold["tf"])):
        g = gpu("time_ev", solver, gold["y0"][k:k + 1], gold["p"][k:k + 1], tf=float(gold["tf"][k]), arith="strict", **opts)
        assert g["status"][0] == 0 and g["t"][0] == gold["t"][k]
        assert rel_err(g["y"][0], gold["y"][k]) <= 1e-12
        assert g["stats"]["ntimeevents"][0] == gold["ntimeevents"][k]
    # an ensemble with different event times per member, both arithmetic modes, against the oracle
    N = 777
    rng = np.arange(N) / (N - 1.0)
    p = np.stack([0.3 + rng, 1.4 + rng, 2.5 + 0.4 * rng, 3.2 + 0.1 * rng, 0.5 + rng, 1.0 - 2.0 * rng], axis=1)
    y0 = np.zeros((N, 1))
    o = O.simulate("time_ev", solver, y0, p, tf=4.0, max_ev=24, **opts)
    for arith in ("strict", "fast"):
        g = gpu("time_ev", solver, y0, p, tf=4.0, arith=arith, **opts)
        assert (g["status"] == 0).all() and np.array_equal(g["t"], o["t"])
        assert rel_err(g["y"], o["y"]) <= (1e-12 if arith == "strict" else 1e-5)
        assert np.array_equal(g["stats"]["ntimeevents"], o["stats"]["ntimeevents"])
        assert (g["stats"]["ntimeevents"] == 4).all() and np.array_equal(g["nev"], o["stats"]["ntimeevents"])
        assert np.array_equal(g["ev_t"][:, :4], o["ev_t"][:, :4])          # time events are hit exactly
        assert (g["ev_info"][:, :4] == (1 << 30)).all()


# ---------------------------------------------------------------------------------------------------------
# ImplicitEuler
# ---------------------------------------------------------------------------------------------------------
IE_KEYS = ("nsteps", "nfcns", "njacs", "nniters", "nnfails")


@pytest.mark.parametrize("usejac", [1, 0])
@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_implicit_euler_vdp(usejac, arith):
    N = 2048
    y0, p = cases.vdp_sweep(N)
    o = O.simulate("vdp", "ImplicitEuler", y0, p, tf=2.0, h=1e-3, usejac=usejac)
    g = gpu("vdp", "ImplicitEuler", y0, p, tf=2.0, h=1e-3, usejac=usejac, arith=arith)
    assert (g["status"] == 0).all() and np.array_equal(g["t"], o["t"])
    assert rel_err(g["y"], o["y"]) <= (1e-10 if arith == "strict" else 1e-7)
    if arith == "strict":
        for k in IE_KEYS:
            assert np.array_equal(g["stats"][k], o["stats"][k]), k
    else:
        assert steps_within(g, o) and abs(int(g["sums"]["nniters"]) - int(o["stats"]["nniters"].sum())) <= 0.02 * o["stats"]["nniters"].sum()
    # the reference's own vectors
    gold = golden("impleuler_vdp_" + ("jac" if usejac else "fd"))
    y0g, pg = cases.vdp_sweep(int(gold["N"]), gold["idx"])
    gg = gpu("vdp", "ImplicitEuler", y0g, pg, tf=2.0, h=1e-3, usejac=usejac, arith=arith)
    assert rel_err(gg["y"], gold["y"]) <= (1e-10 if arith == "strict" else 1e-7)
    if arith == "strict":
        for k in IE_KEYS:
            assert np.array_equal(gg["stats"][k], gold[k]), k


def test_implicit_euler_events_dense_and_failures():
    # state events + post-event partial steps (euler.pyx:167-184,216-219)
    N = 512
    y0, p, sw0 = cases.ball_sweep(N)
    o = O.simulate("ball", "ImplicitEuler", y0, p, sw0, tf=3.0, max_ev=24, h=1e-3)
    g = gpu("ball", "ImplicitEuler", y0, p, sw0, tf=3.0, h=1e-3)
    assert (g["status"] == 0).all() and rel_err(g["y"], o["y"]) <= 1e-10
    assert np.array_equal(g["nev"], o["stats"]["nstateevents"]) and np.array_equal(g["sw"], o["sw"])
    m = np.isfinite(o["ev_t"])
    assert np.max(np.abs(g["ev_t"][m] - o["ev_t"][m])) < 1e-11
    for k in IE_KEYS:
        assert np.array_equal(g["stats"][k], o["stats"][k]), k
    gold = golden("impleuler_ball")
    y0g, pg, swg = cases.ball_sweep(int(gold["N"]), gold["idx"])
    gg = gpu("ball", "ImplicitEuler", y0g, pg, swg, tf=3.0, h=1e-3)
    assert rel_err(gg["y"], gold["y"]) <= 1e-10 and np.array_equal(gg["nev"], gold["nev"])
    # linear dense output on a grid
    grid = np.linspace(0, 2, 41)[1:]
    y0v, pv = cases.vdp_sweep(300)
    o = O.simulate("vdp", "ImplicitEuler", y0v, pv, tf=2.0, h=1e-3, usejac=1, t_grid=grid)
    g = gpu("vdp", "ImplicitEuler", y0v, pv, tf=2.0, h=1e-3, usejac=1, t_grid=grid)
    assert np.nanmax(np.abs(g["dense"] - o["dense"])) <= 1e-9 and not np.isnan(g["dense"]).any()
    # 'Newton iteration failed at 0.000000' on Robertson, as in the reference: status -8, state untouched
    gold = golden("impleuler_robertson")
    y0r, pr = cases.robertson_sweep(int(gold["N"]), gold["idx"])
    gr = gpu("robertson", "ImplicitEuler", y0r, pr, tf=4.0, h=1e-2, rtol=1e-6, atol=np.array(cases.ROBERTSON_ATOL), usejac=1)
    assert (gr["status"] == -8).all() and (gr["t"] == 0.0).all() and np.array_equal(gr["y"], gold["y"])
    for k in ("njacs", "nniters", "nnfails"):
        assert np.array_equal(gr["stats"][k], gold[k]), k


def test_implicit_euler_solver_class():
    from assimulo_b200.gpu import Batched_Explicit_Problem, ImplicitEuler
    y0, mu = cases.vdp_sweep(64)
    sim = ImplicitEuler(Batched_Explicit_Problem(model="vdp", y0=y0, p0=mu))
    assert sim.usejac and sim.newt == 7 and sim.h == 0.01
    sim.h = 1e-3
    t, y = sim.simulate(2.0, 4)
    o = O.simulate("vdp", "ImplicitEuler", y0, mu, tf=2.0, h=1e-3, usejac=1)
    assert len(t) == 5 and rel_err(y[-1], o["y"]) <= 1e-10
    assert sim.statistics["nniters"] == int(o["stats"]["nniters"].sum()) and sim.statistics["nsteps"] == 64 * 2001


# ---------------------------------------------------------------------------------------------------------
# RodasODE
# ---------------------------------------------------------------------------------------------------------
RODAS_KEYS = ("nsteps", "nfcns", "nerrfails", "njacs", "nlus")


@pytest.mark.parametrize("usejac", [1, 0])
@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_rodas_vdp_sweep(usejac, arith):
    N = 2048
    y0, p = cases.vdp_sweep(N)
    o = O.simulate("vdp", "RodasODE", y0, p, tf=2.0, rtol=RTOL, atol=1e-6, usejac=usejac)
    g = gpu("vdp", "RodasODE", y0, p, tf=2.0, rtol=RTOL, atol=1e-6, usejac=usejac, arith=arith)
    assert (g["status"] == 0).all() and np.array_equal(g["t"], o["t"])
    # strict arithmetic reproduces every operation except libm's pow (err ** 0.25 in the controller: device and glibc
    # pow differ by an ulp now and then, hence h, hence the state in its last bits).  With an analytic Jacobian that
    # stays at 1e-13; a finite-difference Jacobian divides such last-bit differences by delt ~ 1e-8 and a Rosenbrock
    # method uses J in the solution itself (not only for convergence like Newton), so there the bound is rtol.
    strict_tol = 1e-10 if usejac else RTOL
    assert rel_err(g["y"], o["y"]) <= (strict_tol if arith == "strict" else 10 * RTOL)
    assert steps_within(g, o)
    if arith == "strict" and usejac:
        for k in RODAS_KEYS:
            assert np.array_equal(g["stats"][k], o["stats"][k]), k
    elif arith == "strict":
        assert np.abs(g["stats"]["nsteps"].astype(int) - o["stats"]["nsteps"]).max() <= 2
    gold = golden("rodas_vdp_" + ("jac" if usejac else "fd"))
    y0g, pg = cases.vdp_sweep(int(gold["N"]), gold["idx"])
    gg = gpu("vdp", "RodasODE", y0g, pg, tf=2.0, rtol=RTOL, atol=1e-6, usejac=usejac, arith=arith)
    assert rel_err(gg["y"], gold["y"]) <= (strict_tol if arith == "strict" else 10 * RTOL)


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_rodas_stiff_events_dense(arith):
    # the reference's known answer (tests/solvers/test_rosenbrock.py:83-91)
    g = gpu("vdp", "RodasODE", [[2.0, -0.6]], [[1e6]], tf=2.0, rtol=1e-4, atol=1e-4, usejac=1, arith=arith)
    assert g["status"][0] == 0 and abs(g["y"][0, 0] - 1.7061680350) < 1e-4
    # Robertson, 2^10 members
    N = 1024
    y0, p = cases.robertson_sweep(N)
    o = O.simulate("robertson", "RodasODE", y0, p, tf=40.0, rtol=1e-6, atol=cases.ROBERTSON_ATOL, usejac=1)
    g = gpu("robertson", "RodasODE", y0, p, tf=40.0, rtol=1e-6, atol=np.array(cases.ROBERTSON_ATOL), usejac=1, arith=arith)
    assert (g["status"] == 0).all() and rel_err(g["y"], o["y"]) <= 1e-5 and steps_within(g, o)
    assert np.max(np.abs(g["y"].sum(axis=1) - 1.0)) < 1e-9
    # bouncing ball: state events through CONTRO
    y0, p, sw0 = cases.ball_sweep(512)
    o = O.simulate("ball", "RodasODE", y0, p, sw0, tf=3.0, max_ev=24, rtol=1e-6, atol=1e-6)
    g = gpu("ball", "RodasODE", y0, p, sw0, tf=3.0, rtol=1e-6, atol=1e-6, arith=arith)
    assert (g["status"] == 0).all() and np.array_equal(g["nev"], o["stats"]["nstateevents"])
    m = np.isfinite(o["ev_t"])
    assert np.max(np.abs(g["ev_t"][m] - o["ev_t"][m])) < 1e-9 and rel_err(g["y"], o["y"]) <= 1e-5
    # dense rows
    grid = np.linspace(0, 2, 41)[1:]
    y0v, pv = cases.vdp_sweep(300)
    o = O.simulate("vdp", "RodasODE", y0v, pv, tf=2.0, t_grid=grid, usejac=1)
    g = gpu("vdp", "RodasODE", y0v, pv, tf=2.0, t_grid=grid, usejac=1, arith=arith)
    assert np.nanmax(np.abs(g["dense"] - o["dense"])) <= (1e-9 if arith == "strict" else 1e-4) and not np.isnan(g["dense"]).any()


def test_rodas_solver_class():
    from assimulo_b200.gpu import Batched_Explicit_Problem, RodasODE, Rodas_Exception
    y0, mu = cases.vdp_sweep(64)
    sim = RodasODE(Batched_Explicit_Problem(model="vdp", y0=y0, p0=mu))
    assert sim.usejac and sim.maxsteps == 10000 and sim.fac2 == 6.0 and sim.inith == 0.01
    with pytest.raises(Rodas_Exception):
        sim.maxh = -1.0
    t, y = sim.simulate(2.0, 4)
    o = O.simulate("vdp", "RodasODE", y0, mu, tf=2.0, usejac=1)
    assert len(t) == 5 and rel_err(y[-1], o["y"]) <= 1e-5
    assert abs(sim.statistics["nsteps"] - int(o["stats"]["nsteps"].sum())) <= 0.02 * o["stats"]["nsteps"].sum()
    assert sim.statistics["nlus"] > 0 and sim.statistics["njacs"] > 0


def test_known_answers_of_the_reference_examples_on_the_device():
    """The pins SURVEY 8(c) lists for the benchmark MODELS, evaluated by the device path itself (both arithmetic builds):
    Robertson y1(0.4) (examples/radau5ode_with_jac_sparse.py:91) and y(4) (examples/cvode_with_parameters.py:89-91),
    gyro y(0.1) (examples/cvode_gyro.py:85-86), bouncing ball first impact sqrt(2*2/9.81) and the impact sequence of
    examples/lsodar_bouncing_ball.py."""
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5, Radau5ODE, RodasODE, RungeKutta34
    for arith in ("strict", "fast"):
        rob = dict(model="robertson", y0=[[1.0, 0.0, 0.0]], p0=[[0.04, 1e4, 3e7]])
        sim = Radau5ODE(Batched_Explicit_Problem(**rob))
        sim.arith, sim.rtol, sim.atol = arith, 1e-4, cases.ROBERTSON_ATOL
        sim.simulate(0.4)
        assert abs(sim.y[0, 0] - 0.9851) < 1e-3
        for cls in (Radau5ODE, RodasODE):
            sim = cls(Batched_Explicit_Problem(**rob))
            sim.arith, sim.rtol, sim.atol = arith, 1e-8, [1e-10, 1e-14, 1e-10]
            sim.simulate(4.0)
            assert np.abs(sim.y[0] - [0.905518032, 2.24046805e-5, 0.0944595637]).max() < 1e-4
        gy = Dopri5(Batched_Explicit_Problem(model="gyro", y0=[np.hstack([[1e4, 5e4, 6e4], np.eye(3).reshape(9)])],
                                             p0=[[1000.0, 5000.0, 6000.0]]))
        gy.arith, gy.rtol, gy.atol = arith, 1e-10, 1e-10
        gy.simulate(0.1)
        # the example's y[0] comes from CVode BDF(maxord 2) and carries a 2.3e-4 error (see test_oracle_golden.py)
        assert abs(gy.y[0, 0] - 692.800241862) < 1e-3 and abs(gy.y[0, 0] - 692.8004700450) < 1e-7
        assert abs(gy.y[0, 8] - 7.08468221e-1) < 1e-6
        for cls in (Dopri5, RungeKutta34):
            ball = cls(Batched_Explicit_Problem(model="ball", y0=[[2.0, 0.0]], p0=[[-9.81, 0.88]], sw0=[[1, 0]]))
            ball.arith = arith
            ball.simulate(3.0)
            te = ball.event_times(0)
            # impacts (every second event: the apex events lie between them), SURVEY 8(c) probe confirmations
            want = [0.638550856814101, 1.200475610810578, 1.762400364807010, 2.256894148324066, 2.751387931840934]
            hits = [t for t in te if any(abs(t - w) < 1e-6 for w in want)]
            assert len(hits) == 5 and abs(hits[0] - np.sqrt(2 * 2 / 9.81)) < 1e-9
