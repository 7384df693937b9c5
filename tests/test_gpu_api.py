"""-m gpu: the reference-facing surface on the device -- solver classes (simulate / ncp / continuation / statistics /
event log), NVRTC-compiled user models, the raw C ABI through ctypes, and full-size property checks."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
import ens_oracle as O
from conftest import ROOT, rel_err

pytestmark = pytest.mark.gpu


def test_simulate_surface_dopri5():
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5
    N = 1000
    y0, mu = cases.vdp_sweep(N)
    sim = Dopri5(Batched_Explicit_Problem(model="vdp", y0=y0, p0=mu, name="vdp sweep"))
    sim.rtol, sim.atol = 1e-6, 1e-6
    t, y = sim.simulate(2.0)
    assert t == [0.0, 2.0] and y.shape == (2, N, 2)
    assert np.array_equal(y[0], y0)
    o = O.simulate("vdp", "Dopri5", y0, mu, tf=2.0, rtol=1e-6, atol=1e-6)
    assert rel_err(y[-1], o["y"]) <= 1e-5
    assert abs(sim.statistics["nsteps"] - o["stats"]["nsteps"].sum()) <= 0.02 * o["stats"]["nsteps"].sum()
    assert sim.statistics["nfcns"] == o["stats"]["nfcns"].sum() or abs(
        sim.statistics["nfcns"] - o["stats"]["nfcns"].sum()) <= 0.02 * o["stats"]["nfcns"].sum()
    assert sim.statistics.per_member("nsteps").shape == (N,)
    assert sim.t == 2.0 and (sim.status == 0).all() and sim.n_not_complete == 0
    sim.print_statistics()


def test_ncp_output_grid_and_continuation():
    """ODE.simulate: rows at linspace(t0, tf, ncp+1)[1:]; a second simulate() continues from the current time
    (src/ode.pyx:195-196,222, examples/euler_basic.py:49-50) and must land on the one-shot result."""
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5, ExplicitEuler
    N = 257
    y0, mu = cases.vdp_sweep(N)
    sim = Dopri5(Batched_Explicit_Problem(model="vdp", y0=y0, p0=mu))
    sim.arith = "strict"
    t, y = sim.simulate(2.0, 20)
    assert len(t) == 21 and np.allclose(t, np.linspace(0, 2, 21)) and y.shape == (21, N, 2)
    grid = np.linspace(0, 2, 21)[1:]
    o = O.simulate("vdp", "Dopri5", y0, mu, tf=2.0, t_grid=grid)
    assert np.nanmax(np.abs(np.transpose(y[1:], (1, 0, 2)) - o["dense"])) <= 1e-9
    # continuation 0 -> 1 -> 2 with a fixed-step solver equals 0 -> 2 exactly when 1.0 is on the step grid
    a = ExplicitEuler(Batched_Explicit_Problem(model="linear", y0=[[4.0]] * 8, p0=[-1.0, 0.0, 0.0]))
    a.h = 0.01
    a.simulate(3.0)
    a.simulate(5.0)
    assert a.t == 5.0
    assert abs(a.y[0, 0] - 0.02628193) < 1e-6                      # examples/euler_basic.py:63
    # ncp_list
    t, y = sim.simulate(3.0, ncp_list=[2.5, 2.75])
    assert t == [2.0, 2.5, 2.75, 3.0]
    # reset
    sim.reset()
    t, y = sim.simulate(2.0)
    assert rel_err(y[-1], o["y"]) <= 1e-9


def test_rk4_ignores_ncp_like_the_reference():
    from assimulo_b200.gpu import Batched_Explicit_Problem, RungeKutta4
    sim = RungeKutta4(Batched_Explicit_Problem(model="linear", y0=[[4.0]], p0=[-1.0, 0.0, 0.0]))
    t, y = sim.simulate(5.0, 100)
    assert t == [0.0, 5.0] and abs(y[-1][0, 0] - 0.02695179) < 1e-6


def test_event_log_and_switches_through_the_solver_class():
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5, Radau5ODE
    N = 64
    y0, p, sw0 = cases.ball_sweep(N)
    o = O.simulate("ball", "Dopri5", y0, p, sw0, tf=3.0, max_ev=16)
    sim = Dopri5(Batched_Explicit_Problem(model="ball", y0=y0, p0=p, sw0=sw0))
    sim.simulate(3.0)
    cnt, te, info = sim.event_log
    assert np.array_equal(cnt, o["stats"]["nstateevents"]) and sim.statistics["nstateevents"] == cnt.sum()
    for m in (0, N // 2, N - 1):
        ev = sim.event_times(m)
        assert len(ev) == cnt[m] and np.all(np.diff(ev) > 0)
        assert np.max(np.abs(ev - o["ev_t"][m, :cnt[m]])) < 1e-9
    assert np.array_equal(sim.sw.astype(np.uint8), o["sw"])
    # ring semantics: only the last `event_slots` events are kept, in slot e % slots
    sim2 = Dopri5(Batched_Explicit_Problem(model="ball", y0=y0, p0=p, sw0=sw0))
    sim2.event_slots = 4
    sim2.simulate(3.0)
    for m in (0, N - 1):
        assert np.allclose(sim2.event_times(m), o["ev_t"][m, max(0, cnt[m] - 4):cnt[m]], atol=1e-9)
    r = Radau5ODE(Batched_Explicit_Problem(model="extended", y0=[0.0, -1.0, 0.0]))
    r.simulate(10.0)
    np.testing.assert_allclose(r.y[0], [8.0, 3.0, 2.0], atol=1e-6)      # test_radau5.py:341-354
    assert r.statistics["nstateevents"] == 1


VDP_SRC = r"""
__device__ void aens_rhs(double t, const double* y, const unsigned char* sw, const double* p, double* yd) {
    yd[0] = y[1];
    yd[1] = p[0] * ((1. - y[0] * y[0]) * y[1] - y[0]);
}
__device__ void aens_jac(double t, const double* y, const unsigned char* sw, const double* p, double* J) {
    J[0] = 0.; J[1] = p[0] * (-2. * y[0] * y[1] - 1.); J[2] = 1.; J[3] = p[0] * (1. - y[0] * y[0]);
}
"""
BALL_SRC = r"""
__device__ void aens_rhs(double t, const double* y, const unsigned char* sw, const double* p, double* yd) {
    yd[0] = y[1]; yd[1] = p[0];
}
__device__ void aens_events(double t, const double* y, const unsigned char* sw, const double* p, double* g) {
    g[0] = sw[0] ? y[0] : 5.; g[1] = sw[1] ? y[1] : 5.;
}
__device__ int aens_handle_event(double t, double* y, unsigned char* sw, const int* info, const double* p) {
    if (info[0] != 0) { sw[0] = 0; sw[1] = 1; y[1] = -p[1] * y[1]; } else { sw[0] = 1; sw[1] = 0; }
    return 0;
}
"""
TERMINATE_SRC = r"""
__device__ void aens_rhs(double t, const double* y, const unsigned char* sw, const double* p, double* yd) { yd[0] = 1.0; }
__device__ void aens_events(double t, const double* y, const unsigned char* sw, const double* p, double* g) { g[0] = y[0] - p[0]; }
__device__ int aens_handle_event(double t, double* y, unsigned char* sw, const int* info, const double* p) { return 1; }
"""


def test_nvrtc_user_model_equals_builtin():
    """A model supplied as a CUDA source string runs through the same kernel templates: identical results."""
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5, Radau5ODE, RungeKutta4
    N = 512
    y0, mu = cases.vdp_sweep(N)
    for cls, kw in ((Dopri5, {}), (RungeKutta4, {"h": 1e-3}), (Radau5ODE, {})):
        res = []
        for prob in (Batched_Explicit_Problem(model="vdp", y0=y0, p0=mu),
                     Batched_Explicit_Problem(source=VDP_SRC, n=2, n_p=1, has_jac=True, y0=y0, p0=mu)):
            sim = cls(prob)
            sim.arith = "strict"
            for k, v in kw.items():
                setattr(sim, k, v)
            sim.simulate(2.0)
            res.append((sim.y.copy(), sim.statistics["nsteps"]))
        assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1], cls.__name__
    y0, p, sw0 = cases.ball_sweep(128)
    a = Dopri5(Batched_Explicit_Problem(model="ball", y0=y0, p0=p, sw0=sw0))
    b = Dopri5(Batched_Explicit_Problem(source=BALL_SRC, n=2, n_p=2, n_g=2, n_sw=2, y0=y0, p0=p, sw0=sw0))
    for s in (a, b):
        s.arith = "strict"          # fast arithmetic too: the second cubin is compiled on demand
        s.simulate(3.0)
    assert np.array_equal(a.y, b.y) and np.array_equal(a.event_log[1], b.event_log[1], equal_nan=True)   # unused slots read NaN
    b.arith = "fast"
    b.reset()
    b.simulate(3.0)
    assert rel_err(b.y, a.y) < 1e-5


def test_handle_event_can_terminate_a_member():
    """TerminateSimulation from handle_event (src/explicit_ode.pyx:260-262) -> per-member status 1."""
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5
    thr = np.array([[0.5], [5.0]])
    sim = Dopri5(Batched_Explicit_Problem(source=TERMINATE_SRC, n=1, n_p=1, n_g=1, n_sw=1, y0=[[0.0], [0.0]],
                                          p0=thr, sw0=[[1], [1]]))
    sim.simulate(2.0)
    assert sim.status.tolist() == [1, 0]
    assert abs(sim.t_members[0] - 0.5) < 1e-9 and sim.t_members[1] == 2.0
    assert abs(sim.y[0, 0] - 0.5) < 1e-9 and abs(sim.y[1, 0] - 2.0) < 1e-9


def test_bad_source_raises_with_compiler_log():
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5, _core
    sim = Dopri5(Batched_Explicit_Problem(source="__device__ void aens_rhs(int x) { nope }", n=1, y0=[[0.0]]))
    with pytest.raises(_core.AensError) as e:
        sim.simulate(1.0)
    assert e.value.code == -3 and "error" in str(e.value)


def test_raw_c_abi_through_ctypes():
    """The boundary is a plain C ABI: drive it with ctypes only (no Cython, no torch)."""
    lib = C.CDLL(os.path.join(ROOT, "assimulo_b200", "lib", "libassimulo_cuda.so"))

    class Opts(C.Structure):
        _fields_ = [("solver", C.c_int), ("arith", C.c_int), ("maxsteps", C.c_int), ("newt", C.c_int),
                    ("usejac", C.c_int), ("reserved", C.c_int), ("rtol", C.c_double), ("atol", C.c_double * 16),
                    ("h", C.c_double), ("inith", C.c_double), ("maxh", C.c_double), ("safe", C.c_double),
                    ("fac1", C.c_double), ("fac2", C.c_double), ("beta", C.c_double), ("thet", C.c_double),
                    ("fnewt", C.c_double), ("quot1", C.c_double), ("quot2", C.c_double)]
    lib.aens_last_error.restype = C.c_char_p
    lib.aens_simulate.argtypes = [C.c_void_p, C.c_double]
    lib.aens_set_ensemble.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
    h = C.c_void_p()
    assert lib.aens_create(0, C.byref(h)) == 0
    assert lib.aens_simulate(h, 1.0) == -4 and b"model" in lib.aens_last_error(h)      # call order enforced
    assert lib.aens_set_model_builtin(h, b"no_such_model") == -1
    assert lib.aens_set_model_builtin(h, b"vdp") == 0
    N = 777
    y0, mu = cases.vdp_sweep(N)
    y0 = np.ascontiguousarray(y0)
    mu = np.ascontiguousarray(mu)
    assert lib.aens_set_ensemble(h, N, y0.ctypes.data, mu.ctypes.data, None, 0.0) == 0
    o = Opts()
    assert lib.aens_default_opts(3, C.byref(o)) == 0
    o.arith = 0
    assert lib.aens_set_opts(h, C.byref(o)) == 0
    assert lib.aens_set_output(h, 0, None, 0) == 0
    assert lib.aens_simulate(h, 2.0) == 0
    t = np.empty(N)
    y = np.empty((N, 2))
    st = np.empty(N, dtype=np.int32)
    ns = np.empty(N, dtype=np.int32)
    assert lib.aens_get_final(h, t.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), None,
                              st.ctypes.data_as(C.c_void_p)) == 0
    assert lib.aens_get_stats(h, 0, ns.ctypes.data_as(C.c_void_p)) == 0
    ref = O.simulate("vdp", "Dopri5", y0, mu, tf=2.0)
    assert np.array_equal(ns, ref["stats"]["nsteps"]) and rel_err(y, ref["y"]) < 1e-10 and (st == 0).all()
    lib.aens_nstats.restype = C.c_int
    sums = (C.c_longlong * lib.aens_nstats())()   # sized by the library, not by a literal
    assert lib.aens_get_stats_sum(h, sums) == 0 and sums[0] == ns.sum()
    assert lib.aens_reset(h) == 0 and lib.aens_simulate(h, 2.0) == 0
    y2 = np.empty((N, 2))
    lib.aens_get_final(h, None, y2.ctypes.data_as(C.c_void_p), None, None)
    assert np.array_equal(y, y2)                                                      # reset + rerun is deterministic
    assert lib.aens_destroy(C.byref(h)) == 0 and not h.value


def test_full_size_properties_cfg2():
    """BASELINE cfg 2 at full size (2^20 members): size-independent properties + a strided oracle sample."""
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5, RungeKutta4, RungeKutta34
    N = 1 << 20
    y0, mu = cases.vdp_sweep(N)
    idx = np.arange(0, N, 4096)
    for cls, oname, kw in ((RungeKutta4, "RungeKutta4", dict(h=1e-3)),
                           (RungeKutta34, "RungeKutta34", dict(rtol=1e-6, atol=1e-6, inith=0.01)),
                           (Dopri5, "Dopri5", dict(rtol=1e-6, atol=1e-6))):
        sim = cls(Batched_Explicit_Problem(model="vdp", y0=y0, p0=mu))
        for k, v in kw.items():
            setattr(sim, k, v)
        sim.options["reuse_buffers"] = True
        sim.simulate(2.0)
        assert not sim.status.any() and np.isfinite(sim.y).all() and (sim.t_members == 2.0).all()
        ns = sim.statistics.per_member("nsteps")
        assert sim.statistics["nsteps"] == ns.astype(np.int64).sum()
        if oname == "RungeKutta4":
            assert (ns == 2001).all()
        else:
            assert np.abs(np.diff(ns.astype(int))).max() <= 3       # smooth sweep: neighbours cost the same
        o = O.simulate("vdp", oname, y0[idx], mu[idx], tf=2.0, **kw)
        tol = 1e-13 if oname == "RungeKutta4" else 1e-5
        assert rel_err(sim.y[idx], o["y"]) <= tol
        assert abs(int(ns[idx].sum()) - int(o["stats"]["nsteps"].sum())) <= 0.02 * o["stats"]["nsteps"].sum()
        # the members are independent: a permuted ensemble gives the permuted result
        perm = np.random.RandomState(0).permutation(N)[:65536]
        sub = cls(Batched_Explicit_Problem(model="vdp", y0=y0[perm], p0=mu[perm]))
        for k, v in kw.items():
            setattr(sub, k, v)
        sub.simulate(2.0)
        assert rel_err(sub.y, sim.y[perm]) <= (0 if oname == "RungeKutta4" else 1e-12)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("N,nchunks", [(1, 1), (33, 4), (1000, 7), (4096, 8), (5000, 64)])
def test_host_pipeline_equals_the_three_call_path(N, nchunks, mode):
    """aens_simulate_host -- mode 1: chunked upload / kernels / download on overlapping streams; mode 2: zero-copy,
    the kernel reads and writes the caller's page-locked host arrays directly -- must give bit-identical states, end
    times, status codes and statistics to set_ensemble + simulate + get_final, for ragged sizes, any chunk count and
    both fetch directions."""
    from assimulo_b200.gpu import _core
    y0, mu = cases.vdp_sweep(max(N, 2))
    y0, mu = y0[:N], mu[:N]
    if mode == 2:
        y0p, mup = _core.pinned_empty((N, 2)), _core.pinned_empty((N, 1))
        y0p[:], mup[:] = y0, mu
        y0, mu = y0p, mup
    for reverse in (0, 1, 2):
        h = _core.Handle(0)
        h.set_host_mode(mode)
        h.set_model_builtin("vdp")
        h.set_opts({"solver": _core.SOLVER_IDS["Dopri5"], "arith": 0, "rtol": 1e-6, "atol": 1e-6})
        h.set_fetch_reverse(reverse)
        h.set_ensemble(y0, mu, None, 0.0)
        h.set_output(None, 16)
        h.simulate(2.0)
        t1, y1, _, st1 = h.get_final()
        s1 = h.get_stats_sum()
        n1 = {k: h.get_stats(k) for k in ("nsteps", "nfcns", "nerrfails")}
        oy, ot, ost = _core.pinned_empty((N, 2)), _core.pinned_empty((N,)), _core.pinned_empty((N,), np.intc)
        oy[:], ot[:], ost[:] = np.nan, np.nan, -99
        s2, nbad = h.simulate_host(y0, mu, None, 0.0, 2.0, out_y=oy, out_t=ot, out_status=ost, nchunks=nchunks)
        assert nbad == 0 and s1 == s2
        assert np.array_equal(oy, y1) and np.array_equal(ot, t1) and np.array_equal(ost, st1)
        for k, v in n1.items():
            assert np.array_equal(h.get_stats(k), v)
        # summary entry point agrees
        s3, nb3 = h.get_summary()
        assert s3 == s2 and nb3 == 0
        # aens_reset restores the pristine copy written by the pipelined upload
        h.reset()
        h.simulate(2.0)
        assert np.array_equal(h.get_final()[1], y1)
        h.close()


def test_host_pipeline_with_events_switches_and_failures():
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5
    N = 3000
    y0, p, sw0 = cases.ball_sweep(N)
    o = O.simulate("ball", "Dopri5", y0, p, sw0, tf=3.0, max_ev=16, rtol=1e-6, atol=1e-6)
    # zero-copy with switches and all outputs, straight through the handle
    from assimulo_b200.gpu import _core
    h = _core.Handle(0)
    h.set_host_mode(2)
    h.set_model_builtin("ball")
    h.set_opts({"solver": _core.SOLVER_IDS["Dopri5"], "arith": 1, "rtol": 1e-6, "atol": 1e-6})
    pin = [_core.pinned_empty((N, 2)), _core.pinned_empty((N, 2)), _core.pinned_empty((N, 2), np.uint8)]
    pin[0][:], pin[1][:], pin[2][:] = y0, p, sw0
    oy, ot = _core.pinned_empty((N, 2)), _core.pinned_empty((N,))
    osw, ost = _core.pinned_empty((N, 2), np.uint8), _core.pinned_empty((N,), np.intc)
    sums, nbad = h.simulate_host(pin[0], pin[1], pin[2], 0.0, 3.0, out_y=oy, out_t=ot, out_sw=osw, out_status=ost)
    assert nbad == 0 and (ost == 0).all() and (ot == 3.0).all()
    assert rel_err(oy, o["y"]) <= 1e-5 and np.array_equal(osw, o["sw"])
    assert sums["nstateevents"] == int(o["stats"]["nstateevents"].sum())
    cnt = h.get_events()[0]
    assert np.array_equal(cnt, o["stats"]["nstateevents"])
    with pytest.raises(_core.AensError):                 # pageable arrays cannot be mapped: mode 2 refuses
        h.simulate_host(y0, p, sw0, 0.0, 3.0, out_y=oy)
    h.close()
    sim = Dopri5(Batched_Explicit_Problem(model="ball", y0=y0, p0=p, sw0=sw0))
    sim.options["host_chunks"] = 5
    sim.order = "descending"
    sim.simulate(3.0)                                   # first call: ensemble is dirty -> pipelined path
    assert sim.elapsed_kernel_ms is None and sim.n_not_complete == 0 and (sim.status == 0).all()
    assert rel_err(sim.y, o["y"]) <= 1e-5 and np.array_equal(sim.sw, o["sw"].astype(bool))
    cnt, te, info = sim.event_log
    assert np.array_equal(cnt, o["stats"]["nstateevents"])
    m = np.isfinite(o["ev_t"])
    assert np.max(np.abs(te[m] - o["ev_t"][m])) < 1e-9
    assert sim.statistics["nstateevents"] == int(o["stats"]["nstateevents"].sum())
    # members that fail are counted by the summary and their codes fetched lazily
    y0v, mu = cases.vdp_sweep(512)
    sim = Dopri5(Batched_Explicit_Problem(model="vdp", y0=y0v, p0=mu))
    sim.maxsteps = 40
    sim.simulate(2.0)
    ora = O.simulate("vdp", "Dopri5", y0v, mu, tf=2.0, rtol=1e-6, atol=1e-6, maxsteps=40)
    bad = ora["status"] != 0
    assert sim.n_not_complete == int(bad.sum()) > 0
    assert ((sim.status == -2) == bad).all() and (sim.t_members[bad] < 2.0).all() and sim.t < 2.0
    assert (sim.t_members[~bad] == 2.0).all()


def test_time_events_through_the_solver_class_and_nvrtc():
    """tests/solvers/test_rungekutta.py:49-82 on the device: built-in model and the same model as NVRTC source."""
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5, RungeKutta34
    src = r"""
__device__ void aens_rhs(double t, const double* y, const unsigned char* sw, const double* p, double* yd) { yd[0] = p[5]; }
__device__ double aens_time_events(double t, const double* y, const unsigned char* sw, const double* p) {
    for (int k = 0; k < 4; ++k) if (t < p[k]) return p[k];
    return AENS_NO_TIME_EVENT;
}
__device__ int aens_handle_time_event(double t, double* y, unsigned char* sw, const double* p) { y[0] += p[4]; return 0; }
"""
    N = 100
    p = np.tile(np.array([1.0, 2.0, 2.5, 3.0, 1.0, 1.0]), (N, 1))
    p[:, 4] = 1.0 + np.arange(N) / N
    for cls in (Dopri5, RungeKutta34):
        a = cls(Batched_Explicit_Problem(model="time_ev", y0=np.zeros((N, 1)), p0=p))
        b = cls(Batched_Explicit_Problem(source=src, n=1, n_p=6, time_events=True, y0=np.zeros((N, 1)), p0=p))
        for sim in (a, b):
            sim.simulate(5.0, 100)                                   # exp_sim(5., 100) as in the reference test
            assert sim.statistics["ntimeevents"] == 4 * N and sim.t == 5.0
            assert np.allclose(sim.y[:, 0], 5.0 + 4.0 * p[:, 4], rtol=0, atol=1e-9)
            cnt, te, info = sim.event_log
            assert (cnt == 4).all() and np.array_equal(te[:, :4], p[:, :4]) and (info[:, :4] == (1 << 30)).all()
        assert np.array_equal(a.y, b.y)


def test_ensemble_statistics_on_device():
    """aens_reduce_final: mean / var / min / max of the final states over the completed members, on the device."""
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5
    N = 5000
    y0, mu = cases.vdp_sweep(N)
    sim = Dopri5(Batched_Explicit_Problem(model="vdp", y0=y0, p0=mu))
    sim.simulate(2.0)
    st = sim.ensemble_statistics()
    y = sim.y
    assert (st["count"] == N).all()
    np.testing.assert_allclose(st["mean"], y.mean(axis=0), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(st["var"], y.var(axis=0), rtol=1e-10)
    assert np.array_equal(st["min"], y.min(axis=0)) and np.array_equal(st["max"], y.max(axis=0))
    # failed members are left out
    sim = Dopri5(Batched_Explicit_Problem(model="vdp", y0=y0, p0=mu))
    sim.maxsteps = 60
    sim.simulate(2.0)
    ok = sim.status == 0
    assert 0 < ok.sum() < N
    st = sim.ensemble_statistics()
    assert (st["count"] == ok.sum()).all()
    np.testing.assert_allclose(st["mean"], sim.y[ok].mean(axis=0), rtol=1e-12, atol=1e-13)
    assert np.array_equal(st["min"], sim.y[ok].min(axis=0)) and np.array_equal(st["max"], sim.y[ok].max(axis=0))


def test_maximum_state_dimension_nvrtc_model_against_analytic_solution():
    """n = AENS_MAX_N = 16 through NVRTC: exercises the wide-system code paths (row-sum Dopri5 stages, rolled LU of
    Radau5ODE / ImplicitEuler in local memory, FD and analytic Jacobians).  y_i' = -lam_i y_i + c, lam_i = p0 (i+1)."""
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5, RungeKutta34, RungeKutta4, Radau5ODE, ImplicitEuler
    n = 16
    src = r"""
__device__ void aens_rhs(double t, const double* y, const unsigned char* sw, const double* p, double* yd) {
    for (int i = 0; i < 16; ++i) yd[i] = -p[0] * (i + 1) * y[i] + p[1];
}
__device__ void aens_jac(double t, const double* y, const unsigned char* sw, const double* p, double* J) {
    for (int k = 0; k < 256; ++k) J[k] = 0.0;
    for (int i = 0; i < 16; ++i) J[i + 16 * i] = -p[0] * (i + 1);
}
"""
    N = 300
    p = np.stack([0.5 + np.arange(N) / N, 0.1 + 0.0 * np.arange(N)], axis=1)
    y0 = np.tile(1.0 + 0.1 * np.arange(n), (N, 1))
    lam = p[:, :1] * (np.arange(n)[None, :] + 1.0)
    tf = 1.5
    exact = (y0 - p[:, 1:2] / lam) * np.exp(-lam * tf) + p[:, 1:2] / lam
    prob = lambda: Batched_Explicit_Problem(source=src, n=n, n_p=2, has_jac=True, y0=y0, p0=p)  # noqa: E731
    for cls, tol, opts in ((Dopri5, 2e-5, {}), (RungeKutta34, 2e-5, {}), (RungeKutta4, 1e-6, dict(h=0.005)),
                           (Radau5ODE, 2e-5, dict(usejac=True)), (Radau5ODE, 2e-5, dict(usejac=False)),
                           (ImplicitEuler, 2e-2, dict(h=0.005, usejac=True)),
                           (ImplicitEuler, 2e-2, dict(h=0.005, usejac=False))):
        sim = cls(prob())
        for k, v in opts.items():
            setattr(sim, k, v)
        sim.simulate(tf)
        assert sim.n_not_complete == 0, cls.__name__
        err = np.max(np.abs(sim.y - exact) / np.maximum(np.abs(exact), 1e-3))
        assert err < tol, (cls.__name__, opts, err)


def test_full_size_properties_cfg3_ball_events():
    """BASELINE cfg 3 at full size (2^20 bouncing balls): every member's event log obeys the physics (impact times
    increase, indicators alternate impact / apex), counts are monotone in the drop height, and a strided sample
    matches the oracle event for event."""
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5
    N = 1 << 20
    y0, p, sw0 = cases.ball_sweep(N)
    sim = Dopri5(Batched_Explicit_Problem(model="ball", y0=y0, p0=p, sw0=sw0))
    sim.options["event_slots"] = 16
    sim.simulate(3.0)
    assert sim.n_not_complete == 0 and sim.t == 3.0
    cnt, te, info = sim.event_log
    assert sim.statistics["nstateevents"] == int(cnt.astype(np.int64).sum()) and cnt.min() >= 3 and cnt.max() <= 16
    first = te[:, 0]
    assert np.allclose(first, np.sqrt(2.0 * y0[:, 0] / 9.81), rtol=0, atol=1e-7)      # first impact: analytic
    k = np.arange(16)[None, :] < cnt[:, None]
    dte = np.diff(te, axis=1)
    assert (dte[k[:, 1:]] > 0).all()                                                   # chronological
    assert (info[:, 0::2][k[:, 0::2]] == 0b01).all() and (info[:, 1::2][k[:, 1::2]] == 0b0100).all()  # impact (g0: + to -), apex (g1: + to -), ...
    assert (np.diff(cnt.astype(int)) <= 0).all()                                       # higher drop -> fewer bounces by t=3
    assert (sim.y[:, 0] > -1e-9).all()                                                 # nobody fell through the floor
    idx = np.arange(0, N, 8192)
    o = O.simulate("ball", "Dopri5", y0[idx], p[idx], sw0[idx], tf=3.0, max_ev=16, rtol=1e-6, atol=1e-6)
    assert np.array_equal(cnt[idx], o["stats"]["nstateevents"]) and rel_err(sim.y[idx], o["y"]) <= 1e-5
    m = np.isfinite(o["ev_t"])
    assert np.max(np.abs(te[idx][m] - o["ev_t"][m])) < 1e-9


def test_full_size_properties_cfg4_robertson_and_cfg5_gyro():
    """cfg 4 (2^18 Robertson, Radau5ODE): mass conservation y1+y2+y3 = 1 and positivity for every member, strided
    oracle sample.  cfg 5 (gyro, Dopri5 rtol 1e-8, dense rows; 2^18 of the 2^22 members to keep the test's host arrays
    small -- bench.py runs the full size): |pi| is conserved and Q stays orthogonal in every dense row."""
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5, Radau5ODE
    N = 1 << 18
    y0, p = cases.robertson_sweep(N)
    sim = Radau5ODE(Batched_Explicit_Problem(model="robertson", y0=y0, p0=p))
    sim.rtol, sim.atol, sim.usejac = 1e-6, cases.ROBERTSON_ATOL, True
    sim.simulate(40.0)
    assert sim.n_not_complete == 0
    assert np.max(np.abs(sim.y.sum(axis=1) - 1.0)) < 1e-9 and (sim.y > -1e-9).all()
    idx = np.arange(0, N, 1024)
    o = O.simulate("robertson", "Radau5ODE", y0[idx], p[idx], tf=40.0, rtol=1e-6, atol=cases.ROBERTSON_ATOL, usejac=1)
    assert rel_err(sim.y[idx], o["y"]) <= 1e-5
    ns = sim.statistics.per_member("nsteps")
    assert abs(int(ns[idx].sum()) - int(o["stats"]["nsteps"].sum())) <= 0.02 * o["stats"]["nsteps"].sum()
    # gyro
    Ng = 1 << 22
    idx = np.arange(0, Ng, 16)
    y0, p = cases.gyro_sweep(Ng, idx)
    sim = Dopri5(Batched_Explicit_Problem(model="gyro", y0=y0, p0=p))
    sim.rtol, sim.atol = 1e-8, 1e-8
    t, y = sim.simulate(0.1, 10)
    assert sim.n_not_complete == 0 and y.shape == (11, len(idx), 12) and np.isfinite(y).all()
    pi2 = (y[:, :, :3] ** 2).sum(axis=2)
    assert np.max(np.abs(pi2 / pi2[0] - 1.0)) < 1e-5                                   # |pi| conserved
    Q = y[:, :, 3:].reshape(11, len(idx), 3, 3)
    QtQ = np.einsum("tmij,tmik->tmjk", Q, Q)
    assert np.max(np.abs(QtQ - np.eye(3))) < 1e-5                                      # Q orthogonal
    sub = np.arange(0, len(idx), 4096)
    o = O.simulate("gyro", "Dopri5", y0[sub], p[sub], tf=0.1, rtol=1e-8, atol=1e-8, t_grid=np.linspace(0, 0.1, 11)[1:])
    assert np.max(np.abs(np.transpose(y[1:, sub], (1, 0, 2)) - o["dense"]) / np.maximum(np.abs(o["dense"]), 1.0)) <= 1e-7


@pytest.mark.parametrize("name,opts", [("Dopri5", dict(rtol=1e-6, atol=1e-6)), ("RungeKutta34", dict(rtol=1e-6, atol=1e-6)),
                                        ("Radau5ODE", dict(rtol=1e-6, atol=1e-6, usejac=True)),
                                        ("ImplicitEuler", dict(h=1e-3, usejac=True))])
def test_continuation_restarts_like_the_reference(name, opts):
    """A second simulate() continues from the device state: the reference re-initialises the solver (HINIT, FACOLD,
    Radau reinit, first-step logic) at the restart point (src/ode.pyx:195-196, explicit_ode.pyx:190-204), so 0 -> 1 -> 2
    must equal the oracle started at t = 1 from the state it reached at t = 1 -- states, times and statistics."""
    import assimulo_b200.gpu as G
    N = 500
    y0, mu = cases.vdp_sweep(N)
    sim = getattr(G, name)(G.Batched_Explicit_Problem(model="vdp", y0=y0, p0=mu))
    sim.arith = "strict"
    for k, v in opts.items():
        setattr(sim, k, v)
    sim.simulate(1.0)
    o1 = O.simulate("vdp", name, y0, mu, tf=1.0, **{k: (int(v) if k == "usejac" else v) for k, v in opts.items()})
    assert rel_err(sim.y, o1["y"]) <= 1e-10 and sim.statistics["nsteps"] == int(o1["stats"]["nsteps"].sum())
    sim.simulate(2.0)
    o2 = O.simulate("vdp", name, o1["y"], mu, t0=1.0, tf=2.0, **{k: (int(v) if k == "usejac" else v) for k, v in opts.items()})
    assert sim.t == 2.0 and rel_err(sim.y, o2["y"]) <= 1e-9
    assert sim.statistics["nsteps"] == int(o2["stats"]["nsteps"].sum())        # statistics restart with every simulate()


def _row_reduction_case(cls, prob_kw, tf, ncp, N_expect=None, **opts):
    """simulate the same ensemble twice -- rows stored per member, rows reduced in the kernel -- and compare the
    device's row statistics with numpy reductions of the stored rows"""
    from assimulo_b200.gpu import Batched_Explicit_Problem
    sims = []
    for red in (False, True):
        sim = cls(Batched_Explicit_Problem(**prob_kw))
        for k, v in opts.items():
            setattr(sim, k, v)
        sim.reduce_rows = red
        t, y = sim.simulate(tf, ncp)
        sims.append((sim, t, y))
    (a, ta, ya), (b, tb, yb) = sims
    rows = np.asarray(ya)[1:]                                    # [ncp, N, n], NaN where a member never got there
    rs = b.row_statistics
    assert a.row_statistics is None and rs is not None
    assert np.allclose(rs["t"], ta[1:]) and tb == [ta[0], ta[-1]] and np.asarray(yb).shape[0] == 2
    assert rel_err(b.y, a.y) <= 1e-12 and np.array_equal(b.status, a.status)   # the integration itself is untouched
    for k in ("nsteps", "nfcns", "nstateevents"):
        assert a.statistics[k] == b.statistics[k]
    reached = ~np.isnan(rows[:, :, 0])
    assert np.array_equal(rs["count"], reached.sum(axis=1))
    for r in range(rows.shape[0]):
        v = rows[r][reached[r]]
        if len(v) == 0:
            assert np.isnan(rs["mean"][r]).all() and np.isnan(rs["min"][r]).all()
            continue
        scale = np.abs(v).max(axis=0) + 1e-300
        if opts.get("arith") == "strict":   # no FMA contraction: both kernels evaluate the interpolant identically
            assert np.array_equal(rs["min"][r], v.min(axis=0)) and np.array_equal(rs["max"][r], v.max(axis=0))
        else:                               # separately compiled kernels may contract the interpolant differently
            assert (np.abs(rs["min"][r] - v.min(axis=0)) <= 1e-14 * scale).all()
            assert (np.abs(rs["max"][r] - v.max(axis=0)) <= 1e-14 * scale).all()
        assert (np.abs(rs["mean"][r] - v.mean(axis=0)) <= 1e-13 * scale).all()          # summation order only
        assert (np.abs(rs["var"][r] - v.var(axis=0)) <= 1e-11 * scale ** 2).all()
    return a, b


def test_row_reduction_dopri5_vdp_ragged_ensemble():
    """ensemble mean/var/min/max of every dense-output row, reduced inside the integration kernel."""
    from assimulo_b200.gpu import Dopri5
    for N in (1, 31, 5000):
        y0, mu = cases.vdp_sweep(max(N, 2))
        _row_reduction_case(Dopri5, dict(model="vdp", y0=y0[:N], p0=mu[:N]), 2.0, 50, rtol=1e-6, atol=1e-6)
    _row_reduction_case(Dopri5, dict(model="vdp", y0=y0, p0=mu), 2.0, 50, rtol=1e-6, atol=1e-6, arith="strict")
    # lanes refilled one by one / in small groups: the warp's lanes are at unrelated rows all the time
    for thr in (1, 8):
        _row_reduction_case(Dopri5, dict(model="vdp", y0=y0, p0=mu), 2.0, 50, rtol=1e-6, atol=1e-6, refill_threshold=thr,
                            order="two-ended")
    rng = np.random.default_rng(5)
    _row_reduction_case(Dopri5, dict(model="vdp", y0=y0, p0=mu), 2.0, 7, rtol=1e-6, atol=1e-6, order=rng.permutation(5000))


@pytest.mark.parametrize("name", ["ExplicitEuler", "RungeKutta34", "Dopri5", "Radau5ODE", "RodasODE"])
def test_row_reduction_with_state_events_every_solver(name):
    """bouncing ball: lanes of one warp reach the rows in different steps and from both emission sites (event branch and
    plain step), which is what the per-warp lock of aens_reduce_grid is for"""
    import assimulo_b200.gpu as G
    y0, p, sw0 = cases.ball_sweep(3001)
    opts = dict(h=1e-3) if name == "ExplicitEuler" else dict(rtol=1e-6, atol=1e-6)
    a, b = _row_reduction_case(getattr(G, name), dict(model="ball", y0=y0, p0=p, sw0=sw0), 3.0, 64, **opts)
    assert a.statistics["nstateevents"] > 3001


def test_row_reduction_gyro_twelve_states_and_failed_members():
    from assimulo_b200.gpu import Dopri5
    y0, p = cases.gyro_sweep(4099)
    _row_reduction_case(Dopri5, dict(model="gyro", y0=y0, p0=p), 0.1, 100, rtol=1e-8, atol=1e-8)
    _row_reduction_case(Dopri5, dict(model="gyro", y0=y0[:999], p0=p[:999]), 0.1, 100, rtol=1e-8, atol=1e-8, arith="strict")
    # members that stop early (maxsteps) contribute to the rows they reached only
    y0, mu = cases.vdp_sweep(5000)
    a, b = _row_reduction_case(Dopri5, dict(model="vdp", y0=y0, p0=mu), 2.0, 40, maxsteps=60)
    assert 0 < b.n_not_complete < 5000
    c = b.row_statistics["count"]
    assert c[0] == 5000 and c[-1] == 5000 - b.n_not_complete and (np.diff(c) <= 0).all()


def test_row_reduction_nvrtc_model_and_mode_switching():
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5
    y0, mu = cases.vdp_sweep(777)
    prob = dict(source=VDP_SRC, n=2, n_p=1, has_jac=True, y0=y0, p0=mu)
    _row_reduction_case(Dopri5, prob, 2.0, 25, rtol=1e-6, atol=1e-6)
    # one solver object: stored rows, then reduced rows, then stored rows again
    sim = Dopri5(Batched_Explicit_Problem(model="vdp", y0=y0, p0=mu))
    t, y = sim.simulate(1.0, 10)
    assert np.asarray(y).shape == (11, 777, 2) and sim.row_statistics is None
    sim.reduce_rows = True
    sim.simulate(2.0, 10)
    assert sim.row_statistics["mean"].shape == (10, 2) and (sim.row_statistics["count"] == 777).all()
    sim.reduce_rows = False
    t, y = sim.simulate(3.0, 10)
    assert np.asarray(y).shape == (11, 777, 2) and sim.row_statistics is None


def test_first_passage_time_statistics_on_device():
    """aens_reduce_end_times: members stopped by handle_event at their first event (TerminateSimulation) -- the
    statistics of their end times are first-passage-time statistics; y' = 1, event at y = p => passage time p."""
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5
    N = 3000
    p = np.linspace(0.1, 2.9, N)[:, None]
    p[::7] = 5.0                                   # these never get there before tfinal = 3
    sim = Dopri5(Batched_Explicit_Problem(source=TERMINATE_SRC, n=1, n_p=1, n_g=1, n_sw=1, y0=np.zeros((N, 1)), p0=p,
                                          sw0=np.ones((N, 1), dtype=bool)))
    sim.simulate(3.0)
    hit = p[:, 0] < 3.0
    assert np.array_equal(sim.status == 1, hit)
    fp = sim.end_time_statistics(status=1)
    assert fp["count"] == hit.sum()
    assert abs(fp["mean"] - p[hit, 0].mean()) <= 1e-9 and abs(fp["var"] - p[hit, 0].var()) <= 1e-9
    assert abs(fp["min"] - p[hit, 0].min()) <= 1e-9 and abs(fp["max"] - p[hit, 0].max()) <= 1e-9
    done = sim.end_time_statistics(status=0)
    assert done["count"] == N - hit.sum() and done["min"] == done["max"] == 3.0
    none = sim.end_time_statistics(status=-2)
    assert none["count"] == 0 and np.isnan(none["mean"])


def test_fused_step_times_rhs_hook_of_user_models():
    """AENS_MODEL_RHS_H: a user model may supply hk = h * f(t, y) (the built-in Van der Pol model does); the fast Dopri5
    build consumes it in its stages, the strict build never calls it.  Same expressions as the built-in model =>
    identical doubles; and the fused form stays within the fast build's tolerance of the unfused one."""
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5
    src = VDP_SRC + r"""
    __device__ void aens_rhs_h(double h, double t, const double* y, const unsigned char* sw, const double* p, double* hk) {
        const double hmu = h * p[0];
        hk[0] = h * y[1];
        hk[1] = hmu * ((1. - y[0] * y[0]) * y[1] - y[0]);
    }
    """
    y0, mu = cases.vdp_sweep(2000)
    res = {}
    for name, prob in (("builtin", dict(model="vdp")),
                       ("fused", dict(source=src, n=2, n_p=1, has_jac=True, fused_rhs=True)),
                       ("plain", dict(source=VDP_SRC, n=2, n_p=1, has_jac=True))):
        sim = Dopri5(Batched_Explicit_Problem(y0=y0, p0=mu, **prob))
        sim.rtol, sim.atol = 1e-6, 1e-6
        sim.simulate(2.0)
        res[name] = (sim.y.copy(), sim.statistics["nsteps"])
    assert np.array_equal(res["fused"][0], res["builtin"][0]) and res["fused"][1] == res["builtin"][1]
    assert rel_err(res["plain"][0], res["builtin"][0]) <= 1e-7
    assert abs(res["plain"][1] - res["builtin"][1]) <= 0.001 * res["builtin"][1]
    # strict arithmetic ignores the hook: reference operation order
    st = []
    for prob in (dict(model="vdp"), dict(source=src, n=2, n_p=1, has_jac=True, fused_rhs=True)):
        sim = Dopri5(Batched_Explicit_Problem(y0=y0, p0=mu, **prob))
        sim.arith = "strict"
        sim.simulate(2.0)
        st.append(sim.y.copy())
    o = O.simulate("vdp", "Dopri5", y0, mu, tf=2.0, rtol=1e-6, atol=1e-6)
    assert np.array_equal(st[0], st[1]) and rel_err(st[0], o["y"]) <= 1e-9   # pow ulps, see DESIGN section 7


def test_graft_entry_smoke_runs():
    """the driver's round-end smoke check must keep working when the solver classes change"""
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    g = importlib.import_module("__graft_entry__")
    g.smoke()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_shared_initial_state_row_equals_the_expanded_array(mode):
    """aens_set_shared_y0: a parameter sweep gives y0 as ONE row; results are identical to passing [N, n] -- through the
    solver class (host pipeline: zero-copy and chunked staging; set_ensemble path with an output grid), reset() and
    re_init() included."""
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5, _core
    N = 3001
    _, mu = cases.vdp_sweep(N)
    mu_p = _core.pinned_empty((N, 1))
    mu_p[:] = mu
    full = Dopri5(Batched_Explicit_Problem(model="vdp", y0=np.tile([2.0, -0.6], (N, 1)), p0=mu_p, detect_shared_y0=False))
    row = Dopri5(Batched_Explicit_Problem(model="vdp", y0=[2.0, -0.6], p0=mu_p))
    assert row.problem.y0_row is not None and full.problem.y0_row is None and row.y.shape == (N, 2)
    for s in (full, row):
        s.options["reuse_buffers"] = True      # pinned result buffers: zero-copy mode is possible
    row.host_mode = ("auto", "staged", "zero-copy")[mode]   # (the expanded array of `full` is pageable: automatic mode)
    yf = np.array(full.simulate(2.0)[1][-1])
    yr = np.array(row.simulate(2.0)[1][-1])
    assert np.array_equal(yf, yr) and full.statistics["nsteps"] == row.statistics["nsteps"]
    row.reset()
    assert row._y_row is not None and np.array_equal(row.y, np.tile([2.0, -0.6], (N, 1)))
    assert np.array_equal(np.array(row.simulate(2.0)[1][-1]), yf)
    # with an output grid the ensemble goes through aens_set_ensemble: the row is expanded on the device
    row.reset(); full.reset()
    tr, dr = row.simulate(2.0, 4)
    tf_, df = full.simulate(2.0, 4)
    assert np.array_equal(np.array(dr)[1:], np.array(df)[1:])
    # re_init with another shared row, and with a full array
    row.re_init(0.0, [1.0, 0.0]); full.re_init(0.0, np.tile([1.0, 0.0], (N, 1)))
    assert np.array_equal(np.array(row.simulate(1.0)[1][-1]), np.array(full.simulate(1.0)[1][-1]))
    # raw handle: 1-D y0 needs something that fixes N
    h = _core.Handle(0)
    h.set_model_builtin("linear")
    with pytest.raises(ValueError):
        h.set_ensemble(np.array([1.0]), None)
    h.set_ensemble(np.array([4.0]), np.tile([-1.0, 0.0, 0.0], (5, 1)))
    assert h.N == 5


def test_kernel_time_history_of_queued_launches():
    """aens_kernel_ms_history: every launch records its own event pair, so launches queued back to back with
    simulate_async can be timed individually afterwards (bench.py's roofline uses the launches of its timed region)."""
    from assimulo_b200.gpu import _core
    y0, mu = cases.vdp_sweep(1 << 16)
    h = _core.Handle(0)
    h.set_model_builtin("vdp")
    h.set_ensemble(y0, mu, None, 0.0)
    h.set_opts({"solver": _core.SOLVER_IDS["Dopri5"], "rtol": 1e-6, "atol": 1e-6})
    assert h.kernel_ms_history() == []
    for _ in range(3):
        h.reset()
        h.simulate_async(2.0)
    h.sync()
    ms = h.kernel_ms_history()
    assert len(ms) == 3 and all(m > 0 for m in ms) and ms[0] == h.last_kernel_ms()
    assert max(ms) < 3 * min(ms)
    assert len(h.kernel_ms_history(2)) == 2


def test_unreached_rows_are_nan_and_survive_a_continuation_call():
    """Rows of the output grid a member has not reached are NaN (written by the member itself when it is stored, no
    pre-fill of the buffer); a second aens_simulate on the same grid continues and fills them, keeping the earlier rows."""
    from assimulo_b200.gpu import _core
    N = 1000
    y0, mu = cases.vdp_sweep(N)
    grid = np.linspace(0.0, 2.0, 9)[1:]
    h = _core.Handle(0)
    h.set_model_builtin("vdp")
    h.set_ensemble(y0, mu, None, 0.0)
    h.set_opts({"solver": _core.SOLVER_IDS["Dopri5"], "rtol": 1e-6, "atol": 1e-6})
    h.set_output(grid, 16)
    h.simulate(1.0)
    d1 = h.get_dense().copy()
    assert np.isfinite(d1[:4]).all() and np.isnan(d1[4:]).all()
    h.simulate(2.0)
    d2 = h.get_dense()
    assert np.isfinite(d2).all() and np.array_equal(d2[:4], d1[:4])
    o = O.simulate("vdp", "Dopri5", y0, mu, tf=2.0, t_grid=grid, rtol=1e-6, atol=1e-6)
    assert rel_err(np.transpose(d2, (1, 0, 2)), o["dense"]) <= 1e-4   # the restart at t = 1 changes the step sequence
    # members that fail early leave NaN rows behind
    h.reset()
    h.set_opts({"solver": _core.SOLVER_IDS["Dopri5"], "rtol": 1e-6, "atol": 1e-6, "maxsteps": 40})
    h.set_output(grid, 16)
    h.simulate(2.0)
    t, _, _, st = h.get_final()
    d3 = h.get_dense()
    assert (st != 0).any() and (st == 0).any()
    for k, tk in enumerate(grid):
        assert np.array_equal(np.isfinite(d3[k][:, 0]), t >= tk)
