"""-m gpu, round 2: the end of the path (the all-gather inside the C ABI), the host-array modes of simulate(), the
time limit, and the edge cases the round-1 review found (grid rows at the segment start in row-reduction mode,
continuation in row-reduction mode, unwritten event-log slots, fetch orders that are not permutations)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases
import ens_oracle as O
from conftest import ROOT, rel_err

pytestmark = pytest.mark.gpu


def _handle(model, solver, y0, p, sw0=None, **opts):
    from assimulo_b200.gpu import _core
    h = _core.Handle(0)
    h.set_model_builtin(model)
    h.set_ensemble(y0, p, sw0, 0.0)
    d = {"solver": _core.SOLVER_IDS[solver]}
    d.update(opts)
    h.set_opts(d)
    return h


# ---------------------------------------------------------------------------------------------------------
# grid rows at / before the segment start (C ABI: nothing restricts t_out[0] > t0)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("solver,opts", [("Dopri5", dict(rtol=1e-6, atol=1e-6)),
                                         ("Radau5ODE", dict(rtol=1e-6, atol=1e-6, usejac=1)),
                                         ("RodasODE", dict(rtol=1e-6, atol=1e-6, usejac=1))])
def test_grid_row_at_t0_stored_and_reduced(solver, opts):
    """t_out[0] == t0: the d1 kernels store y0 in that row (first SOLOUT call), the d2 kernels -- which have no row
    buffer at all -- must feed it to the reduction instead of storing through a NULL pointer."""
    N = 777
    y0, mu = cases.vdp_sweep(N)
    mu = np.minimum(mu, 5.0)
    grid = np.linspace(0.0, 1.0, 6)                       # includes t0 = 0
    h = _handle("vdp", solver, y0, mu, arith=0, **opts)
    h.set_output(grid, 0)
    h.simulate(1.0)
    rows = h.get_dense()                                  # [6, N, 2]
    assert np.array_equal(rows[0], y0) and np.isfinite(rows).all()
    _, y_final, _, st = h.get_final()
    assert (st == 0).all() and np.allclose(rows[-1], y_final, rtol=1e-13, atol=1e-15)   # interpolant at theta = 1
    h.close()
    h = _handle("vdp", solver, y0, mu, arith=0, **opts)
    h.set_row_reduction(True)
    h.set_output(grid, 0)
    h.simulate(1.0)
    rs = h.get_row_reduction()
    assert (rs["count"] == N).all()
    assert np.array_equal(rs["min"], rows.min(axis=1)) and np.array_equal(rs["max"], rows.max(axis=1))
    assert np.allclose(rs["mean"], rows.mean(axis=1), rtol=1e-12, atol=1e-13)
    assert np.array_equal(rs["mean"][0], y0[0])           # the row at t0 is the initial state itself
    h.close()


def test_row_reduction_survives_a_continuation_call():
    """A second aens_simulate on the same grid continues where the rows stopped -- also when they are reduced: the
    partial records of the first call must not be zeroed by the second launch."""
    N = 2100
    y0, mu = cases.vdp_sweep(N)
    grid = np.linspace(0.0, 2.0, 41)[1:]
    one = _handle("vdp", "Dopri5", y0, mu, arith=0, rtol=1e-6, atol=1e-6)
    one.set_row_reduction(True)
    one.set_output(grid, 0)
    one.simulate(2.0)
    a = one.get_row_reduction()
    two = _handle("vdp", "Dopri5", y0, mu, arith=0, rtol=1e-6, atol=1e-6)
    two.set_row_reduction(True)
    two.set_output(grid, 0)
    two.simulate(0.83)
    mid = two.get_row_reduction()
    assert (mid["count"][grid <= 0.83] == N).all() and (mid["count"][grid > 0.83] == 0).all()
    two.simulate(2.0)
    b = two.get_row_reduction()
    assert (b["count"] == N).all()
    # rows of the first call: the very same integration, except for the last steps before 0.83, which the first call
    # clips to its own end point (dopri5.f:411-414) -- rows well before that are identical
    early = grid <= 0.3
    assert early.sum() >= 5
    assert np.array_equal(b["min"][early], a["min"][early]) and np.array_equal(b["max"][early], a["max"][early])
    assert np.allclose(b["mean"][early], a["mean"][early], rtol=1e-13, atol=1e-14)
    # rows of the second call come from a restarted solver (HINIT again at 0.83): close, not identical
    assert np.allclose(b["mean"], a["mean"], rtol=1e-4, atol=1e-5)
    # a reset starts the records afresh
    two.reset()
    two.simulate(2.0)
    c = two.get_row_reduction()
    assert (c["count"] == N).all() and np.array_equal(c["min"], a["min"])
    one.close()
    two.close()


def test_continuation_after_partial_termination_with_rows():
    """simulate() with some members terminated sets solver.t = min(t_members); the next simulate(ncp) then hands the
    members that did complete a grid with rows <= their own t -- stored (d1) and reduced (d2) alike."""
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5
    from test_gpu_api import TERMINATE_SRC
    thr = np.array([[0.5], [50.0], [50.0], [0.75]])
    kw = dict(source=TERMINATE_SRC, n=1, n_p=1, n_g=1, n_sw=1, y0=[[0.0]] * 4, p0=thr, sw0=[[1]] * 4)
    out = []
    for red in (False, True):
        sim = Dopri5(Batched_Explicit_Problem(**kw))
        sim.simulate(2.0)
        assert sim.status.tolist() == [1, 0, 0, 1] and abs(sim.t - 0.5) < 1e-9
        sim.reduce_rows = red
        t, y = sim.simulate(3.0, 5)                       # grid from 0.5: rows at 1.0, 1.5, 2.0 precede t = 2 of members 1, 2
        assert np.isfinite(sim.y).all()
        out.append((sim, t, y))
    (a, ta, ya), (b, tb, yb) = out
    assert rel_err(b.y, a.y) <= 1e-12
    rs = b.row_statistics
    rows = np.asarray(ya)[1:]
    reached = ~np.isnan(rows[:, :, 0])
    assert np.array_equal(rs["count"], reached.sum(axis=1)) and rs["count"].max() > 0
    for r in range(rows.shape[0]):
        v = rows[r][reached[r]]
        if len(v):
            assert np.allclose(rs["min"][r], v.min(axis=0), rtol=1e-14) and np.allclose(rs["max"][r], v.max(axis=0), rtol=1e-14)


# ---------------------------------------------------------------------------------------------------------
# event log, schedule validation
# ---------------------------------------------------------------------------------------------------------
def test_event_log_slots_beyond_the_count_read_nan():
    N = 300
    y0, p, sw0 = cases.ball_sweep(N)
    h = _handle("ball", "Dopri5", y0, p, sw0, rtol=1e-6, atol=1e-6)
    h.set_output(None, 16)
    cnt, te, info = h.get_events()                         # before any launch: nothing written, nothing but NaN / 0
    assert (cnt == 0).all() and np.isnan(te).all() and (info == 0).all()
    h.simulate(1.5)
    cnt, te, info = h.get_events()
    assert cnt.max() < 16 and cnt.min() >= 1
    used = np.arange(16)[None, :] < cnt[:, None]
    assert np.isfinite(te[used]).all() and np.isnan(te[~used]).all() and (info[~used] == 0).all()
    h.reset()                                              # back to t0: the log is empty again
    _, te, info = h.get_events()
    assert np.isnan(te).all() and (info == 0).all()
    h.close()


def test_fetch_order_must_be_a_permutation():
    import ctypes as C
    N = 64
    y0, mu = cases.vdp_sweep(N)
    h = _handle("vdp", "Dopri5", y0, mu)
    for bad in (np.zeros(N, dtype=np.intc), np.arange(N, dtype=np.intc) + 1, np.arange(N - 1, dtype=np.intc)):
        with pytest.raises(ValueError):
            h.set_schedule(32, bad)
    h.set_schedule(32, np.arange(N, dtype=np.intc)[::-1].copy())
    h.simulate(2.0)
    assert (h.get_final()[3] == 0).all()
    h.close()
    # the C ABI itself refuses it too (a caller without the Cython layer)
    lib = C.CDLL(os.path.join(ROOT, "assimulo_b200", "lib", "libassimulo_cuda.so"))
    lib.aens_last_error.restype = C.c_char_p
    hh = C.c_void_p()
    assert lib.aens_create(0, C.byref(hh)) == 0
    assert lib.aens_set_model_builtin(hh, b"vdp") == 0
    y = np.ascontiguousarray(y0)
    m = np.ascontiguousarray(mu)
    lib.aens_set_ensemble.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
    assert lib.aens_set_ensemble(hh, N, y.ctypes.data, m.ctypes.data, None, 0.0) == 0
    dup = np.zeros(N, dtype=np.intc)
    assert lib.aens_set_schedule(hh, 32, dup.ctypes.data_as(C.c_void_p)) == -1
    assert b"permutation" in lib.aens_last_error(hh)
    oob = np.arange(N, dtype=np.intc)
    oob[5] = N
    assert lib.aens_set_schedule(hh, 32, oob.ctypes.data_as(C.c_void_p)) == -1
    lib.aens_destroy(C.byref(hh))


# ---------------------------------------------------------------------------------------------------------
# host-array modes of simulate(): zero-copy, staged, and the two hybrids give the same members
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shared", [False, True])
def test_host_modes_agree(shared):
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5, _core
    N = 70000 + 13
    y0, mu = cases.vdp_sweep(N)
    y0_h, p_h = _core.pinned_empty((N, 2)), _core.pinned_empty((N, 1))
    y0_h[:], p_h[:] = y0, mu
    o = O.simulate("vdp", "Dopri5", y0[::97], mu[::97], tf=2.0, rtol=1e-6, atol=1e-6)
    ref = None
    for mode in ("zero-copy", "staged", "zero-copy-in", "zero-copy-out", "auto"):
        prob = Batched_Explicit_Problem(model="vdp", y0=y0_h, p0=p_h, detect_shared_y0=shared)
        assert (prob.y0_row is not None) == shared
        sim = Dopri5(prob)
        sim.options["reuse_buffers"] = True
        sim.options["host_chunks"] = 5
        sim.host_mode = mode
        sim.order = "two-ended"
        sim.simulate(2.0)
        assert sim.n_not_complete == 0
        y = np.array(sim.y)
        if ref is None:
            ref = (y, sim.statistics["nsteps"], sim.statistics.per_member("nsteps"))
            assert rel_err(y[::97], o["y"]) <= 1e-5
        else:
            assert np.array_equal(y, ref[0]) and sim.statistics["nsteps"] == ref[1]
            assert np.array_equal(sim.statistics.per_member("nsteps"), ref[2])
        # a second pass on the same solver object (the bench's e2e loop): reset + simulate
        sim.reset()
        sim.simulate(2.0)
        assert np.array_equal(np.array(sim.y), ref[0])


def test_host_modes_with_events_and_switches():
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5, _core
    N = 4099
    y0, p, sw0 = cases.ball_sweep(N)
    arrs = [_core.pinned_empty(a.shape, a.dtype) for a in (y0, p, sw0)]
    for d, s_ in zip(arrs, (y0, p, sw0)):
        d[:] = s_
    ref = None
    for mode in ("zero-copy", "staged", "zero-copy-in", "zero-copy-out"):
        sim = Dopri5(Batched_Explicit_Problem(model="ball", y0=arrs[0], p0=arrs[1], sw0=arrs[2]))
        sim.host_mode = mode
        sim.options["host_chunks"] = 3
        sim.simulate(3.0)
        cnt, te, info = sim.event_log
        cur = (np.array(sim.y), np.array(sim.sw), cnt.copy(), np.nan_to_num(te), info.copy())
        if ref is None:
            ref = cur
            o = O.simulate("ball", "Dopri5", y0, p, sw0, tf=3.0, max_ev=16, rtol=1e-6, atol=1e-6)
            assert np.array_equal(cnt, o["stats"]["nstateevents"])
        else:
            for a, b in zip(cur, ref):
                assert np.array_equal(a, b), mode


# ---------------------------------------------------------------------------------------------------------
# time limit (src/ode.pyx:372-392, src/explicit_ode.pyx:290-292)
# ---------------------------------------------------------------------------------------------------------
def test_time_limit_leaves_late_members_untouched_and_raises():
    from assimulo_b200.gpu import Batched_Explicit_Problem, RungeKutta4, TimeLimitExceeded
    N = 1 << 20
    y0, mu = cases.vdp_sweep(N)
    sim = RungeKutta4(Batched_Explicit_Problem(model="vdp", y0=y0, p0=mu))
    sim.h = 1e-4                                            # 20000 steps per member: tens of milliseconds for the ensemble
    sim.time_limit = 2e-3
    with pytest.raises(TimeLimitExceeded) as e:
        sim.simulate(2.0)
    assert "time limit was exceeded" in str(e.value)
    st = sim.status
    late = st == -9
    assert late.any() and (st[~late] == 0).all() and (~late).sum() >= 148 * 128     # the first wave always runs
    assert np.array_equal(sim.y[late], y0[late]) and (sim.t_members[late] == 0.0).all()
    assert (sim.t_members[~late] == 2.0).all()
    assert sim.n_not_complete == int(late.sum())
    # without a limit (and with a generous one) everything completes
    for lim in (0, 600):
        sim2 = RungeKutta4(Batched_Explicit_Problem(model="vdp", y0=y0[:5000], p0=mu[:5000]))
        sim2.h = 1e-2
        sim2.time_limit = lim
        sim2.simulate(2.0)
        assert (sim2.status == 0).all()


# ---------------------------------------------------------------------------------------------------------
# the path's one collective, inside the C ABI
# ---------------------------------------------------------------------------------------------------------
def test_allgather_single_rank_communicator_packs_and_unpacks():
    """aens_comm_init with one rank: the same pack -> ncclAllGather -> unpack code as at 8 GPUs, on any box."""
    from assimulo_b200.gpu import _core
    N = 10007
    y0, p, sw0 = cases.ball_sweep(N)
    h = _handle("ball", "Dopri5", y0, p, sw0, rtol=1e-6, atol=1e-6)
    h.set_output(None, 8)
    h.simulate(3.0)
    h.comm_init(0, 1, _core.comm_unique_id())
    h.allgather_results(0)
    g = h.get_gathered(N, want_events=True)
    assert h.last_gather_ms() > 0.0
    t, y, sw, st = h.get_final()
    cnt, te, info = h.get_events()
    assert np.array_equal(g["t"], t) and np.array_equal(g["y"], y) and np.array_equal(g["status"], st)
    for i, k in enumerate(_core.STAT_NAMES):
        assert np.array_equal(g["stats"][i], h.get_stats(k)), k
    assert np.array_equal(np.nan_to_num(g["ev_t"]), np.nan_to_num(te)) and np.array_equal(g["ev_info"], info)
    # a prefix and a field subset
    h.allgather_results(_core.GATHER_Y | _core.GATHER_STATUS)
    g2 = h.get_gathered(100, want_t=False, want_stats=False)
    assert np.array_equal(g2["y"], y[:100]) and np.array_equal(g2["status"], st[:100])
    with pytest.raises(_core.AensError):
        h.get_gathered(100, want_t=True, want_stats=False)          # t was not part of that gather
    # the library issues the exchange itself, right behind the last kernel of every simulate
    h.set_gather_in_simulate(_core.GATHER_T | _core.GATHER_Y)
    h.reset()
    h.simulate(1.5)
    g3 = h.get_gathered(N, want_status=False, want_stats=False)
    t3, y3, _, _ = h.get_final()
    assert np.array_equal(g3["y"], y3) and np.array_equal(g3["t"], t3) and not np.array_equal(y3, y)
    h.set_gather_in_simulate(0)
    h.comm_destroy()
    with pytest.raises(_core.AensError):
        h.allgather_results(0)
    h.close()
    # ... also in the one-call host modes (chunked: the exchange overlaps the device->host copies of the last chunks)
    y0v, mu = cases.vdp_sweep(N)
    for mode in (1, 2, 3):
        hv = _core.Handle(0)
        hv.set_model_builtin("vdp")
        hv.set_opts({"solver": _core.SOLVER_IDS["Dopri5"]})
        hv.set_host_mode(mode)
        hv.comm_init(0, 1, _core.comm_unique_id())
        hv.set_gather_in_simulate(_core.GATHER_Y | _core.GATHER_STATUS)
        y0_h, p_h, out = _core.pinned_empty((N, 2)), _core.pinned_empty((N, 1)), _core.pinned_empty((N, 2))
        y0_h[:], p_h[:] = y0v, mu
        hv.simulate_host(y0_h, p_h, None, 0.0, 2.0, out_y=out, nchunks=4)
        gv = hv.get_gathered(N, want_t=False, want_stats=False)
        assert np.array_equal(gv["y"], out) and (gv["status"] == 0).all()
        hv.close()


def test_allgather_over_nccl_on_two_gpus():
    """world_size 2 on real hardware: torchrun tools/nccl_gather_check.py (ball ensemble with state events, both
    ranks compare the gathered whole with the CPU oracle)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(29700 + os.getpid() % 200),
                        os.path.join(ROOT, "tools", "nccl_gather_check.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert r.stdout.count("OK") >= 2


def test_hostlink_probe_reports_plausible_rates():
    from assimulo_b200.gpu import _core
    for mode in range(6):
        gbs = _core.measure_hostlink(0, 1 << 26, mode, 0.05)
        assert 1.0 < gbs < 200.0, (mode, gbs)


# ---------------------------------------------------------------------------------------------------------
# sub-warp-per-member Dopri5 (csrc/aens_split.cuh): 4 lanes x 3 states of the gyroscope
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [1, 7, 8, 33, 5000])
def test_split_dopri5_gyro_matches_oracle_and_thread_per_member(N):
    from assimulo_b200.gpu import _core
    y0, p = cases.gyro_sweep(max(N, 2))
    y0, p = y0[:N], p[:N]
    o = O.simulate("gyro", "Dopri5", y0, p, tf=0.1, rtol=1e-8, atol=1e-8)
    res = {}
    for lanes in (1, 4, 0):
        h = _handle("gyro", "Dopri5", y0, p, rtol=1e-8, atol=1e-8)
        h.set_lanes_per_member(lanes)
        h.simulate(0.1)
        assert h.last_lanes_per_member() == (4 if lanes == 4 else 1)      # automatic = one thread per member (the faster one)
        t, y, _, st = h.get_final()
        res[lanes] = (y, h.get_stats("nsteps"), h.get_stats("nfcns"), h.get_stats("nerrfails"))
        assert (st == 0).all() and (t == 0.1).all()
        assert rel_err(y, o["y"]) <= 1e-7                                    # 10 x rtol
        assert abs(int(res[lanes][1].sum()) - int(o["stats"]["nsteps"].sum())) <= max(1, 0.02 * o["stats"]["nsteps"].sum())
        h.close()
    # same arithmetic per component, only the summation order of the norms differs: the two kernels agree far below rtol
    assert rel_err(res[4][0], res[1][0]) <= 1e-10
    assert np.array_equal(res[0][0], res[1][0])


def test_split_dopri5_gyro_dense_rows_order_and_continuation():
    from assimulo_b200.gpu import _core
    N = 1000
    y0, p = cases.gyro_sweep(N)
    grid = np.linspace(0.0, 0.1, 21)
    o = O.simulate("gyro", "Dopri5", y0, p, tf=0.1, t_grid=grid[1:], rtol=1e-8, atol=1e-8)
    rows = {}
    for lanes in (1, 4):
        h = _handle("gyro", "Dopri5", y0, p, rtol=1e-8, atol=1e-8)
        h.set_lanes_per_member(lanes)
        h.set_output(grid, 0)                                # includes t0: the row of the initial state
        h.set_schedule(32, np.random.RandomState(3).permutation(N).astype(np.intc))
        h.simulate(0.04)
        mid = h.get_dense()
        assert np.isfinite(mid[grid <= 0.04]).all() and np.isnan(mid[grid > 0.04]).all()
        h.simulate(0.1)                                      # continuation: the remaining rows
        rows[lanes] = h.get_dense()
        assert np.array_equal(rows[lanes][0], y0) and np.isfinite(rows[lanes]).all()
        assert h.last_lanes_per_member() == lanes
        # orthogonality of the rotation matrix and |pi| in every row (size-independent properties of the model)
        Q = rows[lanes][:, :, 3:].reshape(len(grid), N, 3, 3)
        assert np.max(np.abs(np.einsum("tnij,tnkj->tnik", Q, Q) - np.eye(3))) < 1e-6
        h.close()
    assert np.nanmax(np.abs(np.transpose(rows[4][1:], (1, 0, 2)) - o["dense"]) / np.maximum(np.abs(o["dense"]), 1.0)) <= 1e-6
    assert rel_err(rows[4][:9], rows[1][:9]) <= 1e-10       # rows of the first call: the same steps in both kernels


def test_split_kernel_selection_rules():
    from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5, Radau5ODE, _core
    y0, p = cases.gyro_sweep(64)
    h = _handle("gyro", "Dopri5", y0, p, rtol=1e-8, atol=1e-8, arith=0)     # strict arithmetic: one thread per member
    h.simulate(0.1)
    assert h.last_lanes_per_member() == 1
    h.set_lanes_per_member(4)
    with pytest.raises(_core.AensError):
        h.simulate(0.1)
    h.close()
    y0v, mu = cases.vdp_sweep(64)
    h = _handle("vdp", "Dopri5", y0v, mu)
    h.set_lanes_per_member(4)
    with pytest.raises(_core.AensError):                                      # the model has no cooperative form
        h.simulate(1.0)
    h.close()
    sim = Dopri5(Batched_Explicit_Problem(model="gyro", y0=y0, p0=p))
    sim.rtol = sim.atol = 1e-8
    sim.reduce_rows = True                                                    # row reduction: thread-per-member kernels
    sim.simulate(0.1, 4)
    assert sim._handle.last_lanes_per_member() == 1 and (sim.row_statistics["count"] == 64).all()
    sim2 = Radau5ODE(Batched_Explicit_Problem(model="gyro", y0=y0, p0=p))
    sim2.simulate(0.1)
    assert sim2._handle.last_lanes_per_member() == 1 and (sim2.status == 0).all()


# ---------------------------------------------------------------------------------------------------------
# backward integration (options["backward"]; POSNEG in dopri5.f:380, radau5.c:341, rodas_decsol.f:569)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("solver", ["Dopri5", "Radau5ODE", "RodasODE"])
def test_backward_integration_matches_oracle_and_reference_golden(solver):
    import assimulo_b200.gpu as G
    from conftest import golden
    cls = getattr(G, solver)
    g = golden("bwd_vdp_" + solver.lower())
    N = len(g["mu"])
    y0, mu = np.tile(g["y_start"], (N, 1)), g["mu"].reshape(-1, 1)
    for arith in ("strict", "fast"):
        sim = cls(G.Batched_Explicit_Problem(model="vdp", y0=y0, p0=mu, t0=2.0))
        sim.arith, sim.backward = arith, True
        sim.rtol = sim.atol = 1e-6
        t, y = sim.simulate(0.0, 8)
        assert np.allclose(t, g["t_rows"]) and sim.t == 0.0 and (sim.status == 0).all()
        rows = np.transpose(np.asarray(y), (1, 0, 2))                     # [N, 9, 2]
        if arith == "strict":
            # the time-reversed forward run mirrors the reference's POSNEG run operation for operation
            # (to the ulp-level difference between the device's and glibc's pow in the step-size controllers)
            assert np.max(np.abs(rows - g["y_rows"])) <= 1e-12
            assert np.array_equal(sim.statistics.per_member("nsteps"), g["nsteps_rows"])
        else:
            assert rel_err(rows, g["y_rows"]) <= 1e-5
            assert abs(sim.statistics["nsteps"] - g["nsteps_rows"].sum()) <= max(2, 0.02 * g["nsteps_rows"].sum())
        with pytest.raises(G.AssimuloException):
            sim.simulate(1.0)                                              # tfinal above the current time
    # a larger ensemble against the oracle, continuation 2 -> 1 -> 0, final states only
    N = 3000
    y0, mu = cases.vdp_sweep(N)
    mu = np.minimum(mu, 4.0)
    y0 = np.tile([1.2, -0.8], (N, 1))
    kw = dict(rtol=1e-6, atol=1e-6)
    o1 = O.simulate("vdp", solver, y0, mu, t0=2.0, tf=1.0, **(kw if solver == "Dopri5" else dict(kw, usejac=1)))
    o2 = O.simulate("vdp", solver, o1["y"], mu, t0=1.0, tf=0.0, **(kw if solver == "Dopri5" else dict(kw, usejac=1)))
    sim = cls(G.Batched_Explicit_Problem(model="vdp", y0=y0, p0=mu, t0=2.0))
    sim.arith, sim.backward = "strict", True
    sim.rtol = sim.atol = 1e-6
    sim.simulate(1.0)
    assert rel_err(sim.y, o1["y"]) <= 1e-11 and sim.statistics["nsteps"] == int(o1["stats"]["nsteps"].sum())
    sim.simulate(0.0)
    assert rel_err(sim.y, o2["y"]) <= 1e-10 and sim.statistics["nsteps"] == int(o2["stats"]["nsteps"].sum())


def test_backward_integration_state_events_and_user_source():
    import assimulo_b200.gpu as G
    from conftest import golden
    thr = np.array([0.5, 0.9, 5.0])
    p = np.stack([np.full(3, -1.0), np.zeros(3), thr], axis=1)
    for solver in ("Dopri5", "Radau5ODE", "RodasODE"):
        e = golden("bwd_linear_ev_" + solver.lower())
        sim = getattr(G, solver)(G.Batched_Explicit_Problem(model="linear_ev", y0=np.full((3, 1), 0.2), p0=p, sw0=np.ones((3, 1))))
        sim.arith, sim.backward = "strict", True
        sim.rtol = sim.atol = 1e-6
        sim.simulate(-2.0)
        cnt, te, info = sim.event_log
        assert np.array_equal(cnt, e["nev"]) and (sim.t_members == -2.0).all()
        m = e["nev"] > 0
        assert np.max(np.abs(te[m, 0] - e["ev_t"][m])) <= (1e-12 if solver == "Dopri5" else 1e-7)
        assert (info[m, 0] == 3).all() and np.max(np.abs(sim.y - e["y"])) <= (1e-12 if solver == "Dopri5" else 1e-6)
        assert np.array_equal(sim.sw[:, 0], ~m)                               # handle_event switched the indicator off
    # NVRTC user model with explicit time dependence: y' = -t  =>  y(t) = y0 - (t^2 - t0^2) / 2, backwards from t0 = 2 to -1
    src = """
__device__ void aens_rhs(double t, const double* y, const unsigned char* sw, const double* p, double* yd) { yd[0] = -t * p[0]; }
__device__ void aens_jac(double t, const double* y, const unsigned char* sw, const double* p, double* J) { J[0] = 0.0; }
"""
    c = np.linspace(0.5, 2.0, 40)[:, None]
    for solver in ("Dopri5", "Radau5ODE", "RodasODE"):
        sim = getattr(G, solver)(G.Batched_Explicit_Problem(source=src, n=1, n_p=1, has_jac=True, y0=np.ones((40, 1)), p0=c, t0=2.0))
        sim.backward = True
        sim.rtol = sim.atol = 1e-9
        t, y = sim.simulate(-1.0, 3)
        assert np.allclose(t, [2.0, 1.0, 0.0, -1.0])
        for k, tk in enumerate(t[:-1]):
            assert np.max(np.abs(np.asarray(y)[k][:, 0] - (1.0 - c[:, 0] * (tk * tk - 4.0) / 2.0))) < 1e-7, (solver, tk)
        # (the row AT tfinal is interpolated only by members whose last step lands on tfinal exactly -- x + (xend - x) can
        # miss xend by an ulp, SURVEY A.7, in either direction of time; the final state itself is always there)
        assert (sim.status == 0).all() and np.max(np.abs(sim.y[:, 0] - (1.0 - c[:, 0] * (1.0 - 4.0) / 2.0))) < 1e-7
        assert np.max(np.abs(sim.t_members + 1.0)) < 1e-14
    # what has no direction refuses
    for name in ("RungeKutta34", "RungeKutta4", "ExplicitEuler", "ImplicitEuler"):
        s_ = getattr(G, name)(G.Batched_Explicit_Problem(model="vdp", y0=[[2.0, -0.6]], p0=[[1.0]]))
        with pytest.raises(G.AssimuloException):
            s_.backward = True
    te_ = G.Dopri5(G.Batched_Explicit_Problem(model="time_ev", y0=[[0.0]]))
    with pytest.raises(G.AssimuloException):
        te_.backward = True
