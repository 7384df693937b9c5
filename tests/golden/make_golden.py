#!/usr/bin/env python
"""Generate golden vectors by running the REFERENCE ITSELF (oracle/_ref, built from /root/reference by
oracle/build_ref.py) on small subsets of BASELINE.json's configs.  Run in the build container only:

    python oracle/build_ref.py && python tests/golden/make_golden.py

Writes tests/golden/*.npz (committed).  The reference cannot travel to the GPU box; these files can.
Dopri5 runs the reference's runge_kutta.py::Dopri5 class over oracle/dopri5_c.c (our restatement of the
Fortran core, see that file's header) -- every other solver is 100 % reference code.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, HERE)

import cases  # noqa: E402
import ref_models  # noqa: E402

STAT_KEYS = ["nsteps", "nfcns", "nerrfails", "njacs", "nlus", "nstatefcns", "nstateevents"]


def _solver_cls(name):
    import assimulo.solvers as S
    return getattr(S, name)


def run_member(problem, solver, tf, ncp=0, report_continuously=False, **opts):
    """One reference simulate(); returns dict(t_final, y_final, stats, t_sol, y_sol, events, sw)."""
    sim = _solver_cls(solver)(problem)
    sim.verbosity = 50
    for k, v in opts.items():
        setattr(sim, k, v)
    if report_continuously:
        sim.report_continuously = True
    failed = ""
    try:
        t, y = sim.simulate(tf, ncp)
    except Exception as e:  # per-member failure (e.g. ImplicitEuler 'Newton iteration failed'): recorded, not fatal
        failed = str(e)
        t, y = list(sim.t_sol), list(sim.y_sol)
    stats = {}
    for k in STAT_KEYS:
        try:
            v = sim.statistics[k]
        except KeyError:
            v = -1
        stats[k] = v
    if solver in ("RungeKutta4", "ExplicitEuler") and False:
        pass
    if solver in ("RungeKutta4", "ExplicitEuler"):
        # these keep no counters: accepted steps = len(t_sol)-1 without events
        stats["nsteps"] = len(t) - 1
    extra = {}
    for k in ("nniters", "nnfails"):
        try:
            extra[k] = sim.statistics[k]
        except KeyError:
            extra[k] = -1
    return dict(failed=failed, extra=extra, t_final=float(sim.t), y_final=np.array(sim.y, dtype=float).copy(), stats=stats,
                t_sol=np.array(t), y_sol=np.array(y), events=list(getattr(problem, "events", [])),
                sw=list(sim.sw) if sim.sw is not None else [])


def pack(results, max_ev=24):
    n = len(results[0]["y_final"])
    N = len(results)
    out = {"y": np.zeros((N, n)), "t": np.zeros(N)}
    for k in STAT_KEYS:
        out[k] = np.array([r["stats"][k] for r in results], dtype=np.int64)
    ev_t = np.full((N, max_ev), np.nan)
    ev_info = np.zeros((N, max_ev), dtype=np.int64)
    nev = np.zeros(N, dtype=np.int64)
    for i, r in enumerate(results):
        out["y"][i] = r["y_final"]
        out["t"][i] = r["t_final"]
        nev[i] = len(r["events"])
        for j, (te, info) in enumerate(r["events"][:max_ev]):
            ev_t[i, j] = te
            m = 0
            if isinstance(info, str):  # time event
                m = 1 << 30
            else:
                for c, v in enumerate(info):
                    if v != 0:
                        m |= (1 << (2 * c)) | ((1 if v > 0 else 0) << (2 * c + 1))
            ev_info[i, j] = m
    out["ev_t"], out["ev_info"], out["nev"] = ev_t, ev_info, nev
    return out


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in arrays.items()})


def time_event_cases():
    """Time events (src/explicit_ode.pyx:212-216,233-264): the reference's test_time_event model with every solver,
    plus a variant whose last event sits exactly on tfinal and one with an event 2e-16 before it
    (tests/conftest.py ExplicitTimeEventCloseToFinalTime)."""
    for solver, opts in (("Dopri5", {}), ("RungeKutta34", {}), ("RungeKutta4", dict(h=0.01)),
                         ("ExplicitEuler", dict(h=0.01)), ("Radau5ODE", {})):
        # RungeKutta34 returns ID_PY_OK instead of ID_PY_COMPLETE from the last step of a segment unless state events
        # or report_continuously are active (runge_kutta.py:657-667), so _simulate never sees its time events (probe:
        # 0 events, y(5) = 5): its vectors are generated with report_continuously=True, the mode in which the
        # reference does what it intends.
        rc = dict(report_continuously=True) if solver == "RungeKutta34" else {}
        res = [run_member(ref_models.time_ev(), solver, 5.0, **rc, **opts),
               run_member(ref_models.time_ev((0.5, 1.5, 2.5, 3.0), 0.25, -2.0, 1.0), solver, 3.0, **rc, **opts),
               run_member(ref_models.time_ev((1.0, 2.0, 2.5, 3.0 - 6e-16), 1.0, 1.0, 0.0), solver, 3.0, **rc, **opts)]
        for r in res:
            r["stats"]["ntimeevents"] = len(r["events"])
        out = pack(res)
        out["ntimeevents"] = np.array([len(r["events"]) for r in res])
        out["p"] = np.array([[1.0, 2.0, 2.5, 3.0, 1.0, 1.0], [0.5, 1.5, 2.5, 3.0, 0.25, -2.0],
                             [1.0, 2.0, 2.5, 3.0 - 6e-16, 1.0, 1.0]])
        out["y0"] = np.array([[0.0], [1.0], [0.0]])
        out["tf"] = np.array([5.0, 3.0, 3.0])
        save("time_ev_" + solver.lower(), **out)


def implicit_euler_cases():
    """ImplicitEuler (src/solvers/euler.pyx:29-483): VdP sweep members with analytic and finite-difference
    Jacobians, the stiff Robertson problem, and the bouncing ball / linear_ev problems with state events."""
    N = 4096
    idx = np.arange(0, N, 256)
    y0, p = cases.vdp_sweep(N, idx)
    for tag, usejac in (("jac", True), ("fd", False)):
        res = [run_member(ref_models.vdp(float(mu)), "ImplicitEuler", 2.0, h=1e-3, usejac=usejac) for mu in p[:, 0]]
        save("impleuler_vdp_" + tag, N=N, idx=idx, mu=p[:, 0], **pack_ie(res))
    N = 1 << 18
    idx = np.arange(0, N, N // 8)
    y0, p = cases.robertson_sweep(N, idx)
    res = [run_member(ref_models.robertson(*[float(v) for v in pk], with_jac=True), "ImplicitEuler", 4.0, h=1e-2,
                      rtol=1e-6, atol=cases.ROBERTSON_ATOL) for pk in p]
    save("impleuler_robertson", N=N, idx=idx, p=p, **pack_ie(res))
    N = 1 << 20
    idx = np.arange(0, N, N // 8)
    y0, p, sw0 = cases.ball_sweep(N, idx)
    res = [run_member(ref_models.ball(float(h0)), "ImplicitEuler", 3.0, h=1e-3) for h0 in y0[:, 0]]
    save("impleuler_ball", N=N, idx=idx, h0=y0[:, 0], **pack_ie(res))
    res = [run_member(ref_models.linear_ev(-1.0, 0.0, 0.5, 1.0), "ImplicitEuler", 2.0)]
    save("impleuler_linear_ev", **pack_ie(res))


def rodas_cases():
    """RodasODE (src/solvers/rosenbrock.py:259-465).  The reference's class runs over oracle/rodas_c.c, our C
    restatement of the Fortran RODAS core (no Fortran compiler here): these vectors pin the ensemble oracle and the
    GPU path to that pair, and the pair itself to the reference's one known answer (VdP mu = 1e6)."""
    N = 4096
    idx = np.arange(0, N, 256)
    y0, p = cases.vdp_sweep(N, idx)
    for tag, usejac in (("jac", True), ("fd", False)):
        res = [run_member(ref_models.vdp(float(mu)), "RodasODE", 2.0, rtol=1e-6, atol=1e-6, usejac=usejac) for mu in p[:, 0]]
        save("rodas_vdp_" + tag, N=N, idx=idx, mu=p[:, 0], **pack(res))
    res = [run_member(ref_models.vdp(1e6), "RodasODE", 2.0, rtol=1e-4, atol=1e-4)]   # test_rosenbrock.py:83-91
    save("rodas_vdp_mu1e6", **pack(res))
    N = 1 << 18
    idx = np.arange(0, N, N // 8)
    y0, p = cases.robertson_sweep(N, idx)
    res = [run_member(ref_models.robertson(*[float(v) for v in pk], with_jac=True), "RodasODE", 40.0,
                      rtol=1e-6, atol=cases.ROBERTSON_ATOL) for pk in p]
    save("rodas_robertson", N=N, idx=idx, p=p, **pack(res))
    N = 1 << 20
    idx = np.arange(0, N, N // 8)
    y0, p, sw0 = cases.ball_sweep(N, idx)
    res = [run_member(ref_models.ball(float(h0)), "RodasODE", 3.0, rtol=1e-6, atol=1e-6) for h0 in y0[:, 0]]
    save("rodas_ball", N=N, idx=idx, h0=y0[:, 0], **pack(res))
    grid = np.linspace(0.0, 3.0, 31)[1:]
    r = run_member(ref_models.ball(2.0), "RodasODE", 3.0, ncp=30, report_continuously=True, rtol=1e-6, atol=1e-6)
    t, y = r["t_sol"], r["y_sol"]
    dense = np.full((len(grid), 2), np.nan)
    for k, tg in enumerate(grid):
        hit = np.nonzero(t == tg)[0]
        if len(hit):
            dense[k] = y[hit[0]]
    save("ball_dense_rodasode", grid=grid, dense=dense, **pack([r]))


def pack_ie(results):
    out = pack(results)
    for k in ("nniters", "nnfails"):
        out[k] = np.array([r["extra"][k] for r in results], dtype=np.int64)
    out["failed"] = np.array([1 if r["failed"] else 0 for r in results], dtype=np.int64)
    return out


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "time_events":
        return time_event_cases()
    if len(sys.argv) > 1 and sys.argv[1] == "implicit_euler":
        return implicit_euler_cases()
    if len(sys.argv) > 1 and sys.argv[1] == "rodas":
        return rodas_cases()
    time_event_cases()
    implicit_euler_cases()
    rodas_cases()
    # ---- cfg 1: Dopri5 VdP sweep (subset of the N=4096 law) --------------------------------------
    N = 4096
    idx = np.arange(0, N, 64)
    y0, p = cases.vdp_sweep(N, idx)
    res = [run_member(ref_models.vdp(float(mu)), "Dopri5", 2.0, rtol=1e-6, atol=1e-6) for mu in p[:, 0]]
    save("cfg1_dopri5_vdp", N=N, idx=idx, mu=p[:, 0], **pack(res))

    # ---- cfg 2: RK4 h=1e-3 and RK34 on the 2^20 law ---------------------------------------------
    N = 1 << 20
    idx = np.arange(0, N, N // 16)
    y0, p = cases.vdp_sweep(N, idx)
    res = [run_member(ref_models.vdp(float(mu)), "RungeKutta4", 2.0, h=1e-3) for mu in p[:, 0]]
    save("cfg2_rk4_vdp", N=N, idx=idx, mu=p[:, 0], **pack(res))
    res = [run_member(ref_models.vdp(float(mu)), "ExplicitEuler", 2.0, h=1e-3) for mu in p[:, 0]]
    save("cfg2_euler_vdp", N=N, idx=idx, mu=p[:, 0], **pack(res))
    idx = np.arange(0, N, N // 48)
    y0, p = cases.vdp_sweep(N, idx)
    res = [run_member(ref_models.vdp(float(mu)), "RungeKutta34", 2.0, rtol=1e-6, atol=1e-6, inith=0.01)
           for mu in p[:, 0]]
    save("cfg2_rk34_vdp", N=N, idx=idx, mu=p[:, 0], **pack(res))

    # ---- cfg 3: bouncing ball with state events, ncp=0 / report_continuously=False ---------------
    N = 1 << 20
    idx = np.arange(0, N, N // 16)
    y0, p, sw0 = cases.ball_sweep(N, idx)
    for solver, opts in (("Dopri5", dict(rtol=1e-6, atol=1e-6)), ("RungeKutta34", dict(rtol=1e-6, atol=1e-6)),
                         ("ExplicitEuler", dict(h=1e-3)), ("Radau5ODE", dict(rtol=1e-6, atol=1e-6))):
        res = [run_member(ref_models.ball(float(h0)), solver, 3.0, **opts) for h0 in y0[:, 0]]
        save("cfg3_ball_" + solver.lower(), N=N, idx=idx, h0=y0[:, 0], **pack(res))

    # ---- cfg 4: Radau5ODE Robertson --------------------------------------------------------------
    N = 1 << 18
    idx = np.arange(0, N, N // 16)
    y0, p = cases.robertson_sweep(N, idx)
    for tag, with_jac in (("jac", True), ("fd", False)):
        res = [run_member(ref_models.robertson(*[float(v) for v in pk], with_jac=with_jac), "Radau5ODE", 40.0,
                          rtol=1e-6, atol=cases.ROBERTSON_ATOL) for pk in p]
        save("cfg4_radau5_robertson_" + tag, N=N, idx=idx, p=p, **pack(res))
    # Radau5ODE VdP stiff (the reference's own test case, tests/solvers/test_radau5.py:438-446)
    res = [run_member(ref_models.vdp(1e6), "Radau5ODE", 2.0, rtol=1e-4, atol=1e-4, inith=1e-4)]
    save("radau5_vdp_mu1e6", **pack(res))
    y0, p = cases.vdp_sweep(4096, np.arange(0, 4096, 256))
    res = [run_member(ref_models.vdp(float(mu)), "Radau5ODE", 2.0, rtol=1e-6, atol=1e-6) for mu in p[:, 0]]
    save("radau5_vdp_sweep", mu=p[:, 0], **pack(res))

    # ---- cfg 5: gyro Dopri5 with dense output on 100 grid points ----------------------------------
    N = 1 << 22
    idx = np.array([0, N // 3, N - 1])
    y0, p = cases.gyro_sweep(N, idx)
    rows, res = [], []
    for k in range(len(idx)):
        r = run_member(ref_models.gyro(tuple(y0[k, :3])), "Dopri5", 0.1, ncp=100, rtol=1e-8, atol=1e-8)
        res.append(r)
        rows.append((r["t_sol"], r["y_sol"]))
    nrow = min(len(t) for t, _ in rows)
    save("cfg5_gyro_dopri5", N=N, idx=idx, y0=y0, t_rows=np.array([t[:nrow] for t, _ in rows]),
         y_rows=np.array([y[:nrow] for _, y in rows]), **pack(res))
    # the reference's own formulation (np.dot/curl) for the pinned member
    r = run_member(ref_models.gyro_reference_form(), "Dopri5", 0.1, rtol=1e-10, atol=1e-10)
    save("gyro_reference_form", y=r["y_final"], t=np.array([r["t_final"]]), nsteps=np.array([r["stats"]["nsteps"]]))

    # ---- event problems from the reference's tests ------------------------------------------------
    for solver, opts in (("Dopri5", {}), ("RungeKutta34", {}), ("ExplicitEuler", {}), ("Radau5ODE", {})):
        res = [run_member(ref_models.extended(), solver, 10.0, **opts)]
        save("extended_" + solver.lower(), **pack(res))
        res = [run_member(ref_models.linear_ev(-1.0, 0.0, 0.5, 1.0), solver, 2.0, **opts)]
        save("linear_ev_" + solver.lower(), **pack(res))

    # ---- dense output + events (report_continuously=True, ncp>0): Dopri5 / Radau5ODE / Euler on the ball ----
    grid = np.linspace(0.0, 3.0, 31)[1:]
    for solver, opts in (("Dopri5", dict(rtol=1e-6, atol=1e-6)), ("Radau5ODE", dict(rtol=1e-6, atol=1e-6)),
                         ("ExplicitEuler", dict(h=1e-3))):
        r = run_member(ref_models.ball(2.0), solver, 3.0, ncp=30, report_continuously=True, **opts)
        t, y = r["t_sol"], r["y_sol"]
        dense = np.full((len(grid), 2), np.nan)
        for k, tg in enumerate(grid):
            hit = np.nonzero(t == tg)[0]
            if len(hit):
                dense[k] = y[hit[0]]
        save("ball_dense_" + solver.lower(), grid=grid, dense=dense, **pack([r]))


if __name__ == "__main__":
    main()
