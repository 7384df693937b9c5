#!/usr/bin/env python
"""Exploratory GPU run: parity of every solver/model against the C oracle + first timings.
Not a test and not the bench -- prints a table; used while developing (gpurun -- python tools/gpu_probe.py)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import cases  # noqa: E402
import ens_oracle as O  # noqa: E402
from assimulo_b200.gpu import _core  # noqa: E402

SOLVERS = ["ExplicitEuler", "RungeKutta4", "RungeKutta34", "Dopri5", "Radau5ODE"]


def run_gpu(model, solver, y0, p, sw0, tf, arith, t_grid=None, **opts):
    h = _core.Handle(0)
    h.set_model_builtin(model)
    h.set_ensemble(y0, p, sw0, 0.0)
    d = {"solver": _core.SOLVER_IDS[solver], "arith": arith}
    d.update(opts)
    h.set_opts(d)
    h.set_output(t_grid, 24)
    h.simulate(tf)
    t, y, sw, st = h.get_final()
    out = {"t": t, "y": y, "sw": sw, "status": st, "ms": h.last_kernel_ms(),
           "stats": {k: h.get_stats(k) for k in _core.STAT_NAMES}}
    if t_grid is not None:
        out["dense"] = h.get_dense()
    if h.n_g:
        out["ev"] = h.get_events()
    h.close()
    return out


def report(tag, g, o):
    dy = np.abs(g["y"] - o["y"])
    scale = np.maximum(np.abs(o["y"]), 1.0)
    rel = (dy / scale).max()
    ds = {k: int(np.abs(g["stats"][k].astype(np.int64) - o["stats"][k]).max()) for k in ("nsteps", "nfcns", "nerrfails")}
    nst = o["stats"]["nsteps"].sum()
    gst = g["stats"]["nsteps"].sum()
    print("%-44s relerr %.2e  dt %.1e  nsteps gpu/ora %d/%d (%.3f%%) maxdiff %s status %s  %.3f ms" % (
        tag, rel, np.abs(g["t"] - o["t"]).max(), gst, nst, 100.0 * (gst - nst) / max(nst, 1), ds,
        np.unique(g["status"]).tolist(), g["ms"]))


def main():
    print(_core.version())
    tf_peak, mhz = _core.measure_fp64_peak(0, 1.0)
    print("fp64 DFMA-chain peak: %.2f TFLOP/s (implied SM clock %.0f MHz)" % (tf_peak, mhz))

    N = 4096
    y0, p = cases.vdp_sweep(N)
    for solver, opts in (("ExplicitEuler", dict(h=1e-3)), ("RungeKutta4", dict(h=1e-3)),
                         ("RungeKutta34", dict(rtol=1e-6, atol=1e-6, inith=0.01)),
                         ("Dopri5", dict(rtol=1e-6, atol=1e-6)),
                         ("Radau5ODE", dict(rtol=1e-6, atol=1e-6, usejac=1))):
        o = O.simulate("vdp", solver, y0, p, tf=2.0, **opts)
        for arith in (0, 1):
            g = run_gpu("vdp", solver, y0, p, None, 2.0, arith, **opts)
            report("vdp %s %s" % (solver, "strict" if arith == 0 else "fast"), g, o)

    N = 2048
    y0, p, sw0 = cases.ball_sweep(N)
    for solver, opts in (("ExplicitEuler", dict(h=1e-3)), ("RungeKutta34", dict(rtol=1e-6, atol=1e-6)),
                         ("Dopri5", dict(rtol=1e-6, atol=1e-6)), ("Radau5ODE", dict(rtol=1e-6, atol=1e-6))):
        o = O.simulate("ball", solver, y0, p, sw0, tf=3.0, max_ev=24, **opts)
        for arith in (0, 1):
            g = run_gpu("ball", solver, y0, p, sw0, 3.0, arith, **opts)
            report("ball %s %s" % (solver, "strict" if arith == 0 else "fast"), g, o)
            cnt, te, info = g["ev"]
            same_cnt = (cnt == o["stats"]["nstateevents"]).all()
            m = min(te.shape[1], o["ev_t"].shape[1])
            dte = np.nanmax(np.abs(te[:, :m] - o["ev_t"][:, :m]))
            print("      events: counts equal %s, max |dt_event| %.3e, info equal %s" % (
                same_cnt, dte, (info[:, :m] == o["ev_info"][:, :m]).all()))

    N = 2048
    y0, p = cases.robertson_sweep(N)
    for uj in (1, 0):
        opts = dict(rtol=1e-6, atol=cases.ROBERTSON_ATOL, usejac=uj)
        o = O.simulate("robertson", "Radau5ODE", y0, p, tf=40.0, **opts)
        for arith in (0, 1):
            g = run_gpu("robertson", "Radau5ODE", y0, p, None, 40.0, arith, **opts)
            report("robertson Radau5 usejac=%d %s" % (uj, "strict" if arith == 0 else "fast"), g, o)

    N = 512
    y0, p = cases.gyro_sweep(N)
    grid = np.linspace(0, 0.1, 101)[1:]
    opts = dict(rtol=1e-8, atol=1e-8)
    o = O.simulate("gyro", "Dopri5", y0, p, tf=0.1, t_grid=grid, **opts)
    for arith in (0, 1):
        g = run_gpu("gyro", "Dopri5", y0, p, None, 0.1, arith, t_grid=grid, **opts)
        report("gyro Dopri5 dense %s" % ("strict" if arith == 0 else "fast"), g, o)
        dd = np.abs(np.transpose(g["dense"], (1, 0, 2)) - o["dense"])
        print("      dense rows max abs diff %.3e (nan rows gpu %d)" % (np.nanmax(dd), np.isnan(g["dense"]).sum()))

    for model in ("extended", "linear_ev"):
        n, npar, ng, nsw = O.model_dims(model)
        y0 = np.array([[0.0, -1.0, 0.0]]) if model == "extended" else np.array([[1.0]])
        p = None if model == "extended" else np.array([[-1.0, 0.0, 0.5]])
        sw0 = np.array([[0, 1, 1]], dtype=np.uint8) if model == "extended" else np.array([[1]], dtype=np.uint8)
        tf = 10.0 if model == "extended" else 2.0
        for solver in ("ExplicitEuler", "RungeKutta34", "Dopri5", "Radau5ODE"):
            o = O.simulate(model, solver, y0, p, sw0, tf=tf, max_ev=24)
            g = run_gpu(model, solver, y0, p, sw0, tf, 0)
            report("%s %s strict" % (model, solver), g, o)
            print("      y gpu", g["y"][0], "ora", o["y"][0], "events", g["ev"][0][0], o["stats"]["nstateevents"][0])

    # ---- timing: 1M-member sweeps ----
    N = 1 << 20
    y0, p = cases.vdp_sweep(N)
    for solver, opts in (("Dopri5", dict(rtol=1e-6, atol=1e-6)), ("RungeKutta34", dict(rtol=1e-6, atol=1e-6, inith=0.01)),
                         ("RungeKutta4", dict(h=1e-3)), ("ExplicitEuler", dict(h=1e-3))):
        for arith in (1, 0):
            h = _core.Handle(0)
            h.set_model_builtin("vdp")
            d = {"solver": _core.SOLVER_IDS[solver], "arith": arith}
            d.update(opts)
            best = 1e30
            for rep in range(3):
                h.set_ensemble(y0, p, None, 0.0)
                h.set_opts(d)
                h.set_output(None, 0)
                h.simulate(2.0)
                best = min(best, h.last_kernel_ms())
            s = h.get_stats_sum()
            print("TIMING vdp N=2^20 %-14s %-6s kernel %.3f ms  nsteps %d  -> %.3e accepted steps/s ; attempts %d" % (
                solver, "fast" if arith else "strict", best, s["nsteps"], s["nsteps"] / (best * 1e-3), s["nstepstotal"]))
            # reversed fetch order (most expensive members first)
            if solver == "Dopri5" and arith == 1:
                for thr in (32, 16, 8):
                    h.set_ensemble(y0, p, None, 0.0)
                    h.set_schedule(thr, np.arange(N - 1, -1, -1, dtype=np.intc))
                    h.simulate(2.0)
                    print("       reversed order, refill_threshold=%d: %.3f ms" % (thr, h.last_kernel_ms()))
                for thr in (16, 8, 1):
                    h.set_ensemble(y0, p, None, 0.0)
                    h.set_schedule(thr, None)
                    h.simulate(2.0)
                    print("       natural order, refill_threshold=%d: %.3f ms" % (thr, h.last_kernel_ms()))
            h.close()


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("probe wall %.1f s" % (time.time() - t0))
