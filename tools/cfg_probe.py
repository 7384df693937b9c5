#!/usr/bin/env python
"""BASELINE.json configs 2-5 at full size on one GPU: kernel time, accepted steps/s, status (development probe)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases
from assimulo_b200.gpu import _core

def run(tag, model, solver, y0, p, sw0, tf, t_grid=None, reverse=False, reps=3, arith=None, **opts):
    h = _core.Handle(0)
    h.set_model_builtin(model)
    h.set_ensemble(y0, p, sw0, 0.0)
    d = {"solver": _core.SOLVER_IDS[solver]}
    if arith is not None: d["arith"] = arith
    d.update(opts)
    h.set_opts(d)
    h.set_output(t_grid, 16)
    h.set_fetch_reverse(reverse)
    ms = []
    for _ in range(reps):
        h.reset(); h.simulate(tf); ms.append(h.last_kernel_ms())
    sums, nbad = h.get_summary()
    print("%-34s N=%8d kernel %8.3f ms (min of %s)  nsteps %12d -> %.3e steps/s  attempts %d nbad %d events %d" % (
        tag, y0.shape[0], min(ms), ["%.2f" % m for m in ms], sums["nsteps"], sums["nsteps"] / (min(ms) * 1e-3),
        sums["nstepstotal"], nbad, sums["nstateevents"]), flush=True)
    h.close()

N = 1 << 20
y0, mu = cases.vdp_sweep(N)
run("cfg2 RK4 vdp h=1e-3 strict", "vdp", "RungeKutta4", y0, mu, None, 2.0, h=1e-3, reverse=True)
run("cfg2 RK4 vdp h=1e-3 fast", "vdp", "RungeKutta4", y0, mu, None, 2.0, h=1e-3, arith=1, reverse=True)
run("cfg2 RK34 vdp fast", "vdp", "RungeKutta34", y0, mu, None, 2.0, rtol=1e-6, atol=1e-6, inith=0.01, reverse=True)
run("cfg2 RK34 vdp fast natural", "vdp", "RungeKutta34", y0, mu, None, 2.0, rtol=1e-6, atol=1e-6, inith=0.01)
run("cfg1-like Dopri5 vdp 1M fast", "vdp", "Dopri5", y0, mu, None, 2.0, rtol=1e-6, atol=1e-6, reverse=True)
y0, p, sw0 = cases.ball_sweep(N)
run("cfg3 Dopri5 ball events", "ball", "Dopri5", y0, p, sw0, 3.0, rtol=1e-6, atol=1e-6)
run("cfg3 Dopri5 ball events reverse", "ball", "Dopri5", y0, p, sw0, 3.0, rtol=1e-6, atol=1e-6, reverse=True)
N4 = 1 << 18
y0, p = cases.robertson_sweep(N4)
run("cfg4 Radau5 robertson jac", "robertson", "Radau5ODE", y0, p, None, 40.0, rtol=1e-6, atol=np.array([1e-8, 1e-14, 1e-6]), usejac=1)
run("cfg4 Radau5 robertson fd", "robertson", "Radau5ODE", y0, p, None, 40.0, rtol=1e-6, atol=np.array([1e-8, 1e-14, 1e-6]), usejac=0)
N5 = 1 << 22
y0, p = cases.gyro_sweep(N5)
run("cfg5 Dopri5 gyro final only", "gyro", "Dopri5", y0, p, None, 0.1, rtol=1e-8, atol=1e-8)
grid = np.linspace(0, 0.1, 101)[1:]
run("cfg5 Dopri5 gyro dense 100 rows", "gyro", "Dopri5", y0, p, None, 0.1, t_grid=grid, rtol=1e-8, atol=1e-8, reps=2)
