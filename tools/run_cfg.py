#!/usr/bin/env python
"""Run ONE BASELINE config a few times (for ncu captures):  python tools/run_cfg.py {rk4|rk34|dopri5|dopri5_23|ball|radau|rodas|gyro|gyrodense|gyrored} [reps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases
from assimulo_b200.gpu import _core

which = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = 1 << 20
grid = None
rev = False
if which in ("rk4", "rk34", "dopri5", "dopri5_23"):
    if which == "dopri5_23":
        N = 1 << 23  # the bench headline size
    y0, p = cases.vdp_sweep(N); sw0 = None; model = "vdp"; tf = 2.0; rev = True
    solver, opts = {"rk4": ("RungeKutta4", dict(h=1e-3)), "rk34": ("RungeKutta34", dict(rtol=1e-6, atol=1e-6, inith=0.01)),
                    "dopri5": ("Dopri5", dict(rtol=1e-6, atol=1e-6)),
                    "dopri5_23": ("Dopri5", dict(rtol=1e-6, atol=1e-6))}[which]
elif which == "ball":
    y0, p, sw0 = cases.ball_sweep(N); model = "ball"; tf = 3.0; solver, opts = "Dopri5", dict(rtol=1e-6, atol=1e-6)
elif which == "rodas":
    y0, p = cases.robertson_sweep(1 << 18); sw0 = None; model = "robertson"; tf = 40.0
    solver, opts = "RodasODE", dict(rtol=1e-6, atol=np.array(cases.ROBERTSON_ATOL), usejac=1)
elif which == "radau":
    y0, p = cases.robertson_sweep(1 << 18); sw0 = None; model = "robertson"; tf = 40.0
    solver, opts = "Radau5ODE", dict(rtol=1e-6, atol=np.array(cases.ROBERTSON_ATOL), usejac=1)
else:
    y0, p = cases.gyro_sweep(1 << 22); sw0 = None; model = "gyro"; tf = 0.1
    solver, opts = "Dopri5", dict(rtol=1e-8, atol=1e-8)
    if which in ("gyrodense", "gyrored"):
        grid = np.linspace(0, 0.1, 101)[1:]
h = _core.Handle(0)
h.set_model_builtin(model)
h.set_ensemble(y0, p, sw0, 0.0)
d = {"solver": _core.SOLVER_IDS[solver]}
d.update(opts)
h.set_opts(d)
if which == "gyrored":
    h.set_row_reduction(True)
h.set_output(grid, 16)
h.set_fetch_reverse(rev)
if os.environ.get("AENS_LANES"):
    h.set_lanes_per_member(int(os.environ["AENS_LANES"]))
for _ in range(reps):
    h.reset(); h.simulate(tf)
    print(which, "kernel ms", h.last_kernel_ms(), "lanes", h.last_lanes_per_member(), h.get_summary())
if which == "gyrored":
    import time
    t0 = time.perf_counter(); rs = h.get_row_reduction(); t1 = time.perf_counter()
    print("get_row_reduction %.2f ms; count[-1] %d; mean[-1][:3] %s" % (1e3 * (t1 - t0), rs["count"][-1], rs["mean"][-1][:3]))
