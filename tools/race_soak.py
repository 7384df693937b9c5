#!/usr/bin/env python
"""Race / ordering soak of the kernels whose lanes cooperate through shared memory (the row reduction's per-warp tile and
lock, aens_reduce_grid): the same launch repeated R times under different warp schedules -- varying refill thresholds
and fetch orders change which lanes arrive at the tile together -- must give bit-identical counts, minima and maxima
every time, and means / variances equal to rounding.  A lost update, a torn record or a lane reading a half-written
tile shows up as a different count or extremum.  (compute-sanitizer's racecheck is not available on this pool.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases
from assimulo_b200.gpu import _core

R = int(sys.argv[1]) if len(sys.argv) > 1 else 24
bad = 0
for model, solver, arrs, tf, grid, opts in (
        ("ball", "Dopri5", cases.ball_sweep(20011), 3.0, np.linspace(0, 3, 61)[1:], dict(rtol=1e-6, atol=1e-6)),
        ("vdp", "Radau5ODE", cases.vdp_sweep(20011), 2.0, np.linspace(0, 2, 41)[1:], dict(usejac=1)),
        ("gyro", "Dopri5", cases.gyro_sweep(20011), 0.1, np.linspace(0, 0.1, 51)[1:], dict(rtol=1e-8, atol=1e-8))):
    if model == "vdp":
        arrs = (arrs[0], np.minimum(arrs[1], 30.0))
    ref = None
    rng = np.random.RandomState(1)
    for r in range(R):
        h = _core.Handle(0)
        h.set_model_builtin(model)
        h.set_ensemble(arrs[0], arrs[1], arrs[2] if len(arrs) > 2 else None, 0.0)
        d = {"solver": _core.SOLVER_IDS[solver], "arith": 0}
        d.update(opts)
        h.set_opts(d)
        h.set_row_reduction(True)
        h.set_output(grid, 8)
        thr = [32, 1, 7, 16][r % 4]
        order = None if r % 3 == 0 else rng.permutation(len(arrs[0])).astype(np.intc)
        h.set_schedule(thr, order)
        h.simulate(tf)
        rs = h.get_row_reduction()
        chk = h.check_failures()
        h.close()
        if ref is None:
            ref = rs
            continue
        same = (np.array_equal(rs["count"], ref["count"]) and np.array_equal(rs["min"], ref["min"], equal_nan=True)
                and np.array_equal(rs["max"], ref["max"], equal_nan=True))
        close = (np.allclose(rs["mean"], ref["mean"], rtol=1e-12, atol=1e-13, equal_nan=True)
                 and np.allclose(rs["var"], ref["var"], rtol=1e-9, atol=1e-12, equal_nan=True))
        if not (same and close) or chk:
            bad += 1
            print("MISMATCH", model, solver, "run", r, "threshold", thr, "permuted" if order is not None else "ascending", chk)
    print("%-6s %-10s %d launches under %d schedules: counts / min / max identical, mean / var to rounding: %s"
          % (model, solver, R, R, "OK" if not bad else "FAILED"), flush=True)
sys.exit(1 if bad else 0)
