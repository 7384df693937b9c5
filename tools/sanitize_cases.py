#!/usr/bin/env python
"""Small runs of every kernel family for compute-sanitizer (tools/sanitize.sh): 1024-member ensembles through the C ABI --
d0 (final states), d1 (state events / stored rows) and d2 (rows reduced in the kernel) per solver, the sub-warp-per-member
Dopri5, the one-call host modes, the time limit and the all-gather on a single-rank communicator.  Prints one line per
case; exits non-zero if a member fails that should not."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases
from assimulo_b200.gpu import _core

N = int(os.environ.get("AENS_SAN_N", "1024"))
only = sys.argv[1:] or None
bad = 0


def run(tag, model, solver, arrs, tf, grid=None, reduce=False, lanes=0, arith=None, **opts):
    global bad
    if only and not any(o in tag for o in only):
        return
    h = _core.Handle(0)
    h.set_model_builtin(model)
    h.set_ensemble(arrs[0], arrs[1], arrs[2] if len(arrs) > 2 else None, 0.0)
    d = {"solver": _core.SOLVER_IDS[solver]}
    if arith is not None:
        d["arith"] = arith
    d.update(opts)
    h.set_opts(d)
    h.set_row_reduction(reduce)
    h.set_output(grid, 8)
    h.set_lanes_per_member(lanes)
    h.simulate(tf)
    st = h.get_final()[3]
    sums, nbad = h.get_summary()
    if reduce:
        rs = h.get_row_reduction()
        assert (rs["count"] == N).all()
    chk = h.check_failures()
    print("%-46s steps %8d events %6d not complete %d bounds violation %d" % (tag, sums["nsteps"], sums["nstateevents"], nbad, chk), flush=True)
    bad += int(nbad) + int(chk != 0)
    h.close()


vdp = cases.vdp_sweep(N)
vdp = (vdp[0], np.minimum(vdp[1], 20.0))
ball = cases.ball_sweep(N)
rob = cases.robertson_sweep(N)
gyro = cases.gyro_sweep(N)
g20 = np.linspace(0.0, 1.0, 21)
for arith in (1, 0):
    a = "fast" if arith else "strict"
    run("d0 Dopri5 vdp " + a, "vdp", "Dopri5", vdp, 1.0, arith=arith, rtol=1e-6, atol=1e-6)
    run("d0 RungeKutta34 vdp " + a, "vdp", "RungeKutta34", vdp, 1.0, arith=arith)
    run("d0 Radau5ODE robertson " + a, "robertson", "Radau5ODE", rob, 40.0, arith=arith, rtol=1e-6, atol=np.array(cases.ROBERTSON_ATOL), usejac=1)
    run("d1 RodasODE robertson " + a, "robertson", "RodasODE", rob, 40.0, arith=arith, rtol=1e-6, atol=np.array(cases.ROBERTSON_ATOL), usejac=1)
    run("d1 Dopri5 ball events " + a, "ball", "Dopri5", ball, 2.0, arith=arith)
    run("d1 Radau5ODE ball events " + a, "ball", "Radau5ODE", ball, 2.0, arith=arith)
    run("d1 Dopri5 vdp rows " + a, "vdp", "Dopri5", vdp, 1.0, grid=g20, arith=arith)
    run("d2 Dopri5 vdp rows reduced " + a, "vdp", "Dopri5", vdp, 1.0, grid=g20, reduce=True, arith=arith)
    run("d2 Radau5ODE vdp rows reduced " + a, "vdp", "Radau5ODE", vdp, 1.0, grid=g20, reduce=True, arith=arith, usejac=1)
    run("d2 Dopri5 ball events + rows reduced " + a, "ball", "Dopri5", ball, 2.0, grid=2 * g20, reduce=True, arith=arith)
run("d0 RungeKutta4 vdp", "vdp", "RungeKutta4", vdp, 1.0, h=1e-2)
run("d1 ExplicitEuler ball events", "ball", "ExplicitEuler", ball, 2.0, h=1e-2)
run("d2 ExplicitEuler vdp rows reduced", "vdp", "ExplicitEuler", vdp, 1.0, grid=g20, reduce=True, h=1e-2)
run("d0 ImplicitEuler vdp", "vdp", "ImplicitEuler", vdp, 1.0, h=1e-2, usejac=1)
run("d0 Dopri5 gyro thread-per-member", "gyro", "Dopri5", gyro, 0.1, rtol=1e-8, atol=1e-8)
run("d2 Dopri5 gyro rows reduced", "gyro", "Dopri5", gyro, 0.1, grid=0.1 * g20, reduce=True, rtol=1e-8, atol=1e-8)
run("d0 Dopri5 gyro 4 lanes per member", "gyro", "Dopri5", gyro, 0.1, lanes=4, rtol=1e-8, atol=1e-8)
run("d1 Dopri5 gyro 4 lanes per member rows", "gyro", "Dopri5", gyro, 0.1, grid=0.1 * g20, lanes=4, rtol=1e-8, atol=1e-8)
run("d0 Radau5ODE gyro (n = 12, rolled)", "gyro", "Radau5ODE", gyro, 0.1, rtol=1e-6, atol=1e-6)

if not only or any("host" in o for o in only):
    # one-call host modes (zero-copy reads / writes of mapped host memory) and the all-gather on a 1-rank communicator
    y0_h, p_h = _core.pinned_empty((N, 2)), _core.pinned_empty((N, 1))
    y0_h[:], p_h[:] = vdp
    out = _core.pinned_empty((N, 2))
    for mode in (2, 1, 3, 4):
        h = _core.Handle(0)
        h.set_model_builtin("vdp")
        h.set_opts({"solver": _core.SOLVER_IDS["Dopri5"]})
        h.set_host_mode(mode)
        sums, nbad = h.simulate_host(y0_h, p_h, None, 0.0, 1.0, out_y=out, nchunks=3)
        chk = h.check_failures()
        print("host mode %d: steps %d not complete %d bounds violation %d" % (mode, sums["nsteps"], nbad, chk), flush=True)
        bad += int(nbad) + int(chk != 0)
        if mode == 2:
            h.comm_init(0, 1, _core.comm_unique_id())
            h.allgather_results(0)
            g = h.get_gathered(N)
            assert np.array_equal(g["y"], out)
            print("all-gather (1 rank): %.3f ms" % h.last_gather_ms(), flush=True)
        h.close()
sys.exit(1 if bad else 0)
