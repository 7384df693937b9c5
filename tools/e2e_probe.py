#!/usr/bin/env python
"""Where does the end-to-end time of aens_simulate_host go?  (development probe, not a test)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from assimulo_b200.gpu import _core

N = 1 << 23
mu = 10.0 ** (-1.0 + 3.0 * np.arange(N) / (N - 1))
y0 = _core.pinned_empty((N, 2)); y0[:, 0] = 2.0; y0[:, 1] = -0.6
p = _core.pinned_empty((N, 1)); p[:, 0] = mu
oy = _core.pinned_empty((N, 2))
h = _core.Handle(0)
h.set_model_builtin("vdp")
h.set_opts({"solver": _core.SOLVER_IDS["Dopri5"], "arith": 1, "rtol": 1e-6, "atol": 1e-6})
h.set_fetch_reverse(True)

def timeit(f, n=5):
    f(); f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

# raw PCIe
d = torch.empty(N * 3, dtype=torch.float64, device="cuda")
hsrc = torch.empty(N * 3, dtype=torch.float64).pin_memory()
print("H2D 201 MB: %.2f ms" % timeit(lambda: d.copy_(hsrc, non_blocking=True)))
print("D2H 201 MB: %.2f ms" % timeit(lambda: hsrc.copy_(d, non_blocking=True)))
h.set_host_mode(2)
for rev in (0, 1, 2):
    h.set_fetch_reverse(rev)
    print("zero-copy fetch mode %d: full %.2f ms, kernel %.2f ms" % (
        rev, timeit(lambda: h.simulate_host(y0, p, None, 0.0, 2.0, out_y=oy)), h.last_kernel_ms()))
h.set_fetch_reverse(1)
print("zero-copy: full %.2f ms   no-D2H %.2f ms" % (
    timeit(lambda: h.simulate_host(y0, p, None, 0.0, 2.0, out_y=oy)),
    timeit(lambda: h.simulate_host(y0, p, None, 0.0, 2.0))))
print("   kernel %.2f ms" % h.last_kernel_ms())
h.set_host_mode(1)
for nch in (1, 16):
    print("chunks %2d: full %.2f ms   no-D2H %.2f ms" % (
        nch, timeit(lambda: h.simulate_host(y0, p, None, 0.0, 2.0, out_y=oy, nchunks=nch)),
        timeit(lambda: h.simulate_host(y0, p, None, 0.0, 2.0, nchunks=nch))))
# kernel only, chunked vs not: resident data
h.set_ensemble(y0, p, None, 0.0)
h.set_output(None, 0)
def k():
    h.reset(); h.simulate(2.0)
for rev in (0, 1, 2):
    h.set_fetch_reverse(rev)
    print("fetch mode %d reset+kernel resident: %.2f ms (kernel %.2f)" % (rev, timeit(k), h.last_kernel_ms()))
