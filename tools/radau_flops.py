#!/usr/bin/env python
"""fp64 flops per ACCEPTED Radau5ODE step on the cfg4 Robertson ensemble, counted by the instrumented CPU oracle
(oracle/ens_oracle.c built with -DORA_COUNT_FLOPS; SURVEY 8d convention: add/sub/mul/div/sqrt/pow = 1, a multiply-add
= 2, abs/max/min/compare/negation = 0; rejected attempts, Newton failures, Jacobians and LU decompositions are all
charged to the accepted steps).  Writes profiles/r2_radau_flops.json, which bench.py reads as flops_per_accepted_step
of cfg4 (round 1 used the survey's estimate, 1040)."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import cases
import ens_oracle as O

lib = os.path.join(ROOT, "oracle", "libens_oracle_flops.so")
subprocess.check_call(["gcc", "-DORA_COUNT_FLOPS", "-O2", "-ffp-contract=off", "-fno-strict-aliasing", "-fPIC", "-shared",
                       "-fopenmp", os.path.join(ROOT, "oracle", "ens_oracle.c"), "-o", lib, "-lm"])
O.LIB = lib
O.build = lambda force=False: lib
O._lib = None
take = O.lib().ora_flops_take
take.restype = C.c_double

out = {}
N = 1 << 18
idx = np.arange(0, N, 64)
for name, usejac in (("analytic Jacobian (bench cfg4)", 1), ("finite-difference Jacobian", 0)):
    y0, p = cases.robertson_sweep(N, idx)
    take()
    o = O.simulate("robertson", "Radau5ODE", y0, p, tf=40.0, rtol=1e-6, atol=cases.ROBERTSON_ATOL, usejac=usejac)
    fl = take()
    st = {k: int(v.astype(np.int64).sum()) for k, v in o["stats"].items()}
    out[name] = {"members": len(idx), "flops": fl, "accepted_steps": st["nsteps"], "flops_per_accepted_step": fl / st["nsteps"],
                 "attempts": st["nstepstotal"], "nfcns": st["nfcns"], "njacs": st["njacs"], "nlus": st["nlus"],
                 "nerrfails": st["nerrfails"], "newton_iterations_per_attempt": (st["nfcns"] - st["nsteps"] - len(idx)) / 3.0 / st["nstepstotal"]}
    print(name, json.dumps(out[name]))
take()
y0, p = cases.robertson_sweep(N, idx)
o = O.simulate("robertson", "RodasODE", y0, p, tf=40.0, rtol=1e-6, atol=cases.ROBERTSON_ATOL, usejac=1)
fl = take()
st = {k: int(v.astype(np.int64).sum()) for k, v in o["stats"].items()}
out["RodasODE, analytic Jacobian (bench cfg4')"] = {"members": len(idx), "flops": fl, "accepted_steps": st["nsteps"],
                                                    "flops_per_accepted_step": fl / st["nsteps"], "attempts": st["nstepstotal"],
                                                    "nfcns": st["nfcns"], "njacs": st["njacs"], "nlus": st["nlus"]}
print("RodasODE", json.dumps(out["RodasODE, analytic Jacobian (bench cfg4')"]))
res = {"flops_per_accepted_step": out["analytic Jacobian (bench cfg4)"]["flops_per_accepted_step"],
       "rodas_flops_per_accepted_step": out["RodasODE, analytic Jacobian (bench cfg4')"]["flops_per_accepted_step"],
       "convention": "add/sub/mul/div/sqrt/pow = 1, multiply-add = 2, abs/max/min/compare/negation = 0 (SURVEY 8d)",
       "workload": "cfg4: Radau5ODE, Robertson, t in [0, 40], rtol 1e-6, atol [1e-8, 1e-14, 1e-6], every 64th of 2^18 members",
       "detail": out}
json.dump(res, open(os.path.join(ROOT, "profiles", "r2_radau_flops.json"), "w"), indent=1)
os.remove(lib)
