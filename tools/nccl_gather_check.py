#!/usr/bin/env python
"""Multi-GPU check of the path's only collective (run under torchrun on N GPUs):
every rank integrates its contiguous block of a bouncing-ball ensemble (state events) on its own GPU,
parallel.gather_final all-gathers {t, y, status, stats, event times} with the library's own ncclAllGather
(aens_comm_init / aens_allgather_results / aens_get_gathered: torch.distributed only carries the communicator id),
and every rank compares the assembled result with the CPU oracle.  tests/test_gpu_round2.py runs it on 2 GPUs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

import cases
import ens_oracle as O
from assimulo_b200.gpu import Batched_Explicit_Problem, Dopri5, parallel

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 10007
y0, p, sw0 = cases.ball_sweep(N)
full = Batched_Explicit_Problem(model="ball", y0=y0, p0=p, sw0=sw0)
sim = Dopri5(parallel.shard_problem(full, rank, world), device=local)
sim.simulate(3.0)
g = parallel.gather_final(sim, N)
o = O.simulate("ball", "Dopri5", y0, p, sw0, tf=3.0, max_ev=16, rtol=1e-6, atol=1e-6)
err = np.max(np.abs(g["y"] - o["y"]) / np.maximum(np.abs(o["y"]), 1.0))
m = np.isfinite(o["ev_t"])
ok = (err < 1e-5 and np.array_equal(g["event_count"], o["stats"]["nstateevents"]) and
      np.max(np.abs(g["event_t"][m] - o["ev_t"][m])) < 1e-9 and np.array_equal(g["event_info"][m], o["ev_info"][m]) and
      (g["status"] == 0).all() and g["statistics"]["nstateevents"] == int(o["stats"]["nstateevents"].sum()))
print("rank %d/%d: gather_final over NCCL: N=%d max rel err %.2e, %d events, %s" % (
    rank, world, N, err, int(g["event_count"].sum()), "OK" if ok else "MISMATCH"), flush=True)
dist.barrier()
sys.stdout.flush()
os._exit(0 if ok else 1)   # past the last collective: no communicator teardown while the peers are exiting too
