#!/usr/bin/env python
"""What the host link sustains with 1..N ranks transferring AT ONCE (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py [MiB] [seconds]

Every rank measures the six modes of aens_measure_hostlink on its own GPU with page-locked buffers -- copy engine
H2D / D2H / both, SM-initiated reads / writes / both of mapped host memory (what the zero-copy host mode of
aens_simulate_host does) -- all ranks starting each mode together behind a barrier.  Rank 0 prints one JSON line:
per mode the per-GPU minimum, mean and the sum over ranks in GB/s.  This is the ceiling bench.py's `e2e` is judged
against (profiles/r2_pcie.md)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from assimulo_b200.gpu import _core

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {}
for mode, name in enumerate(_core.HOSTLINK_MODES):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    g = _core.measure_hostlink(local, mib << 20, mode, seconds)
    t = torch.tensor([g], dtype=torch.float64, device="cuda")
    if world > 1:
        lst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
        vals = [float(x[0]) for x in lst]
    else:
        vals = [g]
    out[name] = {"min": min(vals), "mean": sum(vals) / len(vals), "sum": sum(vals), "per_rank": [round(v, 2) for v in vals]}
if rank == 0:
    try:
        topo = os.popen("nvidia-smi topo -m 2>/dev/null | head -12").read()
    except Exception:
        topo = ""
    print(json.dumps({"ranks": world, "buffer_MiB": mib, "seconds_per_mode": seconds, "GBps": out,
                      "cpus": os.cpu_count(), "topo": topo}))
if world > 1:
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)
