"""Helper of tests/test_parallel_gloo.py::test_multi_rank_bench_exit_path: the end of a multi-rank bench.py run, with gloo
standing in for NCCL -- rank 0 is still busy (the CPU baseline leg) when the other ranks are done."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist

import bench

dist.init_process_group("gloo")
rank = dist.get_rank()
dist.barrier()
bench._watchdog(90.0)
if rank == 0:
    time.sleep(2.0)
    print("LINE from rank 0", flush=True)
bench._leave(0)
