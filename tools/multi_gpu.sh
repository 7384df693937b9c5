#!/bin/bash
# Everything measured at N GPUs in one gpurun call:   gpurun --gpus N -- 'bash tools/multi_gpu.sh N [tag]'
#   host-link probe with N concurrent ranks, the headline bench (e2e incl. the all-gather + gather check), and the other
#   BASELINE configurations (one JSON line each).  Outputs: gpurun_out/{pcie,bench,configs}_n<N>_<tag>.json
N=${1:-2}; TAG=${2:-r2}; OUT=gpurun_out; mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$RUN --master-port 29511 tools/pcie_probe.py 256 0.3 > $OUT/pcie_n${N}_$TAG.json 2> $OUT/pcie_n${N}_$TAG.err
$RUN --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err; echo "bench rc $?"
$RUN --master-port 29513 bench.py --gpus $N --config all --steps 3 --warmup 1 > $OUT/configs_n${N}_$TAG.json 2> $OUT/configs_n${N}_$TAG.err; echo "configs rc $?"
python - <<PY
import json
for f in ("pcie","bench","configs"):
    try:
        for l in open("$OUT/%s_n${N}_$TAG.json" % f):
            if not l.startswith("{"): continue
            d = json.loads(l)
            if f == "pcie": print("pcie", {k: round(v["min"],1) for k, v in d["GBps"].items()})
            elif f == "bench": print("bench value %.4g e2e %.4g (%.2f ms, gather %.2f ms, floor %.2f, mode %s) check %s" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["gather_ms"], d["e2e"]["floor"]["floor_ms"], d["e2e"]["host_mode"], d["gather_check"]))
            else: print("cfg %-40s value %.4g  ms %.3f frac %.3f" % (d["config"]["workload"][:40], d["value"], d["ms_per_step"], d["roofline"]["frac"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY
