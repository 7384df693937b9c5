#!/bin/bash
# One gpurun call that produces everything tools/profile_summary.py turns into profiles/<tag>_*:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/profile_all.sh r1'
# Order matters: every command first runs WITHOUT ncu (and must exit 0) before it is profiled.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err || { echo "bench failed"; tail -5 $OUT/bench_$TAG.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_$TAG.json 2>/dev/null
# launch list of a short bench run (per-launch device time, cold cache, serialised)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $OUT/plain_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $OUT/ncu_l.log 2>&1
# one --set full capture per kernel (second launch of each)
# the counters the metric names are not part of --set full: asked for explicitly
FP64=sm__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__sass_thread_inst_executed_op_dmul_pred_on.sum,sm__sass_thread_inst_executed_op_dadd_pred_on.sum
cap() { # name, run_cfg config, kernel regex
  python tools/run_cfg.py $2 2 > $OUT/plain_$1.log 2>&1 &&
    ncu --set full --metrics $FP64 --clock-control none --import-source on -k regex:$3 -s 1 -c 1 -f -o $OUT/prof_$1_$TAG \
        python tools/run_cfg.py $2 2 > $OUT/ncu_$1.log 2>&1
}
cap dopri5_vdp dopri5_23 aens_dopri5_d0_vdp
cap radau radau aens_radau5_d0_robertson
cap ball ball aens_dopri5_d1_ball
if [ -z "${AENS_PROFILE_LEAN:-}" ]; then   # the kernels round 2 did not touch keep their r1 captures in lean mode
  cap rk4 rk4 aens_rk4_d0_vdp
  cap rk34 rk34 aens_rk34_d0_vdp
  cap rodas rodas aens_rodas_d1_robertson
  cap gyrodense gyrodense aens_dopri5_d1_gyro
  cap gyrored gyrored aens_dopri5_d2_gyro
fi
ls -la $OUT/*_$TAG.* | awk '{print $5, $9}'
