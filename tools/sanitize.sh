#!/bin/bash
# compute-sanitizer over small runs of every kernel family:   gpurun -- 'bash tools/sanitize.sh [tag]'
#   memcheck   out-of-bounds / misaligned global, shared and local accesses, leaks of device memory
#   racecheck  shared-memory hazards (the per-warp tile + lock of aens_reduce_grid, the per-thread Radau slabs)
#   synccheck  warp-synchronous primitives with wrong masks (the group shuffles of the sub-warp-per-member kernel)
# Logs: gpurun_out/sanitizer_<tool>_<tag>.log (tools/profile_summary.py copies their summaries into profiles/)
TAG=${1:-r2}; OUT=gpurun_out; mkdir -p $OUT
export AENS_SAN_N=${AENS_SAN_N:-1024}
python tools/sanitize_cases.py > $OUT/sanitizer_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/sanitizer_plain_$TAG.log; exit 1; }
compute-sanitizer --tool memcheck --leak-check full --error-exitcode 9 python tools/sanitize_cases.py > $OUT/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc $?"
AENS_SAN_N=256 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python tools/sanitize_cases.py d2 "d0 Radau5ODE robertson" "d1 RodasODE" > $OUT/sanitizer_racecheck_$TAG.log 2>&1; echo "racecheck rc $?"
AENS_SAN_N=256 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_cases.py lanes d2 > $OUT/sanitizer_synccheck_$TAG.log 2>&1; echo "synccheck rc $?"
for t in memcheck racecheck synccheck; do echo "== $t"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY|Error|hazard|closed" $OUT/sanitizer_${t}_$TAG.log | sort | uniq -c | sort -rn | head -8; done
# Where the pool refuses compute-sanitizer: the library's own bounds-checked build (every index the kernels form, tested
# where it is formed) over the same cases, and the schedule soak of the shared-memory row reduction.
if [ -d assimulo_b200/_variants/checked ]; then
  LD_LIBRARY_PATH=assimulo_b200/_variants/checked python tools/sanitize_cases.py > $OUT/checked_cases_$TAG.log 2>&1; echo "checked build rc $?"
  LD_LIBRARY_PATH=assimulo_b200/_variants/checked python tools/race_soak.py 24 > $OUT/race_soak_$TAG.log 2>&1; echo "race soak rc $?"
  grep -c "bounds violation 0" $OUT/checked_cases_$TAG.log; grep -v "bounds violation 0" $OUT/checked_cases_$TAG.log | head; cat $OUT/race_soak_$TAG.log
fi
