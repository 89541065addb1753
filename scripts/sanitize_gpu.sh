#!/bin/bash
# compute-sanitizer passes over the kernel variants on small grids (race detection).
# memcheck: out-of-bounds / misaligned accesses (ghost rows, bulk-copy windows at the array ends);
# racecheck: shared-memory hazards in the ring pipelines (tiled, tiled2, jacobi2, multistep);
# synccheck: barrier / mbarrier misuse.  Run under gpurun; outputs in gpurun_out/.  Slow (10-100x), so the
# selection is small and every pass is bounded by `timeout`.
O=gpurun_out
mkdir -p $O
: > $O/sanitize_summary.txt
pass() {  # tool, label, timeout, command...
  tool=$1; label=$2; lim=$3; shift; shift; shift
  start=$(date +%s)
  timeout $lim compute-sanitizer --tool $tool --error-exitcode 9 --log-file $O/sanitize_${tool}_${label}.log "$@" > $O/sanitize_${tool}_${label}.out 2>&1
  rc=$?
  errs=$(grep -c "ERROR SUMMARY: 0 errors" $O/sanitize_${tool}_${label}.log)
  echo "$tool $label rc=$rc clean_summaries=$errs seconds=$(( $(date +%s) - start )) :: $(grep 'ERROR SUMMARY' $O/sanitize_${tool}_${label}.log | tail -1) :: $(tail -1 $O/sanitize_${tool}_${label}.out | cut -c1-120)" >> $O/sanitize_summary.txt
}
SMOKE='import __graft_entry__ as g; g.smoke()'
pass memcheck smoke 400 python -c "$SMOKE"
pass racecheck smoke 600 python -c "$SMOKE"
pass synccheck smoke 400 python -c "$SMOKE"
# tiled2 (two steps per pass), jacobi2 (fused solver pairs), 3-D tiled, short / tail multistep: small parity cases
pass racecheck pipelines 900 python -m pytest tests/test_sanitize_cases_gpu.py -q -x -m gpu
pass memcheck pipelines 600 python -m pytest tests/test_sanitize_cases_gpu.py -q -x -m gpu
cat $O/sanitize_summary.txt
