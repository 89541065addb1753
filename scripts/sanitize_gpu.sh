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
# the peer-memory halo exchange kernel as a ring of one rank: all copy widths, split batches, graph replay
pass memcheck peer_exchange 300 python -m pytest tests/test_peer_gpu.py -q -x -m gpu
# control: the textbook single-stage bulk-copy + mbarrier pattern (correct by construction).  A report here means
# racecheck does not model completion through mbarrier::complete_tx
pass racecheck control 300 python -m pytest tests/test_racecheck_control_gpu.py -q -x -m gpu -k single_stage
# control 2: a two-stage ring released through an "empty" mbarrier -- by ONE elected lane per warp after __syncwarp()
# (the pipeline kernels' pattern, cumulative release) vs by EVERY consumer thread
pass racecheck control_ring_elected_lane 300 python -m pytest tests/test_racecheck_control_gpu.py -q -x -m gpu -k ring_elected_lane
pass racecheck control_ring_every_thread 300 python -m pytest tests/test_racecheck_control_gpu.py -q -x -m gpu -k ring_every_thread
# individual hazards (type, thread, address): two steps per pass, fused pairs
for case in two_steps fused_jacobi; do
  NV_COMPUTE_SANITIZER_MAX_RACECHECK_HAZARDS=400000 timeout 900 compute-sanitizer --tool racecheck --racecheck-report hazard --print-limit 8000 \
      --log-file $O/sanitize_racecheck_hazards_$case.log python -m pytest tests/test_sanitize_cases_gpu.py -q -x -m gpu -k "$case" > $O/sanitize_racecheck_hazards_$case.out 2>&1
done
python scripts/racecheck_classify.py $O/sanitize_racecheck_hazards_*.log $O/sanitize_racecheck_control*.log | tee $O/sanitize_racecheck_classes.txt
cat $O/sanitize_summary.txt
