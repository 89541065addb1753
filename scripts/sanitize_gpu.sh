#!/bin/bash
# compute-sanitizer passes over the kernel variants on small grids (race detection).
# memcheck: out-of-bounds / misaligned accesses (ghost rows, bulk-copy windows at the array ends);
# racecheck: shared-memory hazards in the ring pipelines (tiled, tiled2, jacobi2, multistep);
# synccheck: barrier / mbarrier misuse.  Run under gpurun; outputs in gpurun_out/.  Slow (10-50x), so the
# test selection is small and every pass is bounded by `timeout`.
O=gpurun_out
mkdir -p $O
SEL='tests/test_parity_gpu.py -k "golden_single_grid or multistep_tail or oracle_diff2d or golden_overstep"'
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --log-file $O/sanitize_$tool.log \
      python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitize_${tool}_smoke.out 2>&1
  echo "$tool smoke rc=$?" >> $O/sanitize_summary.txt
  eval timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --log-file $O/sanitize_${tool}_tests.log \
      python -m pytest $SEL -q -x -m gpu > $O/sanitize_${tool}_tests.out 2>&1
  echo "$tool tests rc=$?" >> $O/sanitize_summary.txt
done
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $O/sanitize_racecheck_jacobi2.log \
    python -m pytest tests/test_jacobi2.py -q -x -m gpu -k "640 or random" > $O/sanitize_racecheck_jacobi2.out 2>&1
echo "racecheck jacobi2 rc=$?" >> $O/sanitize_summary.txt
cat $O/sanitize_summary.txt
