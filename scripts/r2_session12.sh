#!/bin/bash
# staged copies: streaming (non-temporal) chunk copies vs memcpy; byte-exactness test first
O=gpurun_out
mkdir -p $O
{
echo "== staged copy tests"; timeout 300 python -m pytest tests/test_staged_copy_gpu.py tests/test_edge_cases_gpu.py -m gpu -x -q 2>&1 | tail -2
for nt in 1 0 1 0; do
  echo "== XGB_STAGE_NT=$nt"
  XGB_STAGE_NT=$nt timeout 200 python scripts/e2e_phases.py heat3d 20 2>&1 | grep "^\[warm" | head -2
done
echo "== conv1d 20"; for nt in 1 0; do XGB_STAGE_NT=$nt timeout 100 python scripts/e2e_phases.py conv1d 20 2>&1 | grep "^\[warm" | head -2; done
} 2>&1 | tee $O/r2n_session12.txt
