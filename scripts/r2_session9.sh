#!/bin/bash
# staged host<->device copies: lanes x chunk size, on the 3-D slab's e2e job (8.6 GB up + 1 GB mask, 8.6 GB down)
O=gpurun_out
mkdir -p $O
{
for lanes in 8 12 16; do for mb in 4 8 16; do
  echo "== lanes $lanes chunk ${mb} MiB"
  XGB_STAGE_LANES=$lanes XGB_STAGE_CHUNK_MB=$mb timeout 200 python scripts/e2e_phases.py heat3d 20 2>&1 | grep "^\[warm" | head -2
done; done
} 2>&1 | tee $O/r2k_session9.txt
