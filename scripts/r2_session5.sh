#!/bin/bash
# Experiment: rotating-register variant of the tiled kernel (scripts/experiments/r2_tiled_rot_maskprefetch.patch)
# with 16-row tiles and one CTA per SM (116 registers fit 544 threads), vs the committed kernel.
O=gpurun_out
mkdir -p $O
run() {
  label=$1; shift
  line=$(env "$@" timeout 120 python bench.py --workload heat3d --steps 20 --warmup 5 --no-cpu --no-e2e --no-parity 2>$O/tune_err.txt | tail -1)
  python - "$label" "$line" <<'PY'
import json, sys
label, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    print(f"{label:52s} {d['value']:7.1f} Gpt/s  {d['ms_per_step']:.3f} ms  frac {d['roofline']['frac']:.3f}  clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(f"{label:52s} FAILED {line[:300]}")
PY
}
{
run "committed kernel" A=1
run "committed, TJ=16 smem 200K" XGB_TJ=16 XGB_SMEM=204800
run "committed, TJ=16 smem 200K NSV=2" XGB_TJ=16 XGB_SMEM=204800 XGB_NSV=2
run "committed, TJ=12 smem 200K" XGB_TJ=12 XGB_SMEM=204800
patch -p1 < scripts/experiments/r2_tiled_rot_maskprefetch.patch
run "patched (mask prefetch), no ROT" A=1
run "ROT TJ=16 smem 200K (1 CTA/SM, 17 warps)" XGB_TILED_ROT=1 XGB_TJ=16 XGB_SMEM=204800
run "ROT TJ=16 smem 160K" XGB_TILED_ROT=1 XGB_TJ=16 XGB_SMEM=163840
run "ROT TJ=12 smem 200K" XGB_TILED_ROT=1 XGB_TJ=12 XGB_SMEM=204800
run "ROT TJ=16 smem 200K min_ctas 4096" XGB_TILED_ROT=1 XGB_TJ=16 XGB_SMEM=204800 XGB_MIN_CTAS=4096
run "ROT TJ=16 smem 200K NSV=2" XGB_TILED_ROT=1 XGB_TJ=16 XGB_SMEM=204800 XGB_NSV=2
run "ROT TJ=24 smem 220K NSV=2" XGB_TILED_ROT=1 XGB_TJ=24 XGB_SMEM=225280 XGB_NSV=2
echo "== parity, ROT TJ=16 smem 200K"
XGB_TILED_ROT=1 XGB_TJ=16 XGB_SMEM=204800 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_random_gpu.py -m gpu -x -q -k "heat3d or random or full_size_heat" 2>&1 | tail -3
} 2>&1 | tee $O/r2e_session5.txt
