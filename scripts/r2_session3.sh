#!/bin/bash
# Round 2, GPU session 3: parity of the mask-prefetch tiled kernels (+ rotating-register variant), heat3d sweep, callee/import tests.
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
    print(f"{label:44s} {d['value']:7.1f} Gpt/s  {d['ms_per_step']:.3f} ms  frac {d['roofline']['frac']:.3f}  clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(f"{label:44s} FAILED {line[:300]}")
PY
}
{
echo "== parity (default build)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== parity subset, XGB_TILED_ROT=1"; XGB_TILED_ROT=1 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_random_gpu.py tests/test_edge_cases_gpu.py -m gpu -x -q -k "not 2p24" 2>&1 | tail -4
echo "== parity subset, XGB_TILED_ROT=1 XGB_NSV=2"; XGB_TILED_ROT=1 XGB_NSV=2 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_random_gpu.py -m gpu -x -q -k "heat3d or random or diff2d or oracle" 2>&1 | tail -4
run "default (mask prefetch)" A=1
run "default (repeat)" A=1
run "ROT NSV=4 (116 regs, 1 CTA/SM?)" XGB_TILED_ROT=1
run "ROT NSV=4 minb 2" XGB_TILED_ROT=1 XGB_TILED_MINB=2
run "ROT NSV=2" XGB_TILED_ROT=1 XGB_NSV=2
run "ROT NSV=2 smem 72K" XGB_TILED_ROT=1 XGB_NSV=2 XGB_SMEM=73728
run "ROT NSV=2 min_ctas 16384" XGB_TILED_ROT=1 XGB_NSV=2 XGB_MIN_CTAS=16384
run "min_ctas 16384" XGB_MIN_CTAS=16384
run "smem 72K" XGB_SMEM=73728
run "smem 72K min_ctas 16384" XGB_SMEM=73728 XGB_MIN_CTAS=16384
} 2>&1 | tee $O/r2c_session3.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:heat_3d.*tiled -s 3 -c 1 -o $O/r2c_heat3d_tiled \
   python bench.py --workload heat3d --steps 4 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r2c_ncu.log 2>&1
tail -2 $O/r2c_ncu.log
