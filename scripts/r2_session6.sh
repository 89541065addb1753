#!/bin/bash
# heat3d: batched interior path (all windows of the NSV vectors first) vs per-vector path; parity first.
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
echo "== parity (3-D cases, random programs, sanitize cases)"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_random_gpu.py tests/test_sanitize_cases_gpu.py tests/test_edge_cases_gpu.py -m gpu -x -q 2>&1 | tail -3
run "batched (default)" A=1
run "per-vector (XGB_TILED_BATCH=0)" XGB_TILED_BATCH=0
run "batched (repeat)" A=1
run "per-vector (repeat)" XGB_TILED_BATCH=0
run "batched, min blocks 2" XGB_TILED_MINB=2
run "batched, smem 72K" XGB_SMEM=73728
run "batched, TJ=16 smem 200K" XGB_TJ=16 XGB_SMEM=204800
run "batched, min_ctas 16384" XGB_MIN_CTAS=16384
run "batched, min_ctas 4736" XGB_MIN_CTAS=4736
} 2>&1 | tee $O/r2g_session6.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:heat_3d.*tiled -s 3 -c 1 -o $O/r2g_heat3d_tiled \
   python bench.py --workload heat3d --steps 4 --warmup 3 --no-cpu --no-e2e --no-parity > $O/r2g_ncu.log 2>&1
tail -2 $O/r2g_ncu.log
