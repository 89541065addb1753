#!/bin/bash
O=gpurun_out
mkdir -p $O
run() {
  label=$1; wl=$2; k=$3; shift; shift; shift
  line=$(env "$@" timeout 120 python bench.py --workload $wl --steps $k --warmup $k --no-cpu --no-e2e --no-parity 2>$O/tune_err.txt | tail -1)
  python - "$label" "$line" <<'PY'
import json, sys
label, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    print(f"{label:44s} {d['value']:7.1f} Gpt/s  {d['ms_per_step']:.4f} ms  frac {d['roofline']['frac']:.3f}  launches {d['gpu_launches']} clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(f"{label:44s} FAILED {line[:300]}")
PY
}
{
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== parity (1-D / multistep)"; timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_sanitize_cases_gpu.py tests/test_edge_cases_gpu.py tests/test_zz_ring_protocol_gpu.py tests/test_reference_programs_gpu.py tests/test_import_xgrid_gpu.py -m gpu -x -q 2>&1 | tail -3
run "conv1d 20 steps" conv1d 20 A=1
run "conv1d 20 steps (repeat)" conv1d 20 A=1
run "conv1d 10000 steps" conv1d 10000 A=1
run "conv1d_nl 10000 steps" conv1d_nl 10000 A=1
run "diff1d 10000 steps" diff1d 10000 A=1
} 2>&1 | tee $O/r2m_session11.txt
