#!/bin/bash
# mbarrier.try_wait suspend-time hint: cavity (fused Jacobi pairs are issue-bound, 10 % of their instructions poll) and heat3d
O=gpurun_out
mkdir -p $O
run() {
  label=$1; wl=$2; k=$3; shift; shift; shift
  line=$(env "$@" timeout 200 python bench.py --workload $wl --steps $k --warmup 4 --no-cpu --no-e2e --no-parity 2>$O/tune_err.txt | tail -1)
  python - "$label" "$line" <<'PY'
import json, sys
label, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    print(f"{label:40s} {d['value']:7.1f} Gpt/s  {d['ms_per_step']:.4f} ms  frac {d['roofline']['frac']:.3f}  clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(f"{label:40s} FAILED {line[:300]}")
PY
}
{
for hint in 0 1000 10000 100000; do
  run "cavity hint=$hint" cavity 6 XGB_WAIT_HINT=$hint
done
run "cavity hint=0 (repeat)" cavity 6 XGB_WAIT_HINT=0
for hint in 0 10000; do
  run "heat3d hint=$hint" heat3d 20 XGB_WAIT_HINT=$hint
  run "diff2d two-step hint=$hint" diff2d 50 XGB_WAIT_HINT=$hint
done
} 2>&1 | tee $O/r2p_session13.txt
