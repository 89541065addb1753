#!/bin/bash
# on-GPU tuning sweep: prints workload, env, Gpt/s, roofline fraction
run() { # name shape... -- env
  local w=$1; shift
  out=$(env "$@" timeout 300 python bench.py --workload $w --no-cpu --no-e2e --steps ${STEPS:-100} --warmup 5 ${SHAPE:+--shape $SHAPE} 2>&1 | tail -1)
  echo "$w $* :: $(echo "$out" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("%.1f Gpt/s  frac=%.3f  ms=%.4f"%(d["value"], d["roofline"]["frac"], d["ms_per_step"]))
except Exception as e: print("ERR", e)')"
}
