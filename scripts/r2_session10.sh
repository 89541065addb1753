#!/bin/bash
# heat3d: L2 eviction hints on the producer's bulk copies (rows shared with the neighbouring j-tile evict_last)
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
    print(f"{label:52s} {d['value']:7.1f} Gpt/s  {d['ms_per_step']:.4f} ms  frac {d['roofline']['frac']:.3f}  clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(f"{label:52s} FAILED {line[:300]}")
PY
}
{
run "no hints" XGB_TILED_L2HINT=0
run "shared rows evict_last" XGB_TILED_L2HINT=1
run "shared rows evict_last, others evict_first" XGB_TILED_L2HINT=2
run "no hints (repeat)" XGB_TILED_L2HINT=0
run "shared rows evict_last (repeat)" XGB_TILED_L2HINT=1
run "hint 1, TJ=4 smem 72K (3 CTAs/SM)" XGB_TILED_L2HINT=1 XGB_TJ=4 XGB_SMEM=73728
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
for h in 0 1 2; do
  XGB_TILED_L2HINT=$h timeout 300 ncu --metrics $M --clock-control none -k regex:heat_3d.*tiled -s 3 -c 2 --csv --log-file $O/r2l_l2hint_$h.csv \
      python bench.py --workload heat3d --steps 3 --warmup 3 --no-cpu --no-e2e --no-parity > /dev/null 2>&1
  echo "ncu hint=$h:"; grep -o '"dram__bytes_read.sum","[A-Za-z]*","[0-9.,]*"\|"gpu__time_duration.sum","[a-z]*","[0-9.,]*"\|"lts__t_sector_hit_rate.pct","%","[0-9.]*"' $O/r2l_l2hint_$h.csv | head -3
done
} 2>&1 | tee $O/r2l_session10.txt
