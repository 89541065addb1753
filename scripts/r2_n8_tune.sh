#!/bin/bash
# N=8 heat3d: where do the extra ~0.2 ms per step come from?  NCCL channel count, overlap on/off, per-rank times.
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
run() {
  label=$1; shift
  line=$(env "$@" timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e --no-parity 2>$O/n8_err.txt | tail -1)
  python - "$label" "$line" <<'PY'
import json, sys
label, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    print(f"{label:36s} {d['value']:7.1f} Gpt/s  {d['ms_per_step']:.3f} ms  by rank {d['clocks'].get('ms_per_step_by_rank')}  clk {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(f"{label:36s} FAILED {line[:300]}")
PY
}
{
run "default" A=1
run "NCCL_MAX_NCHANNELS=4" NCCL_MAX_NCHANNELS=4
run "overlap off" XGB_OVERLAP=0
run "default (repeat)" A=1
echo "== cavity 8192^2 over 8 GPUs (fused pairs on slabs)"
timeout 300 $TR bench.py --gpus 8 --workload cavity --shape 1024 8192 --steps 6 --warmup 4 --no-e2e > $O/r2h_n8_cavity.json 2>$O/r2h_n8_cavity.err; tail -c 900 $O/r2h_n8_cavity.json
} 2>&1 | tee $O/r2h_n8_tune.txt
