#!/bin/bash
# A/B of the halo transports on the main line (heat3d slab per GPU), N = $1 (default 2).
N=${1:-2}
O=gpurun_out
mkdir -p $O
export XGB_PEER_TIMEOUT_S=30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
run() {  # label, env...
  L=$1; shift
  env "$@" timeout 300 $TR bench.py --gpus $N --steps 40 --warmup 5 --no-e2e --no-parity --extra none > $O/r2q_n${N}_$L.json 2> $O/r2q_n${N}_$L.err
  python - "$O/r2q_n${N}_$L.json" "$L" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"{sys.argv[2]:>14}: {d['ms_per_step']:.4f} ms/step  {d['value']:.1f} Gpt/s  by rank {d['clocks'].get('ms_per_step_by_rank')}  {d['clocks']['reasons']}  {d.get('halo_transport','')[:12]}")
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
run peer_a XGB_HALO=peer
run nccl_a XGB_HALO=nccl
run peer_b XGB_HALO=peer
run nccl_b XGB_HALO=nccl
run peer_nograph XGB_HALO=peer XGB_BENCH_GRAPHS=0
run nccl_nograph XGB_HALO=nccl XGB_BENCH_GRAPHS=0
run peer_c16 XGB_HALO=peer XGB_PEER_CTAS=16
tail -3 $O/r2q_n${N}_peer_a.err
