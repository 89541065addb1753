#!/bin/bash
N=${1:-2}
export XGB_PEER_TIMEOUT_S=30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571"
for cfg in "XGB_HALO=nccl" "XGB_HALO=peer" "XGB_HALO=peer XGB_PEER_CTAS=16" "XGB_HALO=peer XGB_PEER_CTAS=148"; do
  env $cfg timeout 200 $TR scripts/peer_probe.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -5
done
