#!/bin/bash
# Round 2: halo exchange over peer memory (csrc/xgb_peer.cu) against ncclSend/ncclRecv at N = $1 (default 2):
# the sharded test worker under both transports, then the driver's bench command under both.
N=${1:-2}
O=gpurun_out
mkdir -p $O
export XGB_PEER_TIMEOUT_S=30
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
echo "== tests/test_dist.py + test_peer_gpu.py"
timeout 900 python -m pytest tests/test_dist.py tests/test_peer_gpu.py -m gpu -x -q 2>&1 | tail -15
for H in peer nccl; do
  echo "== bench (driver command), N=$N, XGB_HALO=$H"
  ( time XGB_HALO=$H timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-e2e ) > $O/r2p_n${N}_${H}.json 2> $O/r2p_n${N}_${H}.err
  python scripts/results_table.py $O/r2p_n${N}_${H}.json 2>/dev/null | head -30 || tail -c 1500 $O/r2p_n${N}_${H}.json
  tail -4 $O/r2p_n${N}_${H}.err
done
