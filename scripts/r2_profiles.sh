#!/bin/bash
# Round 2 profiles: launch list of the driver's command, DRAM traffic per launch of every dominant kernel variant.
O=gpurun_out
mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
# launch list of the default command (sub-records off: the cavity records alone add ~10 000 launches)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_bench_launches.csv \
    python bench.py --gpus 1 --steps 20 --warmup 5 --extra none --no-cpu > $O/r2_bench_launches.out 2>&1
t() {  # name, ncu -c count, bench args...
  name=$1; cnt=$2; shift; shift
  timeout 600 ncu --metrics $M --clock-control none -k regex:xg_ -c $cnt --csv --log-file $O/r2_traffic_$name.csv \
      python bench.py "$@" --no-cpu --no-e2e --no-parity > $O/r2_traffic_$name.out 2>&1
}
t heat3d 6 --workload heat3d --steps 3 --warmup 3
t conv1d 6 --workload conv1d --steps 128 --warmup 128
t conv1d_tail 4 --workload conv1d --steps 20 --warmup 20
t conv1d_onepass 8 --workload conv1d --steps 4 --warmup 4 --no-temporal
t conv1d_nl 6 --workload conv1d_nl --steps 128 --warmup 128
t diff1d 6 --workload diff1d --steps 128 --warmup 128
t conv2d 6 --workload conv2d --steps 6 --warmup 6
t conv2d_onepass 6 --workload conv2d --steps 3 --warmup 3 --no-temporal
t diff2d 6 --workload diff2d --steps 6 --warmup 6
t diff2d_onepass 6 --workload diff2d --steps 3 --warmup 3 --no-temporal
t cavity 1700 --workload cavity --steps 1 --warmup 1
ls -la $O/r2_traffic_*.csv | head -20
