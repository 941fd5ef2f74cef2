#!/bin/bash
# usage: tools/final_1gpu.sh <tag>   -> gpurun_out/<tag>_bench_<cfg>.json for every single-GPU workload
T=$1
for w in cfg2 cfg1 cfg3 cfg4 cfg5 cfg5s cfg5full; do
  S=20; [ $w = cfg2 ] && S=50; [ $w = cfg1 ] && S=100; [ $w = cfg5s ] && S=10; [ $w = cfg5full ] && S=8
  SECONDS=0
  timeout 600 python bench.py --workload $w --steps $S --warmup 5 > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err || { echo "FAILED $w"; tail -3 gpurun_out/${T}_bench_$w.err; }
  echo "$w: $SECONDS s"; python tools/show_bench.py gpurun_out/${T}_bench_$w.json
done
