#!/bin/bash
# Round 2, one `gpurun --gpus 8` call: multi-GPU parity at 2/4/8 ranks, the strong-scaling series of config 5 (10M cases in
# total), the plan-sharing A/B, configs 2-4 at 2/4/8 GPUs.  Everything lands in gpurun_out/r02mg_*.
T=r02mg
run() {   # run <n_gpus> <workload> <steps> <tag> [env...]
  local N=$1 W=$2 S=$3 TAG=$4; shift 4
  local O=gpurun_out/${T}_bench_${W}_${N}gpu${TAG}
  if [ "$N" = 1 ]; then
    env "$@" timeout 400 python bench.py --workload $W --steps $S --warmup 3 --no-cpu-baseline > $O.json 2> $O.err || { echo "FAILED $O"; tail -5 $O.err; }
  else
    env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $N --workload $W --steps $S --warmup 3 --no-cpu-baseline --no-unfused > $O.json 2> $O.err || { echo "FAILED $O"; tail -5 $O.err; }
  fi
  python tools/show_bench.py $O.json
}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | head -8
# 1. parity: case-sharded engines in lock step with the oracle
for N in 8 4 2; do
  for EX in p2p nccl; do
    [ $N != 8 ] && [ $EX = nccl ] && continue
    GSGP_EXCHANGE=$EX timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29900 + N)) \
      tests/multigpu_check.py > gpurun_out/${T}_parity_${N}gpu_${EX}.log 2>&1; tail -1 gpurun_out/${T}_parity_${N}gpu_${EX}.log
  done
done
# 2. strong scaling of the 10M-case configuration
for N in 8 4 2 1; do run $N cfg5s 10 ""; done
# 3. plan sharing A/B on the short generations of config 2 (weak scaling), then the other shapes
run 8 cfg2 30 "" GSGP_BENCH_SHARE_PLAN=0
run 8 cfg2 30 _shared GSGP_BENCH_SHARE_PLAN=1
run 2 cfg2 30 "" GSGP_BENCH_SHARE_PLAN=0
run 2 cfg2 30 _shared GSGP_BENCH_SHARE_PLAN=1
run 4 cfg2 30 "" GSGP_BENCH_SHARE_PLAN=0
for N in 8 4 2; do run $N cfg4 8 ""; done
for N in 8 4 2; do run $N cfg3 10 ""; done
