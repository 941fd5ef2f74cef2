#!/bin/bash
# after the host-path / selection-folding changes: multi-GPU parity again (2 and 8 ranks) and the config-2 line at 8 GPUs
T=r02mg2
for N in 8 2; do
  GSGP_EXCHANGE=p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29900 + N)) \
    tests/multigpu_check.py > gpurun_out/${T}_parity_${N}gpu_p2p.log 2>&1; tail -1 gpurun_out/${T}_parity_${N}gpu_p2p.log
done
GSGP_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29931 \
    tests/multigpu_check.py > gpurun_out/${T}_parity_8gpu_nccl.log 2>&1; tail -1 gpurun_out/${T}_parity_8gpu_nccl.log
for N in 8 4 2 1; do
  O=gpurun_out/${T}_bench_cfg2_${N}gpu
  if [ $N = 1 ]; then timeout 400 python bench.py --steps 50 --warmup 5 > $O.json 2> $O.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps 50 --warmup 5 > $O.json 2> $O.err; fi
  python tools/show_bench.py $O.json
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --workload cfg4 --steps 8 --warmup 3 --no-cpu-baseline --no-unfused > gpurun_out/${T}_bench_cfg4_8gpu.json 2> gpurun_out/${T}_bench_cfg4_8gpu.err; python tools/show_bench.py gpurun_out/${T}_bench_cfg4_8gpu.json
