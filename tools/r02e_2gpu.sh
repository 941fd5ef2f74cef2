#!/bin/bash
# two ranks on one box after the gen_kernel change: parity of the sharded path against the oracle, then the config-2 line
T=${1:-r02e}
GSGP_EXCHANGE=p2p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29902 \
    tests/multigpu_check.py > gpurun_out/${T}_parity_2gpu_p2p.log 2>&1; tail -1 gpurun_out/${T}_parity_2gpu_p2p.log
O=gpurun_out/${T}_bench_cfg2_2gpu
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 50 --warmup 5 --no-unfused > $O.json 2> $O.err
python tools/show_bench.py $O.json
