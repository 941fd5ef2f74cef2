#!/bin/bash
# four ranks on one box with the final gen_kernel: parity of the sharded path against the oracle (peer-store exchange)
T=${1:-r02e}
GSGP_EXCHANGE=p2p timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29904 \
    tests/multigpu_check.py > gpurun_out/${T}_parity_4gpu_p2p.log 2>&1; tail -1 gpurun_out/${T}_parity_4gpu_p2p.log
