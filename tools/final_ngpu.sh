#!/bin/bash
# usage: tools/final_ngpu.sh <tag> <n_gpus> <workload> <steps> [exchange] [share]   (run under gpurun --gpus N)
#        share = 1: rank 0 plans for every rank (GSGP_BENCH_SHARE_PLAN) -- compare the e2e value with share = 0
T=$1; N=$2; W=$3; S=$4; EX=${5:-p2p}; SH=${6:-0}
O=gpurun_out/${T}_bench_${W}_${N}gpu_${EX}$([ "$SH" = 1 ] && echo _shared)
GSGP_BENCH_SHARE_PLAN=$SH GSGP_EXCHANGE=$EX timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --workload $W --steps $S --warmup 5 > $O.json 2> $O.err || { echo "FAILED $O"; tail -5 $O.err; }
python - $O.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "n_gpus %d exchange %s value %.4g ms/step %.4g e2e %.4g frac %.3f clocks %s" % (d["n_gpus"], d["engine"]["exchange"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"]))
PY
