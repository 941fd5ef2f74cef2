#!/bin/bash
# Round 2, after the slot-table / 28-warp gen_kernel: block sizes side by side, the GPU suite, the default bench line, one
# ncu --set full capture of gen_kernel (config 2, and config 4's trees) and the launch list of the bench command.
T=${1:-r02c}
timeout 300 python tools/gen_ab.py --workload cfg2 --steps 30 --variants 28,24,20,28 > gpurun_out/${T}_ab_cfg2.jsonl 2> gpurun_out/${T}_ab_cfg2.err
cut -c1-150 gpurun_out/${T}_ab_cfg2.jsonl
timeout 300 python tools/gen_ab.py --workload cfg4 --pop 1000 --steps 6 --warmup 3 --variants 28,24 > gpurun_out/${T}_ab_cfg4.jsonl 2> gpurun_out/${T}_ab_cfg4.err
cut -c1-150 gpurun_out/${T}_ab_cfg4.jsonl
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/${T}_bench_cfg2_1gpu.json 2> gpurun_out/${T}_bench_cfg2_1gpu.err; python tools/show_bench.py gpurun_out/${T}_bench_cfg2_1gpu.json
A="python tools/gen_ab.py --variants 28 --steps 3 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:gen_kernel -s 4 -c 1 -o gpurun_out/${T}_gen_cfg2 timeout 400 $A --workload cfg2 > gpurun_out/${T}_cfg2_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gen_kernel -s 4 -c 1 -o gpurun_out/${T}_gen_cfg4 timeout 400 $A --workload cfg4 --pop 1000 > gpurun_out/${T}_cfg4_ncu.log 2>&1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv timeout 400 $B > gpurun_out/${T}_bench_ncu.log 2>&1
ls gpurun_out/${T}_*
