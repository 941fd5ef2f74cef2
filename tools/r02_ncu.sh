#!/bin/bash
# Round 2 ncu evidence (1 GPU): gen_kernel at config 2, at config 4's shape (depth 8, two trees per individual) and at
# config 5's per-GPU shape (1.25M cases), plan_draw_kernel; plus the launch list of the default bench command.
set -x
A="python tools/gen_ab.py --variants 24 --steps 3 --warmup 3"
timeout 200 $A --workload cfg2 > gpurun_out/r02n_cfg2_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gen_kernel -s 4 -c 1 -o gpurun_out/r02n_gen_cfg2 timeout 400 $A --workload cfg2 > gpurun_out/r02n_cfg2_ncu.log 2>&1
timeout 200 $A --workload cfg4 --pop 1000 > gpurun_out/r02n_cfg4_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gen_kernel -s 4 -c 1 -o gpurun_out/r02n_gen_cfg4 timeout 400 $A --workload cfg4 --pop 1000 > gpurun_out/r02n_cfg4_ncu.log 2>&1
timeout 200 $A --workload cfg5 --pop 500 > gpurun_out/r02n_cfg5_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gen_kernel -s 4 -c 1 -o gpurun_out/r02n_gen_cfg5 timeout 400 $A --workload cfg5 --pop 500 > gpurun_out/r02n_cfg5_ncu.log 2>&1
timeout 200 $A --workload cfg2 > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:plan_draw_kernel -s 2 -c 1 -o gpurun_out/r02n_plan_draw timeout 400 $A --workload cfg2 > gpurun_out/r02n_draw_ncu.log 2>&1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 300 $B > gpurun_out/r02n_bench_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02n_launches.csv timeout 400 $B > gpurun_out/r02n_bench_ncu.log 2>&1
ls -la gpurun_out/r02n_*
