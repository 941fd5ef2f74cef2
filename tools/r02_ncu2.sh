#!/bin/bash
A="python tools/gen_ab.py --variants 24 --steps 3 --warmup 3"
timeout 200 $A --workload cfg2 > gpurun_out/r02v_cfg2_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gen_kernel -s 4 -c 1 -o gpurun_out/r02v_gen_cfg2 timeout 400 $A --workload cfg2 > gpurun_out/r02v_cfg2_ncu.log 2>&1
timeout 200 $A --workload cfg4 --pop 1000 > gpurun_out/r02v_cfg4_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gen_kernel -s 4 -c 1 -o gpurun_out/r02v_gen_cfg4 timeout 400 $A --workload cfg4 --pop 1000 > gpurun_out/r02v_cfg4_ncu.log 2>&1
ls -la gpurun_out/r02v_*
