#!/bin/bash
# Round 2, final gen_kernel (slot table, 28 warps, copies third in the claim order): the GPU suite, the default bench line,
# the other one-GPU shapes, one ncu --set full capture of gen_kernel (config 2, and config 4's trees), the launch list.
T=${1:-r02e}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 400 python bench.py --steps 50 --warmup 5 > gpurun_out/${T}_bench_cfg2_1gpu.json 2> gpurun_out/${T}_bench_cfg2_1gpu.err; python tools/show_bench.py gpurun_out/${T}_bench_cfg2_1gpu.json
for w in cfg3 cfg4 cfg5; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-unfused > gpurun_out/${T}_bench_${w}_1gpu.json 2> gpurun_out/${T}_bench_${w}_1gpu.err; python tools/show_bench.py gpurun_out/${T}_bench_${w}_1gpu.json
done
A="python tools/gen_ab.py --variants 28 --steps 3 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:gen_kernel -s 4 -c 1 -o gpurun_out/${T}_gen_cfg2 timeout 400 $A --workload cfg2 > gpurun_out/${T}_cfg2_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gen_kernel -s 4 -c 1 -o gpurun_out/${T}_gen_cfg4 timeout 400 $A --workload cfg4 --pop 1000 > gpurun_out/${T}_cfg4_ncu.log 2>&1
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv timeout 400 $B > gpurun_out/${T}_bench_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
