#!/bin/bash
# usage: tools/final_1gpu.sh <tag>   -> gpurun_out/<tag>_bench_<cfg>.json for every single-GPU workload
T=$1
for w in cfg1 cfg2 cfg3 cfg4 cfg5 cfg5full; do
  S=20; [ $w = cfg2 ] && S=50; [ $w = cfg1 ] && S=100
  timeout 400 python bench.py --workload $w --steps $S --warmup 5 > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err || echo "FAILED $w"
  python - gpurun_out/${T}_bench_$w.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]; u=r["gs_kernel_unfused"]
print(sys.argv[1], "value %.4g ms/step %.4g e2e %.4g frac %.3f gen_ms %.4g | unfused gs frac %s interp_ms %.4g | cpu %s | clocks %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], r["ms_total"]/max(1,r["launches"]), u["frac"], u["interp_kernel_ms_per_launch"],
    d["cpu_baseline"] and "%.3g (%d cores)" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]), d["clocks"]))
PY
done
