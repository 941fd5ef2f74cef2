#!/bin/bash
# A/B of gen_kernel launch variants in one process (tools/gen_ab.py: warps[:waves]), bit-identical fitness checked
T=${1:-r02d}; V2=${2:-28,24,20,28:12,28}; V4=${3:-28,24}
timeout 300 python tools/gen_ab.py --workload cfg2 --steps 30 --variants $V2 > gpurun_out/${T}_ab_cfg2.jsonl 2> gpurun_out/${T}_ab_cfg2.err
timeout 300 python tools/gen_ab.py --workload cfg4 --pop 1000 --steps 6 --warmup 3 --variants $V4 > gpurun_out/${T}_ab_cfg4.jsonl 2> gpurun_out/${T}_ab_cfg4.err
python -c "
import json
for f in ('cfg2','cfg4'):
    for l in open('gpurun_out/${T}_ab_'+f+'.jsonl'): d=json.loads(l); print(f, d['variant'], round(d['gen_kernel_ms_per_launch'],4), d['bit_identical_to_first'])"
