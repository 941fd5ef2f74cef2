#!/bin/bash
# A/B of gen_kernel launch variants in one process (tools/gen_ab.py: warps[:waves[:tail split]]), then the GPU suite
T=${1:-r02d}; V2=${2:-28,28::2,28::4,28::3,28}; V4=${3:-28,28::2}
timeout 300 python tools/gen_ab.py --workload cfg2 --steps 30 --variants $V2 > gpurun_out/${T}_ab_cfg2.jsonl 2> gpurun_out/${T}_ab_cfg2.err
cut -c1-150 gpurun_out/${T}_ab_cfg2.jsonl; python -c "
import json,sys
for l in open('gpurun_out/${T}_ab_cfg2.jsonl'): d=json.loads(l); print(d['variant'], d['bit_identical_to_first'])"
timeout 300 python tools/gen_ab.py --workload cfg4 --pop 1000 --steps 6 --warmup 3 --variants $V4 > gpurun_out/${T}_ab_cfg4.jsonl 2> gpurun_out/${T}_ab_cfg4.err
cut -c1-150 gpurun_out/${T}_ab_cfg4.jsonl
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
