#!/bin/bash
for w in 2 4 8 12 16 24; do echo -n "waves=$w  "; GSGP_GEN_WAVES=$w timeout 120 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); r=d['roofline']; print('value %.3e ms/step %.4f gen_ms %.4f launches/step %d'%(d['value'],d['ms_per_step'],r['ms_total']/30, r['launches']/30))"; done
