#!/bin/bash
# usage: scripts_sweep.sh "<cpt list>" "<tile list>" [extra bench args]
for cpt in $1; do for tile in $2; do echo "CPT=$cpt TILE=$tile"; GSGP_INTERP_TILE=$tile timeout 180 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('value %.3e e2e %.3e ms/step %.3f gsop_ms %.3f interp_ms %.3f frac %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step'],r['ms_total']/r['launches'],r['interp_kernel']['ms_total']/max(1,r['interp_kernel']['launches']),r['frac']))
    else: print(l.rstrip())
"; done; done
