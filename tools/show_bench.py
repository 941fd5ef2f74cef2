#!/usr/bin/env python
"""One-line summaries of bench.py JSON lines:  python tools/show_bench.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as ex:
        print(path, "unreadable:", repr(ex)[:100])
        continue
    if d.get("impl") == "reference":
        print(f"{path}: reference arm value {d['value']:.4g} ({d['cpu_baseline']['cores']} cores) {d['config']['workload'][:40]}")
        continue
    r, e = d["roofline"], d["e2e"]
    ser = e.get("serial_host_loop") or {}
    cpu = d.get("cpu_baseline") or {}
    print(f"{path}: n_gpus {d['n_gpus']} {d['scaling']} value {d['value']:.4g} ({d['ms_per_step']:.4g} ms/gen) e2e {e['value']:.4g} "
          f"({e['ms_per_step']:.4g} ms; serial loop {ser.get('ms_per_step', float('nan')):.4g} ms) "
          f"gen_kernel {r['ms_per_launch']:.4g} ms frac {r['frac']:.3f} (8d {r['survey_8d']['frac']:.3f}) "
          f"unfused gsop frac {r['gs_kernel_unfused']['frac']} cpu {cpu.get('value', float('nan')):.3g} "
          f"exchange {d['engine']['exchange']} store {d['engine']['store']['buffers']} clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']} "
          f"planner {e['plan_builder'][-40:]}")
