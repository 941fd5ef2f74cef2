#!/usr/bin/env python
"""Summarise an ncu report here (no GPU): details + instruction mix + stall regions."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
keys = ("Duration", "Registers Per", "Theoretical Occ", "Achieved Occ", "Executed Ipc Active", "Issue Slots Busy", "No Eligible",
        "L1/TEX Hit", "Dynamic Shared Memory Per", "Block Limit Reg", "Block Limit Shared", "Warp Cycles Per Issued", "DRAM Throughput",
        "Compute (SM) Throughput", "Memory Throughput", "L2 Hit")
for l in det.splitlines():
    if any(k in l for k in keys): print(l.strip())
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; isrc = h.index('Source'); ie = h.index('Instructions Executed'); iss = h.index('Warp Stall Sampling (All Samples)')
data = [(r[isrc].strip(), int(r[ie]), int(r[iss])) for r in rows[2:] if len(r) > ie]
tot = sum(d[1] for d in data); st = sum(d[2] for d in data)
print('total warp-inst', tot, 'stall samples', st, 'sass lines', len(data))
ops = collections.Counter()
for s_, e, _ in data:
    t = s_.split()
    op = t[0] if not t[0].startswith('@') else t[1]
    ops[op.split('.')[0]] += e
print(' '.join(f'{k}:{v/tot:.3f}' for k, v in ops.most_common(14)))
stall_cols = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
totst = collections.Counter()
for r in rows[2:]:
    if len(r) <= ie: continue
    for i in stall_cols:
        try: totst[h[i]] += int(r[i])
        except ValueError: pass
print(' '.join(f'{k}:{v}' for k, v in totst.most_common(8)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for i in sorted(range(len(data)), key=lambda i: -data[i][2])[:n]:
    print(f'{i:5d} exec={data[i][1]:9d} stall={data[i][2]:6d} {data[i][0][:90]}')
if len(sys.argv) > 3:
    a, b = map(int, sys.argv[3].split(':'))
    for i in range(a, b): print(f'{i:5d} exec={data[i][1]:9d} stall={data[i][2]:6d} {data[i][0][:90]}')
