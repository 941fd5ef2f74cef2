#!/bin/bash
# usage: scripts_prof.sh <tag> <kernel regex> [skip]   -> gpurun_out/<tag>.ncu-rep (+ plain log)
C="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 200 $C > gpurun_out/$1_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$2" -s ${3:-4} -c 1 -o gpurun_out/$1 timeout 400 $C > gpurun_out/$1_ncu.log 2>&1
tail -2 gpurun_out/$1_ncu.log
