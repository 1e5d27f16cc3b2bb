#!/bin/bash
# ncu evidence for the build in the tree (tag = $1): launch list of the bench command + one full capture of the two hot kernels
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --no-sweep"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$1_launches.csv $B > gpurun_out/$1_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:solve_kernel_fast -c 2 -o gpurun_out/prof_$1 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --no-sweep > gpurun_out/prof_$1.log 2>&1
ls -la gpurun_out/prof_$1.ncu-rep
