#!/bin/bash
# round-end check on one GPU (tag = $1): what the driver runs (GPU tests, smoke, both bench arms) + the ncu evidence
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/$1_gputests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/$1_smoke.txt 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/$1_bench_reference.json 2> gpurun_out/$1_bench_reference.err
python bench.py > gpurun_out/$1_bench.json 2> gpurun_out/$1_bench.err
tools/profile_final.sh $1 > /dev/null 2>&1
cat gpurun_out/$1_gputests.txt; tail -2 gpurun_out/$1_smoke.txt
python -c "
import json
r=json.load(open('gpurun_out/$1_bench_reference.json')); d=json.load(open('gpurun_out/$1_bench.json'))
print('reference', r['value'], r['cpu_baseline']['cores'])
print('engine', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['kernel_ms'])
print(d['roofline']); print(d['roofline_fp32']); print(d['roofline_issue']); print(d['cpu_baseline']); print(d['clocks']); print(d['sweep']); print(d['variants'])"
