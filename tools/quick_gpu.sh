#!/bin/bash
# GPU parity suite + kernel-time line of bench.py (no CPU baseline / e2e / sweep); tag = $1
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/$1_gputests.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-sweep > gpurun_out/$1_bench.json 2> gpurun_out/$1_bench.err
cat gpurun_out/$1_gputests.txt
python -c "
import json; d=json.load(open('gpurun_out/$1_bench.json')); print('%.4g' % d['value'], d['kernel_ms'], {k: '%.4g' % v['value'] for k, v in d['variants'].items()})"
