#!/bin/bash
# e2e columns/s of the host-buffer pipeline vs its chunk count (1 GPU)
for c in 10 13 16 20 26 32 40; do
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sweep --no-variants --e2e-chunks $c 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; print('chunks=$c', '%.4g' % e['value'], '%.2f ms' % e['ms_per_step'], e['how'][:40], '| device %.4g' % d['value'])" >> gpurun_out/e2e_chunks.txt
done
cat gpurun_out/e2e_chunks.txt
