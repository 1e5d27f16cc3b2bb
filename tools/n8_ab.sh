#!/bin/bash
# same-box A/B of the gathered update's shortwave chunking at N GPUs ($1)
N=$1
for c in 3 4 3 4; do
  RRTMGP_B200_GATHER_CHUNKS=$c python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-sweep --no-e2e --no-variants 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('chunks=$c', '%.4g' % d['value'], '%.3f ms' % d['ms_per_step'])" >> gpurun_out/n8_ab.txt
done
cat gpurun_out/n8_ab.txt
