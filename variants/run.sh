#!/bin/bash
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --no-sweep"
L=rrtmgp.jl_b200/csrc/librrtmgp_b200.so
cp $L /tmp/orig.so
echo -n "base " >> gpurun_out/whatif.txt
$B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['kernel_ms'])" >> gpurun_out/whatif.txt
for w in $VARIANTS; do
  cp variants/lib_w$w.so $L
  echo -n "whatif=$w " >> gpurun_out/whatif.txt
  $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['value'], d['kernel_ms'])" >> gpurun_out/whatif.txt
done
cp /tmp/orig.so $L
