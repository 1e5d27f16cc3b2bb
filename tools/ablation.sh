#!/bin/bash
# Ablation of the fused kernels (DESIGN.md section 7 "where the time goes"): builds librrtmgp_b200.so variants with one part of
# the kernels removed (-DRB_WHATIF=n, see csrc/solver_fast.cuh) into variants/ and, with `run`, times each of them with
# bench.py on the GPU.  The variants compute wrong fluxes by construction; they exist to measure elapsed-time costs.
#   tools/ablation.sh build "1 2 3 8 4 9 5 6"      (here: nvcc cross-compiles)
#   gpurun -- 'tools/ablation.sh run "1 2 3 8 4 9 5 6"'   -> gpurun_out/ablation.txt
set -e
cd "$(dirname "$0")/.."
C=rrtmgp.jl_b200/csrc
L=$C/librrtmgp_b200.so
if [ "$1" = build ]; then
  mkdir -p variants
  F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 177"
  for w in $2; do
    ( cd $C
      nvcc $F -DRB_WHATIF=$w -DRB_MODE=0 -DRB_NGPT=256 -DRB_NG=1 -DRB_ENTRY=launch_fast_lw_ng1 -c solver_fast_inst.cu -o /tmp/w${w}_lw.o &
      nvcc $F -DRB_WHATIF=$w -DRB_MODE=2 -DRB_NGPT=224 -DRB_NG=1 -DRB_ENTRY=launch_fast_sw_ng1 -c solver_fast_inst.cu -o /tmp/w${w}_sw.o &
      wait
      nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/lib_w$w.so api.o lut.o solver.o solver_tm.o peak.o \
        /tmp/w${w}_lw.o fast_lw_ng2.o /tmp/w${w}_sw.o fast_sw_ng2.o fast_noscat1_ng1.o fast_noscat1_ng2.o fast_noscat4_ng1.o \
        fast_noscat4_ng2.o ws_lw_ng1.o ws_lw_ng2.o ws_sw_ng1.o ws_sw_ng2.o -lcudart -ldl ) &
  done
  wait
  ls -la variants
elif [ "$1" = run ]; then
  B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --no-sweep"
  P='import sys,json; d=json.loads(sys.stdin.readline()); print("%.4g col/s" % d["value"], d["kernel_ms"])'
  cp $L /tmp/orig.so
  trap 'cp /tmp/orig.so '$L EXIT
  echo -n "shipped   " >> gpurun_out/ablation.txt; $B 2>/dev/null | python -c "$P" >> gpurun_out/ablation.txt
  for w in $2; do
    cp variants/lib_w$w.so $L
    echo -n "whatif=$w  " >> gpurun_out/ablation.txt; $B 2>/dev/null | python -c "$P" >> gpurun_out/ablation.txt
  done
else
  echo "usage: $0 build|run \"variants\""; exit 2
fi
