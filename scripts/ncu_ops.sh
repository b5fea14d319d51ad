#!/bin/bash
# usage: ncu_ops.sh <kernel regex> "<skip list>" <tag>   — one full-set capture per skip index (launch order of that kernel)
mkdir -p gpurun_out
for s in $2; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$1 -s $s -c 1 -f -o gpurun_out/$3_$s \
    python bench.py --steps 1 --warmup 1 --batch ${BATCH:-64} --no-cpu-baseline > gpurun_out/ncu_$3_$s.log 2>&1
  tail -1 gpurun_out/ncu_$3_$s.log | cut -c1-100
done
ls -la gpurun_out | head -30
