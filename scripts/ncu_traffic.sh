#!/bin/bash
# One `ncu --set full` capture of the dominant kernel at the bench's own batch / size; prints its DRAM bytes.
# usage: ncu_traffic.sh <kernel regex> <launch index of that kernel within one forward pass> <tag>
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:$1 -s $2 -c 1 -f -o gpurun_out/$3 \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$3.log 2>&1
ncu -i gpurun_out/$3.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; v=r[-1]
d={n:v[i] for i,n in enumerate(h)}
for k in ('Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__grid_size'): print(k, d.get(k))
" | tee gpurun_out/traffic_$3.txt
