#!/bin/bash
# One gpurun call: parity tests, smoke, bench, launch list.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
if [ "${SANITIZE:-0}" = "1" ]; then
  echo "== compute-sanitizer (small forward)"
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/small_forward.py 2>&1 | tail -15 | tee gpurun_out/sanitizer.log
fi
echo "== bench" ; timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 --profile-ops > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json; tail -40 gpurun_out/bench.err
if [ "${NCU:-1}" = "1" ]; then
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
  tail -3 gpurun_out/ncu_bench.log
fi
