#!/bin/bash
# compute-sanitizer memcheck / racecheck / synccheck over the compile-time-plan kernels -> gpurun_out/san/sanitizer.txt
mkdir -p gpurun_out/san
: > gpurun_out/san/sanitizer.txt
for T in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $T python tools/sanitize_static_plans.py" >> gpurun_out/san/sanitizer.txt
  timeout 600 compute-sanitizer --tool $T python tools/sanitize_static_plans.py 2>&1 | grep -v "^=========     " | tail -14 >> gpurun_out/san/sanitizer.txt
done
cat gpurun_out/san/sanitizer.txt | grep -i "summary\|error\|hazard" | head -20
