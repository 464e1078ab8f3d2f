#!/bin/bash
# ncu --set full of one iteration's kernels at the bench configuration; args: tag, kernel regex, extra env (A=1,B=2), bench args
OUT=gpurun_out/${1:-r2ncu}; mkdir -p $OUT
REGEX=${2:-k_}
ENVS=$(echo ${3:-X=0} | tr ',' ' ')
shift; shift; shift
env $ENVS timeout 900 ncu --set full --clock-control none --import-source on -k regex:$REGEX -s 0 -c 12 -o $OUT/prof -f \
    python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-secondary "$@" > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log
ls -la $OUT
