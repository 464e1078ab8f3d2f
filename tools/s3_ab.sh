#!/bin/bash
# session-3 helper: GPU tests (optionally), then A/B bench lines for env settings given as arguments ("A=1,B=2" each; "-" = defaults)
OUT=gpurun_out/${1:-s3}; shift; mkdir -p $OUT
if [ "$1" == "test" ]; then shift; python -m pytest tests -m gpu -q -x -k "not c4_12000" > $OUT/pytest.log 2>&1; tail -4 $OUT/pytest.log; fi
for E in "$@"; do
  for D in mosaic white; do
    env $(echo $E | tr ',' ' ' | sed 's/^-$/PB_NOP=1/') timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --dist $D $BENCH_ARGS > $OUT/o.json 2> $OUT/o.err || tail -5 $OUT/o.err
    python -c "
import json;d=json.load(open('$OUT/o.json'));print('$E','$D',round(d['value']),round(d['ms_per_step'],3),{k:round(v,3) for k,v in d['roofline']['kernels_ms_per_step'].items() if v>0.1})" | tee -a $OUT/ab.log
  done
done
