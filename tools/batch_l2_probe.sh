#!/bin/bash
# Per-image kernel times of the FFT engine against the batch size (is a half spectrum that fits the 126 MB L2 faster?)
OUT=gpurun_out/${1:-l2probe}; mkdir -p $OUT
for B in 1 2 3 4 6 8 32; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --dist mosaic --batch $B > $OUT/o.json 2> $OUT/o.err || tail -5 $OUT/o.err
  python -c "
import json;d=json.load(open('$OUT/o.json'));B=$B;print(B,round(d['value']),round(d['ms_per_step'],3),{k:round(v*1000/B/3,2) for k,v in d['roofline']['kernels_ms_per_step'].items() if v>0.01}, 'us per image-iteration')" | tee -a $OUT/probe.log
done
