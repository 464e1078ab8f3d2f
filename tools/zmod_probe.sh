#!/bin/bash
# Timing of the FFT passes with the half spectrum folded onto two L2-resident slots: build the second library first with
#   make -C polyblur_b200/csrc BUILD=build_zmod2 OUT=../libpb_zmod2.so EXTRA=-DPB_ZMOD=2      (results are wrong by design)
mkdir -p gpurun_out/zmod
for L in polyblur_b200/libpolyblur_sm100.so polyblur_b200/libpb_zmod2.so; do
  PB_LIB_PATH=$PWD/$L timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary --dist mosaic --engine 2 > gpurun_out/zmod/o.json 2> gpurun_out/zmod/o.err || tail -5 gpurun_out/zmod/o.err
  python -c "
import json;d=json.load(open('gpurun_out/zmod/o.json'));print('$L',round(d['value']),round(d['ms_per_step'],3),{k:round(v,3) for k,v in d['roofline']['kernels_ms_per_step'].items() if v>0.01})" | tee -a gpurun_out/zmod/probe.log
done
