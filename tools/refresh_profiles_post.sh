#!/bin/bash
# Condenses gpurun_out/refresh/ (tools/refresh_profiles.sh) into the tracked files under profiles/ (round tag $1).
set -u
R=${1:-r02}; O=gpurun_out/refresh; P=profiles
cp $O/pytest_gpu.log $P/${R}_pytest_gpu.log
for f in bench_mosaic bench_white bench_reference_arm bench_C3 bench_C4 bench_C5; do [ -s $O/$f.json ] && cp $O/$f.json $P/${R}_$f.json; done
for f in config_sweep parity_report fuzz_parity; do [ -s $O/$f.jsonl ] && cp $O/$f.jsonl $P/${R}_$f.jsonl; done
[ -s $O/e2e_probe.json ] && cp $O/e2e_probe.json $P/${R}_e2e_probe.json
[ -s $O/rf_timing.txt ] && cp $O/rf_timing.txt $P/${R}_rf_timing.txt
grep -v "^==" $O/launches_mosaic_b32.csv > $P/${R}_launches_mosaic_b32.csv
python tools/launch_summary.py $O/launches_mosaic_b32.csv $P/${R}_launches_mosaic_b32.md "ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400: python bench.py --steps 2 --warmup 1 --no-graph (C2 mosaic)"
python tools/ncu_summary.py $O/prof_mosaic32.raw.csv $P/${R}_ncu_mosaic_b32.md --traffic mosaic:32x1080x1920 $P/roofline_traffic.json --pipes mosaic:32x1080x1920 $P/roofline_pipes.json --csv $P/${R}_ncu_mosaic_b32.csv
python tools/ncu_summary.py $O/prof_white32.raw.csv $P/${R}_ncu_white_b32.md --traffic white:32x1080x1920 $P/roofline_traffic.json --pipes white:32x1080x1920 $P/roofline_pipes.json --csv $P/${R}_ncu_white_b32.csv
[ -s $O/prof_mosaic_c3.raw.csv ] && python tools/ncu_summary.py $O/prof_mosaic_c3.raw.csv $P/${R}_ncu_mosaic_4k_b8.md --pipes mosaic:8x2160x3840 $P/roofline_pipes.json --csv $P/${R}_ncu_mosaic_4k_b8.csv
[ -s $O/prof_mosaic_c4.raw.csv ] && python tools/ncu_summary.py $O/prof_mosaic_c4.raw.csv $P/${R}_ncu_mosaic_c4.md --pipes mosaic:1x9000x12000 $P/roofline_pipes.json --csv $P/${R}_ncu_mosaic_c4.csv
for rep in prof_mosaic32 prof_white32 prof_mosaic_c3 prof_mosaic_c4; do
  [ -s $O/$rep.sass_mix.md ] && cp $O/$rep.sass_mix.md $P/${R}_sass_mix_${rep#prof_}.md
done
# SASS evidence: mnemonic counts of the shipped library (whole library and the hot kernels)
python tools/sass_listing.py polyblur_b200/libpolyblur_sm100.so $P/${R}_sass_summary.md
ls -la $P | tail -30
