#!/bin/bash
# Re-measures everything profiles/ holds for the current build on one B200 (run through gpurun);
# raw outputs land in gpurun_out/refresh/, tools/refresh_profiles_post.sh condenses them into profiles/ afterwards.
# gpurun brings back at most 64 MiB per call and the ncu reports are ~20-40 MB each: REFRESH_PART=1 (tests, bench lines,
# launch list, sweeps) and REFRESH_PART=2 (the ncu --set full captures) are separate calls.
set -u
O=gpurun_out/refresh
mkdir -p $O
PART=${REFRESH_PART:-1}
if [ "$PART" = "1" ]; then
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_mosaic.json 2> $O/bench_mosaic.err
timeout 600 python bench.py --steps 10 --warmup 3 --dist white --no-cpu-baseline > $O/bench_white.json 2> $O/bench_white.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
for CFG in C3 C4 C5; do
  timeout 900 python bench.py --config $CFG --steps 5 --warmup 3 > $O/bench_$CFG.json 2> $O/bench_$CFG.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file $O/launches_mosaic_b32.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-secondary > $O/launches.log 2>&1
fi
if [ "$PART" = "2" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rows3|k_cols3|k_params|k_fft_rows|k_fft_cols" -s 0 -c 7 -f -o $O/prof_mosaic32 \
    python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-secondary > $O/ncu_mosaic.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_deconv_narrow" -s 0 -c 2 -f -o $O/prof_white32 \
    python bench.py --dist white --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-secondary > $O/ncu_white.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_rows3|k_cols3|k_fft_rows|k_fft_cols" -s 0 -c 7 -f -o $O/prof_mosaic_c3 \
    python bench.py --config C3 --batch 8 --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-secondary > $O/ncu_c3.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_rows3|k_cols3|k_fft_rows|k_fft_cols" -s 0 -c 7 -f -o $O/prof_mosaic_c4 \
    python bench.py --config C4 --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-secondary > $O/ncu_c4.log 2>&1
# the reports themselves are too large to travel together: export the raw pages and the per-instruction mix here
for rep in prof_mosaic32 prof_white32 prof_mosaic_c3 prof_mosaic_c4; do
  ncu -i $O/$rep.ncu-rep --page raw --csv > $O/$rep.raw.csv 2>/dev/null
  ncu -i $O/$rep.ncu-rep --page source --csv --print-source sass 2>/dev/null | python tools/ncu_source_mix.py > $O/$rep.sass_mix.md
done
rm -f $O/prof_white32.ncu-rep $O/prof_mosaic_c3.ncu-rep $O/prof_mosaic_c4.ncu-rep      # prof_mosaic32.ncu-rep (~40 MB) comes back whole
fi
if [ "$PART" = "1" ] && [ -z "${REFRESH_QUICK:-}" ]; then
timeout 600 python tools/config_sweep.py > $O/config_sweep.jsonl 2> $O/config_sweep.err
timeout 600 python tools/parity_report.py > $O/parity_report.jsonl 2> $O/parity_report.err
timeout 900 python tools/fuzz_parity.py > $O/fuzz_parity.jsonl 2> $O/fuzz_parity.err
timeout 300 python tools/e2e_probe.py 32 5 > $O/e2e_probe.json 2> $O/e2e_probe.err
timeout 300 python tools/rf_timing.py > $O/rf_timing.txt 2>&1
fi
ls -la $O
