#!/bin/bash
# Re-measures everything profiles/ holds for the current build on one B200 (run through gpurun);
# raw outputs land in gpurun_out/refresh/, tools/ncu_summary.py condenses the .ncu-rep files afterwards.
set -u
O=gpurun_out/refresh
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_mosaic.json 2> $O/bench_mosaic.err
timeout 600 python bench.py --steps 10 --warmup 3 --dist white > $O/bench_white.json 2> $O/bench_white.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file $O/launches_mosaic_b32.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-secondary > $O/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rows2|k_cols2|k_params|k_fft_rows|k_fft_cols" -s 18 -c 6 -f -o $O/prof_mosaic32 \
    python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-secondary > $O/ncu_mosaic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_rows2|k_cols2|k_params|k_deconv_narrow" -s 15 -c 5 -f -o $O/prof_white32 \
    python bench.py --dist white --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline --no-secondary > $O/ncu_white.log 2>&1
if [ -z "${REFRESH_QUICK:-}" ]; then
timeout 600 python tools/config_sweep.py > $O/config_sweep.jsonl 2> $O/config_sweep.err
timeout 600 python tools/parity_report.py > $O/parity_report.jsonl 2> $O/parity_report.err
timeout 900 python tools/fuzz_parity.py > $O/fuzz_parity.jsonl 2> $O/fuzz_parity.err
fi
ls -la $O
