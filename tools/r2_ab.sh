#!/bin/bash
# FFT-engine parity tests, then A/B of the kernel generations on the bench configuration (one gpurun call)
OUT=gpurun_out/${1:-r2ab}; mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x -k "not c4_12000" > $OUT/pytest.log 2>&1; tail -6 $OUT/pytest.log
shift
bash tools/ab_check.sh "$@" 2>&1 | tee $OUT/ab.log
